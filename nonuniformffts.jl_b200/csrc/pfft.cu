// pfft.cu — K-pfft: pruned oversampled-grid FFT fused with deconvolution, mode truncation and zero padding.
//
// Replaces, for complex plans whose oversampled sizes are powers of two (sigma = 2 on power-of-two mode counts — the
// headline configurations), the reference's three full-grid stages
//   type 1:  FFT of the whole oversampled grid (src/NonuniformFFTs.jl:197-203)  +  copy_deconvolve_to_non_oversampled!
//            (:350-414)
//   type 2:  fill_with_zeros (:260-266) + copy_deconvolve_to_oversampled! (:416-480)  +  backward FFT (:293-314)
// by D one-dimensional passes that never touch modes that are thrown away (type 1) or known to be zero (type 2):
//   type 1:  pass along x reads the N~^3 grid and writes only the K_x kept modes, pass along y reads K_x N~ N~ and writes
//            K_x K_y N~, pass along z writes the user's K_x K_y K_z array — 1 + 1/2 + 1/2 + 1/4 + 1/4 + 1/8 = 2.6 grid
//            volumes of HBM traffic instead of >= 6 (cuFFT, 3 passes) + 2 x 1/8 (deconvolution);
//   type 2:  the same passes in reverse with the zero padding done in shared memory.
// 1 / phihat_d[k_d] is applied in the pass along d (the product is separable), the normalisation factor and the
// uniform callbacks in the pass that touches the user's array.  Other plans (real data, sizes with factors 3 or 5) keep
// cuFFT + K-deconv.
//
// One CTA transforms a tile of TW lines of length L held in shared memory (TW consecutive elements of the contiguous
// dimension -> 128-byte global segments for the strided passes): in-place Stockham radix-8 passes, every thread
// keeps its butterflies in registers across the CTA barrier, twiddles from a table computed in double precision, an
// XOR swizzle of the line index keeps the strided butterfly stores free of bank conflicts.
#include <cmath>
#include <vector>
#include "common.cuh"

namespace nufft {

constexpr int PF_MIN_L = 16, PF_MAX_L = 4096;
#ifndef PF_TILE_BYTES
#define PF_TILE_BYTES 32768          // line data per CTA (tuning knob)
#endif
#ifndef PF_MIN_CTAS
#define PF_MIN_CTAS 4                // resident CTAs per SM the register allocation aims at (tuning knob)
#endif

template <typename T> struct PfftArgs {
    using C2 = typename Vec2<T>::type;
    const C2 *in;
    C2 *out;
    int K;                 // kept modes along the transformed dimension
    int64_t n_lo, n_hi;    // product of the (current) sizes of the faster / slower dimensions
    const int32_t *imap;   // kept index -> oversampled index
    const T *iphihat;      // [K] 1 / phihat
    const C2 *tw;          // [L] exp(-2 pi i k / L)
    T scale;               // extra factor (type-1 normalisation on the last pass)
    // uniform callbacks, only on the pass that touches the user's array (always the slowest dimension, n_hi == 1)
    int user_side;
    int K0, K1;            // kept sizes of the faster dimensions (to split `lo` for the separable factors)
    const T *fsep[3];
    const T *fdense;
    // strided passes of the z-slab pipeline (mgpu.cu): the kept side is split into blocks of `blk` kept indices, one per
    // destination rank of the transpose that follows / source rank of the one that preceded: element (lo, I, hi) lives at
    // lo + n_lo ((I % blk) + blk hi) + (I / blk) blk_stride.  blk = 0: plain layout lo + n_lo (I + K hi).
    int blk;
    int64_t blk_stride;
};

template <typename T> __host__ __device__ constexpr int pf_group() { return 128 / (2 * (int)sizeof(T)); }
template <typename T> __host__ __device__ constexpr int pf_tw(int L)
{
    const int maxel = PF_TILE_BYTES / (2 * (int)sizeof(T));       // <= 32 KiB of line data per CTA: 4 CTAs per SM overlap their phases
    return pf_group<T>() < maxel / L ? pf_group<T>() : (maxel / L > 0 ? maxel / L : 1);
}
template <typename T> __host__ __device__ constexpr int pf_nt(int L)
{
    return pf_tw<T>(L) * L / 8 < 256 ? pf_tw<T>(L) * L / 8 : 256;
}
template <typename T> __host__ __device__ constexpr size_t pf_smem(int L)
{
    return (size_t)(pf_tw<T>(L) * L + 2 * L) * 2 * sizeof(T);       // tile (no padding: lines are XOR-swizzled) + twiddles
}

template <typename C2> __device__ __forceinline__ C2 c_add(C2 a, C2 b) { C2 r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C2> __device__ __forceinline__ C2 c_sub(C2 a, C2 b) { C2 r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename C2> __device__ __forceinline__ C2 c_mul(C2 a, C2 b)
{
    C2 r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
template <typename C2> __device__ __forceinline__ C2 c_mul_mi(C2 a) { C2 r; r.x = a.y; r.y = -a.x; return r; }   // a * (-i)

// Float32: packed 2 x f32 instructions of sm_100a (FADD2 / FMUL2 / FFMA2) — one instruction per complex add, two per
// complex multiplication by a twiddle stored as (c, s, -s, c)
__device__ __forceinline__ unsigned long long pf_pk(float2 a)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 pf_unpk(unsigned long long v)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
template <> __device__ __forceinline__ float2 c_add<float2>(float2 a, float2 b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pf_pk(a)), "l"(pf_pk(b)));
    return pf_unpk(d);
}
template <> __device__ __forceinline__ float2 c_sub<float2>(float2 a, float2 b)
{
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pf_pk(a)), "l"(pf_pk(b)));
    return pf_unpk(d);
}
// a * w with w given as wa = (c, s), wb = (-s, c):  (a.x c - a.y s, a.x s + a.y c) = a.x * wa + a.y * wb
__device__ __forceinline__ float2 c_mul_tw(float2 a, float2 wa, float2 wb)
{
    unsigned long long t, d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pf_pk(make_float2(a.y, a.y))), "l"(pf_pk(wb)));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pf_pk(make_float2(a.x, a.x))), "l"(pf_pk(wa)), "l"(t));
    return pf_unpk(d);
}
__device__ __forceinline__ double2 c_mul_tw(double2 a, double2 wa, double2 wb)
{
    double2 r;
    r.x = a.x * wa.x + a.y * wb.x;
    r.y = a.x * wa.y + a.y * wb.y;
    return r;
}

// forward DFT of R points in registers (R = 2, 4, 8), natural output order
template <typename T, int R> __device__ __forceinline__ void dft(typename Vec2<T>::type (&u)[R])
{
    using C2 = typename Vec2<T>::type;
    if constexpr (R == 2) {
        const C2 a = u[0], b = u[1];
        u[0] = c_add(a, b); u[1] = c_sub(a, b);
    } else if constexpr (R == 4) {
        const C2 e0 = c_add(u[0], u[2]), e1 = c_sub(u[0], u[2]), e2 = c_add(u[1], u[3]), e3 = c_mul_mi(c_sub(u[1], u[3]));
        u[0] = c_add(e0, e2); u[2] = c_sub(e0, e2); u[1] = c_add(e1, e3); u[3] = c_sub(e1, e3);
    } else {
        const T h = (T)0.70710678118654752440;
        C2 a[4], b[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { a[t] = c_add(u[t], u[t + 4]); b[t] = c_sub(u[t], u[t + 4]); }
        // b[t] *= w8^t, w8 = exp(-i pi / 4)
        { C2 v = b[1]; b[1].x = h * (v.x + v.y); b[1].y = h * (v.y - v.x); }
        b[2] = c_mul_mi(b[2]);
        { C2 v = b[3]; b[3].x = h * (v.y - v.x); b[3].y = -h * (v.x + v.y); }
        dft<T, 4>(a);
        dft<T, 4>(b);
#pragma unroll
        for (int m = 0; m < 4; ++m) { u[2 * m] = a[m]; u[2 * m + 1] = b[m]; }
    }
}

// per-pass twiddle tables: the pass with Ns = P > 1 and radix R reads w[(t - 1) * P + k] = exp(-2 pi i t k / (P R)),
// t = 1..R-1, k = 0..P-1 (consecutive lanes -> consecutive k: no bank conflicts); tables are concatenated in pass order
__host__ __device__ constexpr int pf_radix(int L, int P) { return L / P >= 8 ? 8 : L / P; }
__host__ __device__ constexpr int pf_tw_offset(int L, int P)
{
    int off = 0;
    for (int p = 1; p < P; p *= pf_radix(L, p))
        if (p > 1) off += (pf_radix(L, p) - 1) * p;
    return off;
}

// XOR swizzle of the position inside a line (G = elements per 128 bytes): bits [lg G, 2 lg G) are folded onto the low
// bits.  It is linear over sums of numbers with disjoint bits — phys(a + b) = phys(a) ^ phys(b) — so the per-access cost
// is one XOR with a compile-time constant.
template <typename T> __host__ __device__ constexpr int pf_phys(int i)
{
    constexpr int G = pf_group<T>();
    return i ^ ((i / G) & (G - 1));
}
// element (line, i) of the tile: lines are L apart (a power of two, no padding) and additionally rotated by the line
// index inside each 128-byte group, so that the same position of 16 different lines (strided global passes) hits 16
// different banks.  line * L + (phys(i) ^ (line % G)): the XOR constants of the butterflies stay below L.
template <typename T, int L> __device__ __forceinline__ int pf_slot(int line, int physpos)
{
    constexpr int G = pf_group<T>();
    return line * L + (physpos ^ (line & (G - 1) & (L - 1)));
}

// one in-place Stockham pass of radix R with Ns = P already transformed
template <typename T, int L, int P, int R> __device__ __forceinline__ void pf_pass(typename Vec2<T>::type *s, const typename Vec2<T>::type *tw, int tid)
{
    using C2 = typename Vec2<T>::type;
    constexpr int TW = pf_tw<T>(L), NT = pf_nt<T>(L);
    constexpr int TL = L / R;                       // butterflies per line
    constexpr int NB = TW * TL / NT;                // butterflies per thread
    static_assert(NB >= 1 && NB * NT == TW * TL, "pfft thread mapping");
    static_assert(NT % TL == 0 || TL % NT == 0, "pfft thread mapping");
    C2 u[NB][R];
    int line[NB], jj[NB];
#pragma unroll
    for (int bi = 0; bi < NB; ++bi) {               // butterfly b = tid + bi * NT -> (line, j) = (b / TL, b % TL)
        if constexpr (NT % TL == 0) { line[bi] = tid / TL + bi * (NT / TL); jj[bi] = tid % TL; }
        else { line[bi] = (bi * NT) / TL; jj[bi] = tid + (bi * NT) % TL; }
    }
#pragma unroll
    for (int bi = 0; bi < NB; ++bi) {
        const int j = jj[bi], k = j & (P - 1);
        const int pa = pf_slot<T, L>(line[bi], pf_phys<T>(j));     // j and t TL occupy disjoint bits below L
#pragma unroll
        for (int t = 0; t < R; ++t) u[bi][t] = s[pa ^ pf_phys<T>(t * TL)];
        if constexpr (P > 1) {
            constexpr int OFF = pf_tw_offset(L, P);
#pragma unroll
            for (int t = 1; t < R; ++t) {
                const C2 *w2 = tw + 2 * (OFF + (t - 1) * P + k);        // (c, s), (-s, c): one 16-byte (Float32) load
                u[bi][t] = c_mul_tw(u[bi][t], w2[0], w2[1]);
            }
        }
        dft<T, R>(u[bi]);
    }
    __syncthreads();
#pragma unroll
    for (int bi = 0; bi < NB; ++bi) {
        const int j = jj[bi], k = j & (P - 1);
        const int pb = pf_slot<T, L>(line[bi], pf_phys<T>((j - k) * R + k));   // (j - k) R, t P and k occupy disjoint bits
#pragma unroll
        for (int t = 0; t < R; ++t) s[pb ^ pf_phys<T>(t * P)] = u[bi][t];
    }
    __syncthreads();
}

template <typename T, int L, int P> __device__ __forceinline__ void pf_run(typename Vec2<T>::type *s, const typename Vec2<T>::type *tw, int tid)
{
    if constexpr (P < L) {
        constexpr int R = pf_radix(L, P);
        pf_pass<T, L, P, R>(s, tw, tid);
        pf_run<T, L, P * R>(s, tw, tid);
    }
}

// uniform callback factor (src/plan.jl:146-164) of kept index I of line `line`: rare path, kept out of line (arguments
// by value so that the kernel parameters never need an address)
template <typename T, bool CONTIG>
__device__ __noinline__ T pf_callback_factor(const T *fdense, const T *f0, const T *f1, const T *f2, int K, int K0, int K1,
                                             int64_t n_lo, int line, int I, int64_t lo0, int64_t hi0)
{
    T f = (T)1;
    const int64_t lo = CONTIG ? 0 : lo0 + line;
    const int64_t lin = CONTIG ? (int64_t)I + (int64_t)K * (hi0 + line) : lo + n_lo * I;
    if (fdense) f *= fdense[lin];
    if (CONTIG) {
        if (f0) f *= f0[I];
    } else if (K1 == 0) {                // 2-D: lo = i0, I = i1
        if (f0) f *= f0[lo];
        if (f1) f *= f1[I];
    } else {                             // 3-D: lo = i0 + K0 i1, I = i2
        if (f0) f *= f0[lo % K0];
        if (f1) f *= f1[lo / K0];
        if (f2) f *= f2[I];
    }
    return f;
}

// FWD: full lines in, kept modes out (type 1).  !FWD: kept modes in, zero padding, full lines out (type 2; computed as
// conj(FFT(conj x)) so that one set of twiddles serves both directions).  CONTIG: lines along the contiguous dimension
// (the tile is TW consecutive lines = one contiguous block of memory); otherwise TW consecutive elements of the
// contiguous dimension times a strided line.
template <typename T, int L, bool FWD, bool CONTIG>
__global__ void __launch_bounds__(pf_nt<T>(L), PF_MIN_CTAS) pfft_pass_kernel(PfftArgs<T> a)
{
    using C2 = typename Vec2<T>::type;
    constexpr int TW = pf_tw<T>(L), NT = pf_nt<T>(L);
    constexpr int EPT = TW * L / NT;                // elements per thread
    static_assert(NT % TW == 0 && (NT % L == 0 || L % NT == 0), "pfft thread mapping");
    extern __shared__ __align__(16) unsigned char pf_raw[];
    C2 *s = reinterpret_cast<C2 *>(pf_raw);         // [TW][L], swizzled (pf_slot)
    C2 *tw = s + TW * L;                            // concatenated per-pass twiddle tables, two entries per twiddle
    const int tid = threadIdx.x;
    const int K = a.K;
    int64_t lo0 = 0, hi0 = 0;
    int nlines;
    if constexpr (CONTIG) {
        hi0 = (int64_t)blockIdx.x * TW;
        nlines = (int)(a.n_hi - hi0 < TW ? a.n_hi - hi0 : TW);
    } else {
        const unsigned tiles_lo = (unsigned)((a.n_lo + TW - 1) / TW);
        hi0 = blockIdx.x / tiles_lo;
        lo0 = (int64_t)(blockIdx.x % tiles_lo) * TW;
        nlines = (int)(a.n_lo - lo0 < TW ? a.n_lo - lo0 : TW);
    }
    for (int i = tid; i < 2 * pf_tw_offset(L, L); i += NT) tw[i] = a.tw[i];

    // deconvolution factor (and uniform callback, on the user's side only) of kept index I of line `line`
    const bool has_cb = a.user_side && (a.fdense || a.fsep[0] || a.fsep[1] || a.fsep[2]);
    auto factor = [&](int line, int I) -> T {
        T f = a.scale * a.iphihat[I];
        if (has_cb) f *= pf_callback_factor<T, CONTIG>(a.fdense, a.fsep[0], a.fsep[1], a.fsep[2], K, a.K0, a.K1, a.n_lo, line, I, lo0, hi0);
        return f;
    };
    // element n of this thread in a full-length tile: (line, i) and its global offset
    //   CONTIG : e = tid + n NT runs over the contiguous block of TW lines
    //   strided: line = tid % TW is fixed, i = tid / TW + n (NT / TW)
    auto full_line = [&](int n) -> int {
        if constexpr (!CONTIG) return tid % TW;
        else if constexpr (NT % L == 0) return tid / L + n * (NT / L);
        else return (n * NT) / L;
    };
    auto full_pos = [&](int n) -> int {          // swizzled position inside the line
        if constexpr (!CONTIG) return pf_phys<T>(tid / TW) ^ pf_phys<T>(n * (NT / TW));
        else if constexpr (NT % L == 0) return pf_phys<T>(tid % L);
        else return pf_phys<T>(tid) ^ pf_phys<T>((n * NT) % L);
    };
    const int64_t full_base = CONTIG ? (int64_t)L * hi0 + tid : lo0 + tid % TW + a.n_lo * (tid / TW + (int64_t)L * hi0);
    const int64_t full_step = CONTIG ? (int64_t)NT : a.n_lo * (NT / TW);
    // kept elements of this thread: CONTIG: e = tid + n NT over the contiguous block of TW x K; strided: as above with K
    const int64_t kept_base = CONTIG ? (int64_t)K * hi0 + tid : lo0 + tid % TW + a.n_lo * (tid / TW + (int64_t)K * hi0);
    auto kept_off = [&](int I, int n) -> int64_t {       // offset of kept element I (= tid / TW + n NT / TW in strided passes)
        if (CONTIG || a.blk == 0) return kept_base + n * full_step;
        const int b = I / a.blk;
        return lo0 + tid % TW + a.n_lo * ((int64_t)(I - b * a.blk) + (int64_t)a.blk * hi0) + (int64_t)b * a.blk_stride;
    };

    // ---- load (all global loads of a thread are issued before the first shared-memory store) ------------------------
    if constexpr (FWD) {
        C2 r[EPT];
        if (nlines == TW) {                          // full tile (the common case): no per-element predicates
#pragma unroll
            for (int n = 0; n < EPT; ++n) r[n] = a.in[full_base + n * full_step];
        } else {
#pragma unroll
            for (int n = 0; n < EPT; ++n) {
                r[n].x = 0; r[n].y = 0;
                if (full_line(n) < nlines) r[n] = a.in[full_base + n * full_step];
            }
        }
#pragma unroll
        for (int n = 0; n < EPT; ++n) s[pf_slot<T, L>(full_line(n), full_pos(n))] = r[n];
    } else {
        C2 z; z.x = 0; z.y = 0;
        for (int e = tid; e < TW * L; e += NT) s[e] = z;
        C2 r[EPT];
        int pos[EPT];
        int line = CONTIG ? tid / K : tid % TW, I = CONTIG ? tid % K : tid / TW;
#pragma unroll
        for (int n = 0; n < EPT; ++n) {
            pos[n] = -1;
            if (line < nlines && I < K) {
                const C2 v = a.in[kept_off(I, n)];
                const T f = factor(line, I);
                r[n].x = v.x * f; r[n].y = -(v.y * f);
                pos[n] = pf_slot<T, L>(line, pf_phys<T>(a.imap[I]));
            }
            if constexpr (CONTIG) { I += NT; while (I >= K) { I -= K; ++line; } }
            else I += NT / TW;
        }
        __syncthreads();
#pragma unroll
        for (int n = 0; n < EPT; ++n) if (pos[n] >= 0) s[pos[n]] = r[n];
    }
    __syncthreads();

    pf_run<T, L, 1>(s, tw, tid);

    // ---- store --------------------------------------------------------------------------------------------------
    if constexpr (FWD) {
        // rolled loop: K <= L kept modes per line, so up to half of the EPT slots of a thread are empty
        int line = CONTIG ? tid / K : tid % TW, I = CONTIG ? tid % K : tid / TW;
#pragma unroll 4
        for (int n = 0; n < EPT; ++n) {
            if (CONTIG ? line >= nlines : I >= K) break;
            if (CONTIG || line < nlines) {
                C2 v = s[pf_slot<T, L>(line, pf_phys<T>(a.imap[I]))];
                const T f = factor(line, I);
                v.x *= f; v.y *= f;
                a.out[kept_off(I, n)] = v;
            }
            if constexpr (CONTIG) { I += NT; while (I >= K) { I -= K; ++line; } }
            else I += NT / TW;
        }
    } else {
        if (nlines == TW) {
#pragma unroll
            for (int n = 0; n < EPT; ++n) {
                C2 v = s[pf_slot<T, L>(full_line(n), full_pos(n))];
                v.y = -v.y;
                a.out[full_base + n * full_step] = v;
            }
        } else {
#pragma unroll
            for (int n = 0; n < EPT; ++n) {
                if (full_line(n) < nlines) {
                    C2 v = s[pf_slot<T, L>(full_line(n), full_pos(n))];
                    v.y = -v.y;
                    a.out[full_base + n * full_step] = v;
                }
            }
        }
    }
}

template <typename T, int L, bool FWD, bool CONTIG> static int pf_launch_LC(const PfftArgs<T> &a, cudaStream_t st)
{
    constexpr int TW = pf_tw<T>(L), NT = pf_nt<T>(L);
    const size_t smem = pf_smem<T>(L);
    auto kern = pfft_pass_kernel<T, L, FWD, CONTIG>;
    static bool attr_done_dev[64] = {};        // per instantiation and device (function attributes are per device)
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    bool &attr_done = attr_done_dev[dev & 63];
    if (!attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_done = true;
    }
    const int64_t tiles = a.n_lo == 1 ? cdiv(a.n_hi, TW) : cdiv(a.n_lo, TW) * a.n_hi;
    if (tiles <= 0) return NUFFT_SUCCESS;
    if (tiles >= ((int64_t)1 << 31)) { set_error("pruned FFT: too many tiles"); return NUFFT_ERR_UNSUPPORTED; }
    kern<<<(unsigned)tiles, NT, smem, st>>>(a);
    NUFFT_COUNT_LAUNCH();
    return NUFFT_SUCCESS;
}

template <typename T, int L, bool FWD> static int pf_launch_L(const PfftArgs<T> &a, cudaStream_t st)
{
    return a.n_lo == 1 ? pf_launch_LC<T, L, FWD, true>(a, st) : pf_launch_LC<T, L, FWD, false>(a, st);
}

template <typename T, bool FWD> static int pf_launch(int L, const PfftArgs<T> &a, cudaStream_t st)
{
    switch (L) {
    case 16: return pf_launch_L<T, 16, FWD>(a, st);
    case 32: return pf_launch_L<T, 32, FWD>(a, st);
    case 64: return pf_launch_L<T, 64, FWD>(a, st);
    case 128: return pf_launch_L<T, 128, FWD>(a, st);
    case 256: return pf_launch_L<T, 256, FWD>(a, st);
    case 512: return pf_launch_L<T, 512, FWD>(a, st);
    case 1024: return pf_launch_L<T, 1024, FWD>(a, st);
    case 2048: return pf_launch_L<T, 2048, FWD>(a, st);
    case 4096: return pf_launch_L<T, 4096, FWD>(a, st);
    }
    set_error("pruned FFT: unsupported length %d", L);
    return NUFFT_ERR_UNSUPPORTED;
}

bool pfft_eligible(const Plan &p)
{
    if (!p.cplx) return false;
    if (const char *e = getenv("NUFFT_B200_PFFT")) if (atoi(e) == 0) return false;
    for (int d = 0; d < p.D; ++d) {
        const int64_t n = p.Nos[d];
        if (n < PF_MIN_L || n > PF_MAX_L || (n & (n - 1)) != 0) return false;
    }
    return true;
}

template <typename T> static int pfft_init_t(Plan &p)
{
    using C2 = typename Vec2<T>::type;
    for (int d = 0; d < p.D; ++d) {
        const int L = (int)p.Nos[d];
        std::vector<C2> h((size_t)2 * L);               // two entries per twiddle: (c, s), (-s, c)
        for (int P = 1; P < L; P *= pf_radix(L, P)) {
            const int R = pf_radix(L, P);
            if (P == 1) continue;
            const int off = pf_tw_offset(L, P);
            for (int t = 1; t < R; ++t)
                for (int k = 0; k < P; ++k) {
                    const double ang = -2.0 * M_PI * (double)t * (double)k / ((double)P * (double)R);
                    const size_t e = 2 * ((size_t)off + (size_t)(t - 1) * P + k);
                    h[e].x = (T)std::cos(ang);
                    h[e].y = (T)std::sin(ang);
                    h[e + 1].x = -h[e].y;
                    h[e + 1].y = h[e].x;
                }
        }
        {   // 1 / phihat_d in precision T
            const T *ph = reinterpret_cast<const T *>(p.h_phihat[d].data());
            std::vector<T> inv((size_t)p.nk[d]);
            for (int64_t k = 0; k < p.nk[d]; ++k) inv[(size_t)k] = (T)1 / ph[k];
            CUDA_TRY(cudaMalloc(&p.d_pf_iph[d], inv.size() * sizeof(T)));
            CUDA_TRY(cudaMemcpy(p.d_pf_iph[d], inv.data(), inv.size() * sizeof(T), cudaMemcpyHostToDevice));
        }
        CUDA_TRY(cudaMalloc(&p.d_pf_tw[d], h.size() * sizeof(C2)));
        CUDA_TRY(cudaMemcpy(p.d_pf_tw[d], h.data(), h.size() * sizeof(C2), cudaMemcpyHostToDevice));
    }
    // scratch A: [K0][N1][N2] (3-D) or [K0][N1] (2-D); the second 3-D intermediate [K0][K1][N2] lives in the grid itself
    size_t na = 0;
    if (p.D == 2) na = (size_t)p.nk[0] * p.Nos[1];
    if (p.D == 3) na = (size_t)p.nk[0] * p.Nos[1] * (p.slab_nz > 0 ? p.slab_nz : p.Nos[2]);
    if (na) CUDA_TRY(cudaMalloc(&p.d_pf_a, na * sizeof(C2)));
    return NUFFT_SUCCESS;
}

int pfft_init(Plan &p) { return p.f64 ? pfft_init_t<double>(p) : pfft_init_t<float>(p); }

void pfft_free(Plan &p)
{
    for (int d = 0; d < 3; ++d) if (p.d_pf_tw[d]) { cudaFree(p.d_pf_tw[d]); p.d_pf_tw[d] = nullptr; }
    for (int d = 0; d < 3; ++d) if (p.d_pf_iph[d]) { cudaFree(p.d_pf_iph[d]); p.d_pf_iph[d] = nullptr; }
    if (p.d_pf_a) { cudaFree(p.d_pf_a); p.d_pf_a = nullptr; }
}

template <typename T> static PfftArgs<T> pf_args(const Plan &p, int d, const nufft_callbacks *cb, bool user_side)
{
    using C2 = typename Vec2<T>::type;
    PfftArgs<T> a{};
    a.K = (int)p.nk[d];
    a.imap = p.d_imap[d];
    a.iphihat = (const T *)p.d_pf_iph[d];
    a.tw = (const C2 *)p.d_pf_tw[d];
    a.scale = (T)1;
    a.user_side = user_side ? 1 : 0;
    a.K0 = (int)p.nk[0];
    a.K1 = p.D == 3 ? (int)p.nk[1] : 0;
    for (int e = 0; e < 3; ++e) a.fsep[e] = (user_side && cb && cb->u_factor_sep && e < p.D) ? (const T *)cb->u_factor_sep[e] : nullptr;
    a.fdense = (user_side && cb) ? (const T *)cb->u_factor_dense : nullptr;
    return a;
}

// type 1: oversampled grid (spreading output) -> user's kept modes
template <typename T> static int pfft_type1_t(Plan &p, void *const uhat[], const nufft_callbacks *cb)
{
    using C2 = typename Vec2<T>::type;
    double nf = 1.0;
    for (int d = 0; d < p.D; ++d) nf *= 2.0 * M_PI / (double)p.Nos[d];     // src/NonuniformFFTs.jl:181
    const int D = p.D;
    for (int c = 0; c < p.C; ++c) {
        C2 *grid = (C2 *)p.d_us + (int64_t)c * p.ncells;
        C2 *A = (C2 *)p.d_pf_a;
        const C2 *src = grid;
        int64_t n_lo = 1;
        for (int d = 0; d < D; ++d) {
            PfftArgs<T> a = pf_args<T>(p, d, cb, d == D - 1);
            C2 *dst = d == D - 1 ? (C2 *)uhat[c] : (d == 0 ? A : grid);
            a.in = src; a.out = dst;
            a.n_lo = n_lo;
            a.n_hi = 1;
            for (int e = d + 1; e < D; ++e) a.n_hi *= p.Nos[e];
            if (d == D - 1) a.scale = (T)nf;
            NUFFT_TRY((pf_launch<T, true>((int)p.Nos[d], a, p.stream)));
            n_lo *= p.nk[d];
            src = dst;
        }
    }
    CUDA_TRY(cudaGetLastError());
    return NUFFT_SUCCESS;
}

// type 2: user's kept modes -> oversampled grid (interpolation input)
template <typename T> static int pfft_type2_t(Plan &p, const void *const uhat[], const nufft_callbacks *cb)
{
    using C2 = typename Vec2<T>::type;
    const int D = p.D;
    for (int c = 0; c < p.C; ++c) {
        C2 *grid = (C2 *)p.d_us + (int64_t)c * p.ncells;
        C2 *A = (C2 *)p.d_pf_a;
        const C2 *src = (const C2 *)uhat[c];
        for (int d = D - 1; d >= 0; --d) {
            PfftArgs<T> a = pf_args<T>(p, d, cb, d == D - 1);
            // intermediates: 3-D: z pass -> grid memory, y pass -> A, x pass -> grid; 2-D: y pass -> A, x pass -> grid
            C2 *dst = d == 0 ? grid : (d == 1 ? A : grid);
            a.in = src; a.out = dst;
            a.n_lo = 1;
            for (int e = 0; e < d; ++e) a.n_lo *= p.nk[e];
            a.n_hi = 1;
            for (int e = d + 1; e < D; ++e) a.n_hi *= p.Nos[e];
            NUFFT_TRY((pf_launch<T, false>((int)p.Nos[d], a, p.stream)));
            src = dst;
        }
    }
    CUDA_TRY(cudaGetLastError());
    return NUFFT_SUCCESS;
}

// one pruned pass along dimension d over an array with n_lo faster / n_hi slower elements (z-slab pipeline, mgpu.cu)
template <typename T>
static int pfft_pass_t(Plan &p, int d, bool fwd, const void *in, void *out, int64_t n_lo, int64_t n_hi, double scale, int blk,
                       int64_t blk_stride)
{
    using C2 = typename Vec2<T>::type;
    PfftArgs<T> a = pf_args<T>(p, d, nullptr, false);
    a.in = (const C2 *)in;
    a.out = (C2 *)out;
    a.n_lo = n_lo;
    a.n_hi = n_hi;
    a.scale = (T)scale;
    a.blk = blk;
    a.blk_stride = blk_stride;
    if (fwd) NUFFT_TRY((pf_launch<T, true>((int)p.Nos[d], a, p.stream)));
    else NUFFT_TRY((pf_launch<T, false>((int)p.Nos[d], a, p.stream)));
    CUDA_TRY(cudaGetLastError());
    return NUFFT_SUCCESS;
}

int pfft_pass(Plan &p, int d, bool fwd, const void *in, void *out, int64_t n_lo, int64_t n_hi, double scale, int blk, int64_t blk_stride)
{
    return p.f64 ? pfft_pass_t<double>(p, d, fwd, in, out, n_lo, n_hi, scale, blk, blk_stride)
                 : pfft_pass_t<float>(p, d, fwd, in, out, n_lo, n_hi, scale, blk, blk_stride);
}

int pfft_type1_run(Plan &p, void *const uhat[], const nufft_callbacks *cb)
{
    return p.f64 ? pfft_type1_t<double>(p, uhat, cb) : pfft_type1_t<float>(p, uhat, cb);
}

int pfft_type2_run(Plan &p, const void *const uhat[], const nufft_callbacks *cb)
{
    return p.f64 ? pfft_type2_t<double>(p, uhat, cb) : pfft_type2_t<float>(p, uhat, cb);
}

}  // namespace nufft
