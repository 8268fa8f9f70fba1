// window_common.cuh — pieces shared by the column-streaming spreading / interpolation kernels (cs_spread.cuh, cs_interp.cuh):
// the lane map of a warp over the padded (x, y) footprint of a column of 4 x 4 cells, the zero-padded / lane-transposed
// 1-D weight records, and packed 2 x f32 arithmetic.
//
// Fast path for the headline configuration class: D = 3, HalfSupport(4), ComplexF32.  All points of a column touch the same
// padded (x, y) footprint of P x P = 11 x 11 cells, so a warp keeps that footprint in REGISTERS — lane L owns the cells (slots)
//       (x = L % 11, y = L / 11 + 3k), k = 0..3        [121 cells = 32 lanes x 4 slots - 7]
// (the three cells (10, 2), (10, 5), (10, 8) that this map misses are the 4th slot of lanes 22, 23, 24).
// The per-point weights are zero-padded to the footprint: wx_pad[x], wy_pad[y] (x, y in 0..10) vanish outside the
// point's 8 x 8 support, so every lane applies the same code to its fixed slots.
#pragma once
#include "tile_common.cuh"

namespace nufft {
namespace rt {

constexpr int M = 4;
constexpr int W = 8;
constexpr int SB = 4;              // sub-bin edge (cells)
constexpr int P = SB + W - 1;      // padded footprint edge = 11

// packed 2 x f32 arithmetic (sm_100a FFMA2 / FMUL2); pk2(w, w) operands fold into the .F32 broadcast form
using u64 = unsigned long long;
__device__ __forceinline__ u64 pk2(float a, float b)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float2 unpk2(u64 v)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

constexpr int OFF_WY = 16;            // float offset of the lane-transposed wy rows inside a point record (store_y)

struct LaneSlots {
    int x;        // x of slots 0..2 (and of slot 3 unless special)
    int g;        // y of slot k (k < 3) = g + 3k
    int x3, y3;   // slot 3
    int row;      // row of wyT this lane reads
    bool has3;    // slot 3 exists
};

__device__ __forceinline__ LaneSlots lane_slots(int lane)
{
    LaneSlots s;
    s.g = lane / P;
    s.x = lane - P * s.g;
    if (s.g < 2) { s.x3 = s.x; s.y3 = s.g + 9; s.row = s.g; s.has3 = true; }
    else if (lane < 25) { s.x3 = P - 1; s.y3 = 2 + 3 * (lane - 22); s.row = 3 + (lane - 22); s.has3 = true; }
    else { s.x3 = s.x; s.y3 = s.g + 6; s.row = 2; s.has3 = false; }     // dummy slot: weight 0 (row2.w == 0)
    return s;
}

// p[j] = w[j - o] for 0 <= j - o < 8, else 0   (o in 0..3): two-stage barrel shifter, registers only
__device__ __forceinline__ void pad_shift(const float (&w)[W], int o, float (&p)[P])
{
    float t[9];
    const bool b0 = (o & 1) != 0, b1 = (o & 2) != 0;
    t[0] = b0 ? 0.f : w[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) t[j] = b0 ? w[j - 1] : w[j];
    t[8] = b0 ? w[7] : 0.f;
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const float lo = (j >= 2) ? t[j - 2] : 0.f;
        const float hi = (j <= 8) ? t[j] : 0.f;
        p[j] = b1 ? lo : hi;
    }
}

// store wy_pad as the six float4 rows of the record
__device__ __forceinline__ void store_y(float *rec, const float (&py)[P])
{
    float4 *r = reinterpret_cast<float4 *>(rec + OFF_WY);
    r[0] = make_float4(py[0], py[3], py[6], py[9]);
    r[1] = make_float4(py[1], py[4], py[7], py[10]);
    r[2] = make_float4(py[2], py[5], py[8], 0.f);
    r[3] = make_float4(py[2], py[5], py[8], py[2]);
    r[4] = make_float4(py[2], py[5], py[8], py[5]);
    r[5] = make_float4(py[2], py[5], py[8], py[8]);
}

}  // namespace rt
}  // namespace nufft
