// rt_common.cuh — pieces shared by the register-tile spreading / interpolation kernels (rt_spread.cuh, rt_interp.cuh).
//
// Fast path for the headline configuration class: D = 3, HalfSupport(4), Float32 (real or complex data).
// A bin (<= 16 cells along z, multiples of 4 cells) is refined into COLUMNS of 4 x 4 cells; set_points orders the
// points by (bin, column, z block of 4 cells).  All points of a column touch the same padded (x, y) footprint of
// P x P = 11 x 11 cells, so a warp keeps that footprint in REGISTERS — lane L owns the cells (slots)
//       (x = L % 11, y = L / 11 + 3k), k = 0..3        [121 cells = 32 lanes x 4 slots - 7]
// (the three cells (10, 2), (10, 5), (10, 8) that this map misses are the 4th slot of lanes 22, 23, 24) — for a
// stack of z planes, and shared memory is touched once per column instead of once per point:
//   spreading      accumulate points into registers, read-modify-write the tile when the column changes;
//   interpolation  load the column's cells once, then every point is a register dot product.
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
constexpr int NPL = 6;             // z planes per consumer warp in spreading (tile height <= 24)

// Per-point record in shared memory (floats):
//   [ 0..15]  s[4]      float4 per z residue class r = tile z mod 4 (spreading: value x the two wz of the class;
//                        interpolation: [0..7] = wz[0..7])
//   [16..39]  wyT[6]    float4 rows of wy_pad: row0 = (p0,p3,p6,p9) row1 = (p1,p4,p7,p10) row2 = (p2,p5,p8,0)
//                        row3..5 = (p2,p5,p8, p2|p5|p8)   (lanes 22..24: 4th slot at x = 10)
//   [40..50]  wx_pad[11]
//   [51]      meta      bytes: column x, column y, local z start, 0
// 52 floats = 208 bytes: consecutive records start 80 banks apart -> 16-byte stores of 8 threads are conflict-free.
constexpr int REC_F = RT_REC_F;
constexpr int OFF_S = 0, OFF_WY = 16, OFF_WX = 40, OFF_META = 51;
static_assert(REC_F == 52, "record layout");

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
