// viterbi_common.cuh — tile size and back-pointer record layout shared by the Viterbi kernels.
#pragma once
#include <cstdint>

#include "kernels.cuh"

namespace edb {

constexpr int kTile = 16;          // observations per tile (one 128-byte line of an emission row)

// ---- per-tile scratch record ---------------------------------------------------------------------------
// One record per (work item, tile): 32 lanes x 8 bytes of packed back-pointers (16 observations x 4 bits),
// then 16 x 4 bytes of tile maps, one per chain of the warp.  A tile map composes the tile's 16 back-pointer
// steps: nibble e (e = 0..6, or 7 for the reference's "from = -1") holds the state at the observation just
// BEFORE the tile given state e at the tile's last observation.
constexpr int kRecU2 = 40;                 // uint2 per record (32 lanes + 64 bytes of maps)
constexpr int kRecU32 = 2 * kRecU2;
constexpr int kMapOff = 64;                // uint32 offset of the maps inside a record

__device__ __forceinline__ unsigned bp_nibble(const uint2& w, int q) { return ((q < 8 ? w.x : w.y) >> (4 * (q & 7))) & 0xFu; }
__device__ __forceinline__ int chain_tiles(const ChainDesc& cd)
{
    return cd.nobs > 1 ? (int)(((cd.em_off + cd.nobs - 1) >> 4) - ((cd.em_off + 1) >> 4) + 1) : 0;
}
__device__ __forceinline__ int64_t record_base(const ViterbiArgs& a, int chain, int grp, int n_tiles)
{
    return (int64_t)a.bp_tile_base[chain] * a.groups + (int64_t)grp * n_tiles;
}

}  // namespace edb
