// TilePipe: per-CTA software pipeline that streams one batch item's samples through shared
// memory in time order (or reverse time order), tile by tile, with 1-D TMA bulk copies.
//
//   HBM --cp.async.bulk (UBLKCP) + mbarrier--> smem stage --> registers (compute) --> same
//   smem stage (in place) --cp.async.bulk.global.shared--> HBM
//
// A stage holds `nbuf` buffers of `tile_len` floats (e.g. the C channel tiles of an item, or
// x and dL/dy tiles in a backward).  With S stages, tile i+S-2 is being loaded while tile i is
// computed and the store of tile i-1 drains, so the kernels' hot loops contain no LDG/STG.
// When rows are not 16-byte aligned (N % 4 != 0) the same stages are filled/drained with
// plain cooperative loads/stores instead (`bulk == false`): a correctness path, not tuned.
//
// Usage (all threads of the CTA execute this; `Rows` maps buffer index -> global row pointers):
//
//   pipe.init(...);
//   pipe.prologue(ntiles, geom, rows);
//   for (int i = 0; i < ntiles; ++i) {
//     pipe.acquire(i, ntiles, geom, rows);        // prefetch tile i+S-2, wait for tile i
//     ... compute on pipe.buf(i % S, b), write results in place ...
//     pipe.release(i, geom, rows);                // fence, sync, store tile i
//   }
//   pipe.drain();
#pragma once

#include "common.cuh"

namespace dasp {

// maps the sequence index of a tile (the order tiles are processed in) to its sample range
struct TileGeom {
  int64_t n;        // samples per row
  int tile_len;     // samples per tile
  int ntiles;       // ceil(n / tile_len)
  bool reverse;     // process tiles from the last to the first (backward sweeps)
  __device__ __forceinline__ int tile_of(int seq) const { return reverse ? (ntiles - 1 - seq) : seq; }
  __device__ __forceinline__ int64_t pos_of(int seq) const { return (int64_t)tile_of(seq) * tile_len; }
  __device__ __forceinline__ int len_of(int seq) const {
    int64_t rem = n - pos_of(seq);
    return rem < tile_len ? (int)rem : tile_len;
  }
};

template <int S>
struct TilePipe {
  static_assert(S >= 2, "need >= 2 stages");
  // prefetch distance and the number of store groups that may still be reading shared memory when the next
  // load is issued: with 3+ stages the stage being refilled drained two iterations ago; with 2 stages it is
  // the stage whose store was committed at the end of the previous iteration, so that store must finish
  // reading first (sub-microsecond, against ~10 us of compute per tile in the kernels that use S = 2)
  static constexpr int kAhead = (S == 2) ? 1 : S - 2;
  static constexpr int kPendingStores = (S == 2) ? 0 : 1;
  uint64_t* full;     // S mbarriers (shared memory)
  float* stages;      // S * nbuf * tile_len floats (shared memory, 128-byte aligned)
  int nbuf;           // buffers per stage
  int tile_len;       // floats per buffer
  bool bulk;          // TMA path (true) or cooperative fallback (false)

  __device__ __forceinline__ float* buf(int stage, int b) const {
    return stages + ((size_t)stage * nbuf + b) * tile_len;
  }

  __device__ __forceinline__ void init(uint64_t* bars, float* stage_mem, int nbuf_, int tile_len_, bool bulk_) {
    full = bars; stages = stage_mem; nbuf = nbuf_; tile_len = tile_len_; bulk = bulk_;
    if (threadIdx.x == 0) {
      for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
      fence_barrier_init();
    }
    __syncthreads();
  }

  template <class Rows>
  __device__ __forceinline__ void issue_load(int seq, const TileGeom& g, const Rows& rows) {
    // thread 0 only
    const int st = seq % S;
    const int64_t pos = g.pos_of(seq);
    const uint32_t bytes = (uint32_t)g.len_of(seq) * 4u;
    mbar_arrive_expect_tx(&full[st], bytes * (uint32_t)nbuf);
    for (int b = 0; b < nbuf; ++b) tma_load_1d(buf(st, b), rows.src(b) + pos, bytes, &full[st]);
  }

  template <class Rows>
  __device__ __forceinline__ void prologue(const TileGeom& g, const Rows& rows) {
    if (bulk && threadIdx.x == 0) {
      for (int j = 0; j < kAhead && j < g.ntiles; ++j) issue_load(j, g, rows);
    }
  }

  template <class Rows>
  __device__ __forceinline__ void acquire(int seq, const TileGeom& g, const Rows& rows) {
    const int st = seq % S;
    if (bulk) {
      if (threadIdx.x == 0) {
        const int j = seq + kAhead;
        if (j < g.ntiles) {
          // stage j % S was last used by tile j - S; all store groups older than the newest
          // kPendingStores have finished reading shared memory after this wait
          tma_store_wait_read<kPendingStores>();
          issue_load(j, g, rows);
        }
      }
      mbar_wait(&full[st], (uint32_t)((seq / S) & 1));
    } else {
      const int64_t pos = g.pos_of(seq);
      const int len = g.len_of(seq);
      for (int b = 0; b < nbuf; ++b) {
        const float* s = rows.src(b) + pos;
        float* d = buf(st, b);
        for (int i = threadIdx.x; i < len; i += blockDim.x) d[i] = s[i];
      }
      __syncthreads();
    }
  }

  // results were written in place into buffers for which rows.dst(b) != nullptr
  template <class Rows>
  __device__ __forceinline__ void release(int seq, const TileGeom& g, const Rows& rows) {
    const int st = seq % S;
    const int64_t pos = g.pos_of(seq);
    const int len = g.len_of(seq);
    if (bulk) {
      fence_proxy_async_smem();      // my generic-proxy smem writes -> visible to the TMA engine
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int b = 0; b < nbuf; ++b) {
          float* d = rows.dst(b);
          if (d) tma_store_1d(d + pos, buf(st, b), (uint32_t)len * 4u);
        }
        tma_store_commit();
      }
    } else {
      __syncthreads();
      for (int b = 0; b < nbuf; ++b) {
        float* d = rows.dst(b);
        if (!d) continue;
        const float* s = buf(st, b);
        for (int i = threadIdx.x; i < len; i += blockDim.x) d[pos + i] = s[i];
      }
      // the stage is next overwritten S-1 iterations later, behind several __syncthreads
    }
  }

  __device__ __forceinline__ void drain() {
    if (bulk && threadIdx.x == 0) tma_store_wait_all<0>();
  }
};

}  // namespace dasp
