/*
 * px_kernel.cuh — the framed→ADΔER per-pixel integrate / fire / pop kernel for sm_100a.
 *
 * One thread owns one pixel-channel for one frame (the state machine is px_machine.cuh); a CTA owns
 * a tile of 256 consecutive raster indices, i.e. what one chunk iteration of the reference's rayon
 * loop (video.rs:697-731) does for 256 pixels.
 *
 * Events must come out in the reference's order (pixels in raster order, a pixel's events contiguous
 * in push order).  A thread cannot know its output offset before it has run the state machine, so
 * it parks its (d,t) pairs in a shared-memory scratch, the CTA scans the per-thread counts, a
 * decoupled look-back over per-tile status words (one 64-bit word per tile: epoch|flag|count) turns
 * the tile aggregate into a frame-wide exclusive offset in the same pass, and the CTA then writes
 * its records through a shared staging buffer with fully coalesced 32-bit stores.  The same offsets
 * give the per-chunk lengths of the reference's Vec<Vec<Event>>.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "px_machine.cuh"
#include "state_layout.h"

namespace adder {

struct FrameArgs {
  PxParams px;
  const uint8_t* frame; /* P bytes, raster (y,x,c) */
  uint2* hdr;
  uint4* nodes;
  unsigned long long level_stride; /* uint4 elements between levels */
  uint8_t* running;
  uint32_t* ev_words;        /* output records as 3 u32 words each */
  unsigned long long ev_cap; /* records */
  uint32_t* chunk_off;       /* n_chunks+1 exclusive offsets, or null */
  unsigned long long* tile_status;
  uint32_t* ticket;
  uint32_t* err;
  unsigned long long* total_events; /* cumulative counter, or null */
  uint32_t ticket_base, epoch;
  uint32_t P, n_tiles, C, WC, chunk_px, n_chunks;
  uint32_t ecap; /* event slots per thread in the scratch */
};

constexpr uint32_t kStageRecords = 512;
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kFlagAggregate = 1u, kFlagPrefix = 2u;

/* level k of pixel i: one 128-bit access, 512 contiguous bytes per warp */
struct GlobalNodes {
  uint4* p; /* &nodes[i] */
  unsigned long long stride;
  __device__ __forceinline__ Node load(uint32_t k) const {
    const uint4 v = p[(unsigned long long)k * stride];
    Node n;
    n.integ = __uint_as_float(v.x);
    n.dt = __uint_as_float(v.y);
    n.best_dt = __uint_as_float(v.z);
    n.w = v.w;
    return n;
  }
  __device__ __forceinline__ void store(uint32_t k, const Node& n) const {
    p[(unsigned long long)k * stride] = make_uint4(__float_as_uint(n.integ), __float_as_uint(n.dt), __float_as_uint(n.best_dt), n.w);
  }
};

/* per-thread event scratch in shared memory: slot s of thread t at [s][t] */
struct SmemSink {
  uint32_t* t;
  uint8_t* d;
  uint32_t n, cap, overflow;
  __device__ __forceinline__ void push(uint32_t dd, uint32_t tt) {
    if (n < cap) {
      t[n * ADDER_TILE_PX] = tt;
      d[n * ADDER_TILE_PX] = (uint8_t)dd;
      n++;
    } else {
      overflow = 1;
    }
  }
};

__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
}
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

__global__ void __launch_bounds__(ADDER_TILE_PX) integrate_frame_kernel(const FrameArgs a) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  /* dynamic: ev_t[ecap][256] u32 | stage[kStageRecords*3] u32 | ev_d[ecap][256] u8 */
  uint32_t* s_ev_t = reinterpret_cast<uint32_t*>(smem_dyn);
  uint32_t* s_stage = s_ev_t + a.ecap * ADDER_TILE_PX;
  uint8_t* s_ev_d = reinterpret_cast<uint8_t*>(s_stage + kStageRecords * 3u);

  __shared__ __align__(16) uint8_t s_frame[ADDER_TILE_PX];
  __shared__ uint32_t s_tile, s_prefix, s_total;
  __shared__ uint32_t s_warp_off[ADDER_TILE_PX / 32];

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

  /* tiles are handed out in ticket order so that a tile's predecessors are always already running:
   * the look-back below can then never wait on a CTA that has not been scheduled. */
  if (tid == 0) s_tile = atomicAdd(a.ticket, 1u) - a.ticket_base;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t tile_start = tile * ADDER_TILE_PX;
  const uint32_t i = tile_start + tid;
  const bool live = i < a.P;

  /* ---- frame bytes: 128-bit loads of the tile's 256 samples, staged in shared memory ---------- */
  if (tile_start + ADDER_TILE_PX <= a.P && ((reinterpret_cast<uintptr_t>(a.frame) & 15u) == 0)) {
    if (tid < ADDER_TILE_PX / 16)
      reinterpret_cast<uint4*>(s_frame)[tid] = __ldg(reinterpret_cast<const uint4*>(a.frame + tile_start) + tid);
  } else if (live) {
    s_frame[tid] = a.frame[i];
  }

  uint2 hraw = make_uint2(0u, 0u);
  GlobalNodes mem{a.nodes + i, a.level_stride};
  Node n0 = {0.0f, 0.0f, 0.0f, 0u};
  if (live) {
    hraw = a.hdr[i];
    n0 = mem.load(0);
  }
  __syncthreads();

  SmemSink sink{s_ev_t + tid, s_ev_d + tid, 0u, a.ecap, 0u};
  uint32_t errbits = 0;
  if (live) {
    PxHeader h{__uint_as_float(hraw.x), hraw.y};
    uint8_t disp;
    const bool show = px_step(a.px, s_frame[tid], h, n0, mem, sink, errbits, &disp);
    a.hdr[i] = make_uint2(__float_as_uint(h.lf), h.y);
    if (show) a.running[i] = disp;
    if (sink.overflow) errbits |= ADDER_DEVERR_DEPTH;
  }
  const uint32_t nev = sink.n;

  /* ---- ordered compaction: CTA scan, decoupled look-back across tiles, staged coalesced write -- */
  uint32_t incl = nev;
#pragma unroll
  for (uint32_t o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp_off[warp] = incl;
  if (errbits) atomicOr(a.err, errbits);
  __syncthreads();
  if (warp == 0) {
    const uint32_t wt = lane < ADDER_TILE_PX / 32 ? s_warp_off[lane] : 0u;
    uint32_t wincl = wt;
#pragma unroll
    for (uint32_t o = 1; o < ADDER_TILE_PX / 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(kFull, wincl, o);
      if (lane >= o) wincl += t;
    }
    const uint32_t tot = __shfl_sync(kFull, wincl, ADDER_TILE_PX / 32 - 1);
    if (lane < ADDER_TILE_PX / 32) s_warp_off[lane] = wincl - wt;
    const unsigned long long tag = (unsigned long long)a.epoch << 2;
    uint32_t excl = 0;
    if (tile == 0) {
      if (lane == 0) st_status(&a.tile_status[0], ((tag | kFlagPrefix) << 32) | tot);
    } else {
      if (lane == 0) st_status(&a.tile_status[tile], ((tag | kFlagAggregate) << 32) | tot);
      int j = (int)tile - 1;
      for (;;) {
        const int idx = j - (int)lane;
        uint32_t flag = kFlagPrefix, val = 0;
        if (idx >= 0) {
          unsigned long long s;
          uint32_t shi;
          do {
            s = ld_status(&a.tile_status[idx]);
            shi = (uint32_t)(s >> 32);
          } while ((shi >> 2) != a.epoch || (shi & 3u) == 0u);
          flag = shi & 3u;
          val = (uint32_t)s;
        }
        const uint32_t pm = __ballot_sync(kFull, flag == kFlagPrefix);
        const uint32_t first = pm ? (uint32_t)__ffs((int)pm) - 1u : 31u; /* nearest predecessor holding a prefix */
        uint32_t contrib = lane <= first ? val : 0u;
#pragma unroll
        for (uint32_t o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(kFull, contrib, o);
        excl += contrib;
        if (pm) break;
        j -= 32;
      }
      if (lane == 0) st_status(&a.tile_status[tile], ((tag | kFlagPrefix) << 32) | (excl + tot));
    }
    if (lane == 0) {
      s_prefix = excl;
      s_total = tot;
      if (tile == a.n_tiles - 1u) {
        if (a.chunk_off) a.chunk_off[a.n_chunks] = excl + tot;
        if (a.total_events) atomicAdd(a.total_events, (unsigned long long)(excl + tot));
      }
    }
  }
  __syncthreads();
  const uint32_t prefix = s_prefix, total = s_total;
  const uint32_t off = s_warp_off[warp] + incl - nev; /* this pixel's first record, CTA-relative */

  if (a.chunk_off && live) { /* lengths of the reference's Vec<Vec<Event>>, video.rs:677-734 */
    bool boundary;
    if (a.chunk_px >= ADDER_TILE_PX) {
      const uint32_t cb = ((tile_start + a.chunk_px - 1u) / a.chunk_px) * a.chunk_px;
      boundary = i == cb;
    } else {
      boundary = i % a.chunk_px == 0u;
    }
    if (boundary) a.chunk_off[i / a.chunk_px] = prefix + off;
  }

  if (total == 0u) return;
  uint32_t w0 = 0, w1 = 0;
  if (nev) {
    const uint32_t row = i / a.WC, rem = i - row * a.WC;
    const uint32_t x = rem / a.C, c = rem - x * a.C;
    w0 = x | (row << 16);
    w1 = a.C == 1u ? ADDER_C_NONE : c;
  }
  for (uint32_t sbase = 0; sbase < total; sbase += kStageRecords) {
    for (uint32_t s = 0; s < nev; s++) {
      const uint32_t li = off + s - sbase;
      if (li < kStageRecords) {
        s_stage[li * 3u + 0u] = w0;
        s_stage[li * 3u + 1u] = w1 | ((uint32_t)s_ev_d[s * ADDER_TILE_PX + tid] << 8);
        s_stage[li * 3u + 2u] = s_ev_t[s * ADDER_TILE_PX + tid];
      }
    }
    __syncthreads();
    const uint32_t n = total - sbase < kStageRecords ? total - sbase : kStageRecords;
    const unsigned long long first = (unsigned long long)prefix + sbase;
    uint32_t can = 0;
    if (first < a.ev_cap) can = (a.ev_cap - first) < n ? (uint32_t)(a.ev_cap - first) : n;
    if (can < n && tid == 0) atomicOr(a.err, ADDER_DEVERR_CAPACITY);
    uint32_t* dst = a.ev_words + first * 3ull;
    for (uint32_t j = tid; j < can * 3u; j += ADDER_TILE_PX) dst[j] = s_stage[j];
    __syncthreads();
  }
}

/* ---- small state kernels ---------------------------------------------------------------------- */

/* Video::new (video.rs:364-382): every pixel = PixelArena::new(1.0, coord), event_pixel_tree.rs:69-87 */
__global__ void init_state_kernel(uint2* hdr, uint4* level0, uint8_t* running, uint32_t P) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  hdr[i] = make_uint2(0u, HDR_PACK(0, 10, 1, 1, 0, 0));
  level0[i] = make_uint4(0u, 0u, 0u, NODE_PACK(0 /* get_d(1.0) */, 0, 0));
  running[i] = 0;
}

/* update_crf / update_quality_manual (video.rs:1241-1287): c_thresh = baseline, counter = 0;
 * c_thresh_pos (:445-455): c_thresh only. */
__global__ void reset_c_kernel(uint2* hdr, uint32_t P, uint32_t c, int reset_counter) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  uint32_t y = hdr[i].y;
  y = (y & ~0xFF00u) | (c << 8);
  if (reset_counter) y &= ~0xFF0000u;
  hdr[i].y = y;
}

/* handle_roi (video.rs:865-881) / feature radius reset (:1089-1104): c_thresh over a rectangle */
__global__ void rect_c_kernel(uint2* hdr, uint32_t W, uint32_t C, uint32_t x0, uint32_t y0, uint32_t rw, uint32_t rh, uint32_t c) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t per_row = rw * C;
  if (t >= per_row * rh) return;
  uint32_t ry = t / per_row, r = t - ry * per_row;
  uint32_t i = ((y0 + ry) * W + x0) * C + r;
  uint32_t y = hdr[i].y;
  hdr[i].y = (y & ~0xFF00u) | (c << 8);
}

/* set_initial_d (video.rs:780-801) */
__global__ void set_initial_d_kernel(uint2* hdr, uint4* level0, const uint8_t* frame, uint32_t P) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  uint32_t v = frame[i];
  uint32_t d = v == 0u ? ADDER_D_ZERO_INTEGRATION : 31u - __clz(v); /* floor(log2(v)) */
  level0[i].w = (level0[i].w & ~0xFFu) | d;
  hdr[i].y = (hdr[i].y & ~0xFFu) | v;
}

}  // namespace adder
