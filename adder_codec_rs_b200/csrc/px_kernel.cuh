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
  uint32_t row0;  /* added to every event's y: this plane is a row band of a larger frame (multi-GPU sharding) */
  unsigned long long* counters; /* kCount only: [0] node loads [1] node stores [2] display writes [3] events */
};

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kFlagAggregate = 1u, kFlagPrefix = 2u;

/* level k of pixel i: one 128-bit access, 512 contiguous bytes per warp */
struct GlobalNodes {
  uint4* p; /* &nodes[i] */
  unsigned long long stride;
  uint32_t n_loads, n_stores; /* only read by the counting variant of the kernel */
  __device__ __forceinline__ Node load(uint32_t k) {
    n_loads++;
    const uint4 v = p[(unsigned long long)k * stride];
    Node n;
    n.integ = __uint_as_float(v.x);
    n.dt = __uint_as_float(v.y);
    n.best_dt = __uint_as_float(v.z);
    n.w = v.w;
    return n;
  }
  __device__ __forceinline__ void store(uint32_t k, const Node& n) {
    n_stores++;
    p[(unsigned long long)k * stride] = make_uint4(__float_as_uint(n.integ), __float_as_uint(n.dt), __float_as_uint(n.best_dt), n.w);
  }
};

constexpr uint32_t kSlots = 3; /* events per pixel parked in shared memory */

/*
 * Where a pixel parks its events until the tile's output offset is known.  The first kSlots go to
 * shared memory ([slot][pixel-in-tile]).  A pixel that emits more (a deep pop_best in Normal mode,
 * rare) parks event #e >= kSlots in ITS OWN node column at level e: that level is dead by then —
 * pop_best produces event #e at node k >= e, i.e. after level e has been consumed, and after a
 * pop_best only levels 0 and 1 are written again this frame (length becomes 1, at most 2) — so no
 * extra memory is needed and no live state is touched.  Needs e < depth.
 */
struct EventPark {
  uint32_t* t; /* &slot_t[pixel-in-tile] */
  uint8_t* d;  /* &slot_d[pixel-in-tile] */
  uint4* col;  /* &nodes[i] */
  unsigned long long stride;
  uint32_t tile_px, depth;
  uint32_t n, overflow;
  __device__ __forceinline__ void push(uint32_t dd, uint32_t tt) {
    if (n < kSlots) {
      t[n * tile_px] = tt;
      d[n * tile_px] = (uint8_t)dd;
    } else if (n < depth) {
      col[(unsigned long long)n * stride] = make_uint4(tt, dd, 0u, 0u);
    } else {
      overflow = 1;
      return;
    }
    n++;
  }
  __device__ __forceinline__ void get(uint32_t e, uint32_t& dd, uint32_t& tt) const {
    if (e < kSlots) {
      tt = t[e * tile_px];
      dd = d[e * tile_px];
    } else {
      const uint4 v = col[(unsigned long long)e * stride];
      tt = v.x;
      dd = v.y;
    }
  }
};

__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
}
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

constexpr uint32_t kThreads = ADDER_TILE_PX; /* 256 */
constexpr uint32_t kWarps = kThreads / 32;

__host__ __device__ constexpr uint32_t stage_records(uint32_t R) { return R >= 4 ? 1024u : 512u; }
__host__ __device__ constexpr size_t frame_kernel_smem(uint32_t R) {
  /* slot_t[kSlots][TILE] u32 | stage[records*3] u32 | frame[TILE] u8 | slot_d[kSlots][TILE] u8 */
  return (size_t)kSlots * kThreads * R * 4u + (size_t)stage_records(R) * 12u + (size_t)kThreads * R + (size_t)kSlots * kThreads * R;
}

/*
 * A CTA owns a tile of 256*R consecutive raster indices and walks it as R sub-tiles of 256: in
 * sub-tile r thread t owns pixel tile_start + r*256 + t, so every header / node / sample access of
 * a warp is one contiguous run.  The header and root node of sub-tile r+1 are requested before the
 * state machine of sub-tile r runs (register double buffer).  Events are parked (EventPark), then
 * ONE scan + ONE look-back per tile gives the tile its place in the frame's event stream: the
 * serial look-back chain advances 64 tiles = 64*256*R pixels per L2 round trip, which is why R > 1
 * (at R = 1 the chain, not HBM, bounded the kernel: profiles/r01a).
 * kCount = true is the instrumented twin used (untimed) to measure the algorithmic bytes of a
 * workload: it additionally sums node loads / stores, display writes and events into a.counters.
 */
template <int R, bool kCount>
__global__ void __launch_bounds__(ADDER_TILE_PX) integrate_frame_kernel(const FrameArgs a) {
  constexpr uint32_t TILE = kThreads * R;
  constexpr uint32_t kStage = stage_records(R);
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint32_t* s_slot_t = reinterpret_cast<uint32_t*>(smem_dyn);
  uint32_t* s_stage = s_slot_t + kSlots * TILE;
  uint8_t* s_frame = reinterpret_cast<uint8_t*>(s_stage + kStage * 3u);
  uint8_t* s_slot_d = s_frame + TILE;

  __shared__ uint32_t s_tile, s_prefix, s_total;
  __shared__ uint32_t s_wtot[R * kWarps];

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

  /* tiles are handed out in ticket order so that a tile's predecessors are always already running:
   * the look-back below can then never wait on a CTA that has not been scheduled. */
  if (tid == 0) s_tile = atomicAdd(a.ticket, 1u) - a.ticket_base;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t tile_start = tile * TILE;

  /* ---- frame bytes: 128-bit loads of the tile's samples, staged in shared memory -------------- */
  if (tile_start + TILE <= a.P && ((reinterpret_cast<uintptr_t>(a.frame) & 15u) == 0)) {
    for (uint32_t j = tid; j < TILE / 16; j += kThreads)
      reinterpret_cast<uint4*>(s_frame)[j] = __ldg(reinterpret_cast<const uint4*>(a.frame + tile_start) + j);
  } else {
    for (uint32_t j = tid; j < TILE; j += kThreads)
      if (tile_start + j < a.P) s_frame[j] = a.frame[tile_start + j];
  }

  uint2 h_next = make_uint2(0u, 0u);
  uint4 n_next = make_uint4(0u, 0u, 0u, 0u);
  if (tile_start + tid < a.P) {
    h_next = a.hdr[tile_start + tid];
    n_next = a.nodes[tile_start + tid];
  }
  __syncthreads();

  uint32_t errbits = 0;
  unsigned long long cnt_pack = 0; /* events of this thread's pixel in sub-tile r: byte r */
  unsigned long long c_loads = 0, c_stores = 0, c_disp = 0;
#pragma unroll 1
  for (uint32_t r = 0; r < (uint32_t)R; r++) {
    const uint32_t q = r * kThreads + tid; /* pixel-in-tile */
    const uint32_t i = tile_start + q;
    const uint2 hraw = h_next;
    const uint4 nraw = n_next;
    if (r + 1 < (uint32_t)R && i + kThreads < a.P) { /* next sub-tile's header and root */
      h_next = a.hdr[i + kThreads];
      n_next = a.nodes[i + kThreads];
    }
    if (i < a.P) {
      GlobalNodes mem{a.nodes + i, a.level_stride, 1u, 0u};
      EventPark park{s_slot_t + q, s_slot_d + q, a.nodes + i, a.level_stride, TILE, a.px.depth, 0u, 0u};
      PxHeader h{__uint_as_float(hraw.x), hraw.y};
      Node n0{__uint_as_float(nraw.x), __uint_as_float(nraw.y), __uint_as_float(nraw.z), nraw.w};
      uint8_t disp;
      const bool show = px_step(a.px, s_frame[q], h, n0, mem, park, errbits, &disp);
      a.hdr[i] = make_uint2(__float_as_uint(h.lf), h.y);
      if (show) a.running[i] = disp;
      if (park.overflow) errbits |= ADDER_DEVERR_DEPTH;
      cnt_pack |= (unsigned long long)park.n << (8u * r);
      if (kCount) {
        c_loads += mem.n_loads;
        c_stores += mem.n_stores;
        c_disp += show ? 1u : 0u;
      }
    }
  }
  if (kCount) {
    atomicAdd(&a.counters[0], c_loads);
    atomicAdd(&a.counters[1], c_stores);
    atomicAdd(&a.counters[2], c_disp);
  }
  if (errbits) atomicOr(a.err, errbits);

  /* ---- ordered compaction: tile scan, decoupled look-back across tiles, staged coalesced write -- */
  uint32_t off[R]; /* first record of this thread's pixel in sub-tile r, tile-relative (after the scan) */
#pragma unroll
  for (int r = 0; r < R; r++) {
    const uint32_t nev = (uint32_t)(cnt_pack >> (8 * r)) & 0xFFu;
    uint32_t incl = nev;
#pragma unroll
    for (uint32_t o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wtot[r * kWarps + warp] = incl;
    off[r] = incl - nev;
  }
  __syncthreads();
  if (warp == 0) {
    /* exclusive scan of the R*8 (sub-tile, warp) totals, in pixel order */
    constexpr uint32_t kPer = (R * kWarps + 31) / 32;
    uint32_t mine[kPer], sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < kPer; j++) {
      const uint32_t idx = lane * kPer + j;
      mine[j] = idx < R * kWarps ? s_wtot[idx] : 0u;
      sum += mine[j];
    }
    uint32_t incl = sum;
#pragma unroll
    for (uint32_t o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    const uint32_t tot = __shfl_sync(kFull, incl, 31);
    uint32_t run = incl - sum;
#pragma unroll
    for (uint32_t j = 0; j < kPer; j++) {
      const uint32_t idx = lane * kPer + j;
      if (idx < R * kWarps) s_wtot[idx] = run;
      run += mine[j];
    }

    const unsigned long long tag = (unsigned long long)a.epoch << 2;
    uint32_t excl = 0;
    if (tile == 0) {
      if (lane == 0) st_status(&a.tile_status[0], ((tag | kFlagPrefix) << 32) | tot);
    } else {
      if (lane == 0) st_status(&a.tile_status[tile], ((tag | kFlagAggregate) << 32) | tot);
      int j = (int)tile - 1;
      for (;;) { /* 64 predecessors per round trip: lane l looks at j-l and j-32-l */
        uint32_t flag[2], val[2];
#pragma unroll
        for (int w = 0; w < 2; w++) {
          const int idx = j - 32 * w - (int)lane;
          flag[w] = kFlagPrefix;
          val[w] = 0;
          if (idx >= 0) {
            unsigned long long s;
            uint32_t shi;
            do {
              s = ld_status(&a.tile_status[idx]);
              shi = (uint32_t)(s >> 32);
            } while ((shi >> 2) != a.epoch || (shi & 3u) == 0u);
            flag[w] = shi & 3u;
            val[w] = (uint32_t)s;
          }
        }
        const uint32_t pm0 = __ballot_sync(kFull, flag[0] == kFlagPrefix);
        const uint32_t pm1 = __ballot_sync(kFull, flag[1] == kFlagPrefix);
        uint32_t contrib;
        if (pm0) { /* nearest predecessor holding a prefix is in the first window */
          const uint32_t first = (uint32_t)__ffs((int)pm0) - 1u;
          contrib = lane <= first ? val[0] : 0u;
        } else {
          const uint32_t first = pm1 ? (uint32_t)__ffs((int)pm1) - 1u : 31u;
          contrib = val[0] + (lane <= first ? val[1] : 0u);
        }
#pragma unroll
        for (uint32_t o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(kFull, contrib, o);
        excl += contrib;
        if (pm0 | pm1) break;
        j -= 64;
      }
      if (lane == 0) st_status(&a.tile_status[tile], ((tag | kFlagPrefix) << 32) | (excl + tot));
    }
    if (lane == 0) {
      s_prefix = excl;
      s_total = tot;
      if (kCount) atomicAdd(&a.counters[3], (unsigned long long)tot);
      if (tile == a.n_tiles - 1u) {
        if (a.chunk_off) a.chunk_off[a.n_chunks] = excl + tot;
        if (a.total_events) atomicAdd(a.total_events, (unsigned long long)(excl + tot));
      }
    }
  }
  __syncthreads();
  const uint32_t prefix = s_prefix, total = s_total;
#pragma unroll
  for (int r = 0; r < R; r++) off[r] += s_wtot[r * kWarps + warp];

  if (a.chunk_off) { /* lengths of the reference's Vec<Vec<Event>>, video.rs:677-734 */
    if (a.chunk_px >= TILE) { /* at most one chunk starts inside this tile */
      const uint32_t cb = ((tile_start + a.chunk_px - 1u) / a.chunk_px) * a.chunk_px;
      const uint32_t q = cb - tile_start;
      if (cb < a.P && q < TILE && (q & (kThreads - 1u)) == tid) {
        uint32_t o = 0;
#pragma unroll
        for (int r = 0; r < R; r++)
          if ((q >> 8) == (uint32_t)r) o = off[r];
        a.chunk_off[cb / a.chunk_px] = prefix + o;
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; r++) {
        const uint32_t i = tile_start + r * kThreads + tid;
        if (i < a.P && i % a.chunk_px == 0u) a.chunk_off[i / a.chunk_px] = prefix + off[r];
      }
    }
  }

  if (total == 0u) return;
  for (uint32_t sbase = 0; sbase < total; sbase += kStage) {
#pragma unroll
    for (int r = 0; r < R; r++) {
      const uint32_t nev = (uint32_t)(cnt_pack >> (8 * r)) & 0xFFu;
      if (nev && off[r] + nev > sbase && off[r] < sbase + kStage) {
        const uint32_t q = r * kThreads + tid;
        const uint32_t i = tile_start + q;
        const uint32_t row = i / a.WC, rem = i - row * a.WC;
        const uint32_t x = rem / a.C, c = rem - x * a.C;
        const uint32_t w0 = x | ((row + a.row0) << 16);
        const uint32_t w1 = a.C == 1u ? ADDER_C_NONE : c;
        EventPark park{s_slot_t + q, s_slot_d + q, a.nodes + i, a.level_stride, TILE, a.px.depth, nev, 0u};
        for (uint32_t e = 0; e < nev; e++) {
          const uint32_t li = off[r] + e - sbase; /* unsigned wrap: records before this window fail the test too */
          if (li < kStage) {
            uint32_t dd, tt;
            park.get(e, dd, tt);
            s_stage[li * 3u + 0u] = w0;
            s_stage[li * 3u + 1u] = w1 | (dd << 8);
            s_stage[li * 3u + 2u] = tt;
          }
        }
      }
    }
    __syncthreads();
    const uint32_t n = total - sbase < kStage ? total - sbase : kStage;
    const unsigned long long first = (unsigned long long)prefix + sbase;
    uint32_t can = 0;
    if (first < a.ev_cap) can = (a.ev_cap - first) < n ? (uint32_t)(a.ev_cap - first) : n;
    if (can < n && tid == 0) atomicOr(a.err, ADDER_DEVERR_CAPACITY);
    uint32_t* dst = a.ev_words + first * 3ull;
    for (uint32_t j = tid; j < can * 3u; j += kThreads) dst[j] = s_stage[j];
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
