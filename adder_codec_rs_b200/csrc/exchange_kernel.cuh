/*
 * exchange_kernel.cuh — the one exchange step of the row-band sharding: when a single downstream
 * consumer needs a whole frame's events in order (the reference feeds its serial encoder that way,
 * adder-codec-rs/src/transcoder/source/video.rs:736-740; the compressed encoder needs whole ADUs in raster order,
 * adder-codec-core/src/codec/compressed/stream.rs:264-313), every band delivers its compacted records straight into its
 * place in the consumer GPU's frame buffer over NVLink.
 *
 * No staging and no host in the loop.  The consumer's buffers are peer-mapped into every producer (CUDA IPC between
 * processes, plain pointers inside one process).  Band g's push kernel
 *   1. waits until the consumer has released the ring slot of this frame (one word, ld.acquire.sys),
 *   2. publishes its event total for the frame in the consumer's table (st.release.sys),
 *   3. reads the totals of bands 0..g-1 from that table — an inter-GPU look-back of g words — which gives the band's
 *      offset in the frame's stream (rank order == raster order),
 *   4. stores its records at that offset with 128-bit stores whenever source and destination agree modulo 16 bytes
 *      (records are 12 bytes, so the copy is done on the 32-bit word stream), rebases its chunk offsets, and
 *   5. the last CTA to finish fences at system scope and adds one to the slot's arrival counter.
 * The consumer's wait kernel returns when all bands have arrived.  Publishing comes before waiting, so bands never wait
 * for each other in a cycle; a band only needs the producers of lower rank to have launched their push for the frame.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef ADDER_PUSH_UNROLL
#define ADDER_PUSH_UNROLL 8 /* 128-bit stores whose loads a thread of the push kernel keeps in flight */
#endif

namespace adder {

struct ExchangeRing { /* lives in the consumer's memory; all producers see the same addresses through their mappings */
  uint32_t* ev_words;             /* [slots][out_stride * 3] */
  uint32_t* chunk_off;            /* [slots][total_chunks + 1] */
  unsigned long long* totals;     /* [slots][world]: (seq + 1) << 32 | the band's event total of frame seq */
  unsigned long long* arrived;    /* [slots]: bands that have delivered the slot's current frame (reset when the consumer releases it) */
  unsigned long long* released;   /* [1]: frames the consumer has released (slot of frame s is free when released + slots > s) */
  uint32_t slots, world, total_chunks;
  unsigned long long out_stride; /* records per slot */
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
/* A peer that never shows up must not hang the GPU: every wait gives up after this long and raises ADDER_DEVERR_INTERNAL. */
constexpr unsigned long long kExchangeTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct PushArgs {
  ExchangeRing ring;
  const uint32_t* ev_words;  /* this band's records, frame f at + f * ev_stride * 3 */
  unsigned long long ev_stride;
  const uint32_t* chunk_off; /* this band's offsets, frame f at + f * (n_chunks + 1) */
  uint32_t n_chunks, chunk0, band, n_frames;
  unsigned long long seq0;
  uint32_t* local_done; /* [n_frames] zeroed: CTAs of this launch that finished frame f */
  uint32_t* err;        /* ADDER_DEVERR_CAPACITY when a frame does not fit the slot */
};

/* at most 64 registers: a push CTA has to fit the slot an integrate CTA (64 registers x 256 threads) leaves free (launch_grid) */
__global__ void __launch_bounds__(256, 4) exchange_push_kernel(const PushArgs a) {
  __shared__ unsigned long long s_prefix;
  __shared__ uint32_t s_total, s_ok;
  __shared__ unsigned long long s_part[256];
  const ExchangeRing& g = a.ring;
  /* The frames of a push are independent of each other, and a frame of a sparse stream is a few waits over NVLink and
   * a small copy: latency, not bandwidth.  So the CTAs split into as many groups as there are frames (or one CTA per
   * frame when there are fewer CTAs): group `way` takes frames way, way + ways, ...; `rank` of `gsize` is this CTA's
   * share of a frame's copy.  (A frame at a time with all CTAs cost 35 % of the step at 8 GPUs, profiles/r02g n8.) */
  const uint32_t ways = a.n_frames < gridDim.x ? a.n_frames : gridDim.x;
  const uint32_t way = blockIdx.x % ways, rank = blockIdx.x / ways;
  const uint32_t gsize = (gridDim.x - way + ways - 1u) / ways;
  for (uint32_t f = way; f < a.n_frames; f += ways) {
    const unsigned long long seq = a.seq0 + f;
    const uint32_t slot = (uint32_t)(seq % g.slots);
    const uint32_t* off = a.chunk_off + (unsigned long long)f * (a.n_chunks + 1u);
    const unsigned long long t0 = global_ns();
    if (threadIdx.x == 0) {
      bool ok = true;
      while (ok && ld_acquire_sys(g.released) + g.slots <= seq) { /* the slot's previous frame is still being read */
        __nanosleep(200);
        ok = global_ns() - t0 < kExchangeTimeoutNs;
      }
      const uint32_t total = off[a.n_chunks];
      if (ok && rank == 0u) st_release_sys(g.totals + (unsigned long long)slot * g.world + a.band, ((seq + 1ull) << 32) | total);
      s_total = total;
      s_ok = ok ? 1u : 0u;
    }
    /* look-back over the lower bands' totals, one thread per band (the polls overlap) */
    unsigned long long mine = 0;
    bool ok_t = true;
    for (uint32_t b = threadIdx.x; b < a.band; b += blockDim.x) {
      unsigned long long t = 0;
      while (ok_t && ((t = ld_acquire_sys(g.totals + (unsigned long long)slot * g.world + b)) >> 32) != seq + 1ull) {
        __nanosleep(100);
        ok_t = global_ns() - t0 < kExchangeTimeoutNs;
      }
      mine += (uint32_t)t;
    }
    s_part[threadIdx.x] = mine;
    const int all_ok = __syncthreads_and(ok_t ? 1 : 0);
    if (threadIdx.x == 0) {
      unsigned long long prefix = 0;
      const uint32_t np = a.band < blockDim.x ? a.band : blockDim.x;
      for (uint32_t b = 0; b < np; b++) prefix += s_part[b];
      s_prefix = prefix;
      if (!all_ok) s_ok = 0u;
    }
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    const uint32_t total = s_total;
    if (!s_ok) {
      if (threadIdx.x == 0 && rank == 0u) atomicOr(a.err, 4u /* ADDER_DEVERR_INTERNAL: a peer did not show up */);
    } else if (prefix + total > g.out_stride) {
      if (threadIdx.x == 0 && rank == 0u) atomicOr(a.err, 1u /* ADDER_DEVERR_CAPACITY */);
    } else {
      /* the band's words [0, 3 * total) -> the slot's words [3 * prefix, ...): this CTA's share, 128-bit stores in the body */
      const uint32_t* src = a.ev_words + (unsigned long long)f * a.ev_stride * 3ull;
      uint32_t* dst = g.ev_words + (unsigned long long)slot * g.out_stride * 3ull + prefix * 3ull;
      const unsigned long long n_words = 3ull * total;
      const unsigned long long head = n_words ? ((4ull - ((reinterpret_cast<uintptr_t>(dst) >> 2) & 3ull)) & 3ull) : 0ull; /* words up to dst's 16-byte boundary */
      const unsigned long long h = head < n_words ? head : n_words;
      const unsigned long long n_vec = (n_words - h) >> 2;
      if (rank == 0u && threadIdx.x < h) dst[threadIdx.x] = src[threadIdx.x];
      uint4* dst4 = reinterpret_cast<uint4*>(dst + h);
      const uint32_t* s4 = src + h;
      /* the source is only word-aligned relative to dst: four 32-bit loads per 128-bit store; kPushUnroll stores' worth of
       * loads are in flight per thread (a lone load-then-store chain moved 200 GB/s over NVLink, profiles/r02g n2; four in
       * flight 167 GB/s from ONE band's 16 CTAs on a dense stream, profiles/r02u n2 — the copy is bound by bytes in flight) */
      constexpr int kPushUnroll = ADDER_PUSH_UNROLL;
      const unsigned long long T = (unsigned long long)gsize * blockDim.x;
      unsigned long long i = (unsigned long long)rank * blockDim.x + threadIdx.x;
      for (; i + (unsigned long long)(kPushUnroll - 1) * T < n_vec; i += (unsigned long long)kPushUnroll * T) {
        uint4 v[kPushUnroll];
#pragma unroll
        for (int u = 0; u < kPushUnroll; u++) {
          const uint32_t* p = s4 + 4ull * (i + u * T);
          v[u] = make_uint4(__ldcs(p), __ldcs(p + 1), __ldcs(p + 2), __ldcs(p + 3));
        }
#pragma unroll
        for (int u = 0; u < kPushUnroll; u++) dst4[i + u * T] = v[u];
      }
      for (; i < n_vec; i += T) {
        const uint32_t* p = s4 + 4ull * i;
        dst4[i] = make_uint4(p[0], p[1], p[2], p[3]);
      }
      const unsigned long long tail0 = h + 4ull * n_vec;
      if (rank == 0u && tail0 + threadIdx.x < n_words) dst[tail0 + threadIdx.x] = src[tail0 + threadIdx.x];
      /* chunk offsets of the whole frame: this band's rows, rebased; the last band closes the table */
      uint32_t* goff = g.chunk_off + (unsigned long long)slot * (g.total_chunks + 1u) + a.chunk0;
      for (uint32_t k = rank * blockDim.x + threadIdx.x; k < a.n_chunks; k += gsize * blockDim.x) goff[k] = (uint32_t)prefix + off[k];
      if (a.band + 1u == g.world && rank == 0u && threadIdx.x == 0) goff[a.n_chunks] = (uint32_t)prefix + total;
    }
    /* arrival: the last CTA of the frame's group to finish tells the consumer */
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (atomicAdd(a.local_done + f, 1u) + 1u == gsize) {
        __threadfence_system();
        atomicAdd_system(g.arrived + slot, 1ull);
      }
    }
    __syncthreads();
  }
}

/* Consumer: returns when every band has delivered frames seq0 .. seq0 + n - 1 (arrived[slot] counts bands since the slot
 * was released).  One thread. */
__global__ void exchange_wait_kernel(ExchangeRing g, unsigned long long seq0, uint32_t n, uint32_t* err) {
  const unsigned long long t0 = global_ns();
  for (uint32_t f = 0; f < n; f++) {
    const uint32_t slot = (uint32_t)((seq0 + f) % g.slots);
    while (ld_acquire_sys(g.arrived + slot) < g.world) {
      __nanosleep(200);
      if (global_ns() - t0 >= kExchangeTimeoutNs) {
        atomicOr(err, 4u);
        return;
      }
    }
  }
}

/* Consumer: frames up to (not including) `upto` have been read: their slots may be overwritten. */
__global__ void exchange_release_kernel(ExchangeRing g, unsigned long long from, unsigned long long upto) {
  for (unsigned long long s = from; s < upto; s++) g.arrived[s % g.slots] = 0ull;
  __threadfence_system();
  st_release_sys(g.released, upto);
}

}  // namespace adder
