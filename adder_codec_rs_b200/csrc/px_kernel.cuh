/*
 * px_kernel.cuh — the framed→ADΔER per-pixel integrate / fire / pop kernel for sm_100a.
 *
 * One thread owns one pixel-channel for one frame (the state machine is px_machine.cuh).  A CTA of
 * 256 threads is persistent: it takes tiles of 256*R consecutive raster indices by ticket and walks
 * each as R sub-tiles of 256, so every header / node / sample access of a warp is one contiguous run
 * — what one chunk iteration of the reference's rayon loop (video.rs:697-731) does for 256*R pixels.
 *
 * Events must come out in the reference's order (pixels in raster order, a pixel's events contiguous
 * in push order).  A thread cannot know its output offset before it has run the state machine, so
 * it parks its (d,t) pairs in shared memory; warp scans give every pixel its place in the tile; a
 * decoupled look-back over per-tile status words (one 64-bit word per tile: epoch|flag|count) turns
 * the tile aggregate into a frame-wide exclusive offset in the same pass.
 *
 * The CTA is a three-stage software pipeline over its tiles (parks are triple-buffered):
 *   iteration n:  all warps   run the state machines of tile n            -> park[n%3], warp totals
 *                 warp 0      scans the warp totals, publishes tile n's aggregate, draws ticket n+2
 *                 warp 1      looks back for tile n-1 (its predecessors published long ago)
 *                 all warps   write out tile n-2 (prefix known since iteration n-1); each thread
 *                             stores its own records, a warp's records are contiguous in the output
 * with ONE rendezvous per tile, split into two named barriers so that nobody waits for the serial
 * work: the six plain warps only ARRIVE at barrier A when their state machines are done (warps 0/1
 * wait there), and they wait at barrier B for the serial work of the PREVIOUS iteration, which
 * warps 0/1 signalled a whole tile earlier.  The first version of this kernel did scan + look-back +
 * write-out behind two CTA-wide barriers per tile and lost 22 % of its stall samples there
 * (profiles/r01b), the second deferred only the look-back (17 %, profiles/r01c).
 * Samples arrive through cp.async (128-bit, per-warp staging in shared memory, one tile ahead);
 * header, root and level-1 node of the next sub-tile are requested before the current one is worked on.
 * The same offsets give the per-chunk lengths of the reference's Vec<Vec<Event>>.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "px_machine.cuh"
#include "px_offset.cuh"
#include "state_layout.h"

/* build-time switches of the kernel (A/B builds: tools/gpu_ab.sh) */
#ifndef ADDER_WO_PIPE
#define ADDER_WO_PIPE 1 /* write-out: the read of record e+1 overlaps the store of record e (2: e+1 and e+2 in flight, +-0: profiles/r02w_ab_wopipe2.txt) */
#endif
#ifndef ADDER_WO_DENSE
#define ADDER_WO_DENSE 0 /* 1: write-out with one lane per RECORD of a row instead of one lane per pixel looping over its records: parity-green,
                          * -7 % (4K jitter c = 10) to -12 % (noise) on every workload (profiles/r02t_ab_wodense.txt): off */
#endif
#ifndef ADDER_REC_LTC64
#define ADDER_REC_LTC64 1 /* level-record loads of the offset form carry the L2::64B prefetch-size hint (LDG.E.ENL2.LTC64B.256): an isolated
                           * 32-byte record no longer drags a 128-byte line out of DRAM.  Aged 8K stacks: DRAM reads 74.9 -> 63.0 GB per 32-frame
                           * launch, 996 -> 980 us per frame; 4K jitter +-0 (profiles/r02x_ab_ltc64.txt) */
#endif
#ifndef ADDER_NODE_LTC64
#define ADDER_NODE_LTC64 0 /* 1: the eager form's deep-level loads with the L2::64B hint: no change in DRAM bytes or time (prefetch_levels has pulled the lines already; profiles/r02x_ab_eager_ltc64.txt) */
#endif
#ifndef ADDER_ROW_LTC256
#define ADDER_ROW_LTC256 0 /* 1: the row loop's record-0 loads with the L2::256B prefetch-size hint: +-0 on every workload (profiles/r02y_ab_ltc256.txt) */
#endif
#ifndef ADDER_PF_ROLLED
#define ADDER_PF_ROLLED 0 /* 1: -128 instructions of code, more spills, -2 % .. +2 % (profiles/r02p_ab_pfrolled.txt): off */
#endif
#ifndef ADDER_TILE_PF
#define ADDER_TILE_PF 0 /* 1: headers and first records of the CTA's NEXT tile are pulled into L2 a tile ahead, whole lines, instead of
                         * its first row only: -1.5 % on every workload in both state forms (profiles/r02r_ab_prefetch.txt) — the
                         * row loop's own row-ahead loads already hide that latency.  A separate pass over the next tile that asked
                         * for the level records its pixels would read cost 15-30 % (same file) and was removed. */
#endif
#ifndef ADDER_DEEP_PF
#define ADDER_DEEP_PF 2 /* px_step entry: pull levels 2..length-1 towards L1 (1) or L2 (2); 0 = off */
#endif

namespace adder {

struct FrameArgs {
  PxParams px;          /* running_t_prev / running_t / display change per frame: see running_t[] */
  const uint8_t* frame; /* n_frames frames of P bytes, raster (y,x,c), frame_stride bytes apart */
  unsigned long long frame_stride;
  uint32_t n_frames;       /* consecutive frames in this launch: ticket k = frame k / n_tiles, tile k % n_tiles */
  uint32_t status_ring;    /* frames of status words kept (a power of two) */
  const float* running_t;  /* n_frames + 1 entries: PixelArena.running_t before frame f and after it (event_pixel_tree.rs:337) */
  unsigned long long tiles_magic; /* ceil(2^64 / n_tiles), 0 for n_tiles == 1 */
  uint2* hdr;
  uint4* nodes;
  uint2* park_arena;               /* events beyond the shared-memory slots: [CTA][park buffer][slot][pixel-in-tile] */
  uint32_t arena_slots;            /* slots per pixel in the arena (allocated node depth + 2 - S covers the worst case) */
  unsigned long long pair_stride;  /* uint4 elements between the records of levels (2j, 2j+1) and (2j+2, 2j+3): 2 * Ppad */
  uint8_t* running;
  uint32_t* ev_words;        /* output records as 3 u32 words each; frame f's start at f * ev_cap records */
  unsigned long long ev_cap; /* records per frame */
  uint32_t* chunk_off;       /* n_chunks+1 exclusive offsets per frame, or null */
  unsigned long long* tile_status;
  uint32_t* ticket;
  uint32_t* err;
  unsigned long long* total_events; /* cumulative counter, or null */
  uint32_t ticket_base, epoch;
  uint32_t P, n_tiles, C, WC, chunk_px, n_chunks;
  uint32_t row0;  /* added to every event's y: this plane is a row band of a larger frame (multi-GPU sharding) */
  unsigned long long wc_magic; /* ceil(2^64 / WC), 0 for WC == 1: i / WC without a divide */
  unsigned long long chunk_magic; /* the same for chunk_px */
  uint32_t c_magic;            /* ceil(2^32 / C), unused for C == 1: rem / C for rem < 2^24 */
  unsigned long long* counters; /* kCount only: [0] node loads [1] node stores [2] display writes [3] events
                                  * [4] live nodes at frame entry, summed over px [5] the same at frame exit */
};

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kFlagAggregate = 1u, kFlagPrefix = 2u;

/* level k of pixel i: one 128-bit access, 512 contiguous bytes per warp.  kCoherent: the launch spans several frames,
 * so the state a tile reads may have been written by another SM a moment ago — read it where it is coherent (L2). */
#ifndef ADDER_STATE_L1
#define ADDER_STATE_L1 0 /* 1 (experiment): state loads of a multi-frame launch go through L1 as well — every tile acquires its
                          * predecessor's status word first, and ld.acquire.gpu invalidates the SM's L1 (CCTL.IVALL in the SASS) */
#endif
template <bool kCoherent>
__device__ __forceinline__ uint4 ld_state(const uint4* p) { return (kCoherent && !ADDER_STATE_L1) ? __ldcg(p) : *p; }
template <bool kCoherent>
__device__ __forceinline__ uint2 ld_state(const uint2* p) { return (kCoherent && !ADDER_STATE_L1) ? __ldcg(p) : *p; }
/* root and first child of a pixel: one 32-byte record, one 256-bit access */
template <bool kCoherent>
__device__ __forceinline__ void ld_state256(const uint4* p, uint4& a, uint4& b) {
#if ADDER_ROW_LTC256
  if (kCoherent && !ADDER_STATE_L1)
    asm volatile("ld.relaxed.gpu.global.L2::256B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p)
                 : "memory");
  else
    asm volatile("ld.global.L2::256B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p)
                 : "memory");
  return;
#endif
  if (kCoherent && !ADDER_STATE_L1)
    asm volatile("ld.relaxed.gpu.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p)
                 : "memory");
  else
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void st_state256(uint4* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
template <bool kCoherent>
struct GlobalNodes {
  uint4* p; /* the pixel's first record: &nodes[NODE_SLOT(0, i, Ppad)] */
  unsigned long long stride; /* uint4 elements from one of the pixel's records to the next: 2 * Ppad */
  uint32_t n_loads, n_stores; /* only read by the counting variant of the kernel */
  __device__ __forceinline__ uint4* at(uint32_t k) const { return p + (unsigned long long)(k >> 1) * stride + (k & 1u); }
  __device__ __forceinline__ Node load(uint32_t k) {
    n_loads++;
#if ADDER_NODE_LTC64
    /* levels beyond the first record are reached by some pixels of a row only: isolated sectors, fetched with the L2::64B hint
     * like the offset form's level records (ADDER_REC_LTC64) */
    uint4 v;
    if (kCoherent && !ADDER_STATE_L1)
      asm volatile("ld.relaxed.gpu.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(at(k)) : "memory");
    else
      asm volatile("ld.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(at(k)) : "memory");
#else
    const uint4 v = ld_state<kCoherent>(at(k));
#endif
    Node n;
    n.integ = __uint_as_float(v.x);
    n.dt = __uint_as_float(v.y);
    n.best_dt = __uint_as_float(v.z);
    n.w = v.w;
    return n;
  }
  __device__ __forceinline__ void store(uint32_t k, const Node& n) {
    n_stores++;
    *at(k) = make_uint4(__float_as_uint(n.integ), __float_as_uint(n.dt), __float_as_uint(n.best_dt), n.w);
  }
  /* A freshly spawned tail (PixelNode::new).  When it opens a new record (even level) the whole 32 bytes are written — the
   * other half is beyond the pixel's length, so its content is free — and the sector needs no fill from DRAM. */
  __device__ __forceinline__ void store_fresh(uint32_t k, const Node& n) {
    n_stores++;
    const uint4 v = make_uint4(__float_as_uint(n.integ), __float_as_uint(n.dt), __float_as_uint(n.best_dt), n.w);
    if (k & 1u)
      *at(k) = v;
    else
      st_state256(at(k), v, make_uint4(0u, 0u, 0u, 0u));
  }
  /* level 1 is fetched with the root, before the length is known: it counts as algorithmic traffic
   * only when the state machine needed it; a level requested ahead and then dropped does not count */
  /* every record the walk will visit, requested at once (no register, no scoreboard) */
  __device__ __forceinline__ void prefetch_levels(uint32_t len) {
#if ADDER_DEEP_PF
    const uint4* q = p + stride;
#if ADDER_PF_ROLLED
#pragma unroll 1 /* the compiler unrolls this fifteen times at every place px_step is inlined: 160 instructions of the 47 KB kernel */
#endif
    for (uint32_t k = 2; k < len; k += 2, q += stride) {
#if ADDER_DEEP_PF == 1
      asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
#else
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
#endif
    }
#endif
  }
  /* a running pointer through the pixel's levels (ADDER_FAST_WALK): level 1 is the second half of the first record; from an
   * odd level the next one opens the next record, from an even level it is the record's other half */
  __device__ __forceinline__ uint4* cursor_level1() const { return p + 1; }
  __device__ __forceinline__ uint4* next_from_odd(uint4* q) const { return q + (stride - 1ull); }
  __device__ __forceinline__ uint4* next_from_even(uint4* q) const { return q + 1; }
  __device__ __forceinline__ Node load_at(const uint4* q) {
    n_loads++;
    const uint4 v = ld_state<kCoherent>(q);
    Node n;
    n.integ = __uint_as_float(v.x);
    n.dt = __uint_as_float(v.y);
    n.best_dt = __uint_as_float(v.z);
    n.w = v.w;
    return n;
  }
  __device__ __forceinline__ void store_at(uint4* q, const Node& n) {
    n_stores++;
    *q = make_uint4(__float_as_uint(n.integ), __float_as_uint(n.dt), __float_as_uint(n.best_dt), n.w);
  }
  __device__ __forceinline__ void used_preloaded() { n_loads++; }
  __device__ __forceinline__ void set_popped() {}
  __device__ __forceinline__ void unused_load() { n_loads--; }
  /* px_frame falls back to px_step: root and level 1 again, as they are in memory (nothing has been stored yet) */
  __device__ __forceinline__ void reload(Node& n0, Node& n1) {
    uint4 a, b;
    ld_state256<kCoherent>(p, a, b);
    n0.integ = __uint_as_float(a.x), n0.dt = __uint_as_float(a.y), n0.best_dt = __uint_as_float(a.z), n0.w = a.w;
    n1.integ = __uint_as_float(b.x), n1.dt = __uint_as_float(b.y), n1.best_dt = __uint_as_float(b.z), n1.w = b.w;
    n_loads = 1u; /* the root; level 1 counts when px_step uses it */
    n_stores = 0u;
  }
};

/* The level records of a pixel in OFFSET FORM (px_offset.cuh): record k >= 1 = level k, 32 bytes, one 256-bit access; record 0
 * (root + top level) is loaded and stored by the kernel's row loop.  Counts are in records.
 * What these scattered accesses cost, measured on aged 8K stacks (profiles/r02r_*): a record LOAD brings a whole 128-byte
 * line from DRAM (~113 bytes of reads per load), a record store ~44 bytes of traffic; asking for the records a row or two
 * ahead (prefetch.global.L2 from the row loop, three variants) never paid for its instructions, 64-byte record slots
 * written whole changed nothing but the bytes written, and storing a record as two 128-bit halves from a lane pair
 * neither.  What did pay was needing fewer of them: the top level in record 0 (px_offset.cuh). */
template <bool kCoherent>
struct OffNodes {
  uint4* p; /* the pixel's record 0 */
  unsigned long long stride; /* uint4 elements from one of the pixel's records to the next: 2 * Ppad */
  uint32_t n_loads, n_stores;
  __device__ __forceinline__ OffRec load_rec(uint32_t k) {
    n_loads++;
    uint4 a, b;
#if ADDER_REC_LTC64
    /* a level record is one isolated 32-byte sector: ask L2 to fetch 64 bytes around it instead of its default 128 */
    const uint4* q = p + (unsigned long long)k * stride;
    if (kCoherent && !ADDER_STATE_L1)
      asm volatile("ld.relaxed.gpu.global.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(q) : "memory");
    else
      asm volatile("ld.global.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(q) : "memory");
#else
    ld_state256<kCoherent>(p + (unsigned long long)k * stride, a, b);
#endif
    OffRec r;
    r.oi = a.x, r.od = a.y, r.best_dt = __uint_as_float(a.z), r.w = a.w, r.pmin = b.x, r.pk = b.y;
    return r;
  }
  __device__ __forceinline__ void store_rec(uint32_t k, const OffRec& r) {
    n_stores++;
    st_state256(p + (unsigned long long)k * stride, make_uint4(r.oi, r.od, __float_as_uint(r.best_dt), r.w), make_uint4(r.pmin, r.pk, 0u, 0u));
  }
};

/*
 * Where a pixel parks its events until the tile's output offset is known.  The first S go to
 * shared memory ([slot][pixel-in-tile]); a pixel that emits more (a changed pixel pops its whole
 * stack: rare on noise, bursts of 3-6 on slowly varying scenes) parks event #e >= S in the CTA's own
 * arena in global memory, [park buffer][e - S][pixel-in-tile]: written and read back by the same
 * thread two pipeline iterations apart, coalesced across a warp, never shared between CTAs and never
 * part of the pixel state (so nothing of a frame is left in state memory once its tile has been computed).
 */
template <uint32_t S>
struct EventPark {
  uint32_t* t; /* &slot_t[pixel-in-tile] */
  uint8_t* d;  /* &slot_d[pixel-in-tile] */
  uint2* ovf;  /* &arena[cta][buffer][0][pixel-in-tile] */
  uint32_t tile_px, ovf_slots;
  uint32_t n, overflow;
  __device__ __forceinline__ void push(uint32_t dd, uint32_t tt) {
    if (n < S) {
      t[n * tile_px] = tt;
      d[n * tile_px] = (uint8_t)dd;
    } else if (n - S < ovf_slots) {
      ovf[(n - S) * tile_px] = make_uint2(tt, dd);
    } else {
      overflow = 1;
      return;
    }
    n++;
  }
  __device__ __forceinline__ uint32_t mark() const { return n; }
  __device__ __forceinline__ void rewind(uint32_t m) { n = m; overflow = 0; }
  __device__ __forceinline__ void get(uint32_t e, uint32_t& dd, uint32_t& tt) const {
    if (e < S) {
      tt = t[e * tile_px];
      dd = d[e * tile_px];
    } else {
      const uint2 v = ovf[(e - S) * tile_px];
      tt = v.x;
      dd = v.y;
    }
  }
};

/* A tile's status word is published by ONE thread after it has passed the CTA's rendezvous behind the tile's
 * state machines: the release at gpu scope is cumulative over what that thread has observed through the
 * rendezvous (barrier A), i.e. over every warp's state stores of the tile (the pattern of CUTLASS's Semaphore::release).
 * The next frame's tile over the same pixels acquires it before its first state load. */
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v, bool release) {
  if (release)
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  else /* a single-frame launch has no reader of the pixel state inside the launch: the count alone matters */
    *reinterpret_cast<volatile unsigned long long*>(p) = v;
}
__device__ __forceinline__ unsigned long long ld_status_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

constexpr uint32_t kThreads = ADDER_TILE_PX; /* 256 */
constexpr uint32_t kWarps = kThreads / 32;

/*
 * Tile geometry.  A tile is (8R - 2) rows of 32 consecutive pixels: in rounds 0..R-2 every warp takes
 * one row (row = 8*round + warp), in the last round only warps 2..7 do (row = 8*(R-1) + warp - 2).
 * Warps 0 and 1 carry the CTA's serial work (scan + publish, look-back) and would otherwise finish
 * every iteration last, with the other six waiting for them (profiles/r01d).
 */
#ifndef ADDER_EV_STREAM
#define ADDER_EV_STREAM 1 /* event records (and with 2 the display bytes) leave with st.global.cs: written once, not read again by the kernel (-0.6 % time) */
#endif
#ifndef ADDER_NAMED_BARS
#define ADDER_NAMED_BARS 1 /* rendezvous on named barriers (bar.sync parks the warp) instead of polled mbarriers */
#endif
#ifndef ADDER_DUTY_LESS
#define ADDER_DUTY_LESS 1u /* rows fewer for warps 0/1 in the last round (experiments: 0) */
#endif
#ifndef ADDER_MIN_CTAS
#define ADDER_MIN_CTAS 4
#endif
#ifndef ADDER_LEAN
#define ADDER_LEAN 0 /* 1: px_frame (short path + fallback) instead of px_step; measured slower in this kernel (profiles/r02a_ab.txt) */
#endif
#ifndef ADDER_ROW_RUNS
#define ADDER_ROW_RUNS 0 /* 1: every warp works on a contiguous run of rows of the tile instead of every eighth row */
#endif
__host__ __device__ constexpr uint32_t tile_rows(uint32_t R) { return 8u * R - 2u * ADDER_DUTY_LESS; }
__host__ __device__ constexpr uint32_t tile_px(uint32_t R) { return 32u * tile_rows(R); }
/* shared-memory slots per pixel: 1 for the large tile (a pixel's second and later events of a frame go to the
 * CTA's arena in global memory) so that four CTAs with three park buffers each still fit an SM, else 3 */
__host__ __device__ constexpr uint32_t park_slots(uint32_t R) { return R >= 8 ? 1u : 3u; }
constexpr uint32_t kParkBufs = 3;
__host__ __device__ constexpr size_t frame_kernel_smem(uint32_t R) {
  /* frame[2][8 warps][R*32] u8 ; per park buffer: slot_t[S][TILE] u32 | info[TILE] u16 | slot_d[S][TILE] u8 ; three of them */
  return 2u * (size_t)kThreads * R + kParkBufs * ((size_t)park_slots(R) * tile_px(R) * 4u + (size_t)tile_px(R) * 2u + (size_t)park_slots(R) * tile_px(R));
}

/* (ADDER_NAMED_BARS=0 only; the default rendezvous is on named barriers, below.)
 * mbarriers in shared memory: A = "this warp's state machines of tile n are done" (8 arrivals per
 * phase, warps 0/1 wait), B = "serial work of iteration n done" (2 arrivals, the other warps wait for
 * the phase of the PREVIOUS iteration).  Arriving never blocks and waiting does not count as arriving,
 * so a warp waits only for the event it needs — a bar.sync would also make the six plain warps wait
 * for each other (profiles/r01e). */
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) { /* release.cta */
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) { /* acquire.cta */
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t ok;
  do {
    /* the last operand is the suspend-time hint (ns): the warp sleeps in hardware instead of polling */
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(addr), "r"(parity), "r"(1000000u)
                 : "memory");
  } while (!ok);
}
/* The same two signals on named barriers (ADDER_NAMED_BARS): a warp waiting in bar.sync is parked by the hardware,
 * while mbarrier.try_wait comes back every ~12 ns (profiles/r01m); the speed is the same, but compute-sanitizer
 * racecheck models bar.arrive / bar.sync and reports no hazards (profiles/r01o_sanitizer.txt).  A = barrier 1, 256 threads: the six plain warps bar.arrive, warps 0/1 bar.sync.  B = one
 * barrier per plain warp (its warp number), 96 threads: warps 0/1 bar.arrive on each, the plain warp bar.syncs — with
 * a single B the plain warps would also wait for each other. */
__device__ __forceinline__ void nbar_sync(uint32_t id, uint32_t count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nbar_arrive(uint32_t id, uint32_t count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kMaxLaunchFrames = 512; /* frames one launch may span (FrameArgs::n_frames) */

/* pull the line holding *p towards L2 (no register, no scoreboard) */
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

/*
 * Frame-wide exclusive offset of tile `tile` (run by warp 0): sum of the aggregates of the nearest
 * predecessors back to the first one that already knows its own prefix.  128 predecessors per L2
 * round trip (lane l looks at j-l, j-32-l, j-64-l, j-96-l).
 */
__device__ __forceinline__ uint32_t look_back(const unsigned long long* status, uint32_t epoch, uint32_t tile, uint32_t lane) {
  uint32_t excl = 0;
  int j = (int)tile - 1;
  for (;;) {
    uint32_t flag[4], val[4];
#pragma unroll
    for (int w = 0; w < 4; w++) {
      const int idx = j - 32 * w - (int)lane;
      flag[w] = kFlagPrefix;
      val[w] = 0;
      if (idx >= 0) {
        unsigned long long s;
        uint32_t shi;
        do {
          s = ld_status(&status[idx]);
          shi = (uint32_t)(s >> 32);
        } while ((shi >> 2) != epoch || (shi & 3u) == 0u);
        flag[w] = shi & 3u;
        val[w] = (uint32_t)s;
      }
    }
    /* windows are in order of distance: take everything up to and including the nearest prefix */
    uint32_t contrib = 0;
    bool done = false;
#pragma unroll
    for (int w = 0; w < 4; w++) {
      const uint32_t pm = __ballot_sync(kFull, flag[w] == kFlagPrefix);
      if (!done) {
        if (pm) {
          const uint32_t first = (uint32_t)__ffs((int)pm) - 1u;
          contrib += lane <= first ? val[w] : 0u;
          done = true;
        } else {
          contrib += val[w];
        }
      }
    }
#pragma unroll
    for (uint32_t o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(kFull, contrib, o);
    excl += contrib;
    if (done) break;
    j -= 128;
  }
  return excl;
}

/*
 * kCount = true is the instrumented twin used (untimed) to measure the algorithmic bytes of a
 * workload: it additionally sums node loads / stores, display writes and events into a.counters.
 */
/* kDeep = the long-integration variant: the levels below the first two of an unchanged pixel's stack are walked by the
 * lanes of the warp together (px_machine.cuh deep_item / deep_finish).  Stacks of neighbouring pixels differ in depth
 * (2 .. 11 live nodes after a few hundred frames of a static scene), and a per-lane loop runs as long as the deepest of
 * the 32: 36 warp-instructions per pixel on aged 8K stacks against 18 on two-node stacks (profiles/r02h_static_*). */
/* kOff = the node stacks are in offset form (px_offset.cuh): a frame of an unchanged pixel touches the root and one level
 * record whatever the depth of its stack.  Chosen by the host when offset_form_eligible() (Collapse, integral time). */
template <int R, bool kCount, bool kMulti, bool kDeep = false, bool kPlain = false, bool kOff = false>
__global__ void __launch_bounds__(ADDER_TILE_PX, ADDER_MIN_CTAS) integrate_frame_kernel(const FrameArgs a) {
  static_assert(!(kOff && kDeep), "the cooperative deep walk belongs to the eager form");
  constexpr uint32_t ROWS = tile_rows(R), TILE = tile_px(R);
  constexpr uint32_t S = park_slots(R);
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  /* frame[2][kWarps][R*32] u8 | slot_t[3][S][TILE] u32 | info[3][TILE] u16 | slot_d[3][S][TILE] u8 */
  uint8_t* const s_frame = smem_dyn;
  uint32_t* const s_slot_t = reinterpret_cast<uint32_t*>(smem_dyn + 2u * kThreads * R);
  uint16_t* const s_info = reinterpret_cast<uint16_t*>(s_slot_t + kParkBufs * S * TILE);
  uint8_t* const s_slot_d = reinterpret_cast<uint8_t*>(s_info + kParkBufs * TILE);

  __shared__ uint32_t s_ticket[2], s_prefix[2], s_tot[kParkBufs];
  __shared__ uint32_t s_wtot[kParkBufs][ROWS];
  __shared__ __align__(8) unsigned long long s_bar_a, s_bar_b;
  __shared__ uint8_t s_lut[260];
  __shared__ uint32_t s_kf[kDeep ? kThreads : 1]; /* kDeep: per pixel of the row a warp is working on, the shallowest level that fired */
  __shared__ float s_running_t[kMaxLaunchFrames + 1]; /* a.running_t[], read once per tile */

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const bool duty = warp < 2u;
  uint2* const arena = a.park_arena + (unsigned long long)blockIdx.x * kParkBufs * a.arena_slots * TILE; /* this CTA's */
  const uint32_t my_rows = duty ? (uint32_t)R - ADDER_DUTY_LESS : (uint32_t)R;
  const bool frame_aligned = ((reinterpret_cast<uintptr_t>(a.frame) | a.frame_stride) & 15u) == 0;
  const uint32_t n_frames = kMulti ? a.n_frames : 1u; /* kMulti = false: the single-frame form without the cross-frame machinery */
  const uint32_t n_total = n_frames * a.n_tiles; /* tickets of this launch */
  /* ticket -> frame (tile = ticket - frame * n_tiles); the frame's row of status words.
   * (Handing a tile to one CTA for T consecutive frames, so that its state would come back from L2, was built and measured:
   * DRAM bytes and time unchanged for T = 8 and 16 — one generation of 592 tiles of 1984 pixels does not stay in L2 —
   * profiles/r02r_dram_tblock.txt; removed again, its ticket arithmetic cost 2-3 % on every workload.) */
  auto frame_of = [&](uint32_t k) { return !kMulti ? 0u : a.tiles_magic ? mulhi_u32_u64(k, a.tiles_magic) : k; };
  auto status_row = [&](uint32_t fi) { return !kMulti ? a.tile_status : a.tile_status + (unsigned long long)(fi & (a.status_ring - 1u)) * a.n_tiles; };
  /* this warp's row of round r, and the tile-relative index of this thread's pixel in it */
#if ADDER_ROW_RUNS
  /* a warp's rows are one contiguous run of the tile (warps 0/1: R - 1 rows each, then R rows per plain warp): the pixel
   * index, and with it every state / park address, advances by a constant from one row to the next */
  const uint32_t row0w = duty ? warp * ((uint32_t)R - ADDER_DUTY_LESS) : 2u * ((uint32_t)R - ADDER_DUTY_LESS) + (warp - 2u) * (uint32_t)R;
  auto row_of = [&](uint32_t r) { return row0w + r; };
#else
  auto row_of = [&](uint32_t r) { return r + 1u < (uint32_t)R ? 8u * r + warp : 8u * ((uint32_t)R - 1u) + warp - 2u * ADDER_DUTY_LESS; };
#endif

  /* tiles are handed out in ticket order so that a tile's predecessors are always held by CTAs that
   * are already running: the look-back can then never wait on a CTA that has not been scheduled. */
  if (tid == 0) {
    const uint32_t t0 = atomicAdd(a.ticket, 1u) - a.ticket_base;
    s_ticket[0] = t0;
    s_ticket[1] = t0 < n_total ? atomicAdd(a.ticket, 1u) - a.ticket_base : kNone;
    mbar_init(&s_bar_a, kWarps);
    mbar_init(&s_bar_b, 2u);
  }
  for (uint32_t j = tid; j < 257u; j += kThreads) s_lut[j] = a.px.exact_lut[j];
  if (kDeep) s_kf[tid] = kNoFire;
  if (kMulti && a.running_t) {
    for (uint32_t j = tid; j <= n_frames; j += kThreads) s_running_t[j] = a.running_t[j];
  } else if (tid == 0) { /* a single frame: both values came with the arguments */
    s_running_t[0] = a.px.running_t_prev;
    s_running_t[1] = a.px.running_t;
  }
  PxParams px = a.px; /* the display table is read from shared memory */
  px.exact_lut = s_lut;
  __syncthreads();
  uint32_t t_cur = s_ticket[0], t_m1 = kNone, t_m2 = kNone; /* tiles of iteration n, n-1, n-2 */
  uint32_t n = 0, b = 0;                                     /* iteration, n % 3 */

  /* samples of tile t -> this warp's staging area fb (asynchronously when whole and aligned) */
  auto fetch_frame = [&](uint32_t t, uint32_t fb) {
    const uint32_t fi = frame_of(t);
    const uint32_t start = (t - fi * a.n_tiles) * TILE;
    const uint8_t* src = a.frame + (unsigned long long)fi * a.frame_stride;
    uint8_t* dst = s_frame + fb * (kThreads * R) + warp * (R * 32u);
    if (start + TILE <= a.P && frame_aligned) {
      if (lane < 2u * my_rows) cp_async16(dst + lane * 16u, src + start + 32u * row_of(lane >> 1) + (lane & 1u) * 16u);
    } else {
#pragma unroll 1
      for (uint32_t r = 0; r < my_rows; r++) {
        const uint32_t i = start + 32u * row_of(r) + lane;
        if (i < a.P) dst[r * 32u + lane] = src[i];
      }
    }
  };
  /* header, root and level 1 of the next row are requested before the current row is worked on
   * (root and level 1 do not need the header: both are fetched before the length is known) */
  uint2 h_next = make_uint2(0u, 0u);
  uint4 n0_next = make_uint4(0u, 0u, 0u, 0u), n1_next = n0_next;
  auto fetch_px = [&](uint32_t i) {
    if (i < a.P) {
      h_next = ld_state<kMulti>(a.hdr + i);
      ld_state256<kMulti>(a.nodes + 2ull * i, n0_next, n1_next);
    }
  };
  /* First state load of tile t.  A launch spans n_frames frames; the tail of one frame overlaps the head of the
   * next, so the pixels' state of the previous frame must be known to be in place: the same tile of frame f-1
   * (held by another CTA, or this CTA's own previous tile) must have published its status.  Called after this
   * warp has arrived at the rendezvous of the current iteration, so a CTA can wait for its own publication.
   * Frame 0 of a launch follows the previous launch in stream order. */
  auto fetch_tile_head = [&](uint32_t t) {
    const uint32_t fi = frame_of(t), tl = t - fi * a.n_tiles;
    if (kMulti && fi != 0u) {
      const unsigned long long* dep = status_row(fi - 1u) + tl;
      const uint32_t want = a.epoch + fi - 1u;
      uint32_t shi;
      do {
        shi = (uint32_t)(ld_status_acquire(dep) >> 32);
      } while ((shi >> 2) != want || (shi & 3u) == 0u);
    }
    if (my_rows) fetch_px(tl * TILE + 32u * row_of(0u) + lane);
  };
  if (t_cur < n_total) {
    fetch_frame(t_cur, 0u);
    fetch_tile_head(t_cur);
  }

  while (t_cur < n_total || t_m1 < n_total || t_m2 < n_total) {
    const bool have_tile = t_cur < n_total;
    /* ticket of iteration n+2, requested now and needed after the state machines (drawn only while
     * the previous one was a tile: every CTA draws exactly one ticket past the end, which is what the
     * host advances ticket_base by) */
    uint32_t t_next2 = kNone;
    if (tid == 0 && s_ticket[(n + 1u) & 1u] < n_total) t_next2 = atomicAdd(a.ticket, 1u) - a.ticket_base;

    /* ---- stage 1: the state machines of tile n --------------------------------------------------- */
    if (have_tile) {
      const uint32_t fi_cur = frame_of(t_cur);
      const uint32_t tile_start = (t_cur - fi_cur * a.n_tiles) * TILE;
      px.running_t_prev = s_running_t[fi_cur];
      px.running_t = s_running_t[fi_cur + 1u];
      px.display = fi_cur == 0u ? a.px.display : (a.px.display ? 1u : 0u); /* only the launch's first frame can be a forced one */
      uint32_t* const slot_t = s_slot_t + b * (S * TILE);
      uint8_t* const slot_d = s_slot_d + b * (S * TILE);
      uint16_t* const info = s_info + b * TILE;
      const uint8_t* const samples = s_frame + (n & 1u) * (kThreads * R) + warp * (R * 32u);
      cp_async_wait_all();
      __syncwarp();

      uint32_t errbits = 0;
      unsigned long long c_loads = 0, c_stores = 0, c_disp = 0, c_len_in = 0, c_len_out = 0;
#pragma unroll 1
      for (uint32_t r = 0; r < my_rows; r++) {
        const uint32_t row = row_of(r);
        const uint32_t q = 32u * row + lane; /* pixel-in-tile */
        const uint32_t i = tile_start + q;
        const uint2 hraw = h_next;
        const uint4 n0raw = n0_next, n1raw = n1_next;
        if (r + 1u < my_rows) fetch_px(tile_start + 32u * row_of(r + 1u) + lane);
        uint32_t nev = 0;
        PxHeader h{__uint_as_float(hraw.x), hraw.y};
        bool deferred = false; /* kDeep: the walk below level 1 is still to be done */
        const uint32_t sample = samples[r * 32u + lane];
        if constexpr (kOff) if (i < a.P) {
          OffNodes<kMulti> mem{a.nodes + 2ull * i, a.pair_stride, 0u, 0u};
          EventPark<S> park{slot_t + q, slot_d + q, arena + (unsigned long long)b * a.arena_slots * TILE + q, TILE, a.arena_slots, 0u, 0u};
          Node n0{__uint_as_float(n0raw.x), __uint_as_float(n0raw.y), __uint_as_float(n0raw.z), n0raw.w};
          OffTop top{n1raw.x, n1raw.z, __uint_as_float(n1raw.y), n1raw.w}; /* in memory: a, best_dt, b, pmin */
          uint8_t disp;
          const bool show = px_offset<kPlain>(px, sample, h, n0, top, mem, park, errbits, &disp);
          a.hdr[i] = make_uint2(__float_as_uint(h.lf), h.y);
          st_state256(a.nodes + 2ull * i, make_uint4(__float_as_uint(n0.integ), __float_as_uint(n0.dt), __float_as_uint(n0.best_dt), n0.w),
                      make_uint4(top.a, __float_as_uint(top.best_dt), top.b, top.pmin));
          if (show) a.running[i] = disp;
          if (park.overflow) errbits |= ADDER_DEVERR_DEPTH;
          nev = park.n;
          if (kCount) { /* in 16-byte units like the eager form: record 0 (root + meta) and every level record are two each */
            c_loads += 2u + 2u * mem.n_loads;
            c_stores += 2u + 2u * mem.n_stores;
            c_disp += show ? 1u : 0u;
            c_len_in += HDR_LENGTH(hraw.y);
            c_len_out += HDR_LENGTH(h.y);
          }
        }
        if constexpr (!kOff) if (i < a.P) {
          GlobalNodes<kMulti> mem{a.nodes + 2ull * i, a.pair_stride, 1u, 0u};
          EventPark<S> park{slot_t + q, slot_d + q, arena + (unsigned long long)b * a.arena_slots * TILE + q, TILE, a.arena_slots, 0u, 0u};
          const Node n0{__uint_as_float(n0raw.x), __uint_as_float(n0raw.y), __uint_as_float(n0raw.z), n0raw.w};
          const Node n1{__uint_as_float(n1raw.x), __uint_as_float(n1raw.y), __uint_as_float(n1raw.z), n1raw.w};
          uint8_t disp;
#if ADDER_LEAN /* A/B: the short path of px_frame in front of the general state machine */
          const bool show = px_frame(px, sample, h, n0, n1, mem, park, errbits, &disp);
#else
          const bool show = px_step<kDeep, kPlain>(px, sample, h, n0, n1, mem, park, errbits, &disp, &deferred);
#endif
          if (!kDeep) a.hdr[i] = make_uint2(__float_as_uint(h.lf), h.y);
#if ADDER_EV_STREAM >= 2
          if (show) __stcs(a.running + i, disp);
#else
          if (show) a.running[i] = disp;
#endif
          if (park.overflow) errbits |= ADDER_DEVERR_DEPTH;
          nev = park.n;
          if (kCount) {
            c_loads += mem.n_loads;
            c_stores += mem.n_stores;
            c_disp += show ? 1u : 0u;
            c_len_in += HDR_LENGTH(hraw.y);
            if (!kDeep) c_len_out += HDR_LENGTH(h.y);
          }
        }
        if (kDeep) {
          if (__ballot_sync(kFull, deferred)) { /* some pixel of this row has levels left to walk */
            const uint32_t len_in = HDR_LENGTH(hraw.y);
            const uint32_t my_len = deferred ? len_in : 0u;
            const float my_i = (float)sample;
            uint32_t* const kfw = s_kf + warp * 32u; /* all kNoFire between rows */
            const uint32_t maxlen = __reduce_max_sync(kFull, my_len);
            uint32_t k = 2;
            while (k < maxlen) {
              /* one pass = as many whole levels as fit the 32 lanes: level k's pixels, then level k+1's, ... (the number
               * of pixels that reach a level falls with the level), each lane taking one (pixel, level) */
              uint32_t used = 0, lvl = 0, src = lane;
              bool active = false;
              while (k < maxlen) {
                const uint32_t m = __ballot_sync(kFull, my_len > k);
                const uint32_t c = (uint32_t)__popc(m);
                if (used + c > 32u) break;
                if (lane >= used && lane < used + c) {
                  active = true;
                  lvl = k;
                  src = __fns(m, 0u, (int)(lane - used) + 1); /* the (lane - used)-th pixel of the row that has level k */
                }
                used += c;
                k++;
              }
              const float i_src = __shfl_sync(kFull, my_i, src);
              const uint32_t len_src = __shfl_sync(kFull, my_len, src);
              if (active && kfw[src] > lvl) { /* nothing shallower of that pixel has fired so far */
                GlobalNodes<kMulti> dm{a.nodes + 2ull * (i - lane + src), a.pair_stride, 0u, 0u};
                if (deep_item(dm, lvl, len_src, i_src, px.time)) atomicMin(&kfw[src], lvl);
                if (kCount) {
                  c_loads += dm.n_loads;
                  c_stores += dm.n_stores;
                }
              }
              __syncwarp();
            }
            if (deferred) { /* the pixel's own lane finishes the level that fired and learns the new length */
              const uint32_t kf = kfw[lane];
              kfw[lane] = kNoFire;
              GlobalNodes<kMulti> fm{a.nodes + 2ull * i, a.pair_stride, 0u, 0u};
              const uint32_t nl = deep_finish(px, fm, kf, len_in, my_i, errbits);
              h.y = (h.y & ~(0x1Fu << 24)) | (nl << 24);
              if (kCount) {
                c_loads += fm.n_loads;
                c_stores += fm.n_stores;
              }
            }
            __syncwarp();
          }
          if (i < a.P) {
            a.hdr[i] = make_uint2(__float_as_uint(h.lf), h.y);
            if (kCount) c_len_out += HDR_LENGTH(h.y);
          }
        }
        /* place of this pixel's records inside its row's run: two ballots cover 0..2 events per
         * pixel, the shuffle scan is only taken when some pixel of the row emitted more */
        const uint32_t le = 0xFFFFFFFFu >> (31u - lane);
        const uint32_t b1 = __ballot_sync(kFull, nev >= 1u), b2 = __ballot_sync(kFull, nev >= 2u);
        uint32_t incl, wtotal;
        if (__ballot_sync(kFull, nev >= 3u) == 0u) {
          incl = (uint32_t)__popc(b1 & le) + (uint32_t)__popc(b2 & le);
          wtotal = (uint32_t)__popc(b1) + (uint32_t)__popc(b2);
        } else {
          incl = nev;
#pragma unroll
          for (uint32_t o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
          }
          wtotal = __shfl_sync(kFull, incl, 31);
        }
        info[q] = (uint16_t)((incl - nev) | (nev << 10)); /* <= 31*33 = 1023 | <= 33 */
        if (lane == 0) s_wtot[b][row] = wtotal;
      }
      if (kCount) {
        atomicAdd(&a.counters[0], c_loads);
        atomicAdd(&a.counters[1], c_stores);
        atomicAdd(&a.counters[2], c_disp);
        atomicAdd(&a.counters[4], c_len_in);
        atomicAdd(&a.counters[5], c_len_out);
      }
      if (errbits) atomicOr(a.err, errbits);
    }

    /* ---- rendezvous + stage 2 (warps 0 and 1): the serial work ------------------------------------ */
    /* The plain warps wait for B(n-1) BEFORE they arrive at A(n): warps 0/1 can then not complete
     * B(n) (which would alias B(n-1)'s parity) while somebody still waits for B(n-1). */
#if ADDER_NAMED_BARS
    if (!duty) {
      if (n != 0u) nbar_sync(warp, 96u); /* serial work of iteration n-1 (signalled a tile ago) */
      __syncwarp();
      __threadfence_block();
      nbar_arrive(1u, 256u); /* this warp's rows of tile n are parked, their totals stored */
    }
#else
    if (!duty && n != 0u) mbar_wait(&s_bar_b, (n - 1u) & 1u); /* serial work of iteration n-1 (signalled a tile ago) */
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_bar_a); /* this warp's rows of tile n are parked, their totals stored */
#endif
    if (duty) {
#if ADDER_NAMED_BARS
      nbar_sync(1u, 256u); /* every row total of tile n is in shared memory */
#else
      mbar_wait(&s_bar_a, n & 1u); /* every row total of tile n is in shared memory */
#endif
      if (warp == 0) {
        if (have_tile) {
          /* exclusive scan of the row totals, in pixel order; publish the aggregate */
          constexpr uint32_t kPer = (ROWS + 31) / 32;
          uint32_t mine[kPer], sum = 0;
#pragma unroll
          for (uint32_t j = 0; j < kPer; j++) {
            const uint32_t idx = lane * kPer + j;
            mine[j] = idx < ROWS ? s_wtot[b][idx] : 0u;
            sum += mine[j];
          }
          uint32_t incl = sum;
#pragma unroll
          for (uint32_t o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
          }
          const uint32_t tot = __shfl_sync(kFull, incl, 31);
          uint32_t run = incl - sum;
#pragma unroll
          for (uint32_t j = 0; j < kPer; j++) {
            const uint32_t idx = lane * kPer + j;
            if (idx < ROWS) s_wtot[b][idx] = run;
            run += mine[j];
          }
          if (lane == 0) {
            /* tile 0 knows its prefix (0) at once; the others publish their aggregate now and their
             * inclusive prefix after their look-back, one iteration later */
            const uint32_t fi = frame_of(t_cur), tl = t_cur - fi * a.n_tiles;
            const unsigned long long tag = (unsigned long long)(a.epoch + fi) << 2;
            st_status(status_row(fi) + tl, ((tag | (tl == 0u ? kFlagPrefix : kFlagAggregate)) << 32) | tot, kMulti && n_frames > 1u);
            s_tot[b] = tot;
            if (kCount) atomicAdd(&a.counters[3], (unsigned long long)tot);
          }
        }
        if (lane == 0) s_ticket[n & 1u] = t_next2; /* slot (n+2) & 1 */
      } else if (t_m1 < n_total) {
        const uint32_t fi = frame_of(t_m1), tl = t_m1 - fi * a.n_tiles;
        unsigned long long* const row = status_row(fi);
        const uint32_t excl = tl != 0u ? look_back(row, a.epoch + fi, tl, lane) : 0u;
        if (lane == 0) {
          const uint32_t incl_all = excl + s_tot[b == 0u ? 2u : b - 1u];
          const unsigned long long tag = (unsigned long long)(a.epoch + fi) << 2;
          if (tl != 0u) st_status(row + tl, ((tag | kFlagPrefix) << 32) | incl_all, kMulti && n_frames > 1u); /* a dependent may acquire this value instead of the aggregate */
          s_prefix[(n + 1u) & 1u] = excl; /* slot (n-1) & 1 */
          if (a.chunk_off && a.chunk_px >= TILE) {
            /* lengths of the reference's Vec<Vec<Event>>, video.rs:677-734: at most one chunk starts inside this tile.  Its
             * offset is written here, by the one thread that has just learnt the tile's prefix: the row prefixes (warp 0's
             * scan, a tile ago) and the pixel's place in its row are in the park buffer of tile n-1, which no warp rewrites
             * before iteration n+2 — and nobody starts that before this warp has signalled B(n). */
            const uint32_t pbm1 = b == 0u ? 2u : b - 1u;
            const uint32_t ps = tl * TILE;
            const uint32_t t_end = ps + TILE - 1u < a.P - 1u ? ps + TILE - 1u : a.P - 1u;
            const uint32_t ch = a.chunk_magic ? mulhi_u32_u64(t_end, a.chunk_magic) : t_end; /* chunk of the tile's last pixel */
            const uint32_t cb = ch * a.chunk_px;                                              /* its first pixel */
            if (cb >= ps) {
              const uint32_t qq = cb - ps;
              a.chunk_off[(unsigned long long)fi * (a.n_chunks + 1u) + ch] = excl + s_wtot[pbm1][qq >> 5] + ((s_info + pbm1 * TILE)[qq] & 1023u);
            }
          }
          if (tl == a.n_tiles - 1u) {
            if (a.chunk_off) a.chunk_off[(unsigned long long)fi * (a.n_chunks + 1u) + a.n_chunks] = incl_all;
            if (a.total_events) atomicAdd(a.total_events, (unsigned long long)incl_all);
          }
        }
      }
      __syncwarp();
#if ADDER_NAMED_BARS
      __threadfence_block();
#pragma unroll
      for (uint32_t w = 2; w < kWarps; w++) nbar_arrive(w, 96u); /* serial work of iteration n done */
#else
      if (lane == 0) mbar_arrive(&s_bar_b); /* serial work of iteration n done */
#endif
    }

    /* ---- what iteration n+1 will need: its samples and its first row ------------------------------ */
    const uint32_t t_next = s_ticket[(n + 1u) & 1u];
    if (t_next < n_total) {
      fetch_frame(t_next, (n + 1u) & 1u);
#if ADDER_TILE_PF
      { /* the whole next tile's headers and first records towards L2 now (a tile's headers are TILE / 16 lines of 128 bytes,
         * its first records TILE / 4): the row loop's loads then find them there instead of in DRAM a row ahead.  Harmless
         * before the previous frame's tile has stored them: L2 is where those stores land. */
        const uint32_t i0 = (t_next - frame_of(t_next) * a.n_tiles) * TILE;
#pragma unroll 1
        for (uint32_t j = tid; j < TILE / 4u; j += kThreads) {
          const uint32_t i = i0 + 4u * j;
          if (i < a.P) {
            prefetch_l2(a.nodes + 2ull * i);
            if ((j & 3u) == 0u) prefetch_l2(a.hdr + i);
          }
        }
      }
#else
      const uint32_t i = (t_next - frame_of(t_next) * a.n_tiles) * TILE + 32u * row_of(0u) + lane;
      if (my_rows && i < a.P) { /* towards L2 now, into registers after the write-out */
        if ((lane & 15u) == 0u) prefetch_l2(a.hdr + i);
        if ((lane & 3u) == 0u) prefetch_l2(a.nodes + 2ull * i); /* 128-byte lines: four 32-byte records each */
      }
#endif
    }

    /* ---- stage 3: write-out of tile n-2, every thread stores its own pixels' records ------------- */
    if (t_m2 < n_total) {
      const uint32_t pb = b == 2u ? 0u : b + 1u; /* (n-2) % 3 */
      const uint32_t prefix = s_prefix[n & 1u];  /* slot (n-2) & 1 */
      const uint32_t pfi = frame_of(t_m2);
      const uint32_t pstart = (t_m2 - pfi * a.n_tiles) * TILE;
      uint32_t* const ev_out = a.ev_words + (unsigned long long)pfi * a.ev_cap * 3ull;
      uint32_t* const chunk_out = a.chunk_off ? a.chunk_off + (unsigned long long)pfi * (a.n_chunks + 1u) : nullptr;
      const uint32_t* const pslot_t = s_slot_t + pb * (S * TILE);
      const uint8_t* const pslot_d = s_slot_d + pb * (S * TILE);
      const uint16_t* const pinfo = s_info + pb * TILE;
      uint32_t capbits = 0;
      if (chunk_out) { /* chunks smaller than a tile: several may start here, each offset written by the thread that owns the pixel
                        * (a chunk of a tile or more is handled by warp 1 when it learns the tile's prefix) */
        if (a.chunk_px < TILE) {
#pragma unroll 1
          for (uint32_t r = 0; r < my_rows; r++) {
            const uint32_t row = row_of(r), q = 32u * row + lane, i = pstart + q;
            if (i < a.P && i % a.chunk_px == 0u) chunk_out[i / a.chunk_px] = prefix + s_wtot[pb][row] + (pinfo[q] & 1023u);
          }
        }
      }
#if ADDER_WO_DENSE
      /* One lane per RECORD of the row instead of one lane per pixel: a changed pixel gives up its whole stack at once
       * (3-6 records on slowly varying scenes, up to depth + 2), and a loop over a pixel's records runs as long as the
       * longest run of the row with two or three lanes in it, every record beyond the first a dependent read-back from the
       * arena (profiles/r02t_*: 12-20 % of the stall samples).  Record j of the row belongs to the last pixel whose
       * exclusive offset is <= j (offsets are in pixel order; a pixel without records shares its offset with its
       * successor): five shuffle steps.  The row's arena reads are then independent and in flight together. */
#pragma unroll 1
      for (uint32_t r = 0; r < my_rows; r++) {
        const uint32_t row = row_of(r);
        const uint32_t q = 32u * row + lane;
        const uint32_t inf = pinfo[q];
        const uint32_t excl = inf & 1023u;
        const uint32_t wtotal = __shfl_sync(kFull, excl + (inf >> 10), 31);
        if (wtotal == 0u) continue;
#pragma unroll 1
        for (uint32_t j = lane; j - lane < wtotal; j += 32u) {
          uint32_t p = 0;
#pragma unroll
          for (uint32_t s = 16; s; s >>= 1) {
            const uint32_t ex = __shfl_sync(kFull, excl, p + s);
            if (ex <= j) p += s;
          }
          const uint32_t ex_p = __shfl_sync(kFull, excl, p);
          if (j < wtotal) {
            const uint32_t e = j - ex_p, qp = 32u * row + p;
            const uint32_t i = pstart + qp;
            const uint32_t y = a.wc_magic ? mulhi_u32_u64(i, a.wc_magic) : i;
            const uint32_t rem = i - y * a.WC;
            uint32_t x = rem, c_p = ADDER_C_NONE;
            if (a.C != 1u) {
              x = __umulhi(rem, a.c_magic);
              c_p = rem - x * a.C;
            }
            const uint32_t w0_p = x | ((y + a.row0) << 16);
            const uint32_t base = prefix + s_wtot[pb][row];
            uint32_t dd, tt;
            if (e < S) {
              tt = pslot_t[e * TILE + qp];
              dd = pslot_d[e * TILE + qp];
            } else {
              const uint2 v = (arena + (unsigned long long)pb * a.arena_slots * TILE)[(unsigned long long)(e - S) * TILE + qp];
              tt = v.x;
              dd = v.y;
            }
            const unsigned long long rec = (unsigned long long)base + j;
            if (rec < a.ev_cap) {
              uint32_t* dst = ev_out + rec * 3ull;
#if ADDER_EV_STREAM
              __stcs(dst, w0_p);
              __stcs(dst + 1, c_p | (dd << 8));
              __stcs(dst + 2, tt);
#else
              dst[0] = w0_p;
              dst[1] = c_p | (dd << 8);
              dst[2] = tt;
#endif
            } else {
              capbits = ADDER_DEVERR_CAPACITY;
            }
          }
        }
      }
#else
#pragma unroll 1
      for (uint32_t r = 0; r < my_rows; r++) {
        const uint32_t row = row_of(r);
        const uint32_t q = 32u * row + lane;
        const uint32_t inf = pinfo[q];
        const uint32_t nev = inf >> 10;
        if (nev) {
          const uint32_t i = pstart + q;
          const uint32_t first = prefix + s_wtot[pb][row] + (inf & 1023u);
          const uint32_t y = a.wc_magic ? mulhi_u32_u64(i, a.wc_magic) : i;
          const uint32_t rem = i - y * a.WC;
          uint32_t x = rem, c = ADDER_C_NONE;
          if (a.C != 1u) {
            x = __umulhi(rem, a.c_magic);
            c = rem - x * a.C;
          }
          const uint32_t w0 = x | ((y + a.row0) << 16);
          EventPark<S> park{const_cast<uint32_t*>(pslot_t) + q, const_cast<uint8_t*>(pslot_d) + q, arena + (unsigned long long)pb * a.arena_slots * TILE + q, TILE, a.arena_slots, nev, 0u};
#if ADDER_WO_PIPE == 2
          /* two reads in flight: records e+1 and e+2 are requested while record e is stored */
          uint32_t dd, tt, d1 = 0, t1 = 0;
          park.get(0u, dd, tt);
          if (nev > 1u) park.get(1u, d1, t1);
#pragma unroll 1
          for (uint32_t e = 0; e < nev; e++) {
            uint32_t d2 = 0, t2 = 0;
            if (e + 2u < nev) park.get(e + 2u, d2, t2);
            const unsigned long long rec = (unsigned long long)first + e;
            if (rec < a.ev_cap) {
              uint32_t* dst = ev_out + rec * 3ull;
              __stcs(dst, w0);
              __stcs(dst + 1, c | (dd << 8));
              __stcs(dst + 2, tt);
            } else {
              capbits = ADDER_DEVERR_CAPACITY;
            }
            dd = d1, tt = t1;
            d1 = d2, t1 = t2;
          }
#elif ADDER_WO_PIPE
          /* records beyond the shared-memory slot come back from global memory: the read of record
           * e+1 is in flight while record e is stored */
          uint32_t dd, tt;
          park.get(0u, dd, tt);
#pragma unroll 1
          for (uint32_t e = 0; e < nev; e++) {
            uint32_t dn = 0, tn = 0;
            if (e + 1u < nev) park.get(e + 1u, dn, tn);
            const unsigned long long rec = (unsigned long long)first + e;
            if (rec < a.ev_cap) {
              uint32_t* dst = ev_out + rec * 3ull;
#if ADDER_EV_STREAM
              __stcs(dst, w0); /* records are written once and not read again by this kernel */
              __stcs(dst + 1, c | (dd << 8));
              __stcs(dst + 2, tt);
#else
              dst[0] = w0;
              dst[1] = c | (dd << 8);
              dst[2] = tt;
#endif
            } else {
              capbits = ADDER_DEVERR_CAPACITY;
            }
            dd = dn;
            tt = tn;
          }
#else
#pragma unroll 1
          for (uint32_t e = 0; e < nev; e++) {
            uint32_t dd, tt;
            park.get(e, dd, tt);
            const unsigned long long rec = (unsigned long long)first + e;
            if (rec < a.ev_cap) {
              uint32_t* dst = ev_out + rec * 3ull;
#if ADDER_EV_STREAM
              __stcs(dst, w0); /* records are written once and not read again by this kernel */
              __stcs(dst + 1, c | (dd << 8));
              __stcs(dst + 2, tt);
#else
              dst[0] = w0;
              dst[1] = c | (dd << 8);
              dst[2] = tt;
#endif
            } else {
              capbits = ADDER_DEVERR_CAPACITY;
            }
          }
#endif
        }
      }
#endif /* ADDER_WO_DENSE */
      if (capbits) atomicOr(a.err, capbits);
    }

    if (t_next < n_total) fetch_tile_head(t_next);
    t_m2 = t_m1;
    t_m1 = t_cur;
    t_cur = t_next;
    n++;
    b = b == 2u ? 0u : b + 1u;
    __syncwarp(); /* this warp's lanes leave the iteration together (its own s_wtot entries are reused) */
  }
}

/* ---- small state kernels ---------------------------------------------------------------------- */

/* Video::new (video.rs:364-382): every pixel = PixelArena::new(1.0, coord), event_pixel_tree.rs:69-87 */
__global__ void init_state_kernel(uint2* hdr, uint4* nodes, uint8_t* running, uint32_t P) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  hdr[i] = make_uint2(0u, HDR_PACK(0, 10, 1, 1, 0, 0));
  nodes[2ull * i] = make_uint4(0u, 0u, 0u, NODE_PACK(0 /* get_d(1.0) */, 0, 0)); /* the root: first half of the pixel's first record */
  nodes[2ull * i + 1ull] = make_uint4(0u, 0u, 0u, 0u);
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

/* Offset form -> eager form, in place (the launch parameters stopped being eligible, adder_b200.cu choose_form).  The eager
 * record j holds levels 2j and 2j+1, which come from offset records 2j and 2j+1 >= j: ascending j never overwrites a
 * record that is still to be read.  All records of a pixel belong to that pixel alone. */
__global__ void offset_to_eager_kernel(const uint2* hdr, uint4* nodes, uint32_t P, unsigned long long stride) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const uint32_t y = hdr[i].y, len = HDR_LENGTH(y);
  const bool frozen = HDR_POPPED(y) != 0u;
  uint4* const p = nodes + 2ull * i;
  const uint4 root = p[0];
  const uint32_t x = __float2uint_rz(__uint_as_float(root.x)), dt = __float2uint_rz(__uint_as_float(root.y));
  const uint4 topw = p[1];
  auto level = [&](uint32_t k) -> uint4 {
    if (k == 0u) return root;
    if (k + 1u >= len) return make_uint4(0u, 0u, 0u, 0u); /* the implicit fresh tail, and everything beyond the stack */
    uint4 r;
    if (!frozen && k + 2u == len) { /* the top level lives in record 0 */
      const OffRec q = top_unpack(OffTop{topw.x, topw.z, __uint_as_float(topw.y), topw.w});
      r = make_uint4(q.oi, q.od, __float_as_uint(q.best_dt), q.w);
    } else {
      r = p[(unsigned long long)k * stride];
    }
    if (!frozen) {
      r.x = __float_as_uint(__uint2float_rn(x - r.x));
      r.y = __float_as_uint(__uint2float_rn(dt - r.y));
    }
    return r;
  };
  for (uint32_t j = 0; 2u * j < len; j++) {
    const uint4 a = level(2u * j), b = level(2u * j + 1u);
    uint4* q = p + (unsigned long long)j * stride;
    q[0] = a;
    q[1] = b;
  }
}

/* set_initial_d (video.rs:780-801) */
__global__ void set_initial_d_kernel(uint2* hdr, uint4* nodes, const uint8_t* frame, uint32_t P) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  uint32_t v = frame[i];
  uint32_t d = v == 0u ? ADDER_D_ZERO_INTEGRATION : 31u - __clz(v); /* floor(log2(v)) */
  nodes[2ull * i].w = (nodes[2ull * i].w & ~0xFFu) | d;
  hdr[i].y = (hdr[i].y & ~0xFFu) | v;
}

}  // namespace adder
