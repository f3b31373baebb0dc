/*
 * framer_kernel.cuh — the reference's INSTANTANEOUS framer (events -> u8 frames) on the device
 * (SURVEY.md §8(f) #3): what SimulProcessor runs downstream of Framed::consume (utils/simulproc.rs:166-218).
 *
 *   ingest_event_for_chunk     adder-codec-rs/src/framer/driver.rs:984-1133   per event: running timestamp, frames the
 *                                                                              event reaches, first-value-wins fill
 *   ingest_events_events       :564-626
 *   is_frame_filled / pop      :820-927
 *   u8::get_frame_value        adder-codec-rs/src/framer/scale_intensity.rs:58-104
 *
 * The reference keeps, per chunk, a VecDeque of frames of Option<u8> plus filled counts.  Here the frames of the
 * whole plane live in a ring indexed by ABSOLUTE output frame number (deque index i of a chunk is absolute frame
 * frames_written + i), one value byte and one is-some byte per pixel; "filled" is a reduction over the is-some
 * bytes instead of a counter, so the fill needs no atomics.  Per pixel-channel state, packed into one 16-byte record:
 * running timestamp (u64), last filled frame (56 bits, signed), last intensity (u8).
 *
 * One thread handles one RUN of consecutive events of the same pixel-channel (the transcoder's stream keeps a pixel's
 * events of a frame contiguous; that is the precondition of ingest_events): the thread whose event starts a run walks it
 * in order, exactly like the reference's sequential loop over a chunk's events.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "px_machine.cuh"
#include "state_layout.h"

namespace adder {

struct FramerArgs {
  const uint32_t* ev_words;  /* 12-byte records */
  const uint32_t* chunk_off; /* n_chunks + 1 exclusive offsets into ev_words (records) */
  uint32_t n_chunks, chunk_rows, W, H, C;
  /* FrameSequenceState, driver.rs:230-249 */
  long long frames_written;
  uint32_t tpf, ref_interval, source_dtm;
  uint32_t codec_version, framed_source, view_mode, absolute_t;
  float practical_d_max;
  long long buffer_limit; /* < 0: None */
  /* per pixel-channel, one 16-byte record (one 128-bit load and store per run of events instead of three of each):
   * .x/.y = running timestamp (u64), .z/.w = last filled frame (i64) << 8 | last intensity */
  uint4* px_state;
  /* frame ring */
  uint8_t* ring_val;
  uint8_t* ring_some;
  uint32_t ring_frames;
  unsigned long long frame_px; /* W*H*C */
  /* per chunk */
  long long* offset_max;   /* frame_idx_offsets: the furthest frame any pixel of the chunk has reached */
  long long* forced_frame; /* absolute frame whose filled_count was forced to full by buffer_limit, or -1 */
  uint32_t* err;           /* bit 0: an event reaches beyond the ring (the reference would grow its VecDeque) */
  const uint8_t* exact_lut; /* [257] build_exact_lut(ref_interval): the Intensity byte of exactly integral intensities */
  /* ceil(2^64 / d) for the four divisors of the per-event work (ref_magic_of: 0 stands for d == 1): n / d for any n < 2^32 is
   * mulhi_u32_u64(n, magic), exact (px_machine.cuh div_ref, checked in tests/test_px_shortcuts_host.py).  The divisions were
   * 40 % of the ingest kernel's instructions (profiles/r02t_framer_*). */
  unsigned long long tpf_magic, ring_magic, ref_magic, chunk_rows_magic;
};

__host__ __device__ __forceinline__ unsigned long long framer_state_ts(const uint4& s) { return (unsigned long long)s.x | ((unsigned long long)s.y << 32); }
__host__ __device__ __forceinline__ long long framer_state_last_filled(const uint4& s) {
  return (long long)((unsigned long long)s.z | ((unsigned long long)s.w << 32)) >> 8; /* arithmetic shift: -1 stays -1 */
}
__host__ __device__ __forceinline__ uint32_t framer_state_intensity(const uint4& s) { return s.z & 0xFFu; }
__host__ __device__ __forceinline__ uint4 framer_state_pack(unsigned long long ts, long long last_filled, uint32_t intensity) {
  const unsigned long long p = ((unsigned long long)last_filled << 8) | (intensity & 0xFFu);
  return make_uint4((uint32_t)ts, (uint32_t)(ts >> 32), (uint32_t)p, (uint32_t)(p >> 32));
}

/* n / d and n % d for a 64-bit n that almost always fits 32 bits (timestamps, frame numbers): a multiplication by the
 * divisor's magic number there, the 64-bit division (a long dependent sequence) only beyond */
__device__ __forceinline__ unsigned long long udiv_fast(unsigned long long n, uint32_t d, unsigned long long magic) {
  return (n >> 32) == 0ull ? (unsigned long long)(magic ? mulhi_u32_u64((uint32_t)n, magic) : (uint32_t)n) : n / d;
}
__device__ __forceinline__ uint32_t urem_fast(unsigned long long n, uint32_t d, unsigned long long magic) {
  if ((n >> 32) != 0ull) return (uint32_t)(n % d);
  const uint32_t q = magic ? mulhi_u32_u64((uint32_t)n, magic) : (uint32_t)n;
  return (uint32_t)n - q * d;
}

/* <u8 as FrameValue>::get_frame_value, SourceType::U8 (scale_intensity.rs:58-104, :262-270).  The Intensity view is the
 * expression the transcoder evaluates for its display byte: the same exact shortcut (px_machine.cuh frame_value_u8,
 * checked exhaustively in tests/test_px_shortcuts_host.py) instead of an f64 division per event. */
__device__ __forceinline__ uint8_t framer_value_u8(uint32_t view_mode, uint32_t d, uint32_t t, uint32_t ref, const uint8_t* exact_lut,
                                                   float practical_d_max, uint32_t dtm, uint32_t sae_running, uint32_t sae_last) {
  float q;
  switch (view_mode) {
    case 0: {
      PxParams p{};
      p.view_mode = 0u;
      p.ref = ref;
      p.tpf = (double)ref;
      p.tpf_f = (float)ref;
      p.exact_lut = exact_lut;
      return frame_value_u8(p, d, t, 0.0f);
    }
    case 1: q = __fdiv_rn((float)d, practical_d_max); break;
    case 2: q = __fdiv_rn(__uint2float_rn(t), __uint2float_rn(dtm)); break;
    default: q = __fdiv_rn(__uint2float_rn(sae_running - sae_last), __uint2float_rn(dtm)); break;
  }
  const uint32_t u = __float2uint_rz(__fmul_rn(q, 255.0f));
  return (uint8_t)(u > 255u ? 255u : u);
}

/* frame_idx_offsets of a chunk (`offset_max`) is raised to the furthest frame any of its pixels has reached.  Every
 * pixel of a chunk reaches about the same frame, so per-event atomics on the chunk's one word were 70 % of the kernel's
 * time in round 1, and "look first, then atomicMax" still 36 us of 96 (profiles/r02t_framer_knockout.txt): at the start
 * of a call every event of the chunk sees the old value.  The events of a warp are consecutive in the stream, i.e. in one
 * or two chunks: the warp reduces per chunk and one lane per chunk looks and updates.  All 32 lanes call this. */
__device__ __forceinline__ void framer_raise_offset_max(long long* offset_max, bool active, uint32_t chunk, long long reach) {
  const uint32_t lane = threadIdx.x & 31u;
  active = active && reach >= 0;
  uint32_t remaining = __ballot_sync(0xFFFFFFFFu, active);
  while (remaining) { /* warp-uniform */
    const int leader = __ffs((int)remaining) - 1;
    const uint32_t c = __shfl_sync(0xFFFFFFFFu, chunk, leader);
    const bool in = active && chunk == c;
    long long m = in ? reach : -1;
#pragma unroll
    for (uint32_t o = 16; o; o >>= 1) {
      const long long t = __shfl_xor_sync(0xFFFFFFFFu, m, o);
      m = t > m ? t : m;
    }
    if ((int)lane == leader && m > __ldcg(&offset_max[c])) atomicMax(&offset_max[c], m);
    remaining &= ~__ballot_sync(0xFFFFFFFFu, in);
  }
}

/* Measured on 1080p gray noise, 1.93 M events per call (tools/framer_bench.py; profiles/r02t_framer_*): 96 us per call for ingest +
 * refresh at the start of round 2's last session, 55 us now — the chunk's offset_max raised once per warp and chunk instead of
 * looked at by every event (-36 us), the four divisions by runtime constants as multiplications (they were 40 % of the
 * instructions), a grid of exactly the resident CTAs, status + tracker in one launch.  Tried and dropped: 2 or 4 events per
 * thread in flight with phased loads (64 / 100+ registers: 66 / 100 us), the is-some byte of an AbsoluteT event's frame
 * requested together with the state (+-0), 6 CTAs per SM at 40 registers (-6 %). */
__global__ void __launch_bounds__(256) framer_ingest_kernel(const FramerArgs a) {
  const uint32_t total = a.chunk_off[a.n_chunks];
  for (uint32_t jw = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); jw < total; jw += gridDim.x * blockDim.x) { /* warp-uniform */
    const uint32_t j = jw + (threadIdx.x & 31u);
    uint32_t chunk = 0u;
    long long reach = -1;
    bool worked = false;
    if (j < total) do {
    /* this record and the one before it, requested together: the kernel is bound by the latency of dependent loads.
     * A run never crosses a chunk boundary (chunks are disjoint rows), so comparing coordinates is enough to find
     * its first event and its end: the chunk offsets are not needed for that. */
    const uint32_t w0 = a.ev_words[3ull * j], w1 = a.ev_words[3ull * j + 1ull], w2 = a.ev_words[3ull * j + 2ull];
    uint32_t p0 = ~w0, p1 = 0u;
    if (j > 0u) {
      p0 = a.ev_words[3ull * (j - 1u)];
      p1 = a.ev_words[3ull * (j - 1u) + 1ull];
    }
    const uint32_t x = w0 & 0xFFFFu, y = w0 >> 16, cc = w1 & 0xFFu;
    chunk = a.chunk_rows_magic ? mulhi_u32_u64(y, a.chunk_rows_magic) : y;
    if (chunk >= a.n_chunks) break;                    /* malformed event: silently ignored, driver.rs:441-444 */
    if (p0 == w0 && (p1 & 0xFFu) == cc) break;         /* only the first event of a run of same-pixel events works */
    const uint32_t channel = cc == ADDER_C_NONE ? 0u : cc;
    if (x >= a.W || y >= a.H || channel >= a.C) break;
    const unsigned long long gi = ((unsigned long long)y * a.W + x) * a.C + channel;
    const uint4 st = a.px_state[gi];
    unsigned long long running_ts = framer_state_ts(st);
    long long last_filled = framer_state_last_filled(st);
    uint32_t intensity = framer_state_intensity(st);
    bool force = false;
    uint32_t e1 = w1, e2 = w2;
    for (uint32_t e = j;; e++) {
      /* the next record is requested before this one is worked on */
      const bool has_next = e + 1u < total;
      uint32_t n0 = ~w0, n1 = 0u, n2 = 0u;
      if (has_next) {
        n0 = a.ev_words[3ull * (e + 1u)];
        n1 = a.ev_words[3ull * (e + 1u) + 1ull];
        n2 = a.ev_words[3ull * (e + 1u) + 2ull];
      }
      const uint32_t d = (e1 >> 8) & 0xFFu;
      uint32_t t = e2;
      const long long prev_last_filled = last_filled;
      const unsigned long long prev_running_ts = running_ts;
      bool skip = false;
      if (a.codec_version >= 2u && a.absolute_t) { /* :1002-1012 */
        if (prev_running_ts >= (unsigned long long)t) skip = true; else running_ts = t;
      } else {
        running_ts += t;
      }
      if (!skip) {
        const long long fidx = (long long)udiv_fast(running_ts ? running_ts - 1ull : 0ull, a.tpf, a.tpf_magic);
        if (fidx > last_filled) { /* :1014 */
          if (d != ADDER_D_EMPTY) {
            if (a.codec_version >= 2u && a.absolute_t && a.view_mode != 3u) {
              const uint32_t prev32 = (uint32_t)prev_running_ts;
              t = prev32 > t ? 0u : t - prev32; /* saturating_sub :1027 */
            }
            intensity = framer_value_u8(a.view_mode, d, t, a.ref_interval, a.exact_lut, a.practical_d_max, a.source_dtm, (uint32_t)running_ts,
                                        (uint32_t)prev_running_ts);
          }
          last_filled = fidx;
          if (last_filled > reach) reach = last_filled;
          for (long long i = prev_last_filled; i < last_filled; i++) { /* :1078-1091: absolute frame i + 1 */
            const long long fa = i + 1;
            if (fa < a.frames_written) continue;
            if (fa - a.frames_written >= (long long)a.ring_frames) { /* beyond what the ring can hold */
              atomicOr(a.err, 1u);
              break;
            }
            const unsigned long long slot = (unsigned long long)urem_fast((unsigned long long)fa, a.ring_frames, a.ring_magic) * a.frame_px + gi;
            if (!a.ring_some[slot]) {
              a.ring_some[slot] = 1;
              a.ring_val[slot] = (uint8_t)intensity;
            }
          }
        }
        if (a.codec_version >= 1u && a.framed_source) { /* :1094-1113 */
          const uint32_t over = urem_fast(running_ts, a.ref_interval, a.ref_magic);
          if (over) running_ts += a.ref_interval - over; /* = (running_ts / ref_interval + 1) * ref_interval */
        }
        if (a.buffer_limit >= 0 && last_filled > a.frames_written + a.buffer_limit) force = true; /* :1115-1121 */
      }
      if (!(n0 == w0 && (n1 & 0xFFu) == cc)) break; /* the run ends (or the stream does) */
      e1 = n1;
      e2 = n2;
    }
    a.px_state[gi] = framer_state_pack(running_ts, last_filled, intensity);
    worked = true;
    if (force) a.forced_frame[chunk] = a.frames_written; /* deque index 0 == absolute frame frames_written */
    } while (false);
    framer_raise_offset_max(a.offset_max, worked, chunk, reach);
  }
}

/* chunk_filled_tracker bookkeeping + the two predicates the host needs (run by one CTA of 256 threads):
 *   mode 0, after ingest_events_events (driver.rs:598 `*chunk_filled = filled`): only chunks that had events take their status;
 *   mode 1, after a pop (:923-924): every chunk takes the status of the new front frame;
 *   mode 2, flush_frame_buffer (:633-680): all true when some chunk holds more than one frame, else tracker[0] = false.
 * result[0] = is_frame_0_filled() (:851-866), result[1] = is_frame_filled(0) (:820-848), result[2] = any chunk longer than one frame */
__device__ __forceinline__ void framer_tracker_update(uint8_t* tracker, const uint8_t* status, const uint32_t* chunk_off, uint32_t n_chunks, int mode,
                                                      const long long* offset_max, long long frames_written, long long buffer_limit, uint32_t* result) {
  int not_tracked = 0, not_filled = 0, over_limit = 0, multi = 0;
  for (uint32_t k = threadIdx.x; k < n_chunks; k += blockDim.x) {
    const long long om = __ldcg(&offset_max[k]);
    const long long off = om > frames_written ? om : frames_written;
    const long long len = off - frames_written + 1; /* frames the chunk's VecDeque holds */
    if (len > 1) multi = 1;
    const uint8_t st = __ldcg(&status[k]); /* written by other CTAs of this launch */
    if (mode == 0) {
      if (chunk_off[k + 1u] > chunk_off[k]) tracker[k] = st;
    } else if (mode == 1) {
      tracker[k] = st;
    }
    if (buffer_limit >= 0 && len > buffer_limit) over_limit = 1;
    if (!st) not_filled = 1;
  }
  multi = __syncthreads_or(multi);
  if (mode == 2) {
    for (uint32_t k = threadIdx.x; k < n_chunks; k += blockDim.x)
      if (multi) tracker[k] = 1;
    if (!multi && threadIdx.x == 0) tracker[0] = 0;
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < n_chunks; k += blockDim.x)
    if (!tracker[k]) not_tracked = 1;
  not_tracked = __syncthreads_or(not_tracked);
  not_filled = __syncthreads_or(not_filled);
  over_limit = __syncthreads_or(over_limit);
  if (threadIdx.x == 0) {
    result[0] = (over_limit || !not_tracked) ? 1u : 0u;
    result[1] = not_filled ? 0u : 1u;
    result[2] = multi ? 1u : 0u;
  }
}

/* Per-chunk status of absolute frame `fa`: status[k] = 1 when every pixel of chunk k is Some, or the chunk's
 * filled_count was forced to full for that frame (driver.rs:820-848 is_frame_filled, per chunk).  One CTA per chunk, 128-bit
 * loads where the chunk allows; the CTA that finishes last (a counter in device memory, reset for the next launch) runs the
 * tracker update above, so a refresh is one launch instead of two. */
__global__ void __launch_bounds__(256) framer_chunk_status_kernel(const uint8_t* __restrict__ ring_some, uint32_t ring_frames,
                                                                  unsigned long long frame_px, unsigned long long chunk_px,
                                                                  const long long* forced_frame, long long fa, uint8_t* status,
                                                                  uint32_t* done_count, uint8_t* tracker, const uint32_t* chunk_off, int mode,
                                                                  const long long* offset_max, long long buffer_limit, uint32_t* result) {
  const uint32_t k = blockIdx.x, n_chunks = gridDim.x;
  const unsigned long long begin = (unsigned long long)k * chunk_px;
  const unsigned long long end = begin + chunk_px < frame_px ? begin + chunk_px : frame_px;
  const uint8_t* some = ring_some + (unsigned long long)(fa % (long long)ring_frames) * frame_px;
  int empty = 0;
  if ((((uintptr_t)(some + begin)) & 15u) == 0u) { /* is-some bytes are 0 or 1: a 16-byte group is full iff every byte is 1 */
    const unsigned long long n16 = (end - begin) >> 4;
    const uint4* p = reinterpret_cast<const uint4*>(some + begin);
    for (unsigned long long i = threadIdx.x; i < n16; i += blockDim.x) {
      const uint4 v = __ldcg(p + i);
      empty |= (v.x & v.y & v.z & v.w) != 0x01010101u;
    }
    for (unsigned long long i = begin + (n16 << 4) + threadIdx.x; i < end; i += blockDim.x) empty |= !__ldcg(some + i);
  } else {
    for (unsigned long long i = begin + threadIdx.x; i < end; i += blockDim.x) empty |= !__ldcg(some + i);
  }
  const int any_empty = __syncthreads_or(empty);
  __shared__ uint32_t s_last;
  if (threadIdx.x == 0) {
    status[k] = (!any_empty || forced_frame[k] == fa) ? 1 : 0;
    __threadfence();
    s_last = atomicAdd(done_count, 1u) == n_chunks - 1u ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) { /* CTA-uniform */
    __threadfence();
    if (threadIdx.x == 0) *done_count = 0u;
    framer_tracker_update(tracker, status, chunk_off, n_chunks, mode, offset_max, fa, buffer_limit, result);
  }
}

/* write_frame_bytes (driver.rs:936-962) for n consecutive frames: Some(v) -> v, None -> 0; the slots are cleared
 * (a popped frame's storage is reused for a later absolute frame). */
__global__ void __launch_bounds__(256) framer_pop_kernel(uint8_t* __restrict__ ring_val, uint8_t* __restrict__ ring_some, uint32_t ring_frames,
                                                         unsigned long long frame_px, long long first, uint32_t n, uint8_t* __restrict__ out) {
  const unsigned long long total = frame_px * n;
  for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long m = k / frame_px, i = k - m * frame_px;
    const unsigned long long slot = (unsigned long long)((first + (long long)m) % (long long)ring_frames) * frame_px + i;
    out[k] = ring_some[slot] ? ring_val[slot] : 0;
    ring_some[slot] = 0;
    ring_val[slot] = 0;
  }
}

/* flush_frame_buffer (driver.rs:633-680): every empty pixel of the frame at the front takes its last intensity */
__global__ void __launch_bounds__(256) framer_flush_kernel(uint8_t* __restrict__ ring_val, uint8_t* __restrict__ ring_some, uint32_t ring_frames,
                                                           unsigned long long frame_px, long long front, uint4* __restrict__ px_state) {
  const unsigned long long base = (unsigned long long)(front % (long long)ring_frames) * frame_px;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < frame_px; i += (unsigned long long)gridDim.x * blockDim.x) {
    if (!ring_some[base + i]) {
      const uint4 st = px_state[i];
      ring_some[base + i] = 1;
      ring_val[base + i] = (uint8_t)framer_state_intensity(st);
      px_state[i] = framer_state_pack(framer_state_ts(st), framer_state_last_filled(st) + 1, framer_state_intensity(st));
    }
  }
}

__global__ void framer_init_kernel(uint4* px_state, unsigned long long n, long long* offset_max, long long* forced_frame, uint32_t n_chunks) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) px_state[i] = framer_state_pack(0ull, -1, 0u); /* last_filled_frame_ref = -1, driver.rs:349-353 */
  if (i < n_chunks) {
    offset_max[i] = 0;
    forced_frame[i] = -1;
  }
}

}  // namespace adder
