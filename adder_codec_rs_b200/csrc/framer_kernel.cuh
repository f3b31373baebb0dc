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

/* n / d and n % d for a 64-bit n that almost always fits 32 bits (timestamps, frame numbers): the 64-bit division is a
 * long dependent sequence, and this kernel is bound by latency */
__device__ __forceinline__ unsigned long long udiv_fast(unsigned long long n, uint32_t d) {
  return (n >> 32) == 0ull ? (unsigned long long)((uint32_t)n / d) : n / d;
}
__device__ __forceinline__ uint32_t urem_fast(unsigned long long n, uint32_t d) {
  return (n >> 32) == 0ull ? (uint32_t)n % d : (uint32_t)(n % d);
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

__global__ void __launch_bounds__(256) framer_ingest_kernel(const FramerArgs a) {
  const uint32_t total = a.chunk_off[a.n_chunks];
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
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
    const uint32_t chunk = y / a.chunk_rows;
    if (chunk >= a.n_chunks) continue;                 /* malformed event: silently ignored, driver.rs:441-444 */
    if (p0 == w0 && (p1 & 0xFFu) == cc) continue;      /* only the first event of a run of same-pixel events works */
    const uint32_t channel = cc == ADDER_C_NONE ? 0u : cc;
    if (x >= a.W || y >= a.H || channel >= a.C) continue;
    const unsigned long long gi = ((unsigned long long)y * a.W + x) * a.C + channel;
    const uint4 st = a.px_state[gi];
    unsigned long long running_ts = framer_state_ts(st);
    long long last_filled = framer_state_last_filled(st);
    uint32_t intensity = framer_state_intensity(st);
    long long reach = -1;
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
        const long long fidx = (long long)udiv_fast(running_ts ? running_ts - 1ull : 0ull, a.tpf);
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
            const unsigned long long slot = (unsigned long long)urem_fast((unsigned long long)fa, a.ring_frames) * a.frame_px + gi;
            if (!a.ring_some[slot]) {
              a.ring_some[slot] = 1;
              a.ring_val[slot] = (uint8_t)intensity;
            }
          }
        }
        if (a.codec_version >= 1u && a.framed_source) { /* :1094-1113 */
          const uint32_t over = urem_fast(running_ts, a.ref_interval);
          if (over) running_ts += a.ref_interval - over; /* = (running_ts / ref_interval + 1) * ref_interval */
        }
        if (a.buffer_limit >= 0 && last_filled > a.frames_written + a.buffer_limit) force = true; /* :1115-1121 */
      }
      if (!(n0 == w0 && (n1 & 0xFFu) == cc)) break; /* the run ends (or the stream does) */
      e1 = n1;
      e2 = n2;
    }
    a.px_state[gi] = framer_state_pack(running_ts, last_filled, intensity);
    /* every pixel of a chunk reaches about the same frame, so nearly all of these would be atomics that change nothing
     * on one contended word per chunk (they were 70 % of the kernel's time): look first.  The word only grows, so a
     * stale read can at worst cause an atomic that was not needed. */
    if (reach >= 0 && reach > __ldcg(&a.offset_max[chunk])) atomicMax(&a.offset_max[chunk], reach);
    if (force) a.forced_frame[chunk] = a.frames_written; /* deque index 0 == absolute frame frames_written */
  }
}

/* Per-chunk status of absolute frame `fa`: status[k] = 1 when every pixel of chunk k is Some, or the chunk's
 * filled_count was forced to full for that frame (driver.rs:820-848 is_frame_filled, per chunk).  One CTA per chunk. */
__global__ void __launch_bounds__(256) framer_chunk_status_kernel(const uint8_t* __restrict__ ring_some, uint32_t ring_frames,
                                                                  unsigned long long frame_px, unsigned long long chunk_px,
                                                                  const long long* __restrict__ forced_frame, long long fa, uint8_t* status) {
  const uint32_t k = blockIdx.x;
  const unsigned long long begin = (unsigned long long)k * chunk_px;
  const unsigned long long end = begin + chunk_px < frame_px ? begin + chunk_px : frame_px;
  const uint8_t* some = ring_some + (unsigned long long)(fa % (long long)ring_frames) * frame_px;
  int empty = 0;
  for (unsigned long long i = begin + threadIdx.x; i < end; i += blockDim.x) empty |= !some[i];
  const int any_empty = __syncthreads_or(empty);
  if (threadIdx.x == 0) status[k] = (!any_empty || forced_frame[k] == fa) ? 1 : 0;
}

/* chunk_filled_tracker bookkeeping + the two predicates the host needs (one CTA):
 *   mode 0, after ingest_events_events (driver.rs:598 `*chunk_filled = filled`): only chunks that had events take their status;
 *   mode 1, after a pop (:923-924): every chunk takes the status of the new front frame;
 *   mode 2, flush_frame_buffer (:633-680): all true when some chunk holds more than one frame, else tracker[0] = false.
 * result[0] = is_frame_0_filled() (:851-866), result[1] = is_frame_filled(0) (:820-848), result[2] = any chunk longer than one frame */
__global__ void __launch_bounds__(256) framer_tracker_kernel(uint8_t* tracker, const uint8_t* status, const uint32_t* chunk_off, uint32_t n_chunks,
                                                             int mode, const long long* offset_max, long long frames_written,
                                                             long long buffer_limit, uint32_t* result) {
  int not_tracked = 0, not_filled = 0, over_limit = 0, multi = 0;
  for (uint32_t k = threadIdx.x; k < n_chunks; k += blockDim.x) {
    const long long off = offset_max[k] > frames_written ? offset_max[k] : frames_written;
    const long long len = off - frames_written + 1; /* frames the chunk's VecDeque holds */
    if (len > 1) multi = 1;
    if (mode == 0) {
      if (chunk_off[k + 1u] > chunk_off[k]) tracker[k] = status[k];
    } else if (mode == 1) {
      tracker[k] = status[k];
    }
    if (buffer_limit >= 0 && len > buffer_limit) over_limit = 1;
    if (!status[k]) not_filled = 1;
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
