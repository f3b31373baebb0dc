/*
 * raw_kernel.cuh — event records -> raw .adder wire records on the device.
 *
 * What the reference does serially after the parallel section, one event at a time:
 *   Video::integrate_matrix   adder-codec-rs/src/transcoder/source/video.rs:736-740   (encoder.ingest_event per event)
 *   RawOutput::ingest_event   adder-codec-core/src/codec/raw/stream.rs:100-120        (bincode fixint big-endian)
 * A single-channel plane writes EventSingle {x:u16, y:u16, d:u8, t:u32} = 9 bytes, any other plane
 * Event {x, y, c: Some(u8) -> tag 1 + value, d, t} = 11 bytes (codec/header.rs:77-81).
 *
 * One thread serialises one record into shared memory; the CTA then stores its 256 records
 * (2304 or 2816 bytes, a whole number of 32-bit words) with coalesced word stores.  The event count
 * is read from device memory (the integrate kernel's last chunk offset), so the launch needs no
 * host round trip.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gray_math.h"
#include "raw_pack.h"

namespace adder {

constexpr uint32_t kRawThreads = 256;
constexpr uint32_t kRawChunk = 4u * kRawThreads; /* records per CTA iteration of raw_encode_kernel: four per thread */

/* One CTA iteration = 1024 records: the 12 KB of records come in with 128-bit loads, every thread packs four records
 * (raw_pack4) into shared memory, and the 9 or 11 KB of wire bytes leave with 128-bit stores.  (The first version wrote
 * every wire byte with its own shared-memory store and read the records with three strided 32-bit loads per thread:
 * 0.62 of the HBM roofline, profiles/r01q_next_rows.txt.) */
template <uint32_t ESIZE>
__device__ __forceinline__ void raw_encode_body(const uint32_t* __restrict__ ev_words, unsigned long long n, uint8_t* __restrict__ out, uint32_t* s_out) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(ev_words) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0u;
  for (unsigned long long base = (unsigned long long)blockIdx.x * kRawChunk; base < n; base += (unsigned long long)gridDim.x * kRawChunk) {
    const uint32_t cnt = (uint32_t)min((unsigned long long)kRawChunk, n - base);
    const unsigned long long i = base + 4ull * threadIdx.x; /* this thread's four records */
    uint32_t in[12];
    if (i + 4ull <= n && aligned) { /* base is a multiple of 1024 records: 48-byte groups are 16-byte aligned */
      const uint4* p = reinterpret_cast<const uint4*>(ev_words + i * 3ull);
      const uint4 a = __ldcs(p), b = __ldcs(p + 1), c = __ldcs(p + 2);
      in[0] = a.x, in[1] = a.y, in[2] = a.z, in[3] = a.w, in[4] = b.x, in[5] = b.y, in[6] = b.z, in[7] = b.w, in[8] = c.x, in[9] = c.y, in[10] = c.z, in[11] = c.w;
    } else {
#pragma unroll
      for (uint32_t k = 0; k < 12u; k++) in[k] = i * 3ull + k < n * 3ull ? ev_words[i * 3ull + k] : 0u;
    }
    uint32_t o[ESIZE];
#pragma unroll
    for (uint32_t k = 0; k < ESIZE; k++) o[k] = 0u;
    raw_pack4<ESIZE>(in, o);
#pragma unroll
    for (uint32_t k = 0; k < ESIZE; k++) s_out[threadIdx.x * ESIZE + k] = o[k]; /* stride 9 / 11 words: odd, no bank conflicts */
    __syncthreads();
    const uint32_t nbytes = cnt * ESIZE;
    uint8_t* dst = out + base * ESIZE; /* base * ESIZE is a multiple of 16 */
    if (aligned) {
      const uint32_t nvec = nbytes >> 4;
      for (uint32_t j = threadIdx.x; j < nvec; j += kRawThreads) __stcs(reinterpret_cast<uint4*>(dst) + j, reinterpret_cast<const uint4*>(s_out)[j]);
      for (uint32_t j = (nvec << 4) + threadIdx.x; j < nbytes; j += kRawThreads) dst[j] = reinterpret_cast<const uint8_t*>(s_out)[j];
    } else {
      for (uint32_t j = threadIdx.x; j < nbytes; j += kRawThreads) dst[j] = reinterpret_cast<const uint8_t*>(s_out)[j];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kRawThreads) raw_encode_kernel(const uint32_t* __restrict__ ev_words, const uint32_t* __restrict__ n_events_ptr,
                                                                 unsigned long long n_events_max, uint32_t esize /* 9 or 11 */,
                                                                 uint8_t* __restrict__ out) {
  __shared__ __align__(16) uint32_t s_out[kRawThreads * 11];
  const unsigned long long n = min((unsigned long long)*n_events_ptr, n_events_max);
  if (esize == 11u)
    raw_encode_body<11u>(ev_words, n, out, s_out);
  else
    raw_encode_body<9u>(ev_words, n, out, s_out);
}

/*
 * compact_encode_kernel — the frame's records in the compact host form (include/adder_b200.h, "compact form"): what crosses
 * PCIe when the caller asks for integrate_frames_host_compact.  A record's coordinates are implied by raster order, so
 *   dense  (4 * E > P):  P count bytes (events of every pixel-channel, raster order) | E x {d:u8, t:u32 LE}      P + 5 E bytes
 *   sparse (otherwise):  E x {index:u32 LE (flat raster index in this plane), d:u8, t:u32 LE}                      9 E bytes
 * against 12 E bytes of records (noise at 1080p RGB: 5.7 instead of 11.3 bytes per pixel-frame).  The form follows from
 * E and P alone, so the host applies the same rule to the count it already has.  `out` must hold P zero bytes in front
 * (dense form: a pixel without events keeps its 0).  Same staging through shared memory as raw_encode_kernel.
 */
__global__ void __launch_bounds__(kRawThreads) compact_encode_kernel(const uint32_t* __restrict__ ev_words, const uint32_t* __restrict__ n_events_ptr,
                                                                     unsigned long long n_events_max, uint32_t P, uint32_t WC, uint32_t C, uint32_t row0,
                                                                     uint8_t* __restrict__ out) {
  __shared__ __align__(16) uint8_t s_bytes[kRawThreads * 9];
  const unsigned long long n = min((unsigned long long)*n_events_ptr, n_events_max);
  const bool dense = 4ull * n > (unsigned long long)P;
  const uint32_t esize = dense ? 5u : 9u;
  uint8_t* const body = dense ? out + P : out;
  auto index_of = [&](unsigned long long i) {
    const uint32_t w0 = ev_words[i * 3ull], w1 = ev_words[i * 3ull + 1ull];
    const uint32_t x = w0 & 0xFFFFu, y = (w0 >> 16) - row0, c = w1 & 0xFFu;
    return y * WC + x * C + (C == 1u ? 0u : c);
  };
  for (unsigned long long base = (unsigned long long)blockIdx.x * kRawThreads; base < n; base += (unsigned long long)gridDim.x * kRawThreads) {
    const unsigned long long i = base + threadIdx.x;
    if (i < n) {
      const uint32_t w1 = ev_words[i * 3ull + 1ull], t = ev_words[i * 3ull + 2ull];
      const uint32_t d = (w1 >> 8) & 0xFFu, idx = index_of(i);
      uint8_t* p = s_bytes + threadIdx.x * esize;
      if (!dense) {
        p[0] = (uint8_t)idx, p[1] = (uint8_t)(idx >> 8), p[2] = (uint8_t)(idx >> 16), p[3] = (uint8_t)(idx >> 24);
        p += 4;
      } else if (i == 0ull || index_of(i - 1ull) != idx) { /* first record of its pixel: the pixel's count is the length of its run */
        uint32_t run = 1;
        while (i + run < n && index_of(i + run) == idx) run++;
        out[idx] = (uint8_t)run; /* at most depth + 2 <= 33 */
      }
      p[0] = (uint8_t)d;
      p[1] = (uint8_t)t, p[2] = (uint8_t)(t >> 8), p[3] = (uint8_t)(t >> 16), p[4] = (uint8_t)(t >> 24);
    }
    __syncthreads();
    const uint32_t cnt = (uint32_t)min((unsigned long long)kRawThreads, n - base);
    const uint32_t nbytes = cnt * esize;
    uint8_t* dst = body + base * esize;
    /* body starts at out (+ P): word stores when that is word aligned (base is a multiple of 256) */
    if ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0) {
      const uint32_t nwords = nbytes >> 2;
      for (uint32_t j = threadIdx.x; j < nwords; j += kRawThreads) reinterpret_cast<uint32_t*>(dst)[j] = reinterpret_cast<const uint32_t*>(s_bytes)[j];
      for (uint32_t j = (nwords << 2) + threadIdx.x; j < nbytes; j += kRawThreads) dst[j] = s_bytes[j];
    } else {
      for (uint32_t j = threadIdx.x; j < nbytes; j += kRawThreads) dst[j] = s_bytes[j];
    }
    __syncthreads();
  }
}

/*
 * handle_color (adder-codec-rs/src/utils/cv.rs:215-232), the pre-step of Framed::consume for a gray
 * transcode of a colour source (framed.rs:129); the arithmetic is in gray_math.h.
 * Sixteen pixels per thread: three 128-bit loads in, one 128-bit store out (four pixels per thread with 32-bit accesses ran at
 * 0.57 of the HBM roofline, profiles/r01q_next_rows.txt).
 */
__constant__ uint8_t c_gray_diag[256]; /* gray_exact_f64(k, k, k), written once per device by the host */
__device__ __forceinline__ uint32_t gray_of(uint32_t c0, uint32_t c1, uint32_t c2) { return gray_of(c0, c1, c2, c_gray_diag); }
/* four pixels from three words: bytes r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3 */
__device__ __forceinline__ uint32_t gray4(uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t g0 = gray_of(a & 0xFFu, (a >> 8) & 0xFFu, (a >> 16) & 0xFFu);
  const uint32_t g1 = gray_of(a >> 24, b & 0xFFu, (b >> 8) & 0xFFu);
  const uint32_t g2 = gray_of((b >> 16) & 0xFFu, b >> 24, c & 0xFFu);
  const uint32_t g3 = gray_of((c >> 8) & 0xFFu, (c >> 16) & 0xFFu, c >> 24);
  return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
}
constexpr uint32_t kGrayPxPerThread = 16;
__global__ void __launch_bounds__(256) rgb_to_gray_kernel(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ gray, unsigned long long n_px) {
  const unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; /* group of sixteen pixels */
  const unsigned long long i = q * kGrayPxPerThread;
  if (i >= n_px) return;
  if (i + kGrayPxPerThread <= n_px && ((reinterpret_cast<uintptr_t>(rgb) | reinterpret_cast<uintptr_t>(gray)) & 15u) == 0) {
    const uint4* w = reinterpret_cast<const uint4*>(rgb) + q * 3ull;
    const uint4 a = __ldcs(w), b = __ldcs(w + 1), c = __ldcs(w + 2); /* read once */
    uint4 o;
    o.x = gray4(a.x, a.y, a.z);
    o.y = gray4(a.w, b.x, b.y);
    o.z = gray4(b.z, b.w, c.x);
    o.w = gray4(c.y, c.z, c.w);
    reinterpret_cast<uint4*>(gray)[q] = o; /* read again right away by the integrate launch: a plain store */
  } else {
    for (unsigned long long j = i; j < n_px && j < i + kGrayPxPerThread; j++) gray[j] = (uint8_t)gray_of(rgb[3ull * j], rgb[3ull * j + 1ull], rgb[3ull * j + 2ull]);
  }
}

}  // namespace adder
