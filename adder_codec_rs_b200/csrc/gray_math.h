/*
 * gray_math.h — handle_color (adder-codec-rs/src/utils/cv.rs:215-232): gray = (ch0*0.114 + ch1*0.587 + ch2*0.299) as u8,
 * evaluated by the reference in f64, left to right, every product and sum rounded on its own, then truncated and
 * saturated.  Shared by the device kernel and the host tests (tests/host_sim).
 *
 * The f64 pipe and its int<->double conversions are slow on this GPU (the conversion kernel ran at 0.5 TB/s with
 * them), so the byte is taken from 2^24 fixed point whenever that is certain to agree:
 *   S = c0*A0 + c1*A1 + c2*A2 with A = ceil(w * 2^24); every A exceeds w*2^24 by less than 0.42, the three excesses
 *   add up to 1.0002, so S - 255.1 <= X*2^24 <= S for the real-valued X = c0*0.114 + c1*0.587 + c2*0.299
 *   (the f64 constants differ from the decimals by < 2^-56 and the f64 roundings move the result by < 2^-44: both far
 *   below one unit of 2^-24).
 *   If the low 24 bits of S lie in [257, 2^24 - 2], X and the reference's f64 value share the integer part S >> 24.
 * Otherwise (16 in a million on noise, but EVERY gray pixel c0 == c1 == c2 = k, where X is k itself up to rounding):
 * the diagonal comes from a 256-entry table of the reference's own f64 expression, the rest from that expression.
 */
#ifndef ADDER_B200_GRAY_MATH_H
#define ADDER_B200_GRAY_MATH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define ADDER_GRAY_HD __host__ __device__ __forceinline__
#else
#define ADDER_GRAY_HD inline
#endif

namespace adder {

/* the reference's expression, operation by operation */
ADDER_GRAY_HD uint32_t gray_exact_f64(uint32_t c0, uint32_t c1, uint32_t c2) {
#if defined(__CUDA_ARCH__)
  const double s = __dadd_rn(__dadd_rn(__dmul_rn((double)c0, 0.114), __dmul_rn((double)c1, 0.587)), __dmul_rn((double)c2, 0.299));
  const uint32_t u = __double2uint_rz(s);
#else
  volatile double a = (double)c0 * 0.114, b = (double)c1 * 0.587, c = (double)c2 * 0.299;
  volatile double s = a + b;
  s = s + c;
  const uint32_t u = s > 0.0 ? (uint32_t)s : 0u;
#endif
  return u > 255u ? 255u : u;
}

/* diag[k] = gray_exact_f64(k, k, k), built once on the host */
inline void build_gray_diag(uint8_t diag[256]) {
  for (uint32_t k = 0; k < 256u; k++) diag[k] = (uint8_t)gray_exact_f64(k, k, k);
}

ADDER_GRAY_HD uint32_t gray_of(uint32_t c0, uint32_t c1, uint32_t c2, const uint8_t* diag) {
  const uint32_t S = c0 * 1912603u + c1 * 9848226u + c2 * 5016388u; /* <= 255 * 16777217 < 2^32 */
  const uint32_t fr = S & 0xFFFFFFu;
  if (fr >= 257u && fr <= 0xFFFFFEu) return S >> 24;
  if (c0 == c1 && c1 == c2) return diag[c0];
  return gray_exact_f64(c0, c1, c2);
}

}  // namespace adder
#endif
