/*
 * synth.cuh — synthetic input frames generated on the device (bench / test input).
 * Counter-based, so the host (tests/synth.py, numpy) reproduces every byte without shared RNG state:
 *   h(seed, f, i) = splitmix64(seed ^ (f << 40) ^ i) & 0xFF,   i = flat raster index (y*W + x)*C + c
 */
#pragma once
#include <stdint.h>

namespace adder {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline uint32_t synth_hash(uint64_t seed, uint32_t f, uint32_t i) {
  return (uint32_t)(splitmix64(seed ^ ((uint64_t)f << 40) ^ (uint64_t)i) & 0xFFu);
}

/* kind 0: gradient (x + 2y + 3f) & 255 (contains 0 -> exercises D = 128)
 * kind 1: uniform noise h(seed, f, i)
 * kind 2: base h(seed^1, 0, i) plus per-frame jitter (h(seed^2, f, i) % 21) - 10, clamped to 0..255
 * kind 3: static base h(seed^1, 0, i); where h(seed^3, f, i) < 2 the sample is h(seed^4, f, i) for that frame */
__host__ __device__ inline uint8_t synth_value(int kind, uint64_t seed, uint32_t f, uint32_t i, uint32_t W, uint32_t C) {
  switch (kind) {
    case 0: {
      const uint32_t p = i / C, x = p % W, y = p / W;
      return (uint8_t)((x + 2u * y + 3u * f) & 255u);
    }
    case 1: return (uint8_t)synth_hash(seed, f, i);
    case 2: {
      const int v = (int)synth_hash(seed ^ 1ull, 0, i) + (int)(synth_hash(seed ^ 2ull, f, i) % 21u) - 10;
      return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
    default: {
      if (synth_hash(seed ^ 3ull, f, i) < 2u) return (uint8_t)synth_hash(seed ^ 4ull, f, i);
      return (uint8_t)synth_hash(seed ^ 1ull, 0, i);
    }
  }
}

/* i0: flat index of the plane's first sample inside the whole frame (a row band starts at row0 * W * C), so that a band
 * gets exactly the rows of the undivided frame */
__global__ void synth_frame_kernel(uint8_t* out, uint32_t P, uint32_t W, uint32_t C, uint32_t f, int kind, uint64_t seed, uint32_t i0) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) out[i] = synth_value(kind, seed, f, i + i0, W, C);
}

}  // namespace adder
