/*
 * feature_kernel.cuh — the one cross-pixel step of Video::integrate_matrix: handle_features
 * (adder-codec-rs/src/transcoder/source/video.rs:883-1113), run after the integrate kernel of a frame
 * when feature detection is on.
 *
 *   is_feature            adder-codec-rs/src/utils/cv.rs:22-212   asynchronous FAST 9_16 on channel 0 of running_intensities
 *   feature sets          video.rs:894-918   for the last event of every pixel's run in a chunk (e1.coord != e2.coord over
 *                                            circular pairs), channel None/0, not D_EMPTY: insert if it is a feature, else remove
 *   c_thresh reset        video.rs:1077-1104 every pixel within feature_c_radius of a NEWLY inserted feature gets
 *                                            c_thresh = min(c_thresh_baseline, 2)
 *
 * A pixel's events of a frame are contiguous in the stream, so every pixel is examined at most once per
 * frame: the reference's sequential walk has no order dependence and one thread per event reproduces it.
 * The per-chunk HashSet<Coord> becomes one byte per (x, y).  Logging, drawing and DBSCAN clustering of
 * the reference only feed the GUI and are not here.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "state_layout.h"

namespace adder {

__device__ __forceinline__ bool is_feature_dev(const uint8_t* __restrict__ img, uint32_t W, uint32_t H, uint32_t C, uint32_t cx, uint32_t cy) {
  constexpr int kThr = 30, kStreak = 9; /* INTENSITY_THRESHOLD cv.rs:22, STREAK_SIZE :33 */
  if (cx < 3u || cx + 3u >= W || cy < 3u || cy + 3u >= H) return false; /* Coord::is_border(w, h, 3), lib.rs:353-358 */
  const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1}; /* CIRCLE3 :26-31 */
  const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  const long width = (long)W * C;
  const uint8_t* ctr = img + (long)cy * width + (long)cx * C;
  const int candidate = *ctr;
  int px[16];
#pragma unroll
  for (int k = 0; k < 16; k++) px[k] = ctr[dy[k] * width + dx[k] * (long)C];
  auto tab = [&](int v) { return v - candidate < -kThr ? 1 : (v - candidate > kThr ? 2 : 0); }; /* THRESHOLD_TABLE :35-50 */
  int d = tab(px[0]) | tab(px[8]);
  if (d == 0) return false;
  d &= tab(px[2]) | tab(px[10]);
  d &= tab(px[4]) | tab(px[12]);
  d &= tab(px[6]) | tab(px[14]);
  if (d == 0) return false;
  d &= tab(px[1]) | tab(px[9]);
  d &= tab(px[3]) | tab(px[11]);
  d &= tab(px[5]) | tab(px[13]);
  d &= tab(px[7]) | tab(px[15]);
  if (d & 1) { /* dark streak :142-172 */
    const int vt = candidate - kThr;
    int count = 0;
#pragma unroll
    for (int k = 0; k < 25; k++) {
      if (px[k & 15] < vt) {
        if (++count == kStreak) return true;
      } else {
        count = 0;
        if (k == 17) return false; /* :167-169: for the whole function */
      }
    }
  }
  if (d & 2) { /* bright streak :174-205 */
    const int vt = candidate + kThr;
    int count = 0;
#pragma unroll
    for (int k = 0; k < 25; k++) {
      if (px[k & 15] > vt) {
        if (++count == kStreak) return true;
      } else {
        count = 0;
        if (k == 17) return false;
      }
    }
  }
  return false;
}

/* one thread per event of the frame; n_events and the chunk offsets are read from device memory */
__global__ void __launch_bounds__(256) feature_kernel(const uint32_t* __restrict__ ev_words, const uint32_t* __restrict__ chunk_off,
                                                      uint32_t n_chunks, uint32_t chunk_rows, uint32_t row0,
                                                      const uint8_t* __restrict__ running, uint32_t W, uint32_t H, uint32_t C,
                                                      uint8_t* __restrict__ mask, uint32_t* __restrict__ new_xy, uint32_t* n_new,
                                                      uint32_t new_cap, uint8_t* __restrict__ new_mask) {
  const uint32_t total = chunk_off[n_chunks];
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
    const uint32_t w0 = ev_words[3ull * j], w1 = ev_words[3ull * j + 1ull];
    const uint32_t x = w0 & 0xFFFFu, y = (w0 >> 16) - row0, c = w1 & 0xFFu, d = (w1 >> 8) & 0xFFu;
    if (!(c == ADDER_C_NONE || c == 0u) || d == ADDER_D_EMPTY) continue;
    const uint32_t chunk = y / chunk_rows, lo = chunk_off[chunk], hi = chunk_off[chunk + 1u];
    const uint32_t jn = j + 1u < hi ? j + 1u : lo; /* circular_tuple_windows inside the chunk, video.rs:898 */
    const uint32_t v0 = ev_words[3ull * jn], v1 = ev_words[3ull * jn + 1ull];
    if (v0 == w0 && (v1 & 0xFFu) == c) continue; /* e1.coord == e2.coord: not the last event of the pixel's run */
    const uint32_t p = y * W + x;
    if (is_feature_dev(running, W, H, C, x, y)) {
      if (mask[p] == 0) { /* HashSet::insert returned true: a new feature */
        mask[p] = 1;
        if (new_mask) new_mask[p] = 1;
        /* one atomic per group of lanes that got here together (the list is unordered anyway): on noise every seventh
         * pixel is a new feature and they all count on the one word */
        const uint32_t grp = __activemask(), lane = threadIdx.x & 31u, leader = (uint32_t)__ffs((int)grp) - 1u;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(n_new, (uint32_t)__popc(grp));
        base = __shfl_sync(grp, base, (int)leader);
        const uint32_t k = base + (uint32_t)__popc(grp & ((1u << lane) - 1u));
        if (k < new_cap) new_xy[k] = x | (y << 16);
      }
    } else {
      mask[p] = 0;
    }
  }
}

/* The reset of video.rs:1089-1104 costs (2r+1)^2 writes per new feature.  With few features (any real scene) that is
 * the cheapest form; when features are dense (noise: every seventh pixel) the union of the windows is most of the
 * plane and is found instead by dilating the bitmap of new features, first along rows, then along columns.  Both
 * forms are launched; each looks at the count and leaves when it is the other's turn. */
__device__ __forceinline__ bool reset_by_dilation(uint32_t n, int radius, uint32_t W, uint32_t H, int force) {
  if (force) return force > 0;
  const unsigned long long win = (unsigned long long)(2 * radius + 1) * (unsigned long long)(2 * radius + 1);
  return (unsigned long long)n * win >= 16ull * W * H;
}

/* c_thresh := value for every pixel-channel within `radius` of each new feature (video.rs:1089-1104).
 * One CTA per feature (grid-stride), threads over the clipped window. */
__global__ void __launch_bounds__(256) feature_reset_kernel(uint2* __restrict__ hdr, const uint32_t* __restrict__ new_xy,
                                                            const uint32_t* __restrict__ n_new, uint32_t new_cap, uint32_t W, uint32_t H,
                                                            uint32_t C, int radius, uint32_t value, int force) {
  const uint32_t n = min(*n_new, new_cap);
  if (reset_by_dilation(n, radius, W, H, force)) return;
  for (uint32_t k = blockIdx.x; k < n; k += gridDim.x) {
    const int fx = (int)(new_xy[k] & 0xFFFFu), fy = (int)(new_xy[k] >> 16);
    const int r0 = max(fy - radius, 0), r1 = min(fy + radius, (int)H - 1);
    const int c0 = max(fx - radius, 0), c1 = min(fx + radius, (int)W - 1);
    const uint32_t per_row = (uint32_t)(c1 - c0 + 1) * C, cells = per_row * (uint32_t)(r1 - r0 + 1);
    for (uint32_t t = threadIdx.x; t < cells; t += blockDim.x) {
      const uint32_t ry = t / per_row, rr = t - ry * per_row;
      const uint32_t i = (((uint32_t)r0 + ry) * W + (uint32_t)c0) * C + rr;
      const uint32_t y = hdr[i].y;
      hdr[i].y = (y & ~0xFF00u) | (value << 8); /* several features may write the same pixel: the same value */
    }
  }
}

/* dilation, pass 1: new_mask (1 where a feature was inserted this frame) -> row_hit (1 where some new feature of
 * the same row lies within `radius` columns) */
__global__ void __launch_bounds__(256) feature_dilate_rows_kernel(const uint8_t* __restrict__ new_mask, uint8_t* __restrict__ row_hit,
                                                                  const uint32_t* __restrict__ n_new, uint32_t new_cap, uint32_t W, uint32_t H,
                                                                  int radius, int force) {
  if (!reset_by_dilation(min(*n_new, new_cap), radius, W, H, force)) return;
  const uint32_t P = W * H;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    const int y = (int)(p / W), x = (int)(p - (uint32_t)y * W);
    const int c0 = max(x - radius, 0), c1 = min(x + radius, (int)W - 1);
    const uint8_t* row = new_mask + (size_t)y * W;
    uint8_t hit = 0;
    for (int c = c0; c <= c1 && !hit; c++) hit = row[c];
    row_hit[p] = hit;
  }
}
/* pass 2: a pixel is reset when some row within `radius` rows has row_hit at its column */
__global__ void __launch_bounds__(256) feature_dilate_cols_kernel(uint2* __restrict__ hdr, const uint8_t* __restrict__ row_hit,
                                                                  const uint32_t* __restrict__ n_new, uint32_t new_cap, uint32_t W, uint32_t H,
                                                                  uint32_t C, int radius, uint32_t value, int force) {
  if (!reset_by_dilation(min(*n_new, new_cap), radius, W, H, force)) return;
  const uint32_t P = W * H;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    const int y = (int)(p / W), x = (int)(p - (uint32_t)y * W);
    const int r0 = max(y - radius, 0), r1 = min(y + radius, (int)H - 1);
    uint8_t hit = 0;
    for (int r = r0; r <= r1 && !hit; r++) hit = row_hit[(size_t)r * W + x];
    if (hit) {
      for (uint32_t c = 0; c < C; c++) {
        const uint32_t i = p * C + c;
        hdr[i].y = (hdr[i].y & ~0xFF00u) | (value << 8);
      }
    }
  }
}

}  // namespace adder
