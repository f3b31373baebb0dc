/*
 * adder_b200.cu — C ABI (include/adder_b200.h) over the sm_100a kernels in px_kernel.cuh.
 *
 * Host-side counterpart of Video<W>'s transcode state (adder-codec-rs/src/transcoder/source/video.rs:
 * 186-243 VideoState, :322-345 Video, :350-438 new, :493-537 time_parameters, :1241-1287 quality
 * setters, :651-778 integrate_matrix).  No CPU fallback: every entry point that touches pixel state
 * needs a CUDA device and fails with ADDER_ERR_NO_DEVICE / ADDER_ERR_CUDA otherwise.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "../../include/adder_b200.h"
#include "px_kernel.cuh"
#include "feature_kernel.cuh"
#include "framer_kernel.cuh"
#include "raw_kernel.cuh"
#include "exchange_kernel.cuh"
#include "synth.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(_e == cudaErrorMemoryAllocation ? ADDER_ERR_NOMEM : ADDER_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(_e), __FILE__, __LINE__);                                    \
  } while (0)

/* CRF table, adder-codec-core/src/codec/rate_controller.rs:5-18:
 * {c_thresh_baseline, c_thresh_max, c_increase_velocity, feature_c_radius_denom} */
const float kCrf[10][4] = {
    {0.0f, 0.0f, 10.0f, 1E-9f},        {0.0f, 1.0f, 9.0f, 1.0f / 12.0f},  {1.0f, 3.0f, 8.0f, 1.0f / 14.0f},
    {2.0f, 7.0f, 7.0f, 1.0f / 15.0f},  {5.0f, 9.0f, 6.0f, 1.0f / 18.0f},  {6.0f, 10.0f, 5.0f, 1.0f / 20.0f},
    {7.0f, 13.0f, 4.0f, 1.0f / 25.0f}, {8.0f, 16.0f, 3.0f, 1.0f / 30.0f}, {10.0f, 20.0f, 2.0f, 1.0f / 30.0f},
    {15.0f, 25.0f, 1.0f, 1.0f / 30.0f},
};

void crf_lookup(uint8_t crf, uint16_t w, uint16_t h, adder_crf_parameters_t* out) { /* Crf::new :55-70 */
  memset(out, 0, sizeof(*out));
  if (crf > 9) crf = 9;
  const uint16_t min_res = w < h ? w : h;
  out->c_thresh_baseline = (uint8_t)kCrf[crf][0];
  out->c_thresh_max = (uint8_t)kCrf[crf][1];
  out->c_increase_velocity = (uint8_t)kCrf[crf][2];
  const float r = kCrf[crf][3] * (float)min_res;
  out->feature_c_radius = r >= 65535.0f ? 65535 : (uint16_t)r;
}

/* fast-math 0.1 `log2_raw` as published (the crate is not vendored in the reference, Cargo.toml:45):
 * exponent + quadratic in the significand.  Only feeds FramedViewMode::D (scale_intensity.rs:89-91). */
float log2_raw(float x) {
  uint32_t bits;
  memcpy(&bits, &x, 4);
  const int exponent = (int)((bits >> 23) & 0xFF) - 127;
  const uint32_t mb = (bits & 0x007FFFFFu) | 0x3F800000u;
  float m;
  memcpy(&m, &mb, 4);
  volatile float t = (-1.0f / 3.0f) * m;
  t = t + 2.0f;
  t = t * m;
  t = t + (-2.0f / 3.0f);
  return (float)exponent + t;
}

uint32_t f32_as_u32(float x) { /* Rust `as u32` */
  if (!(x > 0.0f)) return 0u;
  if (x >= 4294967296.0f) return 0xFFFFFFFFu;
  return (uint32_t)x;
}

constexpr int kRing = 3; /* host-form pipeline depth */
constexpr uint32_t kMaxFramesPerLaunch = adder::kMaxLaunchFrames; /* frames one integrate launch may span */
constexpr uint32_t kRtSlots = 8;              /* launches whose running_t tables may be in flight before one is reused */

}  // namespace

struct adder_b200_video {
  uint16_t w = 0, h = 0;
  uint8_t c = 0;
  int device = 0;
  int tree_mode = ADDER_MODE_FRAME_PERFECT, multi_mode = ADDER_MULTI_COLLAPSE, time_mode = ADDER_TIME_ABSOLUTE_T,
      view_mode = ADDER_VIEW_INTENSITY;
  uint32_t chunk_rows = 1, n_chunks = 0, in_interval_count = 1, tps = 7650, ref_time = 255, delta_t_max = 7650;
  adder_crf_parameters_t crf{};
  uint32_t want_depth = 0; /* caller's max_depth (0 = derive) */
  uint32_t depth = 0;      /* allocated */
  float running_t = 0.0f;
  uint32_t P = 0, n_tiles = 0;   /* n_tiles: tiles of the smallest shape (upper bound, sizes the status array) */
  uint32_t R = 1, n_tiles_r = 0; /* sub-tiles per tile and the number of tiles that goes with it */
  uint32_t grid = 0;             /* persistent CTAs per launch (the kernel's full occupancy) */
  uint32_t reserve_ctas = 0;     /* CTA slots the integrate launch leaves free for the event exchange's push kernel (set when a comm is bound) */
  uint32_t status_ring = 1;      /* frames of status words (power of two) */
  uint64_t status_words = 0;     /* status_ring * tiles */
  float* d_running_t[kRtSlots] = {}; /* running_t per frame of a launch (kMaxFramesPerLaunch + 1 floats each) */
  float* d_rt_cur = nullptr;
  uint32_t rt_slot = 0;
  bool display_force = true;     /* next frame recomputes every display byte (PxParams::display == 2) */
  uint8_t* d_exact_lut = nullptr; /* [257] display bytes of exactly integral intensities, for lut_ref */
  uint32_t lut_ref = 0;
  uint64_t Ppad = 0;

  /* How the node stacks are stored (state_layout.h): 0 = eager (every level holds its own integration / delta_t), 1 = offset
   * form (px_offset.cuh: levels below the root hold offsets against the root and are touched only when they fire).  A video
   * whose state is pristine takes whichever form the launch's parameters allow; an offset-form state is converted to the
   * eager form when a launch is not eligible any more, and an eager state that has integrated frames stays eager. */
  int form = 0;
  bool pristine = true;    /* nothing integrated since create / reset_state */
  uint32_t records = 0;    /* 32-byte records allocated per pixel */

  uint2* d_hdr = nullptr;
  uint4* d_nodes = nullptr;
  uint2* d_park_arena = nullptr; /* [grid][3 park buffers][arena_slots][tile px]: events beyond the shared-memory slots */
  uint32_t arena_slots = 0;
  uint8_t* d_running = nullptr;
  unsigned long long* d_status = nullptr;
  uint32_t* d_ticket = nullptr;
  uint32_t* d_err = nullptr;
  unsigned long long* d_total = nullptr;
  uint32_t* h_err = nullptr; /* pinned */
  unsigned long long* h_total = nullptr;

  /* host-form resources (allocated on first use) */
  uint8_t* d_frame[kRing] = {nullptr, nullptr, nullptr};
  uint8_t src_c = 0;                                   /* channels of the frames handed in; 0 = the video's own */
  uint8_t* d_rgb[kRing] = {nullptr, nullptr, nullptr}; /* three-channel staging of the host forms (gray transcode of a colour source) */
  uint8_t* d_gray = nullptr;                           /* the same for the device-resident form: gray_frames frames */
  uint32_t gray_frames = 0;
  bool gray_diag_ready = false;
  const uint8_t* d_last_input = nullptr;               /* the (gray) frame the last integrate call worked on */
  /* feature detection (video.rs:202-210) */
  bool feature_detection = false, feature_rate_adjustment = false;
  uint8_t* d_feat_mask = nullptr;  /* (H,W) */
  uint32_t* d_new_xy = nullptr;    /* x | y<<16 of the last frame's new features */
  uint32_t* d_n_new = nullptr;
  uint8_t *d_new_mask = nullptr, *d_row_hit = nullptr; /* (H,W) each: this frame's new features, and their dilation along rows */
  uint32_t* d_feat_off = nullptr;  /* chunk offsets for frames whose caller did not ask for them */
  uint32_t new_cap = 0;
  adder_event_t* d_events[kRing] = {nullptr, nullptr, nullptr};
  uint8_t* d_raw[kRing] = {nullptr, nullptr, nullptr}; /* wire-format copies of d_events (raw host form only) */
  uint32_t* d_chunk_off[kRing] = {nullptr, nullptr, nullptr};
  uint32_t* h_chunk_off[kRing] = {nullptr, nullptr, nullptr}; /* pinned */
  uint64_t events_capacity = 0;                               /* records per slot */
  int last_slot = -1;                                         /* slot holding the last single-frame call's events */
  /* frames integrated by integrate_frames_host[_raw] whose events did not fit the caller's buffer (frames_host_impl) */
  uint32_t pend_n = 0, pend_slot0 = 0;
  int pend_form = 0;

  cudaStream_t stream = nullptr, stream_in = nullptr, stream_out = nullptr;
  cudaEvent_t ev_in[kRing] = {}, ev_k[kRing] = {}, ev_out[kRing] = {}, ev_t0 = nullptr, ev_t1 = nullptr;
  uint64_t launches = 0;
  uint32_t epoch = 0, ticket_base = 0;
  uint32_t row0 = 0;
  bool counting = false;
  unsigned long long* d_counters = nullptr;
};

namespace {

uint32_t derive_depth(const adder_b200_video* v) {
  if (v->want_depth) return std::min<uint32_t>(std::max<uint32_t>(v->want_depth, 2u), ADDER_MAX_DEPTH);
  /* live nodes grow like log2 of the frames a pixel can integrate before Δt_max pops the root
   * (7 at Δt_max/ref = 24, 11 at 4096); +6 leaves two levels of margin over the
   * floor(log2)+4 bound observed there.  The kernel reports ADDER_DEVERR_DEPTH if it is ever exceeded. */
  /* PixelMultiMode::Normal: popped_dtm stays set after the first Δt_max pop (only pop_best_events clears it,
   * event_pixel_tree.rs:283), the root is never popped again and a static pixel's stack grows with the logarithm of
   * the frames integrated since its last change — bounded by nothing but the reference's own 30-iteration guard
   * (:387-389).  Allocate that guard's depth. */
  if (v->multi_mode == ADDER_MULTI_NORMAL) return ADDER_MAX_DEPTH;
  uint32_t ratio = v->delta_t_max / std::max<uint32_t>(v->ref_time, 1u);
  uint32_t lg = 0;
  while ((ratio >> (lg + 1)) != 0) lg++;
  return std::min<uint32_t>(lg + 6u, ADDER_MAX_DEPTH);
}

int set_device(const adder_b200_video* v) {
  CU(cudaSetDevice(v->device));
  return ADDER_OK;
}

/* 32-byte records per pixel: eager form two levels to a record, offset form one level to a record (record 0 = root + meta) */
uint32_t records_for(int form, uint32_t depth) { return form == 1 ? depth : NODE_LEVELS_ALLOC(depth) / 2u; }

int ensure_depth(adder_b200_video* v, uint32_t need) {
  need = std::max(need, v->depth);
  const uint32_t need_records = std::max(records_for(v->form, need), v->records);
  if (need == v->depth && need_records == v->records) return ADDER_OK;
  if (need_records > v->records) {
    uint4* nn = nullptr;
    /* the records of one level (pair) of all pixels are contiguous (state_layout.h), so a larger allocation keeps the old
     * one as its prefix */
    CU(cudaMalloc(&nn, (size_t)need_records * 2u * v->Ppad * sizeof(uint4)));
    if (v->d_nodes) {
      CU(cudaMemcpyAsync(nn, v->d_nodes, (size_t)v->records * 2u * v->Ppad * sizeof(uint4), cudaMemcpyDeviceToDevice, v->stream));
      CU(cudaStreamSynchronize(v->stream));
      CU(cudaFree(v->d_nodes));
    }
    v->d_nodes = nn;
    v->records = need_records;
  }
  if (need == v->depth) return ADDER_OK;
  const bool first = v->depth == 0;
  v->depth = need;
  { /* a pixel emits at most depth + 2 events in one frame (1 + max(L, 2) + 1): the arena takes what the slots do not */
    const uint32_t slots = need + 2u - adder::park_slots(v->R);
    if (v->d_park_arena) {
      CU(cudaStreamSynchronize(v->stream));
      CU(cudaFree(v->d_park_arena));
      v->d_park_arena = nullptr;
    }
    const size_t cta_slots = std::max<uint32_t>(v->grid, 1u);
    CU(cudaMalloc(&v->d_park_arena, cta_slots * adder::kParkBufs * slots * adder::tile_px(v->R) * sizeof(uint2)));
    v->arena_slots = slots;
  }
  if (first) {
    adder::init_state_kernel<<<(v->P + 255) / 256, 256, 0, v->stream>>>(v->d_hdr, v->d_nodes, v->d_running, v->P);
    v->launches++;
    CU(cudaGetLastError());
  }
  /* the host-form event buffers are sized for the worst case of this depth: drop them */
  for (int s = 0; s < kRing; s++) {
    if (v->d_events[s]) {
      CU(cudaStreamSynchronize(v->stream_out));
      CU(cudaFree(v->d_events[s]));
      v->d_events[s] = nullptr;
    }
    if (v->d_raw[s]) {
      CU(cudaFree(v->d_raw[s]));
      v->d_raw[s] = nullptr;
    }
  }
  v->events_capacity = 0;
  return ADDER_OK;
}

int ensure_host_form(adder_b200_video* v) {
  const uint64_t cap = (uint64_t)v->P * (v->depth + 2u); /* 1 + max(L,2) + 1 events per pixel at most, SURVEY §8(a) */
  for (int s = 0; s < kRing; s++) {
    if (!v->d_frame[s]) CU(cudaMalloc(&v->d_frame[s], v->P));
    if (!v->d_events[s]) CU(cudaMalloc(&v->d_events[s], cap * sizeof(adder_event_t)));
  }
  v->events_capacity = cap;
  return ADDER_OK;
}

int realloc_chunks(adder_b200_video* v) {
  v->n_chunks = (v->h + v->chunk_rows - 1) / v->chunk_rows;
  for (int s = 0; s < kRing; s++) {
    if (v->d_chunk_off[s]) CU(cudaFree(v->d_chunk_off[s]));
    if (v->h_chunk_off[s]) CU(cudaFreeHost(v->h_chunk_off[s]));
    CU(cudaMalloc(&v->d_chunk_off[s], ((size_t)v->n_chunks + 1) * sizeof(uint32_t)));
    CU(cudaHostAlloc(&v->h_chunk_off[s], ((size_t)v->n_chunks + 1) * sizeof(uint32_t), cudaHostAllocDefault));
  }
  return ADDER_OK;
}

/* The persistent grid fills every SM to the kernel's occupancy, and at 64 registers x 256 threads x 4 CTAs that is the
 * whole register file: nothing else can start on an SM until a CTA of the launch exits, i.e. until the launch is over.
 * The push kernel of the event exchange is meant to run BESIDE the next integrate launch, so a video with a comm bound
 * to it launches a few CTAs fewer and the push kernel has exactly that many (measured without the reserve at 8 GPUs: the
 * push ran after the next launch instead of beside it and the gather leg lost 35 %, profiles/r02g n8). */
constexpr uint32_t kPushCtas = 16;
uint32_t launch_grid(const adder_b200_video* v) { return v->grid > 8u * v->reserve_ctas ? v->grid - v->reserve_ctas : v->grid; }

/* The default output configuration (TimeMode::AbsoluteT, FramedViewMode::Intensity) runs the instantiation that has both
 * as compile-time constants (kPlain: the large tile, uncounted, per-lane walk); everything else the general one. */
bool use_plain(const adder_b200_video* v) {
  if (const char* e = getenv("ADDER_B200_PLAIN")) return atoi(e) != 0 && v->time_mode == ADDER_TIME_ABSOLUTE_T && v->view_mode == ADDER_VIEW_INTENSITY;
  return v->time_mode == ADDER_TIME_ABSOLUTE_T && v->view_mode == ADDER_VIEW_INTENSITY;
}
template <int R, bool kDeep>
void launch_rd(adder_b200_video* v, const adder::FrameArgs& a, cudaStream_t stream) {
  const size_t smem = adder::frame_kernel_smem(R);
  const uint32_t grid = launch_grid(v);
  if (!kDeep && v->form == 1) { /* offset-form state: the kOff instantiations */
    if (R == 8 && !v->counting && use_plain(v)) {
      if (a.n_frames > 1u)
        adder::integrate_frame_kernel<8, false, true, false, true, true><<<grid, ADDER_TILE_PX, smem, stream>>>(a);
      else
        adder::integrate_frame_kernel<8, false, false, false, true, true><<<grid, ADDER_TILE_PX, smem, stream>>>(a);
    } else if (v->counting) {
      adder::integrate_frame_kernel<R, true, true, false, false, true><<<grid, ADDER_TILE_PX, smem, stream>>>(a);
    } else if (a.n_frames > 1u) {
      adder::integrate_frame_kernel<R, false, true, false, false, true><<<grid, ADDER_TILE_PX, smem, stream>>>(a);
    } else {
      adder::integrate_frame_kernel<R, false, false, false, false, true><<<grid, ADDER_TILE_PX, smem, stream>>>(a);
    }
    return;
  }
  if (R == 8 && !kDeep && !v->counting && use_plain(v)) {
    if (a.n_frames > 1u)
      adder::integrate_frame_kernel<8, false, true, false, true><<<grid, ADDER_TILE_PX, smem, stream>>>(a);
    else
      adder::integrate_frame_kernel<8, false, false, false, true><<<grid, ADDER_TILE_PX, smem, stream>>>(a);
    return;
  }
  if (v->counting)
    adder::integrate_frame_kernel<R, true, true, kDeep><<<grid, ADDER_TILE_PX, smem, stream>>>(a);
  else if (a.n_frames > 1u)
    adder::integrate_frame_kernel<R, false, true, kDeep><<<grid, ADDER_TILE_PX, smem, stream>>>(a);
  else /* one frame: the variant compiled without the cross-frame dependency, fences and L2-only state loads */
    adder::integrate_frame_kernel<R, false, false, kDeep><<<grid, ADDER_TILE_PX, smem, stream>>>(a);
}
template <int R>
void launch_r(adder_b200_video* v, const adder::FrameArgs& a, cudaStream_t stream) {
  launch_rd<R, false>(v, a, stream);
}
/* The variant whose warps walk the levels below the first two together (kDeep).  Parity-green on every case
 * (ADDER_B200_R=8 ADDER_B200_DEEP=1 pytest), but measured 33 % SLOWER than the per-lane walk on aged 8K stacks
 * (1923 vs 1444 us per frame, profiles/r02i_ab_deep.txt): its passes are a chain of dependent L2 round trips with nothing
 * requested ahead, where the per-lane loop always has the next level in flight.  Off unless ADDER_B200_DEEP=1. */
bool use_deep(const adder_b200_video* v) {
  if (v->R != 8 || v->form == 1) return false;
  if (const char* e = getenv("ADDER_B200_DEEP")) return atoi(e) != 0;
  return false;
}
template <int R>
int occupancy_r(int* ctas_per_sm) {
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, adder::integrate_frame_kernel<R, false, true>, ADDER_TILE_PX,
                                                   adder::frame_kernel_smem(R)));
  return ADDER_OK;
}
/* persistent grid: every SM filled to the kernel's occupancy, never more CTAs than tiles */
int choose_grid(adder_b200_video* v) {
  int sms = 0, per_sm = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, v->device));
  int rc;
  switch (v->R) {
    case 1: rc = occupancy_r<1>(&per_sm); break;
    case 2: rc = occupancy_r<2>(&per_sm); break;
    case 4: rc = occupancy_r<4>(&per_sm); break;
    default: rc = occupancy_r<8>(&per_sm); break;
  }
  if (rc) return rc;
  if (v->R == 8) { /* the long-integration variant must fit as many CTAs: the grid is shared */
    int deep_per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&deep_per_sm, adder::integrate_frame_kernel<8, false, true, true>, ADDER_TILE_PX,
                                                     adder::frame_kernel_smem(8)));
    per_sm = std::min(per_sm, deep_per_sm);
  }
  { /* and so must the offset-form instantiations */
    int off_per_sm = 0;
    switch (v->R) {
      case 1: CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&off_per_sm, adder::integrate_frame_kernel<1, false, true, false, false, true>, ADDER_TILE_PX, adder::frame_kernel_smem(1))); break;
      case 2: CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&off_per_sm, adder::integrate_frame_kernel<2, false, true, false, false, true>, ADDER_TILE_PX, adder::frame_kernel_smem(2))); break;
      case 4: CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&off_per_sm, adder::integrate_frame_kernel<4, false, true, false, false, true>, ADDER_TILE_PX, adder::frame_kernel_smem(4))); break;
      default: CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&off_per_sm, adder::integrate_frame_kernel<8, false, true, false, false, true>, ADDER_TILE_PX, adder::frame_kernel_smem(8))); break;
    }
    per_sm = std::min(per_sm, off_per_sm);
  }
  if (per_sm < 1) return fail(ADDER_ERR_INTERNAL, "integrate_frame_kernel does not fit an SM");
  if (const char* e = getenv("ADDER_B200_CTAS_PER_SM")) {
    const int n = atoi(e);
    if (n >= 1 && n < per_sm) per_sm = n;
  }
  v->grid = std::min<uint32_t>(v->n_tiles_r, (uint32_t)sms * (uint32_t)per_sm);
  return ADDER_OK;
}
void launch_variant(adder_b200_video* v, const adder::FrameArgs& a, cudaStream_t stream) {
  switch (v->R) {
    case 1: launch_r<1>(v, a, stream); break;
    case 2: launch_r<2>(v, a, stream); break;
    case 4: launch_r<4>(v, a, stream); break;
    default:
      if (use_deep(v)) launch_rd<8, true>(v, a, stream); else launch_rd<8, false>(v, a, stream);
      break;
  }
}
template <int R>
int set_smem_attr() {
  CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<R, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(R)));
  CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<R, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(R)));
  CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<R, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(R)));
  CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<R, false, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(R)));
  CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<R, false, true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(R)));
  CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<R, true, true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(R)));
  if (R == 8) {
    CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<8, false, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(8)));
    CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<8, false, true, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(8)));
    CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<8, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(8)));
    CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<8, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(8)));
    CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<8, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(8)));
    CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<8, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(8)));
    CU(cudaFuncSetAttribute(adder::integrate_frame_kernel<8, false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adder::frame_kernel_smem(8)));
  }
  return ADDER_OK;
}
/* rounds per tile: 8 (62 rows = 1984 pixels: with one shared-memory park slot per pixel the three park
 * buffers of four CTAs fit an SM) unless the plane is so small that the persistent grid would not get
 * a few tiles per CTA. */
uint32_t choose_r(uint32_t P) {
  if (const char* e = getenv("ADDER_B200_R")) {
    const int r = atoi(e);
    if (r == 1 || r == 2 || r == 4 || r == 8) return (uint32_t)r;
  }
  const uint32_t want_ctas = 148u * 8u;
  for (uint32_t r = 8; r > 1; r >>= 1)
    if (P / adder::tile_px(r) >= want_ctas) return r;
  return 1;
}

/* The form of the node stacks for a launch with the handle's current parameters (see adder_b200_video::form).
 * ADDER_B200_OFFSET=0 keeps everything in the eager form (A/B runs). */
bool offset_wanted(const adder_b200_video* v, float time_spanned) {
  if (const char* e = getenv("ADDER_B200_OFFSET"))
    if (atoi(e) == 0) return false;
  return adder::offset_form_eligible(v->multi_mode == ADDER_MULTI_COLLAPSE, time_spanned, v->delta_t_max);
}
int choose_form(adder_b200_video* v, cudaStream_t stream, float time_spanned) {
  const int want = offset_wanted(v, time_spanned) ? 1 : 0;
  if (want == v->form) return ADDER_OK;
  if (v->pristine) { /* a fresh state reads the same in both forms (one root, no levels below it) */
    v->form = want;
    return ADDER_OK;
  }
  if (v->form == 1) { /* not eligible any more: back to the reference's own representation, in place */
    if (stream != v->stream) CU(cudaStreamSynchronize(v->stream));
    adder::offset_to_eager_kernel<<<(v->P + 255) / 256, 256, 0, stream>>>(v->d_hdr, v->d_nodes, v->P, 2ull * v->Ppad);
    v->launches++;
    CU(cudaGetLastError());
    v->form = 0;
  }
  /* an eager state that has integrated frames stays eager: nothing guarantees the bounds the offset form needs */
  return ADDER_OK;
}

/* Queue one frame on `stream`.  d_frame: P dense bytes on the device. */
/* Queue n_frames consecutive frames (d_frame + f * frame_stride) as ONE launch: frame f's events go to
 * d_events + f * cap, its chunk offsets to d_chunk_off + f * (n_chunks + 1).  Tickets run frame-major over the
 * launch; a tile waits for the same tile of the previous frame through the status words, so the tail of one
 * frame overlaps the head of the next (profiles/r01k_timeline.txt: a third of a single-frame launch is drain). */
int launch_frames(adder_b200_video* v, cudaStream_t stream, const uint8_t* d_frame, size_t frame_stride, uint32_t n_frames,
                  float time_spanned, adder_event_t* d_events, uint64_t cap, uint32_t* d_chunk_off) {
  if (v->tree_mode != ADDER_MODE_FRAME_PERFECT)
    return fail(ADDER_ERR_UNSUPPORTED, "Mode::Continuous is not on the framed path (framed.rs:67 always builds FramePerfect)");
  if (n_frames == 0) return ADDER_OK;
  if (n_frames > 1 && (v->feature_detection || n_frames > kMaxFramesPerLaunch || (uint64_t)n_frames * v->n_tiles_r >= (1ull << 31)))
    return fail(ADDER_ERR_INTERNAL, "launch_frames: batch not split by the caller");
  /* everything that can fail comes before the first side effect (host counters, queued kernels) */
  if (v->feature_detection && v->row0 != 0)
    return fail(ADDER_ERR_UNSUPPORTED, "feature detection on a row band: the FAST neighbourhood would cross bands");
  if (int rc = choose_form(v, stream, time_spanned)) return rc;
  if (int rc = ensure_depth(v, derive_depth(v))) return rc;
  if (v->feature_detection) {
    if (!v->d_feat_mask) {
      const size_t hw = (size_t)v->w * v->h;
      v->new_cap = (uint32_t)std::min<size_t>(hw, 1u << 24);
      CU(cudaMalloc(&v->d_feat_mask, hw));
      CU(cudaMemsetAsync(v->d_feat_mask, 0, hw, stream));
      CU(cudaMalloc(&v->d_new_xy, (size_t)v->new_cap * sizeof(uint32_t)));
      CU(cudaMalloc(&v->d_n_new, sizeof(uint32_t)));
      CU(cudaMalloc(&v->d_new_mask, hw));
      CU(cudaMalloc(&v->d_row_hit, hw));
    }
    if (!d_chunk_off) { /* the feature pass walks the stream chunk by chunk */
      if (!v->d_feat_off) CU(cudaMalloc(&v->d_feat_off, ((size_t)v->h + 1) * sizeof(uint32_t))); /* n_chunks <= h */
      d_chunk_off = v->d_feat_off;
    }
  }

  if (v->in_interval_count == 0) { /* video.rs:656-658 */
    adder::set_initial_d_kernel<<<(v->P + 255) / 256, 256, 0, stream>>>(v->d_hdr, v->d_nodes, d_frame, v->P);
    v->launches++;
  }
  v->in_interval_count += n_frames; /* :662, once per frame */

  if ((uint64_t)v->epoch + n_frames >= 0x3FFFFFF0ull) { /* epoch wrap: forget all status words */
    CU(cudaMemsetAsync(v->d_status, 0, v->status_words * sizeof(unsigned long long), stream));
    v->epoch = 0;
  }
  const uint32_t epoch0 = v->epoch + 1u; /* frame f of this launch carries epoch0 + f */
  v->epoch += n_frames;

  /* running_t before / after every frame of the launch: the same f32 add per frame as event_pixel_tree.rs:337 */
  const float running_t_before = v->running_t;
  if (n_frames == 1) { /* the two values travel in the kernel arguments: no copy in front of the launch */
    v->running_t = v->running_t + time_spanned;
    v->d_rt_cur = nullptr;
  } else {
    std::vector<float> rt(n_frames + 1u);
    rt[0] = v->running_t;
    for (uint32_t f = 0; f < n_frames; f++) rt[f + 1] = rt[f] + time_spanned;
    float* d_rt = v->d_running_t[v->rt_slot];
    v->rt_slot = (v->rt_slot + 1u) % kRtSlots;
    CU(cudaMemcpyAsync(d_rt, rt.data(), rt.size() * sizeof(float), cudaMemcpyHostToDevice, stream)); /* pageable source: staged before the call returns */
    v->running_t = rt[n_frames];
    v->d_rt_cur = d_rt;
  }

  adder::FrameArgs a{};
  a.frame = d_frame;
  a.frame_stride = frame_stride;
  a.n_frames = n_frames;
  a.status_ring = v->status_ring;
  a.running_t = v->d_rt_cur;
  a.tiles_magic = adder::ref_magic_of(v->n_tiles_r);
  a.hdr = v->d_hdr;
  a.nodes = v->d_nodes;
  a.park_arena = v->d_park_arena;
  a.arena_slots = v->arena_slots;
  a.pair_stride = 2ull * v->Ppad;
  a.running = v->d_running;
  a.ev_words = reinterpret_cast<uint32_t*>(d_events);
  a.ev_cap = cap;
  a.chunk_off = d_chunk_off;
  a.tile_status = v->d_status;
  a.ticket = v->d_ticket;
  a.err = v->d_err;
  a.total_events = v->d_total;
  a.ticket_base = v->ticket_base;
  a.epoch = epoch0;
  a.P = v->P;
  a.n_tiles = v->n_tiles_r;
  a.C = v->c;
  a.WC = (uint32_t)v->w * v->c;
  a.chunk_px = v->chunk_rows * a.WC;
  a.n_chunks = v->n_chunks;
  a.row0 = v->row0;
  a.wc_magic = adder::ref_magic_of(a.WC);
  a.chunk_magic = adder::ref_magic_of(a.chunk_px);
  a.c_magic = v->c > 1 ? (uint32_t)((0x100000000ull + v->c - 1u) / v->c) : 0u;
  a.counters = v->d_counters;
  adder::PxParams& p = a.px;
  p.depth = v->depth;
  p.time = time_spanned;
  p.running_t_prev = running_t_before; /* read by single-frame launches; the others take both from a.running_t[] */
  p.running_t = v->running_t;
  p.dtm_f = (float)v->delta_t_max;
  p.ref = v->ref_time;
  p.dtm = v->delta_t_max;
  p.c_max = v->crf.c_thresh_max;
  p.vel_m1 = (uint8_t)(v->crf.c_increase_velocity - 1);
  p.cnt_inc = (uint8_t)(f32_as_u32(time_spanned) / v->ref_time);
  p.collapse = v->multi_mode == ADDER_MULTI_COLLAPSE;
  p.abs_time = v->time_mode == ADDER_TIME_ABSOLUTE_T;
  p.view_mode = (uint32_t)v->view_mode;
  p.display = v->display_force ? 2u : 1u;
  v->display_force = false;
  p.ref_magic = adder::ref_magic_of(v->ref_time);
  p.tpf = (double)v->ref_time; /* video.rs:672 */
  p.tpf_f = (float)v->ref_time;
  if (v->lut_ref != v->ref_time) {
    uint8_t lut[257];
    adder::build_exact_lut(v->ref_time, lut);
    CU(cudaMemcpyAsync(v->d_exact_lut, lut, sizeof(lut), cudaMemcpyHostToDevice, stream)); /* pageable source: staged before the call returns */
    v->lut_ref = v->ref_time;
  }
  p.exact_lut = v->d_exact_lut;
  p.practical_d_max = log2_raw(255.0f * (float)(v->delta_t_max / v->ref_time)); /* :668-670 */

  v->d_last_input = d_frame + (size_t)(n_frames - 1u) * frame_stride;
  v->pristine = false;
  launch_variant(v, a, stream);
  v->ticket_base += n_frames * v->n_tiles_r + launch_grid(v); /* every CTA draws one ticket past the end */
  v->launches++;
  CU(cudaGetLastError());
  if (!v->feature_detection && v->d_n_new) CU(cudaMemsetAsync(v->d_n_new, 0, sizeof(uint32_t), stream)); /* this frame found none */
  if (v->feature_detection) { /* video.rs:744 handle_features */
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, v->device));
    CU(cudaMemsetAsync(v->d_n_new, 0, sizeof(uint32_t), stream));
    const bool adjust = v->feature_rate_adjustment && v->crf.feature_c_radius > 0; /* :1089-1104 */
    if (adjust) CU(cudaMemsetAsync(v->d_new_mask, 0, (size_t)v->w * v->h, stream));
    adder::feature_kernel<<<sms * 8, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(d_events), d_chunk_off, v->n_chunks, v->chunk_rows,
                                                      v->row0, v->d_running, v->w, v->h, v->c, v->d_feat_mask, v->d_new_xy, v->d_n_new,
                                                      v->new_cap, adjust ? v->d_new_mask : nullptr);
    v->launches++;
    if (adjust) {
      /* per feature when they are few, by dilation of the new-feature bitmap when they are dense: decided on the
       * device from the count (ADDER_B200_FEATURE_RESET=dilate|list forces one form, for the tests) */
      int force = 0;
      if (const char* e = getenv("ADDER_B200_FEATURE_RESET")) force = !strcmp(e, "dilate") ? 1 : !strcmp(e, "list") ? -1 : 0;
      const int radius = (int)v->crf.feature_c_radius;
      const uint32_t value = std::min<uint32_t>(v->crf.c_thresh_baseline, 2u);
      adder::feature_reset_kernel<<<sms * 4, 256, 0, stream>>>(v->d_hdr, v->d_new_xy, v->d_n_new, v->new_cap, v->w, v->h, v->c, radius, value, force);
      adder::feature_dilate_rows_kernel<<<sms * 8, 256, 0, stream>>>(v->d_new_mask, v->d_row_hit, v->d_n_new, v->new_cap, v->w, v->h, radius, force);
      adder::feature_dilate_cols_kernel<<<sms * 8, 256, 0, stream>>>(v->d_hdr, v->d_row_hit, v->d_n_new, v->new_cap, v->w, v->h, v->c, radius, value, force);
      v->launches += 3;
    }
    CU(cudaGetLastError());
  }
  return ADDER_OK;
}

int launch_frame(adder_b200_video* v, cudaStream_t stream, const uint8_t* d_frame, float time_spanned, adder_event_t* d_events,
                 uint64_t cap, uint32_t* d_chunk_off) {
  return launch_frames(v, stream, d_frame, v->P, 1u, time_spanned, d_events, cap, d_chunk_off);
}

bool rgb_in(const adder_b200_video* v) { return v->src_c == 3 && v->c == 1; }
size_t in_frame_bytes(const adder_b200_video* v) { return rgb_in(v) ? (size_t)v->P * 3u : (size_t)v->P; }

/* handle_color on the device (utils/cv.rs:215-232): d_rgb (P*3 bytes) -> d_gray (P bytes) */
int launch_gray(adder_b200_video* v, cudaStream_t stream, const uint8_t* d_rgb, uint8_t* d_gray, uint32_t n_frames = 1) {
  if (!v->gray_diag_ready) { /* the table of gray_math.h, once per handle */
    uint8_t diag[256];
    adder::build_gray_diag(diag);
    CU(cudaMemcpyToSymbolAsync(adder::c_gray_diag, diag, sizeof(diag), 0, cudaMemcpyHostToDevice, stream)); /* pageable source: staged before the call returns */
    v->gray_diag_ready = true;
  }
  const uint64_t n_px = (uint64_t)v->P * n_frames; /* frames back to back on both sides */
  const uint64_t groups = (n_px + adder::kGrayPxPerThread - 1u) / adder::kGrayPxPerThread;
  if ((groups + 255u) / 256u > 0x7FFFFFFFull) return fail(ADDER_ERR_INTERNAL, "launch_gray: batch not split by the caller");
  adder::rgb_to_gray_kernel<<<(uint32_t)((groups + 255u) / 256u), 256, 0, stream>>>(d_rgb, d_gray, n_px);
  v->launches++;
  CU(cudaGetLastError());
  return ADDER_OK;
}

int map_deverr(uint32_t e) {
  if (e & ADDER_DEVERR_INTERNAL) return fail(ADDER_ERR_INTERNAL, "a pixel's root was popped with no child (state invariant)");
  if (e & ADDER_DEVERR_DEPTH)
    return fail(ADDER_ERR_ARENA_DEPTH, "a pixel's node stack outgrew the allocated depth (create with a larger max_depth)");
  if (e & ADDER_DEVERR_CAPACITY) return fail(ADDER_ERR_CAPACITY, "device event buffer too small: records beyond capacity were dropped");
  return ADDER_OK;
}

/* Read and clear the device error word (after the stream has been synchronised up to the copy). */
int collect_errors(adder_b200_video* v, cudaStream_t stream) {
  CU(cudaMemcpyAsync(v->h_err, v->d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  CU(cudaStreamSynchronize(stream));
  const uint32_t e = *v->h_err;
  if (e) CU(cudaMemsetAsync(v->d_err, 0, sizeof(uint32_t), stream));
  return map_deverr(e);
}

template <typename F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    return fail(ADDER_ERR_NOMEM, "host allocation failed");
  } catch (...) {
    return fail(ADDER_ERR_INTERNAL, "unexpected exception");
  }
}

}  // namespace

extern "C" {

int adder_b200_abi_version(void) { return ADDER_B200_ABI_VERSION; }
const char* adder_b200_last_error(void) { return g_err; }

int adder_b200_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    fail(ADDER_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return -ADDER_ERR_NO_DEVICE;
  }
  return n;
}

int adder_b200_crf_parameters(uint8_t crf, uint16_t plane_w, uint16_t plane_h, adder_crf_parameters_t* out) {
  if (!out || crf > 9) return fail(ADDER_ERR_BAD_PARAMS, "crf must be 0..=9");
  crf_lookup(crf, plane_w, plane_h, out);
  return ADDER_OK;
}

int adder_b200_video_create(uint16_t width, uint16_t height, uint8_t channels, int pixel_tree_mode, int device,
                            uint32_t max_depth, adder_b200_video** out) {
  return guarded([&]() -> int {
    if (!out) return fail(ADDER_ERR_BAD_PARAMS, "out is NULL");
    *out = nullptr;
    if (width == 0 || height == 0 || channels == 0) /* PlaneSize::new, lib.rs:105-117 */
      return fail(ADDER_ERR_BAD_PARAMS, "plane dimensions must be non-zero");
    if (pixel_tree_mode != ADDER_MODE_FRAME_PERFECT && pixel_tree_mode != ADDER_MODE_CONTINUOUS)
      return fail(ADDER_ERR_BAD_PARAMS, "unknown pixel_tree_mode");
    const uint64_t P64 = (uint64_t)width * height * channels;
    if (P64 * (ADDER_MAX_DEPTH + 2ull) >= (1ull << 32))
      return fail(ADDER_ERR_BAD_PARAMS, "plane too large: per-frame event offsets are 32-bit");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
      return fail(ADDER_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(ADDER_ERR_BAD_PARAMS, "device %d out of range (0..%d)", device, n - 1);

    adder_b200_video* v = new adder_b200_video();
    v->w = width;
    v->h = height;
    v->c = channels;
    v->device = device;
    v->tree_mode = pixel_tree_mode;
    v->want_depth = max_depth;
    crf_lookup(3, width, height, &v->crf); /* EncoderOptions::default -> Crf::new(None) -> quality 3 */
    v->P = (uint32_t)P64;
    v->Ppad = (P64 + 255ull) & ~255ull;
    v->n_tiles = (uint32_t)((P64 + adder::tile_px(1) - 1) / adder::tile_px(1)); /* smallest tile: sizes the status array */
    v->R = choose_r(v->P);
    v->n_tiles_r = (uint32_t)((P64 + adder::tile_px(v->R) - 1) / adder::tile_px(v->R));

    auto build = [&]() -> int {
      CU(cudaSetDevice(device));
      if (const char* e = getenv("ADDER_B200_L2_FETCH")) { /* experiment: L2 fetch granularity in bytes (32 / 64 / 128) */
        const int g = atoi(e);
        if (g == 32 || g == 64 || g == 128) CU(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g));
      }
      CU(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
      CU(cudaStreamCreateWithFlags(&v->stream_in, cudaStreamNonBlocking));
      CU(cudaStreamCreateWithFlags(&v->stream_out, cudaStreamNonBlocking));
      for (int s = 0; s < kRing; s++) {
        CU(cudaEventCreateWithFlags(&v->ev_in[s], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&v->ev_k[s], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&v->ev_out[s], cudaEventDisableTiming));
      }
      CU(cudaEventCreate(&v->ev_t0));
      CU(cudaEventCreate(&v->ev_t1));
      CU(cudaMalloc(&v->d_hdr, (size_t)v->Ppad * sizeof(uint2)));
      CU(cudaMalloc(&v->d_running, (size_t)v->Ppad));
      /* allocated after choose_grid(): see below */
      CU(cudaMalloc(&v->d_ticket, sizeof(uint32_t)));
      CU(cudaMalloc(&v->d_err, sizeof(uint32_t)));
      CU(cudaMalloc(&v->d_total, sizeof(unsigned long long)));
      CU(cudaHostAlloc(&v->h_err, sizeof(uint32_t), cudaHostAllocDefault));
      CU(cudaHostAlloc(&v->h_total, sizeof(unsigned long long), cudaHostAllocDefault));
      CU(cudaMemsetAsync(v->d_ticket, 0, sizeof(uint32_t), v->stream));
      CU(cudaMemsetAsync(v->d_err, 0, sizeof(uint32_t), v->stream));
      CU(cudaMemsetAsync(v->d_total, 0, sizeof(unsigned long long), v->stream));
      if (int rc = set_smem_attr<1>()) return rc;
      if (int rc = set_smem_attr<2>()) return rc;
      if (int rc = set_smem_attr<4>()) return rc;
      if (int rc = set_smem_attr<8>()) return rc;
      if (int rc = choose_grid(v)) return rc;
      { /* Status words for as many frames as can be in flight at once: a CTA holds at most three tickets, so the
         * tickets being worked on span at most 3 * grid / tiles frames; + 3 for the frame being read by look-backs,
         * the one before it (dependencies) and rounding. */
        uint32_t need = (3u * v->grid + v->n_tiles_r - 1u) / v->n_tiles_r + 3u, ring = 1u;
        while (ring < need) ring <<= 1;
        v->status_ring = ring;
        v->status_words = (uint64_t)ring * v->n_tiles_r;
        CU(cudaMalloc(&v->d_status, v->status_words * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(v->d_status, 0, v->status_words * sizeof(unsigned long long), v->stream));
        for (uint32_t k = 0; k < kRtSlots; k++) CU(cudaMalloc(&v->d_running_t[k], (kMaxFramesPerLaunch + 1u) * sizeof(float)));
      }
      CU(cudaMalloc(&v->d_exact_lut, 257));
      CU(cudaMalloc(&v->d_counters, 6 * sizeof(unsigned long long)));
      CU(cudaMemsetAsync(v->d_counters, 0, 6 * sizeof(unsigned long long), v->stream));
      if (int rc = realloc_chunks(v)) return rc;
      v->form = offset_wanted(v, (float)v->ref_time) ? 1 : 0; /* the defaults are eligible; a pristine state changes form for free */
      if (int rc = ensure_depth(v, derive_depth(v))) return rc;
      CU(cudaStreamSynchronize(v->stream));
      return ADDER_OK;
    };
    if (int rc = build()) {
      adder_b200_video_destroy(v);
      return rc;
    }
    *out = v;
    return ADDER_OK;
  });
}

void adder_b200_video_destroy(adder_b200_video* v) {
  if (!v) return;
  cudaSetDevice(v->device);
  if (v->stream) cudaStreamSynchronize(v->stream);
  if (v->stream_in) cudaStreamSynchronize(v->stream_in);
  if (v->stream_out) cudaStreamSynchronize(v->stream_out);
  cudaFree(v->d_hdr);
  cudaFree(v->d_nodes);
  cudaFree(v->d_running);
  cudaFree(v->d_park_arena);
  for (uint32_t k = 0; k < kRtSlots; k++) cudaFree(v->d_running_t[k]);
  cudaFree(v->d_status);
  cudaFree(v->d_ticket);
  cudaFree(v->d_err);
  cudaFree(v->d_total);
  cudaFree(v->d_counters);
  cudaFree(v->d_exact_lut);
  cudaFree(v->d_gray);
  cudaFree(v->d_feat_mask);
  cudaFree(v->d_new_xy);
  cudaFree(v->d_n_new);
  cudaFree(v->d_new_mask);
  cudaFree(v->d_row_hit);
  cudaFree(v->d_feat_off);
  if (v->h_err) cudaFreeHost(v->h_err);
  if (v->h_total) cudaFreeHost(v->h_total);
  for (int s = 0; s < kRing; s++) {
    cudaFree(v->d_frame[s]);
    cudaFree(v->d_events[s]);
    cudaFree(v->d_rgb[s]);
    cudaFree(v->d_raw[s]);
    cudaFree(v->d_chunk_off[s]);
    if (v->h_chunk_off[s]) cudaFreeHost(v->h_chunk_off[s]);
    if (v->ev_in[s]) cudaEventDestroy(v->ev_in[s]);
    if (v->ev_k[s]) cudaEventDestroy(v->ev_k[s]);
    if (v->ev_out[s]) cudaEventDestroy(v->ev_out[s]);
  }
  if (v->ev_t0) cudaEventDestroy(v->ev_t0);
  if (v->ev_t1) cudaEventDestroy(v->ev_t1);
  if (v->stream) cudaStreamDestroy(v->stream);
  if (v->stream_in) cudaStreamDestroy(v->stream_in);
  if (v->stream_out) cudaStreamDestroy(v->stream_out);
  delete v;
}

/* ---- setters ---------------------------------------------------------------------------------- */

int adder_b200_video_chunk_rows(adder_b200_video* v, uint32_t chunk_rows) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (chunk_rows == 0) return fail(ADDER_ERR_BAD_PARAMS, "chunk_rows must be > 0"); /* ndarray chunk size must be non-zero */
  if (int rc = set_device(v)) return rc;
  CU(cudaStreamSynchronize(v->stream));
  v->chunk_rows = chunk_rows;
  return realloc_chunks(v);
}

int adder_b200_video_time_parameters(adder_b200_video* v, uint32_t tps, uint32_t ref_time, uint32_t delta_t_max,
                                     int time_mode, int* applied) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (time_mode > ADDER_TIME_MIXED) return fail(ADDER_ERR_BAD_PARAMS, "unknown time_mode");
  if (time_mode >= 0) v->time_mode = time_mode; /* :500-502 runs before the range checks */
  int ok = 1;
  if (delta_t_max < ref_time) ok = 0; /* :518-523 eprintln + keep */
  if (ref_time == 0) return fail(ADDER_ERR_BAD_PARAMS, "ref_time must be > 0");
  if (ok) {
    v->delta_t_max = delta_t_max;
    v->ref_time = ref_time;
    v->tps = tps;
    v->display_force = true;
  }
  if (applied) *applied = ok;
  return ADDER_OK;
}

int adder_b200_video_write_out(adder_b200_video* v, int time_mode, int pixel_multi_mode) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (time_mode > ADDER_TIME_MIXED || pixel_multi_mode > ADDER_MULTI_COLLAPSE) return fail(ADDER_ERR_BAD_PARAMS, "unknown mode");
  v->multi_mode = pixel_multi_mode < 0 ? ADDER_MULTI_COLLAPSE : pixel_multi_mode; /* unwrap_or_default */
  if (time_mode >= 0) v->time_mode = time_mode;
  return ADDER_OK;
}

static int reset_c(adder_b200_video* v, uint8_t c, int reset_counter) {
  if (int rc = set_device(v)) return rc;
  adder::reset_c_kernel<<<(v->P + 255) / 256, 256, 0, v->stream>>>(v->d_hdr, v->P, c, reset_counter);
  v->launches++;
  CU(cudaGetLastError());
  return ADDER_OK;
}

int adder_b200_video_update_crf(adder_b200_video* v, uint8_t crf) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (crf > 9) return fail(ADDER_ERR_BAD_PARAMS, "crf must be 0..=9");
  crf_lookup(crf, v->w, v->h, &v->crf);
  return reset_c(v, v->crf.c_thresh_baseline, 1);
}

int adder_b200_video_update_quality_manual(adder_b200_video* v, uint8_t c_thresh_baseline, uint8_t c_thresh_max,
                                           uint32_t delta_t_max_multiplier, uint8_t c_increase_velocity,
                                           float feature_c_radius) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  v->crf.c_thresh_baseline = c_thresh_baseline;
  v->crf.c_thresh_max = c_thresh_max;
  v->crf.c_increase_velocity = c_increase_velocity;
  v->crf.feature_c_radius = feature_c_radius >= 65535.0f ? 65535 : (feature_c_radius > 0.0f ? (uint16_t)feature_c_radius : 0);
  v->delta_t_max = delta_t_max_multiplier * v->ref_time;
  v->display_force = true;
  return reset_c(v, c_thresh_baseline, 1);
}

int adder_b200_video_set_crf_parameters(adder_b200_video* v, const adder_crf_parameters_t* params) {
  if (!v || !params) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  v->crf = *params;
  return ADDER_OK;
}

int adder_b200_video_update_delta_t_max(adder_b200_video* v, uint32_t delta_t_max) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  v->delta_t_max = std::max(v->ref_time, delta_t_max); /* video.rs:819-822 */
  v->display_force = true;
  return ADDER_OK;
}

int adder_b200_video_c_thresh_pos(adder_b200_video* v, uint8_t c) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  v->crf.c_thresh_baseline = c;
  return reset_c(v, c, 0);
}

int adder_b200_video_set_c_thresh_rect(adder_b200_video* v, uint16_t x0, uint16_t y0, uint16_t x1, uint16_t y1, uint8_t value) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (x0 >= v->w || y0 >= v->h || x1 < x0 || y1 < y0) return ADDER_OK; /* empty after clipping */
  if (int rc = set_device(v)) return rc;
  const uint32_t rw = std::min<uint32_t>(x1, v->w - 1u) - x0 + 1u, rh = std::min<uint32_t>(y1, v->h - 1u) - y0 + 1u;
  const uint32_t n = rw * v->c * rh;
  adder::rect_c_kernel<<<(n + 255) / 256, 256, 0, v->stream>>>(v->d_hdr, v->w, v->c, x0, y0, rw, rh, value);
  v->launches++;
  CU(cudaGetLastError());
  return ADDER_OK;
}

int adder_b200_video_set_view_mode(adder_b200_video* v, int view_mode) {
  if (!v || view_mode < 0 || view_mode > ADDER_VIEW_SAE) return fail(ADDER_ERR_BAD_PARAMS, "unknown view mode");
  v->view_mode = view_mode;
  v->display_force = true;
  return ADDER_OK;
}

int adder_b200_video_set_in_interval_count(adder_b200_video* v, uint32_t n) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  v->in_interval_count = n;
  return ADDER_OK;
}

int adder_b200_video_update_detect_features(adder_b200_video* v, int detect_features, int feature_rate_adjustment) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  v->feature_detection = detect_features != 0;
  v->feature_rate_adjustment = feature_rate_adjustment != 0;
  return ADDER_OK;
}

int adder_b200_video_new_features(adder_b200_video* v, uint16_t* xy_out, size_t cap, uint32_t* n) {
  if (!v || !n || (cap && !xy_out)) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  *n = 0;
  if (!v->d_n_new) return ADDER_OK; /* detection never ran */
  if (int rc = set_device(v)) return rc;
  uint32_t cnt = 0;
  CU(cudaMemcpyAsync(&cnt, v->d_n_new, sizeof(cnt), cudaMemcpyDeviceToHost, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  if (cnt > v->new_cap) return fail(ADDER_ERR_CAPACITY, "more new features (%u) than the device list holds (%u)", cnt, v->new_cap);
  *n = cnt;
  const uint32_t take = (uint32_t)std::min<size_t>(cnt, cap);
  if (take) {
    std::vector<uint32_t> tmp(take);
    CU(cudaMemcpyAsync(tmp.data(), v->d_new_xy, (size_t)take * sizeof(uint32_t), cudaMemcpyDeviceToHost, v->stream));
    CU(cudaStreamSynchronize(v->stream));
    for (uint32_t k = 0; k < take; k++) {
      xy_out[2 * k] = (uint16_t)(tmp[k] & 0xFFFFu);
      xy_out[2 * k + 1] = (uint16_t)(tmp[k] >> 16);
    }
  }
  return ADDER_OK;
}

int adder_b200_video_feature_mask(adder_b200_video* v, uint8_t* out) {
  if (!v || !out) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  if (int rc = set_device(v)) return rc;
  const size_t hw = (size_t)v->w * v->h;
  if (!v->d_feat_mask) {
    memset(out, 0, hw);
    return ADDER_OK;
  }
  CU(cudaMemcpyAsync(out, v->d_feat_mask, hw, cudaMemcpyDeviceToHost, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  return ADDER_OK;
}

int adder_b200_video_set_source_channels(adder_b200_video* v, uint8_t source_channels) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (source_channels != 0 && source_channels != v->c && !(source_channels == 3 && v->c == 1))
    return fail(ADDER_ERR_BAD_PARAMS, "source_channels must be the video's own channel count, or 3 for a one-channel video");
  v->src_c = source_channels == v->c ? 0 : source_channels;
  return ADDER_OK;
}

int adder_b200_video_input_frame(adder_b200_video* v, uint8_t* out) {
  if (!v || !out) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  if (!v->d_last_input) return fail(ADDER_ERR_BAD_PARAMS, "no frame has been integrated yet");
  if (int rc = set_device(v)) return rc;
  CU(cudaMemcpyAsync(out, v->d_last_input, v->P, cudaMemcpyDeviceToHost, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  return ADDER_OK;
}

int adder_b200_video_set_row_offset(adder_b200_video* v, uint16_t row0) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if ((uint32_t)row0 + v->h > 65536u) return fail(ADDER_ERR_BAD_PARAMS, "row offset + height exceeds the u16 coordinate range");
  v->row0 = row0;
  return ADDER_OK;
}

int adder_b200_video_set_counting(adder_b200_video* v, int on) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (int rc = set_device(v)) return rc;
  v->counting = on != 0;
  CU(cudaMemsetAsync(v->d_counters, 0, 6 * sizeof(unsigned long long), v->stream));
  return ADDER_OK;
}

int adder_b200_video_read_counters(adder_b200_video* v, uint64_t out[6]) {
  if (!v || !out) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  if (int rc = set_device(v)) return rc;
  unsigned long long h[6];
  CU(cudaMemcpyAsync(h, v->d_counters, sizeof(h), cudaMemcpyDeviceToHost, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  for (int i = 0; i < 6; i++) out[i] = h[i];
  return ADDER_OK;
}

int adder_b200_video_get_info(const adder_b200_video* v, adder_b200_video_info_t* out) {
  if (!v || !out) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  memset(out, 0, sizeof(*out));
  out->width = v->w;
  out->height = v->h;
  out->channels = v->c;
  out->pixel_tree_mode = (uint8_t)v->tree_mode;
  out->pixel_multi_mode = (uint8_t)v->multi_mode;
  out->time_mode = (uint8_t)v->time_mode;
  out->view_mode = (uint8_t)v->view_mode;
  out->state_form = (uint8_t)v->form;
  out->chunk_rows = v->chunk_rows;
  out->n_chunks = v->n_chunks;
  out->in_interval_count = v->in_interval_count;
  out->tps = v->tps;
  out->ref_time = v->ref_time;
  out->delta_t_max = v->delta_t_max;
  out->crf = v->crf;
  out->max_depth = v->depth;
  out->device = (uint32_t)v->device;
  out->state_bytes = (uint64_t)v->Ppad * (sizeof(uint2) + 1 + (uint64_t)v->records * 2u * sizeof(uint4)); /* the park arena is scratch, not state */
  out->events_capacity = v->events_capacity;
  return ADDER_OK;
}

int adder_b200_video_reset_state(adder_b200_video* v) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (int rc = set_device(v)) return rc;
  adder::init_state_kernel<<<(v->P + 255) / 256, 256, 0, v->stream>>>(v->d_hdr, v->d_nodes, v->d_running, v->P);
  v->launches++;
  CU(cudaGetLastError());
  v->running_t = 0.0f;
  v->pristine = true;
  v->in_interval_count = 1;
  v->display_force = true;
  v->pend_n = 0; /* events of frames integrated before the reset are dropped with the state */
  if (v->d_feat_mask) CU(cudaMemsetAsync(v->d_feat_mask, 0, (size_t)v->w * v->h, v->stream));
  return ADDER_OK;
}

/* ---- the hot path ----------------------------------------------------------------------------- */

static void offsets_to_counts(const uint32_t* off, uint32_t n_chunks, uint32_t* counts) {
  for (uint32_t i = 0; i < n_chunks; i++) counts[i] = off[i + 1] - off[i];
}

int adder_b200_video_integrate_matrix(adder_b200_video* v, const uint8_t* frame, size_t row_pitch, float time_spanned,
                                      adder_event_t* events_out, size_t events_cap, uint32_t* chunk_counts,
                                      uint64_t* n_events) {
  return guarded([&]() -> int {
    if (!v || !frame) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
    if (v->pend_n) return fail(ADDER_ERR_BAD_PARAMS, "%u frames integrated by integrate_frames_host are waiting to be delivered: resume that call first", v->pend_n);
    if (int rc = set_device(v)) return rc;
    if (int rc = ensure_depth(v, derive_depth(v))) return rc;
    if (int rc = ensure_host_form(v)) return rc;
    const size_t row = (size_t)v->w * (rgb_in(v) ? 3u : v->c);
    if (row_pitch == 0) row_pitch = row;
    if (row_pitch < row) return fail(ADDER_ERR_BAD_PARAMS, "row_pitch smaller than a row");
    const int s = 0;
    if (rgb_in(v)) { /* framed.rs:129: handle_color, here on the device */
      if (!v->d_rgb[s]) CU(cudaMalloc(&v->d_rgb[s], (size_t)v->P * 3u));
      CU(cudaMemcpy2DAsync(v->d_rgb[s], row, frame, row_pitch, row, v->h, cudaMemcpyHostToDevice, v->stream));
      if (int rc = launch_gray(v, v->stream, v->d_rgb[s], v->d_frame[s])) return rc;
    } else {
      CU(cudaMemcpy2DAsync(v->d_frame[s], row, frame, row_pitch, row, v->h, cudaMemcpyHostToDevice, v->stream));
    }
    if (int rc = launch_frame(v, v->stream, v->d_frame[s], time_spanned, v->d_events[s], v->events_capacity, v->d_chunk_off[s]))
      return rc;
    v->last_slot = s;
    CU(cudaMemcpyAsync(v->h_chunk_off[s], v->d_chunk_off[s], ((size_t)v->n_chunks + 1) * sizeof(uint32_t),
                       cudaMemcpyDeviceToHost, v->stream));
    if (int rc = collect_errors(v, v->stream)) return rc;
    return adder_b200_video_fetch_events(v, events_out, events_cap, chunk_counts, n_events);
  });
}

int adder_b200_video_fetch_events(adder_b200_video* v, adder_event_t* events_out, size_t events_cap,
                                  uint32_t* chunk_counts, uint64_t* n_events) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (v->last_slot < 0) return fail(ADDER_ERR_BAD_PARAMS, "no frame has been integrated through the host form yet");
  if (int rc = set_device(v)) return rc;
  const int s = v->last_slot;
  const uint64_t total = v->h_chunk_off[s][v->n_chunks];
  if (n_events) *n_events = total;
  if (chunk_counts) offsets_to_counts(v->h_chunk_off[s], v->n_chunks, chunk_counts);
  if (total > events_cap)
    return fail(ADDER_ERR_CAPACITY, "events_out holds %zu records, the frame produced %llu", events_cap, (unsigned long long)total);
  if (total) {
    if (!events_out) return fail(ADDER_ERR_BAD_PARAMS, "events_out is NULL");
    CU(cudaMemcpyAsync(events_out, v->d_events[s], total * sizeof(adder_event_t), cudaMemcpyDeviceToHost, v->stream));
    CU(cudaStreamSynchronize(v->stream));
  }
  return ADDER_OK;
}

namespace {

uint32_t raw_event_size(const adder_b200_video* v) { return v->c == 1 ? 9u : 11u; } /* codec/header.rs:77-81 */

int launch_raw_encode(adder_b200_video* v, cudaStream_t stream, const adder_event_t* d_events, const uint32_t* d_n, uint64_t n_max,
                      uint8_t* d_out) {
  if (n_max == 0) return ADDER_OK;
  int sms = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, v->device));
  const uint64_t want = (n_max + adder::kRawChunk - 1) / adder::kRawChunk;
  const uint32_t blocks = (uint32_t)std::min<uint64_t>(want, (uint64_t)sms * 8u);
  adder::raw_encode_kernel<<<blocks, adder::kRawThreads, 0, stream>>>(reinterpret_cast<const uint32_t*>(d_events), d_n, n_max,
                                                                      raw_event_size(v), d_out);
  v->launches++;
  CU(cudaGetLastError());
  return ADDER_OK;
}

/* bytes of one frame in the compact host form (raw_kernel.cuh): dense P + 5 E when 4 E > P, else sparse 9 E */
uint64_t compact_frame_bytes(uint64_t P, uint64_t n_events) { return 4ull * n_events > P ? P + 5ull * n_events : 9ull * n_events; }

int launch_compact_encode(adder_b200_video* v, cudaStream_t stream, const adder_event_t* d_events, const uint32_t* d_n, uint64_t n_max, uint8_t* d_out) {
  int sms = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, v->device));
  CU(cudaMemsetAsync(d_out, 0, v->P, stream)); /* the dense form's count bytes: pixels without events stay 0 */
  const uint64_t want = (n_max + adder::kRawThreads - 1) / adder::kRawThreads;
  const uint32_t blocks = (uint32_t)std::max<uint64_t>(std::min<uint64_t>(want, (uint64_t)sms * 8u), 1u);
  adder::compact_encode_kernel<<<blocks, adder::kRawThreads, 0, stream>>>(reinterpret_cast<const uint32_t*>(d_events), d_n, n_max, v->P,
                                                                          (uint32_t)v->w * v->c, v->c, v->row0, d_out);
  v->launches++;
  CU(cudaGetLastError());
  return ADDER_OK;
}

/* n_frames consecutive calls of integrate_matrix with H2D / kernels / D2H overlapped on three streams.
 * raw = false: 12-byte records to out; raw = true: the wire bytes of the same events.
 *
 * Capacity.  A frame's event count is known only after its kernel has run, and up to kRing frames are in flight, so when
 * `out` fills up some frames have been integrated whose events no longer fit.  Their events are NOT lost and the frames are
 * NOT integrated twice: they stay in the handle's ring ("pending"), the call returns ADDER_ERR_CAPACITY with *frames_done =
 * the frames delivered, and the next call — which the caller makes with frames + frames_done * stride, as documented —
 * delivers the pending frames first and skips integrating them. */
int frames_host_impl(adder_b200_video* v, const uint8_t* frames, size_t frame_stride, uint32_t n_frames, float time_spanned,
                     void* out, size_t out_cap /* records, or bytes when form != 0 */, int form /* 0 records, 1 raw wire bytes, 2 compact */, uint64_t* frame_counts,
                     uint32_t* chunk_counts, uint64_t* n_out, uint32_t* frames_done) {
  if (!v || (!frames && n_frames)) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  if (n_out) *n_out = 0;
  if (frames_done) *frames_done = 0;
  const bool raw = form != 0; /* the frame leaves through d_raw (wire bytes or the compact form) */
  static const char* const kFormName[3] = {"record", "raw", "compact"};
  if (v->pend_n && v->pend_form != form)
    return fail(ADDER_ERR_BAD_PARAMS, "%u integrated frames are waiting to be delivered in the %s form: resume with the same call",
                v->pend_n, kFormName[v->pend_form]);
  if (v->pend_n > n_frames)
    return fail(ADDER_ERR_BAD_PARAMS, "%u integrated frames are waiting to be delivered: resume with at least that many frames", v->pend_n);
  if (int rc = set_device(v)) return rc;
  if (int rc = ensure_depth(v, derive_depth(v))) return rc;
  if (int rc = ensure_host_form(v)) return rc;
  const size_t unit = form == 1 ? raw_event_size(v) : sizeof(adder_event_t);
  auto frame_units = [&](uint64_t total) -> uint64_t { /* what one frame takes of `out`, in its units */
    return form == 0 ? total : form == 1 ? total * unit : compact_frame_bytes(v->P, total);
  };
  if (raw)
    for (int s = 0; s < kRing; s++)
      if (!v->d_raw[s]) CU(cudaMalloc(&v->d_raw[s], v->events_capacity * 11u));
  const size_t in_bytes = in_frame_bytes(v);
  if (rgb_in(v))
    for (int s = 0; s < kRing; s++)
      if (!v->d_rgb[s]) CU(cudaMalloc(&v->d_rgb[s], in_bytes));
  if (frame_stride == 0) frame_stride = in_bytes;
  if (frame_stride < in_bytes) return fail(ADDER_ERR_BAD_PARAMS, "frame_stride smaller than a frame");
  CU(cudaStreamSynchronize(v->stream)); /* setters queued on the main stream come first */

  uint64_t written = 0; /* in units */
  /* Frame f of this call uses ring slot (slot0 + f) % kRing; the pending frames of the previous call are frames
   * 0 .. pend_n-1 of this one and already sit in their slots, kernels finished. */
  const uint32_t slot0 = v->pend_n ? v->pend_slot0 : 0u;
  uint32_t submitted = v->pend_n, delivered = 0;
  v->pend_n = 0;
  int rc_final = ADDER_OK;
  /* Frame f's kernel waits for its H2D copy (ev_in) and for the D2H copy that last used the slot (ev_out); its D2H
   * copy is issued once the host knows the count. */
  auto deliver = [&](uint32_t f) -> int {
    const int s = (int)((slot0 + f) % kRing);
    CU(cudaEventSynchronize(v->ev_k[s]));
    const uint32_t* off = v->h_chunk_off[s];
    const uint64_t total = off[v->n_chunks];
    const uint64_t need = frame_units(total);
    if (written + need > out_cap)
      return fail(ADDER_ERR_CAPACITY, "output holds %zu %s; frame %u needs %llu more than fit (its events are kept: call again from frames_done)",
                  out_cap, raw ? "bytes" : "records", f, (unsigned long long)(written + need - out_cap));
    if (total) {
      const void* src = raw ? (const void*)v->d_raw[s] : (const void*)v->d_events[s];
      CU(cudaMemcpyAsync((uint8_t*)out + written * (raw ? 1 : unit), src, raw ? need : total * unit, cudaMemcpyDeviceToHost, v->stream_out));
    }
    CU(cudaEventRecord(v->ev_out[s], v->stream_out));
    if (frame_counts) frame_counts[f] = total;
    if (chunk_counts) offsets_to_counts(off, v->n_chunks, chunk_counts + (size_t)f * v->n_chunks);
    written += need;
    return ADDER_OK;
  };
  auto submit = [&](uint32_t f) -> int {
    const int s = (int)((slot0 + f) % kRing);
    CU(cudaStreamWaitEvent(v->stream_in, v->ev_k[s], 0)); /* previous kernel on this slot has read its frame */
    CU(cudaMemcpyAsync(rgb_in(v) ? v->d_rgb[s] : v->d_frame[s], frames + (size_t)f * frame_stride, in_bytes, cudaMemcpyHostToDevice,
                       v->stream_in));
    CU(cudaEventRecord(v->ev_in[s], v->stream_in));
    CU(cudaStreamWaitEvent(v->stream, v->ev_in[s], 0));
    CU(cudaStreamWaitEvent(v->stream, v->ev_out[s], 0));
    if (rgb_in(v))
      if (int rc = launch_gray(v, v->stream, v->d_rgb[s], v->d_frame[s])) return rc;
    if (int rc = launch_frame(v, v->stream, v->d_frame[s], time_spanned, v->d_events[s], v->events_capacity, v->d_chunk_off[s]))
      return rc;
    if (form == 1)
      if (int rc = launch_raw_encode(v, v->stream, v->d_events[s], v->d_chunk_off[s] + v->n_chunks, v->events_capacity, v->d_raw[s]))
        return rc;
    if (form == 2)
      if (int rc = launch_compact_encode(v, v->stream, v->d_events[s], v->d_chunk_off[s] + v->n_chunks, v->events_capacity, v->d_raw[s]))
        return rc;
    CU(cudaMemcpyAsync(v->h_chunk_off[s], v->d_chunk_off[s], ((size_t)v->n_chunks + 1) * sizeof(uint32_t),
                       cudaMemcpyDeviceToHost, v->stream));
    CU(cudaEventRecord(v->ev_k[s], v->stream));
    return ADDER_OK;
  };

  for (uint32_t f = submitted; f < n_frames && rc_final == ADDER_OK; f++) {
    while (rc_final == ADDER_OK && submitted - delivered >= (uint32_t)kRing) { /* the slot's previous tenant goes first */
      if (int rc = deliver(delivered)) rc_final = rc; else delivered++;
    }
    if (rc_final != ADDER_OK) break;
    if (int rc = submit(f)) rc_final = rc; else submitted++;
  }
  while (rc_final == ADDER_OK && delivered < submitted) {
    if (int rc = deliver(delivered)) rc_final = rc; else delivered++;
  }
  /* every path drains the three streams: no copy into the caller's buffers is in flight when the call returns */
  cudaError_t e1 = cudaStreamSynchronize(v->stream_in), e2 = cudaStreamSynchronize(v->stream), e3 = cudaStreamSynchronize(v->stream_out);
  v->last_slot = -1;
  if (n_out) *n_out = written;
  if (frames_done) *frames_done = delivered;
  if (rc_final == ADDER_ERR_CAPACITY && delivered < submitted) { /* integrated, not delivered: kept for the resuming call */
    v->pend_n = submitted - delivered;
    v->pend_slot0 = (slot0 + delivered) % kRing;
    v->pend_form = form;
  }
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
    return fail(ADDER_ERR_CUDA, "draining the streams: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
  if (int rc = collect_errors(v, v->stream)) return rc;
  return rc_final;
}

}  // namespace

int adder_b200_video_integrate_frames_host(adder_b200_video* v, const uint8_t* frames, size_t frame_stride,
                                           uint32_t n_frames, float time_spanned, adder_event_t* events_out,
                                           size_t events_cap, uint64_t* frame_counts, uint32_t* chunk_counts,
                                           uint64_t* n_events, uint32_t* frames_done) {
  return guarded([&]() -> int {
    return frames_host_impl(v, frames, frame_stride, n_frames, time_spanned, events_out, events_cap, 0, frame_counts,
                            chunk_counts, n_events, frames_done);
  });
}

int adder_b200_video_integrate_frames_host_raw(adder_b200_video* v, const uint8_t* frames, size_t frame_stride,
                                               uint32_t n_frames, float time_spanned, uint8_t* bytes_out, size_t bytes_cap,
                                               uint64_t* frame_counts, uint32_t* chunk_counts, uint64_t* n_bytes,
                                               uint32_t* frames_done) {
  return guarded([&]() -> int {
    return frames_host_impl(v, frames, frame_stride, n_frames, time_spanned, bytes_out, bytes_cap, 1, frame_counts,
                            chunk_counts, n_bytes, frames_done);
  });
}

int adder_b200_video_integrate_frames_host_compact(adder_b200_video* v, const uint8_t* frames, size_t frame_stride,
                                                   uint32_t n_frames, float time_spanned, uint8_t* bytes_out, size_t bytes_cap,
                                                   uint64_t* frame_counts, uint32_t* chunk_counts, uint64_t* n_bytes,
                                                   uint32_t* frames_done) {
  return guarded([&]() -> int {
    return frames_host_impl(v, frames, frame_stride, n_frames, time_spanned, bytes_out, bytes_cap, 2, frame_counts,
                            chunk_counts, n_bytes, frames_done);
  });
}

uint64_t adder_b200_compact_frame_bytes(uint64_t n_px, uint64_t n_events) { return compact_frame_bytes(n_px, n_events); }

/* Host side of the compact form: the frame's 12-byte records back, on n_threads host threads (the work is a memory-bound
 * scatter: 12 bytes written per event). */
int adder_b200_expand_compact(uint16_t width, uint16_t rows, uint8_t channels, uint16_t row0, const uint8_t* block, uint64_t n_events,
                              adder_event_t* events_out, uint32_t n_threads) {
  return guarded([&]() -> int {
    if (!width || !rows || !channels || (n_events && (!block || !events_out))) return fail(ADDER_ERR_BAD_PARAMS, "bad argument");
    const uint64_t P = (uint64_t)width * rows * channels, WC = (uint64_t)width * channels;
    const bool dense = 4ull * n_events > P;
    const uint32_t T = std::max<uint32_t>(1u, std::min<uint32_t>(n_threads ? n_threads : 1u, 256u));
    /* P < 2^40 in principle (u16 x u16 x u8), but an index of the sparse form is a u32 and the dense form walks pixels with
     * running coordinates, so no 64-bit division is on either path */
    const uint32_t WC32 = (uint32_t)WC, C32 = channels;
    auto put = [&](adder_event_t& e, uint32_t idx, const uint8_t* dt) {
      const uint32_t y = idx / WC32, rem = idx - y * WC32, x = rem / C32;
      e.x = (uint16_t)x;
      e.y = (uint16_t)(y + row0);
      e.c = channels == 1 ? (uint8_t)ADDER_C_NONE : (uint8_t)(rem - x * C32);
      e.d = dt[0];
      e.reserved = 0;
      memcpy(&e.t, dt + 1, 4);
    };
    std::vector<std::thread> pool;
    std::vector<int> bad(T, 0);
    if (!dense) { /* E x {index, d, t}: any split of the events works */
      auto work = [&](uint32_t t) {
        const uint64_t a = n_events * t / T, b = n_events * (t + 1) / T;
        for (uint64_t k = a; k < b; k++) {
          uint32_t idx;
          memcpy(&idx, block + 9ull * k, 4);
          if (idx >= P) { bad[t] = 1; return; }
          put(events_out[k], idx, block + 9ull * k + 4);
        }
      };
      for (uint32_t t = 1; t < T; t++) pool.emplace_back(work, t);
      work(0);
      for (auto& th : pool) th.join();
    } else { /* count bytes | E x {d, t}: a thread takes a run of pixels; its first event is the sum of the counts before it */
      std::vector<uint64_t> first(T + 1, 0);
      auto count = [&](uint32_t t) {
        const uint64_t a = P * t / T, b = P * (t + 1) / T;
        uint64_t sum = 0;
        for (uint64_t i = a; i < b; i++) sum += block[i];
        first[t + 1] = sum;
      };
      for (uint32_t t = 1; t < T; t++) pool.emplace_back(count, t);
      count(0);
      for (auto& th : pool) th.join();
      pool.clear();
      for (uint32_t t = 0; t < T; t++) first[t + 1] += first[t];
      if (first[T] != n_events) return fail(ADDER_ERR_BAD_PARAMS, "compact block: the count bytes add up to %llu events, not %llu",
                                            (unsigned long long)first[T], (unsigned long long)n_events);
      const uint8_t* body = block + P;
      auto work = [&](uint32_t t) { /* raster walk with running coordinates: one division per thread, none per event */
        const uint64_t a = P * t / T, b = P * (t + 1) / T;
        uint64_t k = first[t];
        uint32_t y = (uint32_t)(a / WC), x = (uint32_t)((a - (uint64_t)y * WC) / channels), c = (uint32_t)(a - (uint64_t)y * WC - (uint64_t)x * channels);
        static_assert(sizeof(adder_event_t) == 12, "a record is three little-endian words: x | y << 16, c | d << 8, t");
        for (uint64_t i = a; i < b; i++) {
          if (const uint32_t n_here = block[i]) {
            const uint32_t w0 = x | (((y + row0) & 0xFFFFu) << 16), cw = channels == 1 ? (uint32_t)ADDER_C_NONE : c;
            for (uint32_t j = n_here; j; j--, k++) {
              const uint8_t* dt = body + 5ull * k;
              uint32_t* rec = reinterpret_cast<uint32_t*>(events_out + k); /* 4-byte aligned: three word stores */
              uint32_t t;
              memcpy(&t, dt + 1, 4);
              rec[0] = w0;
              rec[1] = cw | ((uint32_t)dt[0] << 8);
              rec[2] = t;
            }
          }
          if (++c == C32) {
            c = 0;
            if (++x == width) x = 0, y++;
          }
        }
      };
      for (uint32_t t = 1; t < T; t++) pool.emplace_back(work, t);
      work(0);
      for (auto& th : pool) th.join(); /* before `first` and the closures go out of scope */
    }
    for (uint32_t t = 0; t < T; t++)
      if (bad[t]) return fail(ADDER_ERR_BAD_PARAMS, "compact block: a pixel index lies outside the plane");
    return ADDER_OK;
  });
}

int adder_b200_video_raw_event_size(const adder_b200_video* v) { return v ? (int)raw_event_size(v) : -ADDER_ERR_BAD_PARAMS; }

int adder_b200_video_raw_encode_device(adder_b200_video* v, const adder_event_t* d_events, const uint32_t* d_n_events,
                                       uint64_t n_events_max, uint8_t* d_out) {
  if (!v || !d_n_events || (n_events_max && (!d_events || !d_out))) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  if ((reinterpret_cast<uintptr_t>(d_out) & 3u) != 0) return fail(ADDER_ERR_BAD_PARAMS, "d_out must be 4-byte aligned");
  if (int rc = set_device(v)) return rc;
  return launch_raw_encode(v, v->stream, d_events, d_n_events, n_events_max, d_out);
}

static uint8_t* put_be16(uint8_t* p, uint16_t x) {
  p[0] = (uint8_t)(x >> 8);
  p[1] = (uint8_t)x;
  return p + 2;
}
static uint8_t* put_be32(uint8_t* p, uint32_t x) {
  p[0] = (uint8_t)(x >> 24);
  p[1] = (uint8_t)(x >> 16);
  p[2] = (uint8_t)(x >> 8);
  p[3] = (uint8_t)x;
  return p + 4;
}

int adder_b200_video_raw_header(const adder_b200_video* v, uint8_t version, uint32_t source_camera, uint32_t adu_interval,
                                uint8_t* out, size_t cap, size_t* n_bytes) {
  if (!v || !out || !n_bytes) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  if (version > 3) return fail(ADDER_ERR_BAD_PARAMS, "codec version must be 0..=3 (CodecError::BadFile, encoder.rs:228)");
  const size_t need = 25u + 4u * version;
  if (cap < need) return fail(ADDER_ERR_CAPACITY, "header needs %zu bytes", need);
  uint8_t* p = out;
  memcpy(p, "adder", 5); /* MAGIC_RAW, header.rs:5 */
  p += 5;
  *p++ = version;
  *p++ = 'b';
  p = put_be16(p, v->w);
  p = put_be16(p, v->h);
  p = put_be32(p, v->tps);
  p = put_be32(p, v->ref_time);
  p = put_be32(p, v->delta_t_max);
  *p++ = (uint8_t)raw_event_size(v);
  *p++ = v->c;
  if (version >= 1) p = put_be32(p, source_camera);
  if (version >= 2) p = put_be32(p, (uint32_t)v->time_mode);
  if (version >= 3) p = put_be32(p, adu_interval);
  *n_bytes = (size_t)(p - out);
  return ADDER_OK;
}

int adder_b200_raw_eof(uint8_t* out, size_t cap, size_t* n_bytes) {
  if (!out || !n_bytes) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  if (cap < ADDER_RAW_EOF_BYTES) return fail(ADDER_ERR_CAPACITY, "the EOF event needs 11 bytes");
  static const uint8_t eof[11] = {0xFF, 0xFF, 0xFF, 0xFF, 1, 0, 0, 0, 0, 0, 0}; /* x = y = 0xFFFF, c = Some(0), d = 0, t = 0 */
  memcpy(out, eof, 11);
  *n_bytes = 11;
  return ADDER_OK;
}

int adder_b200_video_running_intensities(adder_b200_video* v, uint8_t* out) {
  if (!v || !out) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  if (int rc = set_device(v)) return rc;
  CU(cudaMemcpyAsync(out, v->d_running, v->P, cudaMemcpyDeviceToHost, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  return ADDER_OK;
}

int adder_b200_video_integrate_frames_device(adder_b200_video* v, const uint8_t* d_frames, size_t frame_stride,
                                             uint32_t n_frames, float time_spanned, adder_event_t* d_events,
                                             size_t events_stride, uint32_t* d_chunk_offsets) {
  return guarded([&]() -> int {
    if (!v || (!d_frames && n_frames) || (!d_events && events_stride)) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
    if (v->pend_n) return fail(ADDER_ERR_BAD_PARAMS, "%u frames integrated by integrate_frames_host are waiting to be delivered: resume that call first", v->pend_n);
    if (int rc = set_device(v)) return rc;
    if (frame_stride == 0) frame_stride = in_frame_bytes(v);
    /* a colour source feeding a gray transcode is converted into a scratch run of gray frames first (up to 512 MB of
     * them), so that the run still goes through one integrate launch */
    const uint32_t gray_run = rgb_in(v) ? (uint32_t)std::min<uint64_t>(std::max<uint64_t>((512ull << 20) / v->P, 1u), kMaxFramesPerLaunch) : 0u;
    if (rgb_in(v) && v->gray_frames < std::min(gray_run, n_frames)) {
      CU(cudaStreamSynchronize(v->stream)); /* the old scratch may still be read */
      cudaFree(v->d_gray);
      v->d_gray = nullptr;
      v->gray_frames = std::min(gray_run, n_frames);
      CU(cudaMalloc(&v->d_gray, (size_t)v->P * v->gray_frames));
    }
    /* One launch per run of frames when nothing has to happen between frames: no feature pass, no set_initial_d for
     * the first frame. */
    uint32_t f = 0;
    while (f < n_frames) {
      uint32_t* off = d_chunk_offsets ? d_chunk_offsets + (size_t)f * (v->n_chunks + 1) : nullptr;
      const uint8_t* d_in = d_frames + (size_t)f * frame_stride;
      size_t in_stride = frame_stride;
      uint32_t batch = 1;
      if (!v->feature_detection && v->in_interval_count != 0) {
        const uint64_t by_tickets = ((1ull << 31) - 1u) / std::max<uint32_t>(v->n_tiles_r, 1u);
        batch = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(n_frames - f, kMaxFramesPerLaunch), std::max<uint64_t>(by_tickets, 1u));
      }
      if (rgb_in(v)) {
        batch = std::min(batch, v->gray_frames);
        if (frame_stride == (size_t)v->P * 3u) {
          if (int rc = launch_gray(v, v->stream, d_in, v->d_gray, batch)) return rc;
        } else {
          for (uint32_t k = 0; k < batch; k++)
            if (int rc = launch_gray(v, v->stream, d_in + (size_t)k * frame_stride, v->d_gray + (size_t)k * v->P)) return rc;
        }
        d_in = v->d_gray;
        in_stride = v->P;
      }
      if (int rc = launch_frames(v, v->stream, d_in, in_stride, batch, time_spanned, d_events + (size_t)f * events_stride,
                                 events_stride, off))
        return rc;
      f += batch;
    }
    return ADDER_OK;
  });
}

int adder_b200_video_sync(adder_b200_video* v) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (int rc = set_device(v)) return rc;
  return collect_errors(v, v->stream);
}

void* adder_b200_video_stream(adder_b200_video* v) { return v ? (void*)v->stream : nullptr; }
uint64_t adder_b200_video_launch_count(const adder_b200_video* v) { return v ? v->launches : 0; }

int adder_b200_video_events_emitted(adder_b200_video* v, uint64_t* out) {
  if (!v || !out) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  if (int rc = set_device(v)) return rc;
  CU(cudaMemcpyAsync(v->h_total, v->d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  *out = *v->h_total;
  return ADDER_OK;
}

int adder_b200_video_read_px(adder_b200_video* v, size_t index, adder_b200_px_state_t* out) {
  if (!v || !out || index >= v->P) return fail(ADDER_ERR_BAD_PARAMS, "bad argument");
  if (int rc = set_device(v)) return rc;
  memset(out, 0, sizeof(*out));
  CU(cudaStreamSynchronize(v->stream));
  uint2 h;
  CU(cudaMemcpy(&h, v->d_hdr + index, sizeof(h), cudaMemcpyDeviceToHost));
  memcpy(&out->last_fired_t, &h.x, 4);
  out->running_t = v->running_t;
  out->base_val = HDR_BASE(h.y);
  out->c_thresh = HDR_CTHRESH(h.y);
  out->c_increase_counter = HDR_COUNTER(h.y);
  out->length = HDR_LENGTH(h.y);
  out->dtm_reached = HDR_DTM_REACHED(h.y);
  out->popped_dtm = HDR_POPPED(h.y);
  out->time_mode = (uint8_t)v->time_mode;
  uint4 root = make_uint4(0u, 0u, 0u, 0u);
  for (uint32_t k = 0; k < out->length && k < v->depth; k++) {
    uint4 n;
    if (v->form == 1) { /* offset form -> the reference's node (px_offset.cuh) */
      if (k == 0u) {
        CU(cudaMemcpy(&root, v->d_nodes + 2ull * index, sizeof(root), cudaMemcpyDeviceToHost));
        n = root;
      } else if (k + 1u >= out->length) {
        n = make_uint4(0u, 0u, 0u, 0u); /* the tail is always a node that has not integrated yet, and is not stored */
      } else {
        if (!out->popped_dtm && k + 2u == out->length) { /* the top level lives in the second half of record 0 */
          uint4 t;
          CU(cudaMemcpy(&t, v->d_nodes + 2ull * index + 1ull, sizeof(t), cudaMemcpyDeviceToHost));
          adder::OffTop top;
          top.a = t.x, top.b = t.z, top.pmin = t.w; /* in memory: a, best_dt, b, pmin */
          memcpy(&top.best_dt, &t.y, 4);
          const adder::OffRec q = adder::top_unpack(top);
          n.x = q.oi, n.y = q.od, n.w = q.w;
          memcpy(&n.z, &q.best_dt, 4);
        } else {
          CU(cudaMemcpy(&n, v->d_nodes + 2ull * index + (size_t)k * 2ull * v->Ppad, sizeof(n), cudaMemcpyDeviceToHost));
        }
        if (!out->popped_dtm) {
          float rx, rdt;
          memcpy(&rx, &root.x, 4);
          memcpy(&rdt, &root.y, 4);
          const float fx = (float)((uint32_t)rx - n.x), fdt = (float)((uint32_t)rdt - n.y);
          memcpy(&n.x, &fx, 4);
          memcpy(&n.y, &fdt, 4);
        }
      }
    } else
    CU(cudaMemcpy(&n, v->d_nodes + NODE_SLOT(k, index, v->Ppad), sizeof(n), cudaMemcpyDeviceToHost));
    memcpy(&out->nodes[k].integration, &n.x, 4);
    memcpy(&out->nodes[k].delta_t, &n.y, 4);
    memcpy(&out->nodes[k].best_delta_t, &n.z, 4);
    out->nodes[k].d = NODE_D(n.w);
    out->nodes[k].best_d = NODE_BEST_D(n.w);
    out->nodes[k].has_best = NODE_HAS_BEST(n.w);
  }
  return ADDER_OK;
}

/* ---- plumbing --------------------------------------------------------------------------------- */

int adder_b200_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(ADDER_ERR_BAD_PARAMS, "out is NULL");
  CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return ADDER_OK;
}
int adder_b200_host_free(void* p) {
  if (p) CU(cudaFreeHost(p));
  return ADDER_OK;
}
int adder_b200_device_alloc(adder_b200_video* v, size_t bytes, void** out) {
  if (!v || !out) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  if (int rc = set_device(v)) return rc;
  CU(cudaMalloc(out, bytes ? bytes : 1));
  return ADDER_OK;
}
int adder_b200_device_free(adder_b200_video* v, void* p) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (int rc = set_device(v)) return rc;
  CU(cudaStreamSynchronize(v->stream));
  CU(cudaFree(p));
  return ADDER_OK;
}
int adder_b200_copy_to_device(adder_b200_video* v, void* dst, const void* src, size_t bytes) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (int rc = set_device(v)) return rc;
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  return ADDER_OK;
}
int adder_b200_copy_to_host(adder_b200_video* v, void* dst, const void* src, size_t bytes) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (int rc = set_device(v)) return rc;
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, v->stream));
  CU(cudaStreamSynchronize(v->stream));
  return ADDER_OK;
}
int adder_b200_video_timer_start(adder_b200_video* v) {
  if (!v) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (int rc = set_device(v)) return rc;
  CU(cudaEventRecord(v->ev_t0, v->stream));
  return ADDER_OK;
}
int adder_b200_video_timer_stop(adder_b200_video* v, float* ms) {
  if (!v || !ms) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  if (int rc = set_device(v)) return rc;
  CU(cudaEventRecord(v->ev_t1, v->stream));
  CU(cudaEventSynchronize(v->ev_t1));
  CU(cudaEventElapsedTime(ms, v->ev_t0, v->ev_t1));
  return ADDER_OK;
}

int adder_b200_synth_frames(adder_b200_video* v, uint8_t* d_frames, size_t frame_stride, uint32_t f0, uint32_t n_frames,
                            int kind, uint64_t seed) {
  if (!v || !d_frames) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
  if (kind < 0 || kind > 3) return fail(ADDER_ERR_BAD_PARAMS, "unknown synthetic kind");
  if (int rc = set_device(v)) return rc;
  if (frame_stride == 0) frame_stride = v->P;
  const uint32_t blocks = (v->P + 255) / 256;
  for (uint32_t f = 0; f < n_frames; f++) {
    adder::synth_frame_kernel<<<blocks, 256, 0, v->stream>>>(d_frames + (size_t)f * frame_stride, v->P, v->w, v->c, f0 + f, kind, seed,
                                                             v->row0 * (uint32_t)v->w * v->c);
    v->launches++;
  }
  CU(cudaGetLastError());
  return ADDER_OK;
}

}  /* extern "C" */

/* ================================ framer (include/adder_b200.h, framer section) ================================ */

struct adder_b200_framer {
  uint16_t w = 0, h = 0;
  uint8_t c = 0;
  int device = 0;
  uint32_t chunk_rows = 1, n_chunks = 0;
  int64_t frames_written = 0;
  uint32_t tpf = 0, tps = 0, ref_interval = 0, source_dtm = 0, source_camera = 0, ring_frames = 0;
  uint8_t codec_version = 3;
  int view_mode = 0, time_mode = 0;
  int64_t buffer_limit = -1;
  float practical_d_max = 0.0f;
  uint64_t frame_px = 0;
  cudaStream_t stream = nullptr;
  uint4* d_px_state = nullptr; /* running timestamp | last filled frame << 8 | last intensity, framer_kernel.cuh */
  uint8_t *d_ring_val = nullptr, *d_ring_some = nullptr;
  long long *d_offset_max = nullptr, *d_forced = nullptr;
  uint8_t *d_tracker = nullptr, *d_status = nullptr;
  uint32_t *d_result = nullptr, *d_err = nullptr, *d_off_stage = nullptr;
  adder_event_t* d_ev_stage = nullptr;
  uint64_t ev_stage_cap = 0;
  uint8_t* d_out_stage = nullptr;
  uint32_t out_stage_frames = 0;
  uint32_t* h_result = nullptr; /* pinned: [0..2] predicates, [3] error word */
  uint8_t* d_exact_lut = nullptr; /* [257] build_exact_lut(ref_interval) */
  uint32_t ingest_grid = 0;       /* CTAs of framer_ingest_kernel (resident CTAs per SM x SMs), set at the first call */
};

namespace {

int framer_set_device(const adder_b200_framer* f) {
  CU(cudaSetDevice(f->device));
  return ADDER_OK;
}

/* status of the front frame -> tracker update (mode): queued on the framer's stream, nothing waits */
int framer_refresh_queue(adder_b200_framer* f, int mode, const uint32_t* d_chunk_off) {
  const uint64_t chunk_px = (uint64_t)f->chunk_rows * f->w * f->c;
  /* one launch: per-chunk status, then the tracker update by the CTA that finishes last (d_result[4] counts them) */
  adder::framer_chunk_status_kernel<<<f->n_chunks, 256, 0, f->stream>>>(f->d_ring_some, f->ring_frames, f->frame_px, chunk_px, f->d_forced,
                                                                        f->frames_written, f->d_status, f->d_result + 4, f->d_tracker, d_chunk_off,
                                                                        mode, f->d_offset_max, f->buffer_limit, f->d_result);
  CU(cudaGetLastError());
  return ADDER_OK;
}
/* the predicates of the last refresh (and the error word) on the host: the one synchronisation */
int framer_collect(adder_b200_framer* f) {
  CU(cudaMemcpyAsync(f->h_result, f->d_result, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaMemcpyAsync(f->h_result + 3, f->d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  if (f->h_result[3]) {
    CU(cudaMemsetAsync(f->d_err, 0, sizeof(uint32_t), f->stream));
    return fail(ADDER_ERR_CAPACITY, "an event reaches more than %u output frames ahead of the last one written (create the framer with a larger ring_frames)",
                f->ring_frames);
  }
  return ADDER_OK;
}
int framer_refresh(adder_b200_framer* f, int mode, const uint32_t* d_chunk_off) {
  if (int rc = framer_refresh_queue(f, mode, d_chunk_off)) return rc;
  return framer_collect(f);
}

int framer_ingest(adder_b200_framer* f, const adder_event_t* d_events, const uint32_t* d_chunk_off, int* frame_ready, bool wait = true) {
  adder::FramerArgs a{};
  a.ev_words = reinterpret_cast<const uint32_t*>(d_events);
  a.chunk_off = d_chunk_off;
  a.n_chunks = f->n_chunks;
  a.chunk_rows = f->chunk_rows;
  a.W = f->w;
  a.H = f->h;
  a.C = f->c;
  a.frames_written = f->frames_written;
  a.tpf = f->tpf;
  a.ref_interval = f->ref_interval;
  a.source_dtm = f->source_dtm;
  a.codec_version = f->codec_version;
  a.framed_source = f->source_camera <= 5u ? 1u : 0u; /* FramedU8..FramedF64, lib.rs:35-47 */
  a.view_mode = (uint32_t)f->view_mode;
  a.absolute_t = f->time_mode == ADDER_TIME_ABSOLUTE_T ? 1u : 0u;
  a.practical_d_max = f->practical_d_max;
  a.exact_lut = f->d_exact_lut;
  a.tpf_magic = adder::ref_magic_of(f->tpf);
  a.ring_magic = adder::ref_magic_of(f->ring_frames);
  a.ref_magic = adder::ref_magic_of(f->ref_interval);
  a.chunk_rows_magic = adder::ref_magic_of(f->chunk_rows);
  a.buffer_limit = f->buffer_limit;
  a.px_state = f->d_px_state;
  a.ring_val = f->d_ring_val;
  a.ring_some = f->d_ring_some;
  a.ring_frames = f->ring_frames;
  a.frame_px = f->frame_px;
  a.offset_max = f->d_offset_max;
  a.forced_frame = f->d_forced;
  a.err = f->d_err;
  if (!f->ingest_grid) { /* a persistent grid of exactly the resident CTAs: the grid-stride loop then runs as one wave */
    int sms = 0, per_sm = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adder::framer_ingest_kernel, 256, 0));
    f->ingest_grid = (uint32_t)(sms * std::max(per_sm, 1));
  }
  adder::framer_ingest_kernel<<<f->ingest_grid, 256, 0, f->stream>>>(a);
  CU(cudaGetLastError());
  if (int rc = framer_refresh_queue(f, 0, d_chunk_off)) return rc;
  if (!wait) return ADDER_OK; /* the caller asks later (adder_b200_framer_frame_ready) */
  if (int rc = framer_collect(f)) return rc;
  if (frame_ready) *frame_ready = (int)f->h_result[0];
  return ADDER_OK;
}

}  // namespace

extern "C" {

int adder_b200_framer_create(uint16_t width, uint16_t height, uint8_t channels, uint32_t chunk_rows, uint8_t codec_version, int time_mode,
                             uint32_t tps, uint32_t ref_interval, uint32_t delta_t_max, float output_fps, int view_mode,
                             uint32_t source_camera, int64_t buffer_limit, uint32_t ring_frames, int device, adder_b200_framer** out) {
  return guarded([&]() -> int {
    if (!out) return fail(ADDER_ERR_BAD_PARAMS, "out is NULL");
    *out = nullptr;
    if (!width || !height || !channels || !chunk_rows || !ref_interval) return fail(ADDER_ERR_BAD_PARAMS, "plane, chunk_rows and ref_interval must be non-zero");
    if (view_mode < 0 || view_mode > ADDER_VIEW_SAE || time_mode < 0 || time_mode > ADDER_TIME_MIXED) return fail(ADDER_ERR_BAD_PARAMS, "unknown mode");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(ADDER_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(ADDER_ERR_BAD_PARAMS, "device %d out of range", device);
    adder_b200_framer* f = new adder_b200_framer();
    f->w = width;
    f->h = height;
    f->c = channels;
    f->device = device;
    f->chunk_rows = chunk_rows;
    f->n_chunks = (height + chunk_rows - 1) / chunk_rows;
    f->tpf = output_fps > 0.0f ? (uint32_t)((float)tps / output_fps) : ref_interval; /* driver.rs:355-359 */
    if (f->tpf == 0) {
      delete f;
      return fail(ADDER_ERR_BAD_PARAMS, "ticks per output frame would be 0");
    }
    f->tps = tps;
    f->ref_interval = ref_interval;
    f->source_dtm = delta_t_max;
    f->codec_version = codec_version;
    f->source_camera = source_camera;
    f->view_mode = view_mode;
    f->time_mode = time_mode;
    f->buffer_limit = buffer_limit;
    f->practical_d_max = log2_raw(255.0f * (float)(delta_t_max / ref_interval)); /* driver.rs:1020-1021, u8::max_f32() = 255 */
    f->frame_px = (uint64_t)width * height * channels;
    /* a lagging pixel can hold the front frame back by delta_t_max while its neighbours run delta_t_max ahead */
    f->ring_frames = ring_frames ? ring_frames : 4u * (delta_t_max / f->tpf) + 64u;
    auto build = [&]() -> int {
      CU(cudaSetDevice(device));
      CU(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
      const uint64_t P = f->frame_px;
      CU(cudaMalloc(&f->d_px_state, P * sizeof(uint4)));
      CU(cudaMalloc(&f->d_ring_val, P * f->ring_frames));
      CU(cudaMalloc(&f->d_ring_some, P * f->ring_frames));
      CU(cudaMalloc(&f->d_offset_max, f->n_chunks * sizeof(long long)));
      CU(cudaMalloc(&f->d_forced, f->n_chunks * sizeof(long long)));
      CU(cudaMalloc(&f->d_tracker, f->n_chunks));
      CU(cudaMalloc(&f->d_status, f->n_chunks));
      CU(cudaMalloc(&f->d_result, 5 * sizeof(uint32_t))); /* [0..2] predicates, [4] the refresh kernel's CTA counter */
      CU(cudaMemsetAsync(f->d_result, 0, 5 * sizeof(uint32_t), f->stream));
      CU(cudaMalloc(&f->d_err, sizeof(uint32_t)));
      CU(cudaMalloc(&f->d_off_stage, ((size_t)f->n_chunks + 1) * sizeof(uint32_t)));
      CU(cudaHostAlloc(&f->h_result, 4 * sizeof(uint32_t), cudaHostAllocDefault));
      CU(cudaMalloc(&f->d_exact_lut, 257));
      {
        uint8_t lut[257];
        adder::build_exact_lut(ref_interval, lut);
        CU(cudaMemcpyAsync(f->d_exact_lut, lut, sizeof(lut), cudaMemcpyHostToDevice, f->stream)); /* the build ends with a stream sync */
      }
      CU(cudaMemsetAsync(f->d_ring_val, 0, P * f->ring_frames, f->stream));
      CU(cudaMemsetAsync(f->d_ring_some, 0, P * f->ring_frames, f->stream));
      CU(cudaMemsetAsync(f->d_tracker, 0, f->n_chunks, f->stream));
      CU(cudaMemsetAsync(f->d_err, 0, sizeof(uint32_t), f->stream));
      const uint64_t nn = std::max<uint64_t>(P, f->n_chunks);
      adder::framer_init_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, f->stream>>>(f->d_px_state, P, f->d_offset_max, f->d_forced, f->n_chunks);
      CU(cudaGetLastError());
      CU(cudaStreamSynchronize(f->stream));
      return ADDER_OK;
    };
    if (int rc = build()) {
      adder_b200_framer_destroy(f);
      return rc;
    }
    *out = f;
    return ADDER_OK;
  });
}

void adder_b200_framer_destroy(adder_b200_framer* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  if (f->stream) cudaStreamSynchronize(f->stream);
  cudaFree(f->d_exact_lut);
  cudaFree(f->d_px_state);
  cudaFree(f->d_ring_val);
  cudaFree(f->d_ring_some);
  cudaFree(f->d_offset_max);
  cudaFree(f->d_forced);
  cudaFree(f->d_tracker);
  cudaFree(f->d_status);
  cudaFree(f->d_result);
  cudaFree(f->d_err);
  cudaFree(f->d_off_stage);
  cudaFree(f->d_ev_stage);
  cudaFree(f->d_out_stage);
  if (f->h_result) cudaFreeHost(f->h_result);
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
}

int adder_b200_framer_ingest_events_device(adder_b200_framer* f, const adder_event_t* d_events, const uint32_t* d_chunk_offsets, int* frame_ready) {
  return guarded([&]() -> int {
    if (!f || !d_chunk_offsets) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
    if (int rc = framer_set_device(f)) return rc;
    return framer_ingest(f, d_events, d_chunk_offsets, frame_ready);
  });
}

int adder_b200_framer_ingest_events_device_async(adder_b200_framer* f, const adder_event_t* d_events, const uint32_t* d_chunk_offsets) {
  return guarded([&]() -> int {
    if (!f || !d_chunk_offsets) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
    if (int rc = framer_set_device(f)) return rc;
    return framer_ingest(f, d_events, d_chunk_offsets, nullptr, false);
  });
}

int adder_b200_framer_frame_ready(adder_b200_framer* f, int* frame_ready) {
  return guarded([&]() -> int {
    if (!f) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
    if (int rc = framer_set_device(f)) return rc;
    if (int rc = framer_collect(f)) return rc;
    if (frame_ready) *frame_ready = (int)f->h_result[0];
    return ADDER_OK;
  });
}

int adder_b200_framer_ingest_events_host(adder_b200_framer* f, const adder_event_t* events, const uint32_t* chunk_counts, int* frame_ready) {
  return guarded([&]() -> int {
    if (!f || !chunk_counts) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
    if (int rc = framer_set_device(f)) return rc;
    std::vector<uint32_t> off(f->n_chunks + 1u, 0u);
    for (uint32_t k = 0; k < f->n_chunks; k++) off[k + 1] = off[k] + chunk_counts[k];
    const uint64_t total = off[f->n_chunks];
    if (total && !events) return fail(ADDER_ERR_BAD_PARAMS, "events is NULL");
    if (total > f->ev_stage_cap) {
      if (f->d_ev_stage) CU(cudaFree(f->d_ev_stage));
      f->d_ev_stage = nullptr;
      f->ev_stage_cap = std::max<uint64_t>(total, 4096);
      CU(cudaMalloc(&f->d_ev_stage, f->ev_stage_cap * sizeof(adder_event_t)));
    }
    CU(cudaMemcpyAsync(f->d_off_stage, off.data(), off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, f->stream));
    if (total) CU(cudaMemcpyAsync(f->d_ev_stage, events, total * sizeof(adder_event_t), cudaMemcpyHostToDevice, f->stream));
    CU(cudaStreamSynchronize(f->stream)); /* `off` and the caller's buffers may go away */
    return framer_ingest(f, f->d_ev_stage, f->d_off_stage, frame_ready);
  });
}

int adder_b200_framer_write_multi_frame_bytes(adder_b200_framer* f, uint8_t* frames_out, uint32_t max_frames, uint32_t* n_frames) {
  return guarded([&]() -> int {
    if (!f || !n_frames || (max_frames && !frames_out)) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
    *n_frames = 0;
    if (int rc = framer_set_device(f)) return rc;
    if (!f->d_out_stage) {
      f->out_stage_frames = 4;
      CU(cudaMalloc(&f->d_out_stage, f->frame_px * f->out_stage_frames));
    }
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device));
    /* while self.is_frame_filled(0)? { write_frame_bytes } — one frame per pass: the status of the next front frame is only
     * known once the previous one has been popped */
    if (int rc = framer_refresh(f, 3, f->d_off_stage)) return rc;
    while (f->h_result[1] && *n_frames < max_frames) {
      adder::framer_pop_kernel<<<sms * 4, 256, 0, f->stream>>>(f->d_ring_val, f->d_ring_some, f->ring_frames, f->frame_px, f->frames_written, 1u,
                                                             f->d_out_stage);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(frames_out + (size_t)*n_frames * f->frame_px, f->d_out_stage, f->frame_px, cudaMemcpyDeviceToHost, f->stream));
      f->frames_written += 1; /* driver.rs:959 */
      *n_frames += 1;
      if (int rc = framer_refresh(f, 1, f->d_off_stage)) return rc; /* pop_next_frame_for_chunk :923-924 */
    }
    return ADDER_OK;
  });
}

int adder_b200_framer_flush_frame_buffer(adder_b200_framer* f, int* frame_ready) {
  return guarded([&]() -> int {
    if (!f) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
    if (int rc = framer_set_device(f)) return rc;
    if (int rc = framer_refresh(f, 3, f->d_off_stage)) return rc; /* query: does any chunk hold more than one frame? */
    if (f->h_result[2]) {
      int sms = 0;
      CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device));
      adder::framer_flush_kernel<<<sms * 4, 256, 0, f->stream>>>(f->d_ring_val, f->d_ring_some, f->ring_frames, f->frame_px, f->frames_written,
                                                               f->d_px_state);
      CU(cudaGetLastError());
    }
    if (int rc = framer_refresh(f, 2, f->d_off_stage)) return rc;
    if (frame_ready) *frame_ready = (int)f->h_result[0];
    return ADDER_OK;
  });
}

int adder_b200_framer_state(const adder_b200_framer* f, int64_t* frames_written, uint32_t* tpf) {
  if (!f) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  if (frames_written) *frames_written = f->frames_written;
  if (tpf) *tpf = f->tpf;
  return ADDER_OK;
}

}  /* extern "C" */

/* ================================ event exchange between row bands (include/adder_b200.h, comm section) ================================ */

struct adder_b200_comm {
  adder_b200_video* v = nullptr; /* the band (or the consumer's video) whose device and stream order this object follows */
  int device = 0;
  bool owner = false;        /* allocated the ring (the consumer) */
  bool ipc_mapped = false;   /* the ring was opened from another process's handle */
  void* base = nullptr;      /* one allocation: records | chunk offsets | totals | arrived | released */
  size_t bytes = 0;
  adder::ExchangeRing ring{};
  cudaStream_t stream = nullptr; /* pushes / waits run here, behind the video's stream, so that they overlap its next launch */
  cudaEvent_t ev = nullptr;
  cudaEvent_t ev_push[2] = {nullptr, nullptr}; /* completion of the last two push_frames calls */
  uint64_t n_push = 0;
  uint32_t* d_local_done = nullptr; /* [kMaxFramesPerLaunch] */
  uint32_t* d_err = nullptr;
  uint32_t* h_err = nullptr;
  unsigned long long released = 0; /* consumer: frames released so far */
};

namespace {

struct CommBlob { /* what crosses the process boundary: ADDER_COMM_BLOB_BYTES */
  char magic[8];
  cudaIpcMemHandle_t handle; /* 64 bytes */
  uint64_t bytes, out_stride;
  uint32_t slots, world, total_chunks, device;
};
static_assert(sizeof(CommBlob) <= ADDER_COMM_BLOB_BYTES, "blob too large");

size_t align256(size_t x) { return (x + 255u) & ~(size_t)255u; }

/* offsets of the ring's parts inside its one allocation */
void ring_layout(void* base, uint32_t slots, uint32_t world, uint32_t total_chunks, uint64_t out_stride, adder::ExchangeRing* r, size_t* bytes) {
  size_t o = 0;
  uint8_t* b = (uint8_t*)base;
  r->ev_words = (uint32_t*)(b + o);
  o += align256((size_t)slots * out_stride * sizeof(adder_event_t));
  r->chunk_off = (uint32_t*)(b + o);
  o += align256((size_t)slots * (total_chunks + 1u) * sizeof(uint32_t));
  r->totals = (unsigned long long*)(b + o);
  o += align256((size_t)slots * world * sizeof(unsigned long long));
  r->arrived = (unsigned long long*)(b + o);
  o += align256((size_t)slots * sizeof(unsigned long long));
  r->released = (unsigned long long*)(b + o);
  o += 256;
  r->slots = slots;
  r->world = world;
  r->total_chunks = total_chunks;
  r->out_stride = out_stride;
  if (bytes) *bytes = o;
}

int comm_common_init(adder_b200_comm* c) {
  CU(cudaSetDevice(c->device));
  c->v->reserve_ctas = kPushCtas; /* from now on this video's integrate launches leave room for the push kernel */
  int lo = 0, hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CU(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi)); /* small kernels that should not queue behind the persistent one */
  CU(cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_push[0], cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_push[1], cudaEventDisableTiming));
  CU(cudaMalloc(&c->d_local_done, kMaxFramesPerLaunch * sizeof(uint32_t)));
  CU(cudaMalloc(&c->d_err, sizeof(uint32_t)));
  CU(cudaMemsetAsync(c->d_err, 0, sizeof(uint32_t), c->stream));
  CU(cudaHostAlloc(&c->h_err, sizeof(uint32_t), cudaHostAllocDefault));
  return ADDER_OK;
}

}  // namespace

extern "C" {

int adder_b200_comm_create(adder_b200_video* v, uint32_t world, uint32_t total_chunks, uint32_t slots, size_t out_stride,
                           adder_b200_comm** out) {
  return guarded([&]() -> int {
    if (!v || !out || world == 0 || slots == 0 || total_chunks == 0 || out_stride == 0) return fail(ADDER_ERR_BAD_PARAMS, "bad argument");
    *out = nullptr;
    adder_b200_comm* c = new adder_b200_comm();
    c->v = v;
    c->device = v->device;
    c->owner = true;
    auto build = [&]() -> int {
      if (int rc = comm_common_init(c)) return rc;
      ring_layout(nullptr, slots, world, total_chunks, out_stride, &c->ring, &c->bytes);
      CU(cudaMalloc(&c->base, c->bytes));
      ring_layout(c->base, slots, world, total_chunks, out_stride, &c->ring, nullptr);
      /* totals / arrived / released start at zero: no frame has been published */
      CU(cudaMemsetAsync(c->ring.totals, 0, (uint8_t*)c->base + c->bytes - (uint8_t*)c->ring.totals, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      return ADDER_OK;
    };
    if (int rc = build()) {
      adder_b200_comm_destroy(c);
      return rc;
    }
    *out = c;
    return ADDER_OK;
  });
}

int adder_b200_comm_export(adder_b200_comm* c, uint8_t* blob, size_t cap) {
  if (!c || !blob || cap < ADDER_COMM_BLOB_BYTES) return fail(ADDER_ERR_BAD_PARAMS, "blob must hold ADDER_COMM_BLOB_BYTES");
  if (!c->owner) return fail(ADDER_ERR_BAD_PARAMS, "only the consumer's comm can be exported");
  CU(cudaSetDevice(c->device));
  CommBlob b{};
  memcpy(b.magic, "ADDRXCH1", 8);
  CU(cudaIpcGetMemHandle(&b.handle, c->base));
  b.bytes = c->bytes;
  b.out_stride = c->ring.out_stride;
  b.slots = c->ring.slots;
  b.world = c->ring.world;
  b.total_chunks = c->ring.total_chunks;
  b.device = (uint32_t)c->device;
  memset(blob, 0, ADDER_COMM_BLOB_BYTES);
  memcpy(blob, &b, sizeof(b));
  return ADDER_OK;
}

int adder_b200_comm_open(adder_b200_video* v, const uint8_t* blob, size_t blob_bytes, adder_b200_comm** out) {
  return guarded([&]() -> int {
    if (!v || !blob || !out || blob_bytes < sizeof(CommBlob)) return fail(ADDER_ERR_BAD_PARAMS, "bad argument");
    *out = nullptr;
    CommBlob b;
    memcpy(&b, blob, sizeof(b));
    if (memcmp(b.magic, "ADDRXCH1", 8) != 0) return fail(ADDER_ERR_BAD_PARAMS, "not an exchange blob");
    adder_b200_comm* c = new adder_b200_comm();
    c->v = v;
    c->device = v->device;
    auto build = [&]() -> int {
      if (int rc = comm_common_init(c)) return rc;
      CU(cudaIpcOpenMemHandle(&c->base, b.handle, cudaIpcMemLazyEnablePeerAccess)); /* maps the consumer's ring over NVLink */
      c->ipc_mapped = true;
      c->bytes = b.bytes;
      ring_layout(c->base, b.slots, b.world, b.total_chunks, b.out_stride, &c->ring, nullptr);
      return ADDER_OK;
    };
    if (int rc = build()) {
      adder_b200_comm_destroy(c);
      return rc;
    }
    *out = c;
    return ADDER_OK;
  });
}

int adder_b200_comm_attach(adder_b200_video* v, adder_b200_comm* consumer, adder_b200_comm** out) {
  return guarded([&]() -> int {
    if (!v || !consumer || !out || !consumer->owner) return fail(ADDER_ERR_BAD_PARAMS, "bad argument");
    *out = nullptr;
    adder_b200_comm* c = new adder_b200_comm();
    c->v = v;
    c->device = v->device;
    auto build = [&]() -> int {
      if (int rc = comm_common_init(c)) return rc;
      if (c->device != consumer->device) { /* same process, another GPU: plain peer access */
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, c->device, consumer->device));
        if (!can) return fail(ADDER_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", c->device, consumer->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(consumer->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
        (void)cudaGetLastError();
      }
      c->base = consumer->base;
      c->bytes = consumer->bytes;
      c->ring = consumer->ring;
      return ADDER_OK;
    };
    if (int rc = build()) {
      adder_b200_comm_destroy(c);
      return rc;
    }
    *out = c;
    return ADDER_OK;
  });
}

void adder_b200_comm_destroy(adder_b200_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->ipc_mapped && c->base) cudaIpcCloseMemHandle(c->base);
  if (c->owner) cudaFree(c->base);
  cudaFree(c->d_local_done);
  cudaFree(c->d_err);
  if (c->h_err) cudaFreeHost(c->h_err);
  if (c->ev) cudaEventDestroy(c->ev);
  if (c->ev_push[0]) cudaEventDestroy(c->ev_push[0]);
  if (c->ev_push[1]) cudaEventDestroy(c->ev_push[1]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int adder_b200_comm_push_frames(adder_b200_comm* c, uint32_t band, uint32_t chunk0, const adder_event_t* d_events, size_t events_stride,
                                const uint32_t* d_chunk_offsets, uint32_t n_frames, uint64_t frame_seq0) {
  return guarded([&]() -> int {
    if (!c || !d_chunk_offsets || (!d_events && events_stride)) return fail(ADDER_ERR_BAD_PARAMS, "NULL argument");
    if (band >= c->ring.world) return fail(ADDER_ERR_BAD_PARAMS, "band %u of %u", band, c->ring.world);
    if (chunk0 + c->v->n_chunks > c->ring.total_chunks) return fail(ADDER_ERR_BAD_PARAMS, "the band's chunks do not fit the frame's");
    CU(cudaSetDevice(c->device));
    /* behind whatever the band's stream has queued (the integrate launch that produces these frames) */
    CU(cudaEventRecord(c->ev, c->v->stream));
    CU(cudaStreamWaitEvent(c->stream, c->ev, 0));
    for (uint32_t f0 = 0; f0 < n_frames; f0 += kMaxFramesPerLaunch) {
      const uint32_t n = std::min<uint32_t>(kMaxFramesPerLaunch, n_frames - f0);
      CU(cudaMemsetAsync(c->d_local_done, 0, n * sizeof(uint32_t), c->stream));
      adder::PushArgs a{};
      a.ring = c->ring;
      a.ev_words = reinterpret_cast<const uint32_t*>(d_events) + (size_t)f0 * events_stride * 3u;
      a.ev_stride = events_stride;
      a.chunk_off = d_chunk_offsets + (size_t)f0 * (c->v->n_chunks + 1u);
      a.n_chunks = c->v->n_chunks;
      a.chunk0 = chunk0;
      a.band = band;
      a.n_frames = n;
      a.seq0 = frame_seq0 + f0;
      a.local_done = c->d_local_done;
      a.err = c->d_err;
      /* as many CTAs as the band's integrate launches leave room for (launch_grid): they run beside the next launch */
      adder::exchange_push_kernel<<<kPushCtas, 256, 0, c->stream>>>(a);
      CU(cudaGetLastError());
    }
    /* Two batches may be outstanding (the caller alternates two sets of buffers): whatever the band's stream is given
     * after this call — the integrate launch that overwrites the buffers of the call before this one — waits for that
     * earlier push, while this one still overlaps it. */
    CU(cudaEventRecord(c->ev_push[c->n_push & 1u], c->stream));
    if (c->n_push >= 1) CU(cudaStreamWaitEvent(c->v->stream, c->ev_push[(c->n_push - 1u) & 1u], 0));
    c->n_push++;
    return ADDER_OK;
  });
}

int adder_b200_comm_wait_frames(adder_b200_comm* c, uint64_t frame_seq0, uint32_t n_frames) {
  if (!c || !c->owner) return fail(ADDER_ERR_BAD_PARAMS, "only the consumer waits for frames");
  if (n_frames > c->ring.slots) return fail(ADDER_ERR_BAD_PARAMS, "more frames than ring slots");
  CU(cudaSetDevice(c->device));
  adder::exchange_wait_kernel<<<1, 1, 0, c->stream>>>(c->ring, frame_seq0, n_frames, c->d_err);
  CU(cudaGetLastError());
  return ADDER_OK;
}

int adder_b200_comm_release_frames(adder_b200_comm* c, uint64_t upto_seq) {
  if (!c || !c->owner) return fail(ADDER_ERR_BAD_PARAMS, "only the consumer releases frames");
  if (upto_seq < c->released) return fail(ADDER_ERR_BAD_PARAMS, "frames up to %llu were already released", (unsigned long long)c->released);
  if (upto_seq - c->released > c->ring.slots) return fail(ADDER_ERR_BAD_PARAMS, "more frames than ring slots");
  CU(cudaSetDevice(c->device));
  adder::exchange_release_kernel<<<1, 1, 0, c->stream>>>(c->ring, c->released, upto_seq);
  CU(cudaGetLastError());
  c->released = upto_seq;
  return ADDER_OK;
}

int adder_b200_comm_frame(adder_b200_comm* c, uint64_t frame_seq, adder_event_t** d_events, uint32_t** d_chunk_offsets) {
  if (!c || !c->owner) return fail(ADDER_ERR_BAD_PARAMS, "only the consumer holds frames");
  const size_t slot = (size_t)(frame_seq % c->ring.slots);
  if (d_events) *d_events = reinterpret_cast<adder_event_t*>(c->ring.ev_words) + slot * c->ring.out_stride;
  if (d_chunk_offsets) *d_chunk_offsets = c->ring.chunk_off + slot * (c->ring.total_chunks + 1u);
  return ADDER_OK;
}

int adder_b200_comm_sync(adder_b200_comm* c) {
  if (!c) return fail(ADDER_ERR_BAD_PARAMS, "NULL handle");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  const uint32_t e = *c->h_err;
  if (e) CU(cudaMemsetAsync(c->d_err, 0, sizeof(uint32_t), c->stream));
  if (e & 4u) return fail(ADDER_ERR_INTERNAL, "event exchange: a peer did not show up within the time limit");
  if (e & 1u) return fail(ADDER_ERR_CAPACITY, "event exchange: a frame's events do not fit the consumer's slot");
  return ADDER_OK;
}

void* adder_b200_comm_stream(adder_b200_comm* c) { return c ? (void*)c->stream : nullptr; }

}  /* extern "C" */
