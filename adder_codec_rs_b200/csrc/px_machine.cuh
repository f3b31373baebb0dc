/*
 * px_machine.cuh — one pixel-channel, one frame: the FramePerfect state machine of the reference,
 * written once for the device (px_kernel.cuh) and, for local checking without a GPU, for the host
 * (tests/host_sim builds this header with g++ and compares it with the oracle; the product never
 * runs it on the CPU).
 *
 * It restates
 *   integrate_for_px            adder-codec-rs/src/transcoder/source/video.rs:1317-1380
 *   PixelArena::pop_best_events adder-codec-rs/src/transcoder/event_pixel_tree.rs:213-287
 *   PixelArena::integrate       :317-413,  integrate_main :418-479
 *   PixelArena::pop_top_event   :139-210,  delta_t_to_absolute_t :113-137
 *   u8::get_frame_value         adder-codec-rs/src/framer/scale_intensity.rs:58-104
 * as ONE streaming pass over the pixel's node stack: node k is read once (Mem::load), updated in
 * registers and written once to level k - shift (Mem::store), where shift = 1 iff the root is popped
 * this frame — the reference's arena shift (:201-204) becomes a change of destination level instead
 * of extra traffic.  Events go to Sink::push in the reference's push order.
 *
 * All f32/f64 arithmetic goes through the rn_* wrappers: explicit round-to-nearest intrinsics on the
 * device (never contracted into an FMA), plain operators on the host (built with -ffp-contract=off).
 *
 * Three things are done differently from a literal transcription, each with identical results:
 *  - `u / ref_time` in delta_t_to_absolute_t is a multiply by ceil(2^64 / ref) (exact for every u32);
 *  - the Intensity display byte floor(fl64(fl64(2^d / t) * ref)) is first estimated in f32 (good to
 *    ~1e-4 absolute below 256, the f64 value to ~1e-13, so away from integers both truncate alike);
 *    within 2^-10 of an integer the side is decided by an exact u64 comparison, and the exactly
 *    integral case by a 257-entry table of the reference's own f64 expression (frame_value_u8);
 *  - the display byte is not recomputed while the root's best event is the one it was computed
 *    from (same inputs -> same byte); PxParams::display == 2 forces it after a parameter change.
 */
#pragma once
#include <stdint.h>
#include <string.h>

#include "state_layout.h"

#ifndef ADDER_FAST_WALK
#define ADDER_FAST_WALK 0 /* 1: the unshifted walk of px_step through a running pointer and two node registers used in turn:
                           * half the instructions per level in the SASS, parity-green, and 7 % SLOWER on every workload
                           * (profiles/r02m_ab_fastwalk.txt) — like every variant of this kernel that adds code */
#endif
#ifndef ADDER_HOIST_FIRE
#define ADDER_HOIST_FIRE 0 /* 1: the firing node's arithmetic after the level walk instead of inside it (+2 % on noise and
                            * jitter c = 5, -3 % on aged 8K stacks, profiles/r02k_ab_hoist.txt: off) */
#endif
#ifndef ADDER_D_MAX
#define ADDER_D_MAX 127u
#define ADDER_D_ZERO_INTEGRATION 128u
#define ADDER_D_EMPTY 255u
#define ADDER_C_NONE 0xFFu
#endif

#if defined(__CUDACC__)
#define ADDER_HD __host__ __device__ __forceinline__
#else
#define ADDER_HD inline
#endif

namespace adder {

#if defined(__CUDA_ARCH__)
ADDER_HD float rn_add(float a, float b) { return __fadd_rn(a, b); }
ADDER_HD float rn_sub(float a, float b) { return __fsub_rn(a, b); }
ADDER_HD float rn_mul(float a, float b) { return __fmul_rn(a, b); }
ADDER_HD float rn_div(float a, float b) { return __fdiv_rn(a, b); }
ADDER_HD double rn_ddiv(double a, double b) { return __ddiv_rn(a, b); }
ADDER_HD double rn_dmul(double a, double b) { return __dmul_rn(a, b); }
ADDER_HD uint32_t f2u(float x) { return __float2uint_rz(x); }   /* Rust `as u32`: truncating, saturating, NaN -> 0 */
ADDER_HD uint32_t d2u(double x) { return __double2uint_rz(x); }
ADDER_HD float u2f(uint32_t x) { return __uint2float_rn(x); }
ADDER_HD uint32_t clz32(uint32_t x) { return (uint32_t)__clz((int)x); }
ADDER_HD uint32_t f_bits(float x) { return __float_as_uint(x); }
ADDER_HD float bits_f(uint32_t x) { return __uint_as_float(x); }
ADDER_HD double bits_d(uint64_t x) { return __longlong_as_double((long long)x); }
ADDER_HD float fast_div(float a, float b) { return __fdividef(a, b); } /* estimate only, never a result */
ADDER_HD uint32_t mulhi_u32_u64(uint32_t u, uint64_t m) { /* floor(u * m / 2^64) */
  const uint64_t t = (uint64_t)u * (uint32_t)m;
  const uint64_t s = (uint64_t)u * (uint32_t)(m >> 32) + (t >> 32);
  return (uint32_t)(s >> 32);
}
#else
ADDER_HD float rn_add(float a, float b) { volatile float r = a + b; return r; }
ADDER_HD float rn_sub(float a, float b) { volatile float r = a - b; return r; }
ADDER_HD float rn_mul(float a, float b) { volatile float r = a * b; return r; }
ADDER_HD float rn_div(float a, float b) { volatile float r = a / b; return r; }
ADDER_HD double rn_ddiv(double a, double b) { volatile double r = a / b; return r; }
ADDER_HD double rn_dmul(double a, double b) { volatile double r = a * b; return r; }
ADDER_HD uint32_t f2u(float x) { if (!(x > 0.0f)) return 0u; if (x >= 4294967296.0f) return 0xFFFFFFFFu; return (uint32_t)x; }
ADDER_HD uint32_t d2u(double x) { if (!(x > 0.0)) return 0u; if (x >= 4294967296.0) return 0xFFFFFFFFu; return (uint32_t)x; }
ADDER_HD float u2f(uint32_t x) { return (float)x; }
ADDER_HD uint32_t clz32(uint32_t x) { return x ? (uint32_t)__builtin_clz(x) : 32u; }
ADDER_HD uint32_t f_bits(float x) { uint32_t b; memcpy(&b, &x, 4); return b; }
ADDER_HD float bits_f(uint32_t x) { float f; memcpy(&f, &x, 4); return f; }
ADDER_HD double bits_d(uint64_t x) { double f; memcpy(&f, &x, 8); return f; }
#ifdef ADDER_HOST_SIM /* tests perturb the estimate by a few ulp to stand for __fdividef's error */
extern int g_fast_div_ulps;
ADDER_HD float fast_div(float a, float b) {
  float r = a / b;
  uint32_t bits = f_bits(r);
  if (r > 0.0f && r < 3.0e38f) bits = (uint32_t)((int32_t)bits + g_fast_div_ulps);
  return bits_f(bits);
}
#else
ADDER_HD float fast_div(float a, float b) { return a / b; }
#endif
ADDER_HD uint32_t mulhi_u32_u64(uint32_t u, uint64_t m) { return (uint32_t)(((unsigned __int128)u * m) >> 64); }
#endif

/* per-frame constants of the state machine (VideoStateParams video.rs:160-182, CrfParameters
 * rate_controller.rs:40-53, and what integrate_matrix derives per frame, video.rs:660-672) */
struct PxParams {
  float time, running_t_prev, running_t, dtm_f;
  uint32_t ref, dtm;
  uint64_t ref_magic; /* ceil(2^64 / ref); 0 stands for ref == 1 (see ref_magic_of) */
  uint32_t c_max, vel_m1, cnt_inc;
  uint32_t collapse, abs_time, view_mode;
  uint32_t display; /* 0 off, 1 on, 2 on and recompute every pixel whose root holds a best event */
  uint32_t depth;
  double tpf;
  float tpf_f; /* (float)tpf, for the display estimate only */
  const uint8_t* exact_lut; /* [257]: min(255, trunc(fl64(fl64(k / ref) * ref))), build_exact_lut() */
  float practical_d_max;
};

inline uint64_t ref_magic_of(uint32_t ref) { /* host side */
  if (ref <= 1u) return 0ull;
  const uint64_t q = ~0ull / ref; /* floor((2^64 - 1) / ref) */
  return q + 1ull;                /* = ceil(2^64 / ref) for every ref >= 2 (also powers of two) */
}
/* The Intensity display byte of an event whose 2^d / t * ref is exactly the integer k (host side, f64
 * like the reference: scale_intensity.rs:262-270 then :58-68). */
inline void build_exact_lut(uint32_t ref, uint8_t out[257]) {
  for (uint32_t k = 0; k <= 256u; k++) {
    volatile double q = (double)k / (double)ref;
    volatile double r = q * (double)ref;
    out[k] = (uint8_t)(r >= 255.0 ? 255u : (uint32_t)r);
  }
}
/* u / ref for any u32 (error term u*e/(ref*2^64) < 2^-32 < 1/ref, so the floor is exact) */
ADDER_HD uint32_t div_ref(const PxParams& a, uint32_t u) { return a.ref_magic ? mulhi_u32_u64(u, a.ref_magic) : u; }

struct Node {
  float integ, dt, best_dt;
  uint32_t w; /* d | best_d<<8 | has_best<<16 */
};

struct PxHeader {
  float lf;   /* last_fired_t */
  uint32_t y; /* packed, state_layout.h */
};

ADDER_HD uint32_t get_d_from_intensity(float x) { /* event_pixel_tree.rs:482-499 */
  if (x < 1.0f) return ADDER_D_ZERO_INTEGRATION;
  uint32_t d = ((f_bits(x) >> 23) & 0xFFu) - 127u; /* x >= 1 (or NaN/inf): exponent >= 127 */
  return d > ADDER_D_MAX ? ADDER_D_MAX : d;
}
ADDER_HD float d_shift_f32(uint32_t d) { /* D_SHIFT_F32, lib.rs:229-235: [128] = 0 */
  return d >= 128u ? 0.0f : bits_f((d + 127u) << 23);
}
ADDER_HD Node fresh_node(float intensity) { /* PixelNode::new :502-514 */
  Node n;
  n.integ = 0.0f;
  n.dt = 0.0f;
  n.best_dt = 0.0f;
  n.w = NODE_PACK(get_d_from_intensity(intensity), 0, 0);
  return n;
}

/* integrate_main, FramePerfect arm (:418-479).  Returns true when the node fires. */
ADDER_HD bool integrate_main(Node& n, float intensity, float time) {
  const uint32_t d = NODE_D(n.w);
  const float sum = rn_add(n.integ, intensity);
  if (sum >= d_shift_f32(d)) {
    const uint32_t nd = get_d_from_intensity(sum);
    float prop = rn_div(rn_sub(d_shift_f32(nd), n.integ), intensity);
    if (nd == ADDER_D_ZERO_INTEGRATION || d == ADDER_D_ZERO_INTEGRATION || intensity < 1.1920929e-07f) prop = 1.0f;
    n.best_dt = rn_add(n.dt, rn_mul(time, prop)); /* :445, two roundings */
    uint32_t d_after = nd;
    if (nd < ADDER_D_MAX) { /* :449-461; the D_SHIFT walk always ends at nd+1 because 2^(nd+1) > sum */
      n.integ = sum;
      n.dt = rn_add(n.dt, time);
      d_after = nd + 1u;
    }
    n.w = NODE_PACK(d_after, nd, 1);
    return true;
  }
  n.integ = sum;
  n.dt = rn_add(n.dt, time);
  return false;
}

/* integrate_main in two halves, for walks whose lanes fire at different levels: integrate_accumulate does the non-firing
 * arm and only REPORTS a fire (node untouched); integrate_fire is the firing arm, run once after the walk by all lanes
 * that reported one — inside the level loop a warp would run its division once per distinct firing level of its 32 pixels. */
ADDER_HD bool integrate_accumulate(Node& n, float intensity, float time) {
  const float sum = rn_add(n.integ, intensity);
  if (sum >= d_shift_f32(NODE_D(n.w))) return true;
  n.integ = sum;
  n.dt = rn_add(n.dt, time);
  return false;
}
ADDER_HD void integrate_fire(Node& n, float intensity, float time) {
  const uint32_t d = NODE_D(n.w);
  const float sum = rn_add(n.integ, intensity);
  const uint32_t nd = get_d_from_intensity(sum);
  float prop = rn_div(rn_sub(d_shift_f32(nd), n.integ), intensity);
  if (nd == ADDER_D_ZERO_INTEGRATION || d == ADDER_D_ZERO_INTEGRATION || intensity < 1.1920929e-07f) prop = 1.0f;
  n.best_dt = rn_add(n.dt, rn_mul(time, prop)); /* :445, two roundings */
  uint32_t d_after = nd;
  if (nd < ADDER_D_MAX) {
    n.integ = sum;
    n.dt = rn_add(n.dt, time);
    d_after = nd + 1u;
  }
  n.w = NODE_PACK(d_after, nd, 1);
}

/* u8::get_frame_value, SourceType::U8 arm (scale_intensity.rs:58-104, :262-270) */
/* kPlain = the default output configuration, TimeMode::AbsoluteT with FramedViewMode::Intensity, as compile-time
 * constants: the other view modes and the DeltaT arm drop out of the kernel (-8 % code, +4.6 % on noise,
 * profiles/r02n_ab_fixview.txt). */
#define ADDER_VIEW_OF(a) (kPlain ? 0u : (a).view_mode)
#define ADDER_ABS_OF(a) (kPlain ? 1u : (a).abs_time)
template <bool kPlain = false>
ADDER_HD uint8_t frame_value_u8(const PxParams& a, uint32_t d, uint32_t t, float lf) {
  float q;
  switch (ADDER_VIEW_OF(a)) {
    case 0: { /* Intensity: f64 in the reference */
      if (d >= 128u) return 0; /* D_SHIFT_F64[128] = 0; d >= 129 -> 0 (:262-270) */
      const uint32_t tt = t == 0u ? 1u : t; /* t == 0 -> the intensity is 2^d itself */
      const float est = rn_mul(fast_div(bits_f((d + 127u) << 23), u2f(tt)), a.tpf_f);
      if (!(est < 256.5f)) return 255; /* also +inf; the exact value is > 256 */
      const uint32_t k = f2u(est);
      const float fr = rn_sub(est, u2f(k));
      if (fr > 0.0009765625f && fr < 0.9990234375f) return (uint8_t)(k > 255u ? 255u : k);
      if (d < 32u) {
        /* X = 2^d*ref/tt is within 2^-9 of the integer kc.  If it is not kc itself it is at least
         * 1/tt >= 2^-32 away, far more than the 2^-44 the two f64 roundings can move it, so the
         * side it lies on decides the byte; if it IS kc the reference computes fl(fl(kc/ref)*ref),
         * which depends on (kc, ref) only and is tabulated by the host in f64. */
        const uint32_t kc = fr < 0.5f ? k : k + 1u;
        const uint64_t n = (uint64_t)a.ref << d, m = (uint64_t)tt * kc;
        if (n == m) return a.exact_lut[kc];
        const uint32_t r = n > m ? kc : kc - 1u;
        return (uint8_t)(r > 255u ? 255u : r);
      }
      const double p = bits_d((uint64_t)(d + 1023u) << 52); /* D_SHIFT_F64[d] */
      const double inten = t == 0u ? p : rn_ddiv(p, (double)t);
      const uint32_t u = d2u(rn_dmul(inten, a.tpf)); /* saturating, NaN -> 0 */
      return (uint8_t)(u > 255u ? 255u : u);
    }
    case 1: q = rn_div((float)d, a.practical_d_max); break;
    case 2: q = rn_div(u2f(t), u2f(a.dtm)); break;
    default: { /* SAE */
      const uint32_t diff = f2u(a.running_t) - f2u(lf);
      q = rn_div(u2f(diff), u2f(a.dtm));
      break;
    }
  }
  const uint32_t u = f2u(rn_mul(q, 255.0f));
  return (uint8_t)(u > 255u ? 255u : u);
}

/* delta_t_to_absolute_t, FramePerfect (:113-137), then Sink::push(d, t) */
template <bool kPlain = false, class Sink>
ADDER_HD void emit_abs(const PxParams& a, Sink& sink, float& lf, uint32_t d, float dt) {
  if (ADDER_ABS_OF(a)) {
    dt = rn_add(dt, lf);
    const uint32_t u = f2u(dt);
    const uint32_t m = div_ref(a, u) * a.ref;
    lf = u2f(m == u ? u : m + a.ref);
    sink.push(d, u);
  } else {
    sink.push(d, f2u(dt));
  }
}

/* One node of pop_best_events (:223-247): its best event, or a zero event, if it has one to give. */
template <bool kPlain = false, class Sink>
ADDER_HD bool pop_node(const PxParams& a, Sink& sink, float& lf, Node& nk) {
  uint32_t ed;
  float edt;
  if (NODE_HAS_BEST(nk.w)) {
    ed = NODE_BEST_D(nk.w);
    edt = nk.best_dt;
  } else if (nk.dt > 0.0f && nk.integ == 0.0f) { /* get_zero_event(idx, None) :96-111 */
    ed = ADDER_D_ZERO_INTEGRATION;
    edt = nk.dt;
    nk.dt = 0.0f;
  } else {
    return false;
  }
  emit_abs<kPlain>(a, sink, lf, ed, edt);
  return true;
}

/*
 * One pixel, one frame.  `v` is the u8 sample, `h` the pixel's header (updated in place), `n0` its
 * root node and `n1` its level-1 node as loaded (n1 is only looked at when the stack has a level 1;
 * the kernel fetches both before it knows the length).  Returns true and sets *disp when
 * running_intensities must be written.
 */
/* kDefer: an unshifted, fully integrating walk stops after level 1 and reports *deferred = true (header written with the
 * old length): levels 2.. are then walked by deep_item / deep_finish below — in the kernel by the lanes of the warp
 * together, because the lanes' stacks differ in depth and a per-lane loop runs as long as the deepest of 32. */
template <bool kDefer = false, bool kPlain = false, class Mem, class Sink>
ADDER_HD bool px_step(const PxParams& a, uint32_t v, PxHeader& h, Node n0, Node n1, Mem& mem, Sink& sink, uint32_t& errbits,
                      uint8_t* disp, bool* deferred = nullptr) {
  const float intensity = (float)v; /* matrix.mapv(f32::from), video.rs:665 */
  const float time = a.time;
  float lf = h.lf;
  uint32_t base = HDR_BASE(h.y), cth = HDR_CTHRESH(h.y), cnt = HDR_COUNTER(h.y);
  uint32_t len = HDR_LENGTH(h.y);
  uint32_t popped = HDR_POPPED(h.y);
  mem.prefetch_levels(len);
  bool root_new = false; /* the root's best event after this frame is not the one of the last frame */

  /* ---- video.rs:1338-1358: the pixel changed by more than c_thresh -> pop_best_events -------- */
  const uint32_t lo = base > cth ? base - cth : 0u;          /* saturating_sub */
  const uint32_t hi = base + cth > 255u ? 255u : base + cth; /* saturating_add */
  if (v < lo || v > hi) {
    /* Collapse after a Δt_max pop keeps only the first event (:249-265): later nodes would only
     * touch last_fired_t, which is overwritten below, and a root that is replaced. */
    const bool first_only = popped && a.collapse;
    bool any = pop_node<kPlain>(a, sink, lf, n0);
    if (len > 1u && !(first_only && any)) {
      mem.used_preloaded(); /* n1 */
      any |= pop_node<kPlain>(a, sink, lf, n1);
      n0 = n1; /* the tail so far */
      for (uint32_t k = 2; k < len && !(first_only && any); k++) {
        n0 = mem.load(k);
        any |= pop_node<kPlain>(a, sink, lf, n0);
      }
    }
    if (first_only && any) {
      lf = a.running_t_prev; /* running_t before this frame's `+= time` (:337 runs later) */
      sink.push(ADDER_D_EMPTY, f2u(a.running_t_prev));
      n0 = fresh_node(intensity);
    } /* else :267-270: the tail (now in n0) becomes the root */
    len = 1;
    popped = 0;
    base = v;
    root_new = true;
  }

  /* ---- integrate (:317-413), node 0 ---------------------------------------------------------- */
  if (len == 1u && n0.dt == 0.0f && n0.integ == 0.0f) /* :332-335, the tail is the root */
    n0.w = (n0.w & ~0xFFu) | get_d_from_intensity(intensity);
  const bool fired0 = integrate_main(n0, intensity, time);
  const bool only_root = popped && a.collapse;             /* :360-362 */
  const uint32_t dtm_reached = n0.dt >= a.dtm_f ? 1u : 0u; /* :394 */
  const bool need_pop = NODE_D(n0.w) == ADDER_D_MAX || (dtm_reached && !popped);

  if (cth < a.c_max) { /* :402-412 */
    if (cnt >= a.vel_m1) {
      cth = cth + 1u > 255u ? 255u : cth + 1u;
      cnt = 0;
    } else {
      cnt = cnt + a.cnt_inc > 255u ? 255u : cnt + a.cnt_inc;
    }
  }

  /* ---- the rest of the stack + pop_top_event (video.rs:1371-1374, :139-210) ------------------ */
  bool disp_has = false;
  uint32_t disp_d = 0;
  float disp_dt = 0.0f;
  uint32_t new_len = 1;
  if (fired0) { /* :344-355: child seeded from this intensity, deeper nodes dropped (FramePerfect :366) */
    root_new = true;
    if (need_pop) {
      emit_abs<kPlain>(a, sink, lf, NODE_BEST_D(n0.w), n0.best_dt);
      popped = 1;
      mem.store_fresh(0, fresh_node(intensity));
    } else {
      mem.store(0, n0);
      if (a.depth > 1u) mem.store_fresh(1, fresh_node(intensity)); else errbits |= ADDER_DEVERR_DEPTH;
      new_len = 2;
      disp_has = true;
      disp_d = NODE_BEST_D(n0.w);
      disp_dt = n0.best_dt;
    }
  } else {
    uint32_t shift = 0;
    bool cut = false;
    if (need_pop) {
      popped = 1;
      mem.set_popped(); /* offset form (OffsetAdaptor below): the levels stored from here on are never integrated again under Collapse */
      root_new = true;
      if (!NODE_HAS_BEST(n0.w)) {
        if (n0.integ == 0.0f && n0.dt > 0.0f) { /* zero event, :155-160 */
          emit_abs<kPlain>(a, sink, lf, ADDER_D_ZERO_INTEGRATION, n0.dt);
          n0.dt = 0.0f;
          n0.w = (n0.w & ~0xFFu) | get_d_from_intensity(intensity);
          mem.store(0, n0);
        } else { /* :164-193 synthesise a best event, then pop it: the root becomes a fresh node */
          const uint32_t sd = n0.integ < 1.0f ? ADDER_D_ZERO_INTEGRATION : 31u - clz32(f2u(n0.integ));
          emit_abs<kPlain>(a, sink, lf, sd, n0.dt);
          mem.store_fresh(0, fresh_node(intensity));
          cut = true;
        }
      } else {
        emit_abs<kPlain>(a, sink, lf, NODE_BEST_D(n0.w), n0.best_dt);
        shift = 1;
        if (len < 2u) errbits |= ADDER_DEVERR_INTERNAL;
      }
    } else {
      mem.store(0, n0);
      if (NODE_HAS_BEST(n0.w)) {
        disp_has = true;
        disp_d = NODE_BEST_D(n0.w);
        disp_dt = n0.best_dt;
      }
    }
    if (!cut) {
      new_len = len;
      /* Collapse after a Δt_max pop integrates the root only (:360-362): unless the stack moves up
       * a level, nothing below the root changes and nothing below it is read. */
      const uint32_t k_end = (only_root && !shift) ? 1u : len;
      Node nk = n1; /* level 1 is already here; level k+1 is requested before level k is worked on */
      if (k_end > 1u) mem.used_preloaded(); /* n1 */
      const bool defer_deep = kDefer && !shift && !only_root;
      uint32_t kf = 0; /* the level whose node fired (its arithmetic is done after the walk: ADDER_HOIST_FIRE) */
#if ADDER_FAST_WALK
      /* The common walk — nothing moves up a level, every level integrates — written for the instruction count of one
       * level (the general loop below costs 65-95 per level: address re-computation, a four-register rotation and the
       * tail test on every level): a running pointer stepped by the record layout, two node registers used in turn so
       * that nothing is copied, the next level requested before this one is worked on, the fresh-tail fix only at the
       * tail, and the firing node finished once after the walk. */
      if (!kDefer && !shift && !only_root) {
        if (len > 1u) {
          const uint32_t last = len - 1u;
          Node na = nk, nb = nk;
          auto qa = mem.cursor_level1(), qb = qa, qf = qa;
          uint32_t k = 1;
          for (;;) {
            /* an odd level, in na */
            if (k != last) {
              qb = mem.next_from_odd(qa);
              nb = mem.load_at(qb);
            } else if (na.dt == 0.0f && na.integ == 0.0f) {
              na.w = (na.w & ~0xFFu) | get_d_from_intensity(intensity); /* :332-335 */
            }
            if (integrate_accumulate(na, intensity, time)) {
              kf = k;
              nk = na;
              qf = qa;
              if (k != last) mem.unused_load();
              break;
            }
            mem.store_at(qa, na);
            if (k == last) break;
            k++;
            /* an even level, in nb */
            if (k != last) {
              qa = mem.next_from_even(qb);
              na = mem.load_at(qa);
            } else if (nb.dt == 0.0f && nb.integ == 0.0f) {
              nb.w = (nb.w & ~0xFFu) | get_d_from_intensity(intensity);
            }
            if (integrate_accumulate(nb, intensity, time)) {
              kf = k;
              nk = nb;
              qf = qb;
              if (k != last) mem.unused_load();
              break;
            }
            mem.store_at(qb, nb);
            if (k == last) break;
            k++;
          }
          if (kf) { /* its best event, its fresh child; whatever lay deeper is dropped (:344-366) */
            integrate_fire(nk, intensity, time);
            mem.store_at(qf, nk);
            if (kf + 1u < a.depth) mem.store_fresh(kf + 1u, fresh_node(intensity)); else errbits |= ADDER_DEVERR_DEPTH;
            new_len = kf + 2u;
            kf = 0;
          }
        }
      } else
#endif
      for (uint32_t k = 1; k < k_end; k++) {
        if (defer_deep && k == 2u) { /* level 1 did not fire: the rest of the walk is done by deep_item / deep_finish */
          *deferred = true;
          break;
        }
        Node nxt = nk;
        if (k + 1u < k_end && !defer_deep) nxt = mem.load(k + 1u);
        bool fired = false;
        if (!only_root) {
          if (k == len - 1u && nk.dt == 0.0f && nk.integ == 0.0f) /* :332-335 */
            nk.w = (nk.w & ~0xFFu) | get_d_from_intensity(intensity);
#if ADDER_HOIST_FIRE
          if (integrate_accumulate(nk, intensity, time)) { /* fires: finished below, by all such lanes of the warp together */
            kf = k;
            if (k + 1u < k_end && !defer_deep) mem.unused_load(); /* the level requested ahead is dropped with the rest (:366) */
            break;
          }
#else
          fired = integrate_main(nk, intensity, time);
#endif
        }
        mem.store(k - shift, nk);
        if (k == 1u && shift && NODE_HAS_BEST(nk.w)) {
          disp_has = true;
          disp_d = NODE_BEST_D(nk.w);
          disp_dt = nk.best_dt;
        }
        if (fired) {
          if (k + 1u - shift < a.depth) mem.store_fresh(k + 1u - shift, fresh_node(intensity)); else errbits |= ADDER_DEVERR_DEPTH;
          new_len = k + 2u;
          if (k + 1u < k_end && !defer_deep) mem.unused_load(); /* the level requested ahead is dropped with the rest (:366) */
          break;
        }
        nk = nxt;
      }
#if ADDER_HOIST_FIRE
      if (kf) { /* the node in nk fired at level kf: its best event, its fresh child, the rest of the stack dropped (:344-366) */
        integrate_fire(nk, intensity, time);
        mem.store(kf - shift, nk);
        if (kf == 1u && shift) { /* it is the new root (always has a best event now) */
          disp_has = true;
          disp_d = NODE_BEST_D(nk.w);
          disp_dt = nk.best_dt;
        }
        if (kf + 1u - shift < a.depth) mem.store_fresh(kf + 1u - shift, fresh_node(intensity)); else errbits |= ADDER_DEVERR_DEPTH;
        new_len = kf + 2u;
      }
#endif
      new_len -= shift;
      if (new_len == 0u) new_len = 1u; /* flagged INTERNAL above */
      if (new_len > a.depth) new_len = a.depth;
    }
  }

  h.lf = lf;
  h.y = HDR_PACK(base, cth, cnt, new_len, dtm_reached, popped);
  /* video.rs:713-730.  An unchanged root best event gives the byte already in running_intensities. */
  if (disp_has && a.display && (root_new || a.display == 2u || ADDER_VIEW_OF(a) == 3u)) {
    *disp = frame_value_u8<kPlain>(a, disp_d, f2u(disp_dt), lf);
    return true;
  }
  return false;
}


/* ---- the deferred part of a walk (px_step<true>): levels 2 .. len-1 of an unchanged, unshifted stack -------------------
 * Each level is independent of the others until one fires (integrate_main reads only its own node, event_pixel_tree.rs:
 * 418-479); the FIRST level that fires gets its fresh child and everything deeper is dropped (:344-366).  So the levels
 * can be integrated in any order, by any lane: a level that does not fire is stored integrated (harmless if a shallower
 * level turns out to have fired: it then lies beyond the new length), a level that fires is left untouched and reported;
 * the owner of the pixel finishes the shallowest reported level.  Results are those of the sequential walk. */
constexpr uint32_t kNoFire = 0xFFu;
template <class Mem>
ADDER_HD bool deep_item(Mem& mem, uint32_t k, uint32_t len, float intensity, float time) {
  Node n = mem.load(k);
  uint32_t w = n.w;
  if (k == len - 1u && n.dt == 0.0f && n.integ == 0.0f) w = (w & ~0xFFu) | get_d_from_intensity(intensity); /* :332-335 */
  const float sum = rn_add(n.integ, intensity);
  if (sum >= d_shift_f32(NODE_D(w))) return true; /* fires: left as it is for deep_finish */
  n.integ = sum;
  n.dt = rn_add(n.dt, time);
  n.w = w;
  mem.store(k, n);
  return false;
}
/* kf = the shallowest level deep_item reported, or kNoFire.  Returns the stack's new length. */
template <class Mem>
ADDER_HD uint32_t deep_finish(const PxParams& a, Mem& mem, uint32_t kf, uint32_t len, float intensity, uint32_t& errbits) {
  if (kf == kNoFire) return len;
  Node n = mem.load(kf);
  if (kf == len - 1u && n.dt == 0.0f && n.integ == 0.0f) n.w = (n.w & ~0xFFu) | get_d_from_intensity(intensity);
  integrate_main(n, intensity, a.time); /* fires: deep_item saw the same node and the same sum */
  mem.store(kf, n);
  if (kf + 1u < a.depth) mem.store_fresh(kf + 1u, fresh_node(intensity)); else errbits |= ADDER_DEVERR_DEPTH;
  const uint32_t new_len = kf + 2u;
  return new_len > a.depth ? a.depth : new_len;
}

/*
 * px_frame — what the kernel calls: the common cases of px_step on a short path, everything else through px_step.
 *
 * px_step serves every corner of the reference's state machine in one body (arena shift on a Δt_max pop, Collapse after
 * a pop, synthesised and zero events ...) and pays for that generality in every level of every pixel: the profile of
 * round 1 shows ~95 instructions per node level, most of them address re-computation, register moves and the branches
 * that select among the rare variants (profiles/r01n_*).  The short path covers a pixel-frame in which
 *   - the pixel is not in the "Collapse after a Δt_max pop" regime (popped_dtm && Collapse: :249-265, :360-362), and
 *   - integrating the root does not raise need_to_pop_top (:394-396), i.e. no pop_top_event this frame,
 * which is every pixel-frame of a noise / jitter / static workload except the one frame per delta_t_max in which an
 * unchanged pixel pops its root.  Under these conditions the reference's frame is: [changed: pop_best_events over all
 * nodes, the tail becomes the root (:267-270)]; integrate nodes 0, 1, ... until the first one fires, give it a fresh child
 * and drop the rest (:340-390, FramePerfect :366).  The firing node's arithmetic (the division) is done once, after the
 * walk, by all lanes of a warp together instead of inside the divergent level loop.
 *
 * The short path decides whether it applies BEFORE its first state store; if not, the events it parked are rewound
 * (Sink::rewind) and px_step runs from the untouched state (Mem::reload re-reads the header's nodes: nothing was stored).
 */
template <class Mem, class Sink>
ADDER_HD bool px_frame(const PxParams& a, uint32_t v, PxHeader& h, const Node& n0_in, const Node& n1_in, Mem& mem, Sink& sink,
                       uint32_t& errbits, uint8_t* disp) {
  constexpr bool kPlain = false; /* (ADDER_VIEW_OF / ADDER_ABS_OF: the general output configuration) */
  const uint32_t hy = h.y;
  uint32_t len = HDR_LENGTH(hy);
  const uint32_t popped_in = HDR_POPPED(hy);
  bool lean = !(popped_in && a.collapse);
  if (lean) {
    mem.prefetch_levels(len);
    const float intensity = (float)v;
    const float time = a.time;
    float lf = h.lf;
    uint32_t base = HDR_BASE(hy), cth = HDR_CTHRESH(hy), cnt = HDR_COUNTER(hy);
    uint32_t popped = popped_in;
    const uint32_t d_i = get_d_from_intensity(intensity); /* the d of a fresh node, :332-335 / :502-514 */
    const uint32_t lo = base > cth ? base - cth : 0u;
    const uint32_t hi = base + cth > 255u ? 255u : base + cth;
    const auto mark = sink.mark();
    Node cur = n0_in;
    bool root_new = false;
    if (v < lo || v > hi) { /* video.rs:1338-1358 -> pop_best_events :213-287, the plain arm */
      pop_node(a, sink, lf, cur);
      if (len > 1u) {
        mem.used_preloaded();
        cur = n1_in;
        pop_node(a, sink, lf, cur);
        for (uint32_t k = 2; k < len; k++) {
          cur = mem.load(k);
          pop_node(a, sink, lf, cur);
        }
      }
      len = 1; /* :267-270: the tail (in cur) is the new root */
      popped = 0;
      base = v;
      root_new = true;
    }
    /* ---- integrate (:317-413): walk the levels up to the first node that fires -------------------------------- */
    /* node 0 */
    uint32_t w = cur.w;
    if (len == 1u && cur.dt == 0.0f && cur.integ == 0.0f) w = (w & ~0xFFu) | d_i;
    float sum = rn_add(cur.integ, intensity);
    bool fire = sum >= d_shift_f32(NODE_D(w));
    uint32_t kf = 0;
    bool have_fire = fire;
    uint32_t dtm_reached = 0;
    bool disp_has = false;
    uint32_t disp_d = 0;
    float disp_dt = 0.0f;
    if (!fire) {
      cur.integ = sum;
      cur.dt = rn_add(cur.dt, time);
      cur.w = w;
      dtm_reached = cur.dt >= a.dtm_f ? 1u : 0u;
      if (NODE_D(w) == ADDER_D_MAX || (dtm_reached && !popped)) {
        lean = false; /* pop_top_event this frame */
      } else {
        mem.store(0, cur);
        if (NODE_HAS_BEST(w)) {
          disp_has = true;
          disp_d = NODE_BEST_D(w);
          disp_dt = cur.best_dt;
        }
        if (len > 1u) {
          mem.used_preloaded();
          cur = n1_in;
          for (uint32_t k = 1;;) {
            Node nxt = cur;
            if (k + 1u < len) nxt = mem.load(k + 1u);
            w = cur.w;
            if (k == len - 1u && cur.dt == 0.0f && cur.integ == 0.0f) w = (w & ~0xFFu) | d_i;
            sum = rn_add(cur.integ, intensity);
            if (sum >= d_shift_f32(NODE_D(w))) {
              kf = k;
              have_fire = true;
              if (k + 1u < len) mem.unused_load();
              break;
            }
            cur.integ = sum;
            cur.dt = rn_add(cur.dt, time);
            cur.w = w;
            mem.store(k, cur);
            if (++k == len) break;
            cur = nxt;
          }
        }
      }
    }
    uint32_t new_len = len;
    if (lean && have_fire) {
      /* integrate_main's firing arm (:424-466) for the node in cur / w / sum, once per pixel */
      const uint32_t d_old = NODE_D(w);
      const uint32_t nd = get_d_from_intensity(sum);
      float prop = rn_div(rn_sub(d_shift_f32(nd), cur.integ), intensity);
      if (nd == ADDER_D_ZERO_INTEGRATION || d_old == ADDER_D_ZERO_INTEGRATION || intensity < 1.1920929e-07f) prop = 1.0f;
      cur.best_dt = rn_add(cur.dt, rn_mul(time, prop));
      uint32_t d_after = nd;
      if (nd < ADDER_D_MAX) {
        cur.integ = sum;
        cur.dt = rn_add(cur.dt, time);
        d_after = nd + 1u;
      }
      cur.w = NODE_PACK(d_after, nd, 1);
      if (kf == 0u) {
        dtm_reached = cur.dt >= a.dtm_f ? 1u : 0u;
        if (d_after == ADDER_D_MAX || (dtm_reached && !popped)) lean = false; /* pop_top_event this frame */
        root_new = true;
        disp_has = true;
        disp_d = nd;
        disp_dt = cur.best_dt;
      }
      if (lean) {
        mem.store(kf, cur);
        if (kf + 1u < a.depth) mem.store_fresh(kf + 1u, fresh_node(intensity)); else errbits |= ADDER_DEVERR_DEPTH;
        new_len = kf + 2u;
        if (new_len > a.depth) new_len = a.depth;
      }
    }
    if (lean) {
      if (cth < a.c_max) { /* :402-412 */
        if (cnt >= a.vel_m1) {
          cth = cth + 1u > 255u ? 255u : cth + 1u;
          cnt = 0;
        } else {
          cnt = cnt + a.cnt_inc > 255u ? 255u : cnt + a.cnt_inc;
        }
      }
      h.lf = lf;
      h.y = HDR_PACK(base, cth, cnt, new_len, dtm_reached, popped);
      if (disp_has && a.display && (root_new || a.display == 2u || ADDER_VIEW_OF(a) == 3u)) {
        *disp = frame_value_u8(a, disp_d, f2u(disp_dt), lf);
        return true;
      }
      return false;
    }
    sink.rewind(mark); /* nothing was stored: start over on the general path */
  }
#ifdef ADDER_LEAN_ONLY /* experiment: the cost of the short path alone (pixels that need px_step are flagged, not served) */
  errbits |= ADDER_DEVERR_INTERNAL;
  return false;
#endif
  Node n0 = n0_in, n1 = n1_in;
  mem.reload(n0, n1);
  return px_step(a, v, h, n0, n1, mem, sink, errbits, disp);
}

}  // namespace adder
