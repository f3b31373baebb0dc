/*
 * adder_oracle.c — CPU restatement of the reference's framed→ADΔER per-pixel path.
 * TEST INFRASTRUCTURE ONLY — see adder_oracle.h for the rules and for how parity is pinned.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (f32 results must equal the Rust ones bit for
 * bit: separate IEEE mul/add/div, never an FMA).  Citations are relative to /root/reference/.
 */
#include "adder_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ---- Rust `as` casts (saturating, NaN -> 0, truncation toward zero) --------------------------- */
static inline uint32_t f32_as_u32(float x) {
  if (!(x > 0.0f)) return 0u; /* NaN, negatives, zero */
  if (x >= 4294967296.0f) return 0xFFFFFFFFu;
  return (uint32_t)x;
}
static inline u128 f32_as_u128(float x) {
  if (!(x > 0.0f)) return 0;
  if (isinf(x)) return ~(u128)0; /* every finite f32 is < 2^128 */
  return (u128)x;
}
static inline uint8_t f64_as_u8(double x) {
  if (!(x > 0.0)) return 0;
  if (x >= 255.0) return 255;
  return (uint8_t)x;
}
static inline uint8_t f32_as_u8(float x) {
  if (!(x > 0.0f)) return 0;
  if (x >= 255.0f) return 255;
  return (uint8_t)x;
}
static inline uint8_t sat_add_u8(uint8_t a, uint8_t b) { unsigned s = (unsigned)a + b; return s > 255 ? 255 : (uint8_t)s; }
static inline uint8_t sat_sub_u8(uint8_t a, uint8_t b) { return a > b ? (uint8_t)(a - b) : 0; }

/* ---- D tables: adder-codec-core/src/lib.rs:220-235 -------------------------------------------
 * D_SHIFT[n] = 1<<n for n in 0..=127, and 0 for n = 128 (u128 / f64 / f32 flavours).           */
static inline u128 d_shift_u128(unsigned n) { return n == 128 ? (u128)0 : ((u128)1 << n); }
static inline float d_shift_f32(unsigned n) { return n == 128 ? 0.0f : ldexpf(1.0f, (int)n); }
static inline double d_shift_f64(unsigned n) { return n == 128 ? 0.0 : ldexp(1.0, (int)n); }

static void oracle_panic(const char* what) {
  fprintf(stderr, "adder_oracle: reference would panic: %s\n", what);
  abort();
}

static inline unsigned clz_u128(u128 v) {
  uint64_t hi = (uint64_t)(v >> 64), lo = (uint64_t)v;
  if (hi) return (unsigned)__builtin_clzll(hi);
  if (lo) return 64u + (unsigned)__builtin_clzll(lo);
  return 128u;
}

/* get_d_from_intensity, event_pixel_tree.rs:482-499 */
uint8_t oracle_get_d_from_intensity(float intensity) {
  if (intensity < 1.0f) return ADDER_D_ZERO_INTEGRATION;
  unsigned v = 128u - clz_u128(f32_as_u128(intensity)) - 1u;
  return (uint8_t)(v < ADDER_D_MAX ? v : ADDER_D_MAX);
}

/* PixelNode::new, event_pixel_tree.rs:502-514 */
static inline oracle_node node_new(float start_intensity) {
  oracle_node n;
  n.d = oracle_get_d_from_intensity(start_intensity);
  n.integration = 0.0f;
  n.delta_t = 0.0f;
  n.has_best = 0;
  n.best_d = 0;
  n.best_delta_t = 0.0f;
  n.alt = 0;
  return n;
}

static inline oracle_node* arena(oracle_px* px) { return px->heap ? px->heap : px->inl; }
static inline const oracle_node* arena_c(const oracle_px* px) { return px->heap ? px->heap : px->inl; }

/* SmallVec::push */
static void arena_push(oracle_px* px, oracle_node n) {
  if (px->arena_len == px->arena_cap) {
    uint32_t ncap = px->arena_cap * 2;
    oracle_node* nh = (oracle_node*)malloc(sizeof(oracle_node) * ncap);
    if (!nh) oracle_panic("out of memory");
    memcpy(nh, arena(px), sizeof(oracle_node) * px->arena_len);
    free(px->heap);
    px->heap = nh;
    px->arena_cap = ncap;
  }
  arena(px)[px->arena_len++] = n;
}

/* PixelArena::new, event_pixel_tree.rs:69-87 */
void oracle_px_init(oracle_px* px, float start_intensity, uint16_t x, uint16_t y, uint8_t c) {
  memset(px, 0, sizeof(*px));
  px->x = x;
  px->y = y;
  px->c = c;
  px->length = 1;
  px->time_mode = ADDER_TIME_ABSOLUTE_T; /* TimeMode::default(), lib.rs:72-83 */
  px->last_fired_t = 0.0f;
  px->running_t = 0.0f;
  px->base_val = 0;
  px->need_to_pop_top = 0;
  px->c_thresh = 10;
  px->c_increase_counter = 1;
  px->dtm_reached = 0;
  px->popped_dtm = 0;
  px->heap = NULL;
  px->arena_cap = ORACLE_INLINE_NODES;
  px->arena_len = 0;
  arena_push(px, node_new(start_intensity));
}
void oracle_px_free(oracle_px* px) {
  free(px->heap);
  px->heap = NULL;
}
oracle_px* oracle_px_new(float start_intensity, uint16_t x, uint16_t y, uint8_t c) {
  oracle_px* px = (oracle_px*)malloc(sizeof(oracle_px));
  if (px) oracle_px_init(px, start_intensity, x, y, c);
  return px;
}
void oracle_px_delete(oracle_px* px) {
  if (px) {
    oracle_px_free(px);
    free(px);
  }
}
/* PixelArena::time_mode, :89-93 */
void oracle_px_time_mode(oracle_px* px, int time_mode) {
  if (time_mode >= 0) px->time_mode = (uint8_t)time_mode;
}
const oracle_node* oracle_px_node(const oracle_px* px, uint32_t idx) { return &arena_c(px)[idx]; }
uint32_t oracle_px_length(const oracle_px* px) { return px->length; }
int oracle_px_need_to_pop_top(const oracle_px* px) { return px->need_to_pop_top; }

/* Event32: :15-21.  The f32 time is kept until delta_t_to_absolute_t truncates it. */
typedef struct event32 {
  uint8_t d;
  float delta_t;
} event32;

static inline adder_event_t make_event(const oracle_px* px, uint8_t d, uint32_t t) {
  adder_event_t e;
  e.x = px->x;
  e.y = px->y;
  e.c = px->c;
  e.d = d;
  e.reserved = 0;
  e.t = t;
  return e;
}

/* get_zero_event, :96-111 */
static event32 get_zero_event(oracle_px* px, uint32_t idx, int has_next, float next_intensity) {
  oracle_node* node = &arena(px)[idx];
  event32 ret = {ADDER_D_ZERO_INTEGRATION, node->delta_t};
  node->delta_t = 0.0f;
  if (has_next) node->d = oracle_get_d_from_intensity(next_intensity);
  return ret;
}

/* delta_t_to_absolute_t, :113-137 */
static adder_event_t delta_t_to_absolute_t(oracle_px* px, event32* event, int mode, uint32_t ref_time) {
  if (px->time_mode == ADDER_TIME_ABSOLUTE_T) {
    event->delta_t += px->last_fired_t;
    px->last_fired_t = event->delta_t;
    if (mode == ADDER_MODE_FRAME_PERFECT) {
      uint32_t lf = f32_as_u32(px->last_fired_t);
      if (lf % ref_time == 0) {
        px->last_fired_t = (float)lf;
      } else {
        px->last_fired_t = (float)(uint32_t)(((lf / ref_time) + 1u) * ref_time); /* u32 wrapping mul in release */
      }
    }
  }
  return make_event(px, event->d, f32_as_u32(event->delta_t));
}

/* pop_top_event_recursive, :151-210 */
static event32 pop_top_event_recursive(oracle_px* px, float next_intensity) {
  px->need_to_pop_top = 0;
  oracle_node* root = &arena(px)[0];
  if (!root->has_best) {
    if (root->integration == 0.0f && root->delta_t > 0.0f) {
      return get_zero_event(px, 0, 1, next_intensity);
    }
    /* :164-185 synthesise a best event from the running integration */
    root->has_best = 1;
    if (root->integration < 1.0f) {
      root->best_d = ADDER_D_ZERO_INTEGRATION;
    } else {
      uint32_t iv = f32_as_u32(root->integration); /* to_int_unchecked::<u32> (UB beyond 2^32 in the reference) */
      root->best_d = (uint8_t)(32u - (unsigned)__builtin_clz(iv) - 1u);
    }
    root->best_delta_t = root->delta_t;
    if (px->arena_len > 1) { /* :187-193 */
      arena(px)[1] = node_new(next_intensity);
      px->length = 2;
    } else {
      arena_push(px, node_new(next_intensity));
      px->length += 1;
    }
    return pop_top_event_recursive(px, next_intensity);
  }
  event32 ev = {root->best_d, root->best_delta_t};
  oracle_node* a = arena(px);
  for (uint32_t i = 0; i + 1 < px->length; i++) a[i] = a[i + 1]; /* :201-203 */
  px->length -= 1;
  return ev;
}

/* pop_top_event, :139-148 */
adder_event_t oracle_px_pop_top_event(oracle_px* px, float next_intensity, int mode, uint32_t ref_time) {
  event32 ev = pop_top_event_recursive(px, next_intensity);
  px->popped_dtm = 1;
  return delta_t_to_absolute_t(px, &ev, mode, ref_time);
}

static void evec_push(oracle_evec* v, adder_event_t e) {
  if (v->len == v->cap) {
    size_t ncap = v->cap ? v->cap * 2 : 4;
    adder_event_t* nd = (adder_event_t*)realloc(v->data, ncap * sizeof(adder_event_t));
    if (!nd) oracle_panic("out of memory");
    v->data = nd;
    v->cap = ncap;
  }
  v->data[v->len++] = e;
}

/* pop_best_events, :213-287 */
static void pop_best_events(oracle_px* px, oracle_evec* buffer, int mode, int multi_mode, uint32_t ref_time,
                            float intensity) {
  adder_event_t local_small[40];
  adder_event_t* local = local_small;
  size_t nlocal = 0, lcap = 40;
  adder_event_t* local_heap = NULL;
  for (uint32_t node_idx = 0; node_idx < px->length; node_idx++) {
    oracle_node* node = &arena(px)[node_idx];
    adder_event_t out;
    int have = 0;
    if (!node->has_best) {
      if (node->delta_t > 0.0f && node->integration == 0.0f) {
        event32 e = get_zero_event(px, node_idx, 0, 0.0f);
        out = delta_t_to_absolute_t(px, &e, mode, ref_time);
        have = 1;
      }
    } else {
      event32 e = {node->best_d, node->best_delta_t}; /* `Some(mut event)` is a copy: the node keeps its value */
      out = delta_t_to_absolute_t(px, &e, mode, ref_time);
      have = 1;
    }
    if (have) {
      if (nlocal == lcap) {
        lcap *= 2;
        adder_event_t* nh = (adder_event_t*)malloc(lcap * sizeof(adder_event_t));
        if (!nh) oracle_panic("out of memory");
        memcpy(nh, local, nlocal * sizeof(adder_event_t));
        free(local_heap);
        local_heap = nh;
        local = nh;
      }
      local[nlocal++] = out;
    }
  }

  if (px->popped_dtm && multi_mode == ADDER_MULTI_COLLAPSE && nlocal != 0) { /* :249-265 */
    evec_push(buffer, local[0]);
    px->last_fired_t = px->running_t;
    evec_push(buffer, make_event(px, ADDER_D_EMPTY, f32_as_u32(px->running_t)));
    arena(px)[0] = node_new(intensity);
  } else { /* :266-270 */
    for (size_t i = 0; i < nlocal; i++) evec_push(buffer, local[i]);
    oracle_node* a = arena(px);
    oracle_node tmp = a[0];
    a[0] = a[px->length - 1];
    a[px->length - 1] = tmp;
  }
  free(local_heap);
  px->length = 1;
  px->need_to_pop_top = 0;
  px->dtm_reached = 0;
  px->popped_dtm = 0;
}

int oracle_px_pop_best_events(oracle_px* px, adder_event_t* out, size_t cap, int mode, int multi_mode,
                              uint32_t ref_time, float intensity) {
  oracle_evec v = {NULL, 0, 0};
  pop_best_events(px, &v, mode, multi_mode, ref_time, intensity);
  int n = (int)v.len;
  if (v.len > cap) n = -1; else if (v.len) memcpy(out, v.data, v.len * sizeof(adder_event_t));
  free(v.data);
  return n;
}

/* set_d_for_continuous, :289-312 */
int oracle_px_set_d_for_continuous(oracle_px* px, float next_intensity, uint32_t ref_time, adder_event_t* ev) {
  oracle_node* head = &arena(px)[0];
  if (head->has_best) oracle_panic("set_d_for_continuous: best_event must be None");
  uint8_t next_d = oracle_get_d_from_intensity(next_intensity);
  int produced = 0;
  if (next_d < head->d && head->delta_t > 0.0f) {
    event32 r = {ADDER_D_EMPTY, head->delta_t};
    *ev = delta_t_to_absolute_t(px, &r, ADDER_MODE_CONTINUOUS, ref_time);
    head = &arena(px)[0];
    head->delta_t = 0.0f;
    head->integration = 0.0f;
    produced = 1;
  }
  head->d = next_d;
  return produced;
}

/* integrate_main, :418-479.  Returns 1 when the node fires; (*ni,*nt) = what is left for the children. */
static int integrate_main(oracle_px* px, uint32_t index, float intensity, float time, int mode, float* ni, float* nt) {
  oracle_node* node = &arena(px)[index];
  unsigned d_usize = node->d;
  if (node->integration + intensity >= d_shift_f32(d_usize)) {
    uint8_t new_d = oracle_get_d_from_intensity(node->integration + intensity);
    float prop = (d_shift_f32(new_d) - node->integration) / intensity;
    if (new_d == ADDER_D_ZERO_INTEGRATION || d_usize == ADDER_D_ZERO_INTEGRATION || intensity < 1.1920929e-07f /* f32::EPSILON */) {
      prop = 1.0f;
    }
    node->d = new_d;
    d_usize = new_d;
    node->has_best = 1;
    node->best_d = node->d;
    {
      float scaled = time * prop; /* two roundings: mul, then add (:445) */
      node->best_delta_t = node->delta_t + scaled;
    }
    if (node->d < ADDER_D_MAX) { /* :449-461 */
      node->integration += intensity;
      node->delta_t += time;
      for (;;) {
        d_usize += 1;
        if (d_usize > 128) oracle_panic("D_SHIFT index out of bounds");
        if (d_shift_u128(d_usize) > f32_as_u128(node->integration)) break;
      }
      node->d = (uint8_t)d_usize;
    }
    {
      float used = intensity * prop;
      if (intensity - used >= 0.0f) { /* :463-472 */
        if (mode == ADDER_MODE_FRAME_PERFECT) {
          *ni = 0.0f;
          *nt = 0.0f;
        } else {
          float tused = time * prop;
          *ni = intensity - used;
          *nt = time - tused;
        }
        return 1;
      }
    }
    *ni = 0.0f;
    *nt = 0.0f;
    return 1;
  }
  node->integration += intensity;
  node->delta_t += time;
  return 0;
}

/* integrate, :317-413 */
void oracle_px_integrate(oracle_px* px, float intensity, float time, int mode, uint32_t dtm, uint32_t ref_time,
                         uint8_t c_thresh_max, uint8_t c_increase_velocity, int multi_mode) {
  float start_time = time;
  {
    oracle_node* tail = &arena(px)[px->length - 1];
    if (tail->delta_t == 0.0f && tail->integration == 0.0f) tail->d = oracle_get_d_from_intensity(intensity);
  }
  px->running_t += time;

  uint32_t idx = 0;
  int count = 0;
  for (;;) {
    count += 1;
    float ni = 0.0f, nt = 0.0f;
    int filled = integrate_main(px, idx, intensity, time, mode, &ni, &nt);
    if (filled) { /* :344-355 — the child is seeded with the ORIGINAL intensity of this iteration */
      if (px->arena_len > idx + 1) {
        arena(px)[idx + 1] = node_new(intensity);
      } else {
        arena_push(px, node_new(intensity));
      }
      px->length = idx + 2;
      arena(px)[idx].alt = 1;
      intensity = ni;
      time = nt;
    }
    idx += 1;

    if (px->popped_dtm && multi_mode == ADDER_MULTI_COLLAPSE && idx > 0) break; /* :360-362 */

    if (filled) {
      if (mode == ADDER_MODE_FRAME_PERFECT) break; /* :366 */
      if (time > (float)ref_time) arena(px)[idx].d = oracle_get_d_from_intensity(intensity); /* :371-373 */
      if (intensity == 0.0f) break;
    }
    if (idx >= px->length) break;
    if (count > 30) oracle_panic("Infinite loop detected"); /* :387-389 */
  }
  if (px->length == 0) oracle_panic("length == 0");

  px->dtm_reached = arena(px)[0].delta_t >= (float)dtm; /* :394 */
  px->need_to_pop_top = arena(px)[0].d == ADDER_D_MAX || (px->dtm_reached && !px->popped_dtm);

  if (px->c_thresh < c_thresh_max) { /* :402-412 */
    if (px->c_increase_counter >= (uint8_t)(c_increase_velocity - 1)) {
      px->c_thresh = sat_add_u8(px->c_thresh, 1);
      px->c_increase_counter = 0;
    } else {
      uint8_t inc = (uint8_t)(f32_as_u32(start_time) / ref_time);
      px->c_increase_counter = sat_add_u8(px->c_increase_counter, inc);
    }
  }
}

/* integrate_for_px, video.rs:1317-1380 */
int oracle_integrate_for_px(oracle_px* px, uint8_t* base_val, uint8_t frame_val, float intensity, float time_spanned,
                            oracle_evec* buffer, int pixel_tree_mode, int pixel_multi_mode, uint32_t delta_t_max,
                            uint32_t ref_time, uint8_t c_thresh_max, uint8_t c_increase_velocity) {
  int grew = 0;
  if (px->need_to_pop_top) {
    evec_push(buffer, oracle_px_pop_top_event(px, intensity, pixel_tree_mode, ref_time));
    grew = 1;
  }
  *base_val = px->base_val;
  if (frame_val < sat_sub_u8(*base_val, px->c_thresh) || frame_val > sat_add_u8(*base_val, px->c_thresh)) {
    pop_best_events(px, buffer, pixel_tree_mode, pixel_multi_mode, ref_time, intensity);
    grew = 1;
    px->base_val = frame_val;
    if (pixel_tree_mode == ADDER_MODE_CONTINUOUS) {
      adder_event_t ev;
      if (oracle_px_set_d_for_continuous(px, intensity, ref_time, &ev)) evec_push(buffer, ev);
    }
  }
  oracle_px_integrate(px, intensity, time_spanned, pixel_tree_mode, delta_t_max, ref_time, c_thresh_max,
                      c_increase_velocity, pixel_multi_mode);
  if (px->need_to_pop_top) {
    evec_push(buffer, oracle_px_pop_top_event(px, intensity, pixel_tree_mode, ref_time));
    grew = 1;
  }
  return grew;
}

/* fast-math 0.1 `log2_raw` (crate not vendored in the reference; restated from its published
 * source: split exponent/significand, quadratic  (-1/3 m + 2) m - 2/3  on m in [1,2)).  UNPINNED. */
float oracle_log2_raw(float x) {
  uint32_t bits;
  memcpy(&bits, &x, 4);
  int exponent = (int)((bits >> 23) & 0xFF) - 127;
  uint32_t mbits = (bits & 0x007FFFFFu) | 0x3F800000u;
  float m;
  memcpy(&m, &mbits, 4);
  const float a = -1.0f / 3.0f, b = 2.0f, c = -2.0f / 3.0f;
  float t = a * m;
  t = t + b;
  t = t * m;
  t = t + c;
  return (float)exponent + t;
}

/* event_to_intensity, scale_intensity.rs:262-270 */
static inline double event_to_intensity(uint8_t d, uint32_t t) {
  if (d >= 129) return 0.0;
  if (t == 0) return d_shift_f64(d);
  return d_shift_f64(d) / (double)t;
}

/* <u8 as FrameValue>::get_frame_value, scale_intensity.rs:58-104 (SourceType::U8 arm) */
uint8_t oracle_get_frame_value_u8(uint8_t d, uint32_t t, double tpf, float practical_d_max, uint32_t delta_t_max,
                                  int view_mode, uint32_t sae_running_t, uint32_t sae_last_fired_t) {
  switch (view_mode) {
    case ADDER_VIEW_INTENSITY: {
      double intensity = event_to_intensity(d, t);
      return f64_as_u8(intensity * tpf);
    }
    case ADDER_VIEW_D: {
      float q = (float)d / practical_d_max;
      return f32_as_u8(q * 255.0f);
    }
    case ADDER_VIEW_DELTA_T: {
      float q = (float)t / (float)delta_t_max;
      return f32_as_u8(q * 255.0f);
    }
    case ADDER_VIEW_SAE: {
      uint32_t diff = sae_running_t - sae_last_fired_t; /* u32, wrapping in release */
      float q = (float)diff / (float)delta_t_max;
      return f32_as_u8(q * 255.0f);
    }
    default:
      return 0;
  }
}

/* CRF table, adder-codec-core/src/codec/rate_controller.rs:5-18 */
static const float CRF_TABLE[10][4] = {
    {0.0f, 0.0f, 10.0f, 1E-9f},         {0.0f, 1.0f, 9.0f, 1.0f / 12.0f},  {1.0f, 3.0f, 8.0f, 1.0f / 14.0f},
    {2.0f, 7.0f, 7.0f, 1.0f / 15.0f},   {5.0f, 9.0f, 6.0f, 1.0f / 18.0f},  {6.0f, 10.0f, 5.0f, 1.0f / 20.0f},
    {7.0f, 13.0f, 4.0f, 1.0f / 25.0f},  {8.0f, 16.0f, 3.0f, 1.0f / 30.0f}, {10.0f, 20.0f, 2.0f, 1.0f / 30.0f},
    {15.0f, 25.0f, 1.0f, 1.0f / 30.0f},
};

/* Crf::new, rate_controller.rs:55-70 */
void oracle_crf_parameters(uint8_t crf, uint16_t w, uint16_t h, adder_crf_parameters_t* out) {
  memset(out, 0, sizeof(*out));
  if (crf > 9) crf = 9;
  uint16_t min_res = w < h ? w : h;
  out->c_thresh_baseline = (uint8_t)CRF_TABLE[crf][0];
  out->c_thresh_max = (uint8_t)CRF_TABLE[crf][1];
  out->c_increase_velocity = (uint8_t)CRF_TABLE[crf][2];
  float r = CRF_TABLE[crf][3] * (float)min_res;
  out->feature_c_radius = r >= 65535.0f ? 65535 : (uint16_t)r;
}

/* ---- Video ------------------------------------------------------------------------------------ */
struct oracle_video {
  uint16_t w, h;
  uint8_t c;
  int pixel_tree_mode, pixel_multi_mode, view_mode;
  uint32_t delta_t_max, ref_time, tps, chunk_rows, in_interval_count;
  adder_crf_parameters_t crf;
  oracle_px* px;       /* (H,W,C) */
  uint8_t* running;    /* running_intensities (H,W,C) */
  float* matrix_f32;   /* matrix.mapv(f32::from), video.rs:665 */
  uint32_t n_chunks;
  oracle_evec* chunks; /* big_buffer: one Vec<Event> per chunk */
  uint64_t live_entry, live_exit;
  uint32_t max_live, max_px_events;
  /* feature detection, VideoState.{feature_detection, feature_rate_adjustment, features} video.rs:202-210 */
  int feature_detection, feature_rate_adjustment;
  uint8_t* feature_mask;   /* (H,W): coordinate is in its chunk's HashSet<Coord> (state.features) */
  uint16_t* new_features;  /* [x,y] pairs newly inserted by the last frame */
  size_t n_new_features, cap_new_features;
};

static uint32_t n_chunks_of(uint32_t h, uint32_t chunk_rows) { return (h + chunk_rows - 1) / chunk_rows; }

static void video_realloc_chunks(oracle_video* v) {
  if (v->chunks) {
    for (uint32_t i = 0; i < v->n_chunks; i++) free(v->chunks[i].data);
    free(v->chunks);
  }
  v->n_chunks = n_chunks_of(v->h, v->chunk_rows);
  v->chunks = (oracle_evec*)calloc(v->n_chunks, sizeof(oracle_evec));
}

/* Video::new, video.rs:350-438 with VideoState::default :226-243 and VideoStateParams::default :173-182 */
oracle_video* oracle_video_new(uint16_t w, uint16_t h, uint8_t c, int pixel_tree_mode) {
  if (w == 0 || h == 0 || c == 0) return NULL; /* PlaneSize::new, lib.rs:105-117 */
  oracle_video* v = (oracle_video*)calloc(1, sizeof(oracle_video));
  v->w = w;
  v->h = h;
  v->c = c;
  v->pixel_tree_mode = pixel_tree_mode;
  v->pixel_multi_mode = ADDER_MULTI_COLLAPSE;
  v->view_mode = ADDER_VIEW_INTENSITY;
  v->delta_t_max = 7650;
  v->ref_time = 255;
  v->tps = 7650;
  v->chunk_rows = 1;
  v->in_interval_count = 1;
  oracle_crf_parameters(3, w, h, &v->crf); /* EncoderOptions::default -> Crf::new(None) -> quality 3 */
  size_t n = (size_t)w * h * c;
  v->px = (oracle_px*)malloc(n * sizeof(oracle_px));
  v->running = (uint8_t*)calloc(n, 1);
  v->feature_mask = (uint8_t*)calloc((size_t)w * h, 1);
  v->matrix_f32 = (float*)malloc(n * sizeof(float));
  size_t i = 0;
  for (uint32_t y = 0; y < h; y++)
    for (uint32_t x = 0; x < w; x++)
      for (uint32_t ch = 0; ch < c; ch++)
        oracle_px_init(&v->px[i++], 1.0f, (uint16_t)x, (uint16_t)y, c == 1 ? ADDER_C_NONE : (uint8_t)ch);
  video_realloc_chunks(v);
  return v;
}

void oracle_video_delete(oracle_video* v) {
  if (!v) return;
  size_t n = (size_t)v->w * v->h * v->c;
  for (size_t i = 0; i < n; i++) oracle_px_free(&v->px[i]);
  free(v->px);
  free(v->running);
  free(v->feature_mask);
  free(v->new_features);
  free(v->matrix_f32);
  for (uint32_t i = 0; i < v->n_chunks; i++) free(v->chunks[i].data);
  free(v->chunks);
  free(v);
}

void oracle_video_chunk_rows(oracle_video* v, uint32_t chunk_rows) {
  v->chunk_rows = chunk_rows;
  video_realloc_chunks(v);
}

/* Video::time_parameters, video.rs:493-537 */
int oracle_video_time_parameters(oracle_video* v, uint32_t tps, uint32_t ref_time, uint32_t dtm, int time_mode) {
  size_t n = (size_t)v->w * v->h * v->c;
  for (size_t i = 0; i < n; i++) oracle_px_time_mode(&v->px[i], time_mode);
  /* `x > f32::MAX as u32` can never hold for a u32 (f32::MAX as u32 saturates to u32::MAX) */
  if (dtm < ref_time) return 0;
  v->delta_t_max = dtm;
  v->ref_time = ref_time;
  v->tps = tps;
  return 1;
}

/* Video::write_out side effects, video.rs:546-636 */
void oracle_video_write_out(oracle_video* v, int time_mode, int pixel_multi_mode) {
  v->pixel_multi_mode = pixel_multi_mode < 0 ? ADDER_MULTI_COLLAPSE : pixel_multi_mode;
  size_t n = (size_t)v->w * v->h * v->c;
  for (size_t i = 0; i < n; i++) oracle_px_time_mode(&v->px[i], time_mode);
}

static void video_reset_c(oracle_video* v, uint8_t c_base) {
  size_t n = (size_t)v->w * v->h * v->c;
  for (size_t i = 0; i < n; i++) {
    v->px[i].c_thresh = c_base;
    v->px[i].c_increase_counter = 0;
  }
}

void oracle_video_update_crf(oracle_video* v, uint8_t crf) {
  oracle_crf_parameters(crf, v->w, v->h, &v->crf);
  video_reset_c(v, v->crf.c_thresh_baseline);
}

void oracle_video_update_quality_manual(oracle_video* v, uint8_t c_base, uint8_t c_max, uint32_t dtm_mult,
                                        uint8_t velocity, float radius) {
  v->crf.c_thresh_baseline = c_base;
  v->crf.c_thresh_max = c_max;
  v->crf.c_increase_velocity = velocity;
  v->crf.feature_c_radius = radius >= 65535.0f ? 65535 : (radius > 0.0f ? (uint16_t)radius : 0);
  v->delta_t_max = dtm_mult * v->ref_time;
  video_reset_c(v, c_base);
}

/* Video::update_encoder_options (video.rs:1289-1291) / write_out's encoder_options (:553, :634) */
void oracle_video_set_crf_parameters(oracle_video* v, const adder_crf_parameters_t* p) { v->crf = *p; }

void oracle_video_update_delta_t_max(oracle_video* v, uint32_t dtm) {
  v->delta_t_max = v->ref_time > dtm ? v->ref_time : dtm;
}

void oracle_video_c_thresh_pos(oracle_video* v, uint8_t c) {
  size_t n = (size_t)v->w * v->h * v->c;
  for (size_t i = 0; i < n; i++) v->px[i].c_thresh = c;
  v->crf.c_thresh_baseline = c;
}

void oracle_video_set_c_thresh_rect(oracle_video* v, uint16_t x0, uint16_t y0, uint16_t x1, uint16_t y1, uint8_t value) {
  for (uint32_t y = y0; y <= y1 && y < v->h; y++)
    for (uint32_t x = x0; x <= x1 && x < v->w; x++)
      for (uint32_t ch = 0; ch < v->c; ch++) v->px[((size_t)y * v->w + x) * v->c + ch].c_thresh = value;
}

void oracle_video_set_view_mode(oracle_video* v, int view_mode) { v->view_mode = view_mode; }
void oracle_video_set_in_interval_count(oracle_video* v, uint32_t n) { v->in_interval_count = n; }
uint32_t oracle_video_in_interval_count(const oracle_video* v) { return v->in_interval_count; }
uint32_t oracle_video_n_chunks(const oracle_video* v) { return v->n_chunks; }

/* set_initial_d, video.rs:780-801 */
static void set_initial_d(oracle_video* v, const uint8_t* frame) {
  size_t n = (size_t)v->w * v->h * v->c;
  for (size_t i = 0; i < n; i++) {
    uint8_t fv = frame[i];
    uint8_t d_start = fv == 0 ? ADDER_D_ZERO_INTEGRATION : (uint8_t)floorf(log2f((float)fv));
    arena(&v->px[i])[0].d = d_start;
    v->px[i].base_val = fv;
  }
}

/* ---- feature detection: the one cross-pixel step of integrate_matrix ------------------------- */

/* is_feature, adder-codec-rs/src/utils/cv.rs:22-212: asynchronous FAST 9_16 on channel 0 of the image */
int oracle_is_feature(const uint8_t* img, uint16_t w, uint16_t h, uint8_t channels, uint16_t cx, uint16_t cy, uint8_t cc) {
  static const int CIRCLE3[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                                     {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}}; /* :26-31 */
  const int INTENSITY_THRESHOLD = 30, STREAK_SIZE = 9;                                                          /* :22, :33 */
  /* Coord::is_border(w, h, 3), lib.rs:353-358; c must be 0 (None counts as 0: c_usize) */
  if (cx < 3 || (int)cx >= (int)w - 3 || cy < 3 || (int)cy >= (int)h - 3 || cc != 0) return 0; /* :69-71 */
  const long c = channels, width = (long)w * c, y = cy, x = cx;
  const int candidate = img[y * width + x * c];
#define PX(k) ((int)img[(y + CIRCLE3[(k)][1]) * width + (x + CIRCLE3[(k)][0]) * c])
#define TAB(v) ((v) - candidate < -INTENSITY_THRESHOLD ? 1 : ((v) - candidate > INTENSITY_THRESHOLD ? 2 : 0)) /* THRESHOLD_TABLE :35-50 */
  int d = TAB(PX(0)) | TAB(PX(8)); /* :92-96 */
  if (d == 0) return 0;
  d &= TAB(PX(2)) | TAB(PX(10));
  d &= TAB(PX(4)) | TAB(PX(12));
  d &= TAB(PX(6)) | TAB(PX(14));
  if (d == 0) return 0; /* :117-119 */
  d &= TAB(PX(1)) | TAB(PX(9));
  d &= TAB(PX(3)) | TAB(PX(11));
  d &= TAB(PX(5)) | TAB(PX(13));
  d &= TAB(PX(7)) | TAB(PX(15));
  if (d & 1) { /* dark streak :142-172 */
    const int vt = candidate - INTENSITY_THRESHOLD;
    int count = 0;
    for (int k = 0; k < 16; k++) {
      if (PX(k) < vt) {
        if (++count == STREAK_SIZE) return 1;
      } else {
        count = 0;
      }
    }
    for (int k = 16; k < 25; k++) {
      if (PX(k - 16) < vt) {
        if (++count == STREAK_SIZE) return 1;
      } else {
        count = 0;
        if (k == 17) return 0; /* :167-169: returns for the whole function, the bright test is not reached */
      }
    }
  }
  if (d & 2) { /* bright streak :174-205 */
    const int vt = candidate + INTENSITY_THRESHOLD;
    int count = 0;
    for (int k = 0; k < 16; k++) {
      if (PX(k) > vt) {
        if (++count == STREAK_SIZE) return 1;
      } else {
        count = 0;
      }
    }
    for (int k = 16; k < 25; k++) {
      if (PX(k - 16) > vt) {
        if (++count == STREAK_SIZE) return 1;
      } else {
        count = 0;
        if (k == 17) return 0;
      }
    }
  }
#undef PX
#undef TAB
  return 0;
}

static int coord_eq(const adder_event_t* a, const adder_event_t* b) { return a->x == b->x && a->y == b->y && a->c == b->c; }

/* Video::handle_features, video.rs:883-1113, without logging, drawing and clustering (display only) */
static void handle_features(oracle_video* v) {
  v->n_new_features = 0;
  if (!v->feature_detection) return; /* :885-887 */
  for (uint32_t ci = 0; ci < v->n_chunks; ci++) {
    const oracle_evec* ev = &v->chunks[ci];
    for (size_t j = 0; j < ev->len; j++) { /* events.iter().circular_tuple_windows() :898 */
      const adder_event_t* e1 = &ev->data[j];
      const adder_event_t* e2 = &ev->data[j + 1 < ev->len ? j + 1 : 0];
      if ((e1->c == ADDER_C_NONE || e1->c == 0) && !coord_eq(e1, e2) && e1->d != ADDER_D_EMPTY) { /* :899-903 */
        const size_t p = (size_t)e1->y * v->w + e1->x;
        if (oracle_is_feature(v->running, v->w, v->h, v->c, e1->x, e1->y, 0)) {
          if (!v->feature_mask[p]) { /* feature_set.insert(..) returned true :908-910 */
            v->feature_mask[p] = 1;
            if (v->n_new_features == v->cap_new_features) {
              v->cap_new_features = v->cap_new_features ? v->cap_new_features * 2 : 64;
              v->new_features = (uint16_t*)realloc(v->new_features, v->cap_new_features * 2 * sizeof(uint16_t));
            }
            v->new_features[2 * v->n_new_features] = e1->x;
            v->new_features[2 * v->n_new_features + 1] = e1->y;
            v->n_new_features++;
          }
        } else {
          v->feature_mask[p] = 0; /* feature_set.remove :912 */
        }
      }
    }
  }
  /* :1077-1104: c_thresh of every pixel within the radius of a new feature */
  if (v->feature_rate_adjustment && v->crf.feature_c_radius > 0) {
    const int radius = (int)v->crf.feature_c_radius;
    const uint8_t value = v->crf.c_thresh_baseline < 2 ? v->crf.c_thresh_baseline : 2;
    for (size_t k = 0; k < v->n_new_features; k++) {
      const int fx = v->new_features[2 * k], fy = v->new_features[2 * k + 1];
      const int r0 = fy - radius > 0 ? fy - radius : 0, r1 = fy + radius < (int)v->h - 1 ? fy + radius : (int)v->h - 1;
      const int c0 = fx - radius > 0 ? fx - radius : 0, c1 = fx + radius < (int)v->w - 1 ? fx + radius : (int)v->w - 1;
      for (int row = r0; row <= r1; row++)
        for (int col = c0; col <= c1; col++)
          for (int ch = 0; ch < v->c; ch++) v->px[((size_t)row * v->w + col) * v->c + ch].c_thresh = value;
    }
  }
}

/* Video::update_detect_features, video.rs:825-837 (show_features and feature_cluster only affect drawing) */
void oracle_video_update_detect_features(oracle_video* v, int detect_features, int feature_rate_adjustment) {
  v->feature_detection = detect_features;
  v->feature_rate_adjustment = feature_rate_adjustment;
}
size_t oracle_video_new_features(const oracle_video* v, uint16_t* xy_out, size_t cap) {
  const size_t n = v->n_new_features < cap ? v->n_new_features : cap;
  if (n) memcpy(xy_out, v->new_features, n * 2 * sizeof(uint16_t));
  return v->n_new_features;
}
const uint8_t* oracle_video_feature_mask(const oracle_video* v) { return v->feature_mask; }

/* Video::integrate_matrix, video.rs:651-778 (up to and including the parallel section) */
size_t oracle_video_integrate_matrix(oracle_video* v, const uint8_t* frame, float time_spanned, int n_threads) {
  if (v->in_interval_count == 0) set_initial_d(v, frame);
  const adder_crf_parameters_t parameters = v->crf;
  v->in_interval_count += 1;

  const size_t n = (size_t)v->w * v->h * v->c;
  for (size_t i = 0; i < n; i++) v->matrix_f32[i] = (float)frame[i]; /* matrix.mapv(f32::from) */

  /* (u32 / u32) as f32, then log2_raw: only feeds FramedViewMode::D */
  const float practical_d_max = oracle_log2_raw(255.0f * (float)(v->delta_t_max / v->ref_time));
  const double tpf = (double)v->ref_time;
  const size_t chunk_px = (size_t)v->chunk_rows * v->w * v->c;
  uint64_t live_entry = 0, live_exit = 0;
  uint32_t max_live = v->max_live, max_px_events = v->max_px_events;

#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads > 0 ? n_threads : 1) \
    reduction(+ : live_entry, live_exit) reduction(max : max_live, max_px_events) if (n_threads > 1)
#endif
  for (uint32_t ci = 0; ci < v->n_chunks; ci++) {
    oracle_evec* buffer = &v->chunks[ci];
    free(buffer->data); /* a fresh Vec::with_capacity(10) per chunk per frame, video.rs:693 */
    buffer->data = (adder_event_t*)malloc(10 * sizeof(adder_event_t));
    buffer->cap = 10;
    buffer->len = 0;
    uint8_t base_val = 0;
    size_t begin = (size_t)ci * chunk_px;
    size_t end = begin + chunk_px;
    if (end > n) end = n;
    for (size_t i = begin; i < end; i++) {
      oracle_px* px = &v->px[i];
      float input = v->matrix_f32[i];
      size_t before = buffer->len;
      live_entry += px->length;
      oracle_integrate_for_px(px, &base_val, (uint8_t)input, input, time_spanned, buffer, v->pixel_tree_mode,
                              v->pixel_multi_mode, v->delta_t_max, v->ref_time, parameters.c_thresh_max,
                              parameters.c_increase_velocity);
      live_exit += px->length;
      if (px->length > max_live) max_live = px->length;
      if (buffer->len - before > max_px_events) max_px_events = (uint32_t)(buffer->len - before);
      const oracle_node* root = &arena_c(px)[0];
      if (root->has_best) { /* video.rs:713-730 */
        v->running[i] = oracle_get_frame_value_u8(root->best_d, f32_as_u32(root->best_delta_t), tpf, practical_d_max,
                                                 v->delta_t_max, v->view_mode, f32_as_u32(px->running_t),
                                                 f32_as_u32(px->last_fired_t));
      }
    }
  }
  v->live_entry = live_entry;
  v->live_exit = live_exit;
  v->max_live = max_live;
  v->max_px_events = max_px_events;
  handle_features(v); /* video.rs:744 (after the parallel section and the encoder loop) */
  size_t total = 0;
  for (uint32_t ci = 0; ci < v->n_chunks; ci++) total += v->chunks[ci].len;
  return total;
}

void oracle_video_chunk_counts(const oracle_video* v, uint32_t* counts) {
  for (uint32_t ci = 0; ci < v->n_chunks; ci++) counts[ci] = (uint32_t)v->chunks[ci].len;
}

size_t oracle_video_copy_events(const oracle_video* v, adder_event_t* out, size_t cap) {
  size_t total = 0;
  for (uint32_t ci = 0; ci < v->n_chunks; ci++) {
    size_t len = v->chunks[ci].len;
    if (total + len <= cap && len) memcpy(out + total, v->chunks[ci].data, len * sizeof(adder_event_t));
    total += len;
  }
  return total;
}

const uint8_t* oracle_video_running_intensities(const oracle_video* v) { return v->running; }

void oracle_video_stats(const oracle_video* v, uint64_t* live_nodes_entry, uint64_t* live_nodes_exit,
                        uint32_t* max_live_nodes, uint32_t* max_px_events) {
  if (live_nodes_entry) *live_nodes_entry = v->live_entry;
  if (live_nodes_exit) *live_nodes_exit = v->live_exit;
  if (max_live_nodes) *max_live_nodes = v->max_live;
  if (max_px_events) *max_px_events = v->max_px_events;
}

const oracle_px* oracle_video_px(const oracle_video* v, size_t index) { return &v->px[index]; }

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* The host cores this process may run on, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its
 * workers, which made the round-1 reference arm run on one core at N >= 2). */
int oracle_num_procs(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

/* ---- bench / test input --------------------------------------------------------------------------
 * The synthetic frames of SURVEY.md §8(d), byte for byte what tests/synth.py (numpy) and the product's device generator
 * produce; here only so that the CPU arm of bench.py can make 8K frames at memory speed.
 *   h(seed, f, i) = splitmix64(seed ^ (f << 40) ^ i) & 0xFF,  i = flat raster index ((y*W + x)*C + c) of the whole plane */
static inline uint64_t oracle_splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint32_t oracle_synth_hash(uint64_t seed, uint32_t f, uint64_t i) {
  return (uint32_t)(oracle_splitmix64(seed ^ ((uint64_t)f << 40) ^ i) & 0xFFu);
}
/* kind 0 gradient, 1 noise, 2 base +-10 jitter, 3 static base with one-frame blips; n pixels-channels starting at flat
 * index i0 of a plane `w` wide with `c` channels */
void oracle_synth_frame(int kind, uint64_t seed, uint32_t f, uint32_t w, uint32_t c, uint64_t i0, uint64_t n, uint8_t* out,
                        int n_threads) {
  (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
#endif
  for (int64_t k = 0; k < (int64_t)n; k++) {
    const uint64_t i = i0 + (uint64_t)k;
    uint32_t v;
    switch (kind) {
      case 0: {
        const uint64_t p = i / c, x = p % w, y = p / w;
        v = (uint32_t)((x + 2u * y + 3u * (uint64_t)f) & 255u);
        break;
      }
      case 1: v = oracle_synth_hash(seed, f, i); break;
      case 2: {
        const int t = (int)oracle_synth_hash(seed ^ 1ull, 0, i) + (int)(oracle_synth_hash(seed ^ 2ull, f, i) % 21u) - 10;
        v = (uint32_t)(t < 0 ? 0 : (t > 255 ? 255 : t));
        break;
      }
      default:
        v = oracle_synth_hash(seed ^ 3ull, f, i) < 2u ? oracle_synth_hash(seed ^ 4ull, f, i) : oracle_synth_hash(seed ^ 1ull, 0, i);
        break;
    }
    out[k] = (uint8_t)v;
  }
}

/* ---- raw .adder wire format ------------------------------------------------------------------- */

static uint8_t* put_u16(uint8_t* p, uint16_t v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; return p + 2; }
static uint8_t* put_u32(uint8_t* p, uint32_t v) {
  p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
  return p + 4;
}

size_t oracle_raw_header(uint8_t* out, int compressed, uint8_t version, uint16_t width, uint16_t height, uint32_t tps,
                         uint32_t ref_interval, uint32_t delta_t_max, uint8_t channels, uint32_t source_camera,
                         uint32_t time_mode, uint32_t adu_interval) {
  static const uint8_t magic_raw[5] = {97, 100, 100, 101, 114}; /* header.rs:5 'adder' */
  static const uint8_t magic_cmp[5] = {97, 100, 100, 101, 99};  /* header.rs:6 'addec' */
  if (version > 3) return 0; /* encoder.rs:228 BadFile */
  uint8_t* p = out;
  memcpy(p, compressed ? magic_cmp : magic_raw, 5); /* [u8;5]: no length prefix */
  p += 5;
  *p++ = version;
  *p++ = 98; /* 'b' header.rs:68 */
  p = put_u16(p, width);
  p = put_u16(p, height);
  p = put_u32(p, tps);
  p = put_u32(p, ref_interval);
  p = put_u32(p, delta_t_max);
  *p++ = channels == 1 ? 9 : 11; /* header.rs:77-81 */
  *p++ = channels;
  /* extension V0 is an empty struct: no bytes (encoder.rs:193-197) */
  if (version >= 1) p = put_u32(p, source_camera); /* enum as u32 variant index, :199-207 */
  if (version >= 2) p = put_u32(p, time_mode);     /* :209-217 */
  if (version >= 3) p = put_u32(p, adu_interval);  /* :219-227 */
  return (size_t)(p - out);
}

size_t oracle_raw_encode(const adder_event_t* ev, size_t n, uint8_t channels, uint8_t* out) {
  uint8_t* p = out;
  for (size_t i = 0; i < n; i++) {
    p = put_u16(p, ev[i].x);
    p = put_u16(p, ev[i].y);
    if (channels != 1) { /* Option<u8>: tag 1 then the value (raw/stream.rs:115-117) */
      *p++ = 1;
      *p++ = ev[i].c;
    }
    *p++ = ev[i].d;
    p = put_u32(p, ev[i].t);
  }
  return (size_t)(p - out);
}

size_t oracle_raw_eof(uint8_t* out) {
  static const uint8_t eof[11] = {0xFF, 0xFF, 0xFF, 0xFF, 1, 0, 0, 0, 0, 0, 0}; /* x = y = EOF_PX_ADDRESS, c = Some(0), d = 0, t = 0 */
  memcpy(out, eof, 11);
  return 11;
}

/* ---- handle_color (utils/cv.rs:215-232) ------------------------------------------------------- */
void oracle_handle_color(const uint8_t* rgb, size_t n_px, uint8_t* out) {
  for (size_t i = 0; i < n_px; i++) {
    volatile double a = (double)rgb[3 * i] * 0.114;
    volatile double b = (double)rgb[3 * i + 1] * 0.587;
    volatile double c = (double)rgb[3 * i + 2] * 0.299;
    volatile double s = a + b;
    s = s + c;
    out[i] = s >= 255.0 ? 255 : (s > 0.0 ? (uint8_t)s : 0); /* `as u8`: truncating, saturating */
  }
}
