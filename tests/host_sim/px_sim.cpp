/*
 * px_sim.cpp — TEST INFRASTRUCTURE: runs the product's per-pixel state machine
 * (adder_codec_rs_b200/csrc/px_machine.cuh, the very header the CUDA kernel compiles) serially on
 * the host, over the same SoA state layout, so its restructuring of the reference algorithm
 * (streaming node pass, shift-on-pop, lazy tail fix) can be checked against the oracle in this
 * GPU-less container.  Nothing in the product links or calls this.
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#define ADDER_HOST_SIM 1
#include "../../include/adder_b200.h"
#include "../../adder_codec_rs_b200/csrc/px_machine.cuh"
#include "../../adder_codec_rs_b200/csrc/px_offset.cuh"
#include "../../adder_codec_rs_b200/csrc/gray_math.h"
#include "../../adder_codec_rs_b200/csrc/raw_pack.h"

namespace adder {
int g_fast_div_ulps = 0;
}
static int g_entry = 0; /* 4: px_offset on offset-form state where offset_form_eligible() allows it (else px_step); 0: px_step (what the kernel calls by default), 1: px_frame (short path + fallback), 2: px_step<true> + deferred deep walk, 3: px_step<false, true> where it applies */
namespace {
struct HostNodes {
  adder::Node* p;
  size_t stride;
  adder::Node load(uint32_t k) const { return p[(size_t)k * stride]; }
  void store(uint32_t k, const adder::Node& n) const { p[(size_t)k * stride] = n; }
  void store_fresh(uint32_t k, const adder::Node& n) const { p[(size_t)k * stride] = n; }
  void used_preloaded() const {}
  void set_popped() const {}
  void prefetch_levels(uint32_t) const {}
  void unused_load() const {}
  adder::Node* cursor_level1() const { return p + stride; }
  adder::Node* next_from_odd(adder::Node* q) const { return q + stride; }
  adder::Node* next_from_even(adder::Node* q) const { return q + stride; }
  adder::Node load_at(const adder::Node* q) const { return *q; }
  void store_at(adder::Node* q, const adder::Node& n) const { *q = n; }
  void reload(adder::Node& n0, adder::Node& n1) const { n0 = p[0]; n1 = p[stride]; }
};
/* offset-form records of one pixel (px_offset.cuh): record k of pixel i at recs[k * P + i] */
struct HostRecs {
  adder::OffRec* p;
  size_t stride;
  uint32_t n_loads = 0, n_stores = 0;
  adder::OffRec load_rec(uint32_t k) { n_loads++; return p[(size_t)k * stride]; }
  void store_rec(uint32_t k, const adder::OffRec& r) { n_stores++; p[(size_t)k * stride] = r; }
};
struct VecSink {
  std::vector<adder_event_t>* out;
  uint16_t x, y;
  uint8_t c;
  void push(uint32_t d, uint32_t t) {
    adder_event_t e;
    e.x = x; e.y = y; e.c = c; e.d = (uint8_t)d; e.reserved = 0; e.t = t;
    out->push_back(e);
  }
  size_t mark() const { return out->size(); }
  void rewind(size_t m) { out->resize(m); }
};
}  // namespace

struct sim_video {
  uint32_t w, h, c, depth;
  size_t P;
  std::vector<adder::PxHeader> hdr;
  std::vector<adder::Node> nodes;
  std::vector<uint8_t> running;
  /* offset form (g_entry == 4): root nodes stay in nodes[0 * P + i] */
  std::vector<adder::OffRec> recs;
  std::vector<adder::OffTop> meta; /* the top level, second half of record 0 */
  int form = -1; /* -1 undecided, 0 eager, 1 offset */
  uint64_t rec_loads = 0, rec_stores = 0, px_frames = 0;
  std::vector<adder_event_t> events;
  float running_t;
  uint32_t err;
  int force_display;
};

extern "C" {

sim_video* sim_new(uint32_t w, uint32_t h, uint32_t c, uint32_t depth) {
  sim_video* v = new sim_video();
  v->w = w; v->h = h; v->c = c; v->depth = depth;
  v->P = (size_t)w * h * c;
  v->hdr.assign(v->P, adder::PxHeader{0.0f, HDR_PACK(0, 10, 1, 1, 0, 0)});
  v->nodes.assign(v->P * depth, adder::Node{0.0f, 0.0f, 0.0f, NODE_PACK(0, 0, 0)});
  v->running.assign(v->P, 0);
  v->running_t = 0.0f;
  v->err = 0;
  v->force_display = 1;
  return v;
}
void sim_delete(sim_video* v) { delete v; }

void sim_reset_c(sim_video* v, uint32_t cth, int reset_counter) {
  for (auto& h : v->hdr) {
    h.y = (h.y & ~0xFF00u) | (cth << 8);
    if (reset_counter) h.y &= ~0xFF0000u;
  }
}

size_t sim_integrate(sim_video* v, const uint8_t* frame, float time, uint32_t ref, uint32_t dtm, uint32_t c_max,
                     uint32_t vel, int collapse, int abs_time, int view_mode, float practical_d_max) {
  adder::PxParams p{};
  p.time = time;
  p.running_t_prev = v->running_t;
  v->running_t = v->running_t + time;
  p.running_t = v->running_t;
  p.dtm_f = (float)dtm;
  p.ref = ref;
  p.dtm = dtm;
  p.c_max = c_max;
  p.vel_m1 = (uint8_t)(vel - 1);
  p.cnt_inc = (uint8_t)(adder::f2u(time) / ref);
  p.collapse = collapse;
  p.abs_time = abs_time;
  p.view_mode = view_mode;
  p.display = v->force_display ? 2 : 1;
  v->force_display = 0;
  p.depth = v->depth;
  p.tpf = (double)ref;
  p.tpf_f = (float)ref;
  p.ref_magic = adder::ref_magic_of(ref);
  p.practical_d_max = practical_d_max;
  uint8_t lut[257];
  adder::build_exact_lut(ref, lut);
  p.exact_lut = lut;
  v->events.clear();
  if (g_entry == 4) {
    const bool ok = adder::offset_form_eligible(collapse != 0, time, dtm);
    if (v->form < 0) {
      v->form = ok ? 1 : 0;
      if (ok) {
        v->recs.assign(v->P * v->depth, adder::OffRec{0, 0, 0.0f, 0, 0, 0});
        v->meta.assign(v->P, adder::OffTop{0, 0, 0.0f, 0});
      }
    } else if (v->form == 1 && !ok) {
      v->err |= 0x80000000u; /* the test cases never change eligibility mid-stream (the library converts the state) */
    }
  }
  for (size_t i = 0; i < v->P; i++) {
    if (g_entry == 4 && v->form == 1) {
      HostRecs mem{v->recs.data() + i, v->P};
      const uint32_t ch = (uint32_t)(i % v->c), x = (uint32_t)((i / v->c) % v->w), y = (uint32_t)(i / ((size_t)v->c * v->w));
      VecSink sink{&v->events, (uint16_t)x, (uint16_t)y, (uint8_t)(v->c == 1 ? ADDER_C_NONE : ch)};
      uint8_t disp = 0;
      bool show;
      if (abs_time && view_mode == 0)
        show = adder::px_offset<true>(p, frame[i], v->hdr[i], v->nodes[i], v->meta[i], mem, sink, v->err, &disp);
      else
        show = adder::px_offset<false>(p, frame[i], v->hdr[i], v->nodes[i], v->meta[i], mem, sink, v->err, &disp);
      if (show) v->running[i] = disp;
      v->rec_loads += mem.n_loads;
      v->rec_stores += mem.n_stores;
      v->px_frames++;
      continue;
    }
    HostNodes mem{v->nodes.data() + i, v->P};
    const uint32_t ch = (uint32_t)(i % v->c), x = (uint32_t)((i / v->c) % v->w), y = (uint32_t)(i / ((size_t)v->c * v->w));
    VecSink sink{&v->events, (uint16_t)x, (uint16_t)y, (uint8_t)(v->c == 1 ? ADDER_C_NONE : ch)};
    uint8_t disp = 0;
    /* like the kernel: level 1 is fetched before the length is known */
    const adder::Node n1 = v->depth > 1 ? mem.load(1) : mem.load(0);
    bool show;
    if (g_entry == 2) { /* the kernel's long-integration variant: deferred deep walk, here visited deepest level first */
      const uint32_t len_in = HDR_LENGTH(v->hdr[i].y);
      bool deferred = false;
      show = adder::px_step<true>(p, frame[i], v->hdr[i], mem.load(0), n1, mem, sink, v->err, &disp, &deferred);
      if (deferred) {
        uint32_t kf = adder::kNoFire;
        for (uint32_t k = len_in; k-- > 2u;)
          if (adder::deep_item(mem, k, len_in, (float)frame[i], p.time) && k < kf) kf = k;
        const uint32_t nl = adder::deep_finish(p, mem, kf, len_in, (float)frame[i], v->err);
        v->hdr[i].y = (v->hdr[i].y & ~(0x1Fu << 24)) | (nl << 24);
      }
    } else if (g_entry == 3 && abs_time && view_mode == 0) { /* the kernel's kPlain instantiation (AbsoluteT + Intensity view as constants) */
      show = adder::px_step<false, true>(p, frame[i], v->hdr[i], mem.load(0), n1, mem, sink, v->err, &disp);
    } else {
      show = g_entry == 1 ? adder::px_frame(p, frame[i], v->hdr[i], mem.load(0), n1, mem, sink, v->err, &disp)
                          : adder::px_step(p, frame[i], v->hdr[i], mem.load(0), n1, mem, sink, v->err, &disp);
    }
    if (show) v->running[i] = disp;
  }
  return v->events.size();
}
/* unit hooks for the two exact-by-construction shortcuts of px_machine.cuh */
void sim_set_fast_div_ulps(int n) { adder::g_fast_div_ulps = n; }
void sim_set_entry(int e) { g_entry = e; }
uint32_t sim_div_ref(uint32_t u, uint32_t ref) {
  adder::PxParams p{};
  p.ref = ref;
  p.ref_magic = adder::ref_magic_of(ref);
  return adder::div_ref(p, u);
}
uint32_t sim_frame_value_intensity(uint32_t d, uint32_t t, uint32_t ref) {
  adder::PxParams p{};
  p.view_mode = 0;
  p.tpf = (double)ref;
  p.tpf_f = (float)ref;
  p.ref = ref;
  uint8_t lut[257];
  adder::build_exact_lut(ref, lut);
  p.exact_lut = lut;
  return adder::frame_value_u8(p, d, t, 0.0f);
}
/* gray_math.h: the fixed-point shortcut against the reference's f64 expression over all 2^24 colours;
 * returns the number of disagreements and how many colours took the shortcut */
uint64_t sim_gray_check(uint64_t* n_fast) {
  uint8_t diag[256];
  adder::build_gray_diag(diag);
  uint64_t bad = 0, fast = 0;
  for (uint32_t c0 = 0; c0 < 256u; c0++)
    for (uint32_t c1 = 0; c1 < 256u; c1++)
      for (uint32_t c2 = 0; c2 < 256u; c2++) {
        const uint32_t S = c0 * 1912603u + c1 * 9848226u + c2 * 5016388u, fr = S & 0xFFFFFFu;
        fast += fr >= 257u && fr <= 0xFFFFFEu;
        bad += adder::gray_of(c0, c1, c2, diag) != adder::gray_exact_f64(c0, c1, c2);
      }
  *n_fast = fast;
  return bad;
}
uint32_t sim_gray_of(uint32_t c0, uint32_t c1, uint32_t c2) {
  uint8_t diag[256];
  adder::build_gray_diag(diag);
  return adder::gray_of(c0, c1, c2, diag);
}
/* raw_pack.h: n records (12-byte, n a multiple of 4) -> wire bytes, four at a time like a thread of raw_encode_kernel */
void sim_raw_pack(const uint32_t* words, size_t n, uint32_t esize, uint8_t* out) {
  for (size_t i = 0; i + 4 <= n; i += 4) {
    uint32_t o[11] = {0};
    if (esize == 11u) adder::raw_pack4<11u>(words + 3 * i, o); else adder::raw_pack4<9u>(words + 3 * i, o);
    memcpy(out + i * esize, o, 4 * esize);
  }
}
void sim_force_display(sim_video* v) { v->force_display = 1; }
const adder_event_t* sim_events(const sim_video* v) { return v->events.data(); }
const uint8_t* sim_running(const sim_video* v) { return v->running.data(); }
uint32_t sim_err(const sim_video* v) { return v->err; }
int sim_form(const sim_video* v) { return v->form; }
/* offset form: level records read / written and pixel-frames so far */
void sim_rec_traffic(const sim_video* v, uint64_t out[3]) { out[0] = v->rec_loads; out[1] = v->rec_stores; out[2] = v->px_frames; }
/* header words + node k of pixel i, for state comparison */
void sim_px(const sim_video* v, size_t i, float* lf, uint32_t* y) { *lf = v->hdr[i].lf; *y = v->hdr[i].y; }
void sim_node(const sim_video* v, size_t i, uint32_t k, float* integ, float* dt, float* best_dt, uint32_t* w) {
  if (v->form == 1 && k > 0) { /* offset form -> the reference's node (what adder_b200_video_read_px does in the library) */
    const uint32_t len = HDR_LENGTH(v->hdr[i].y);
    if (k + 1u >= len) { *integ = 0.0f; *dt = 0.0f; *best_dt = 0.0f; *w = 0u; return; }
    const bool frozen = HDR_POPPED(v->hdr[i].y) != 0u;
    const adder::OffRec r = (!frozen && k + 2u == len) ? adder::top_unpack(v->meta[i]) : v->recs[(size_t)k * v->P + i];
    const adder::Node& root = v->nodes[i];
    if (frozen) { *integ = adder::bits_f(r.oi); *dt = adder::bits_f(r.od); }
    else { *integ = adder::u2f(adder::f2u(root.integ) - r.oi); *dt = adder::u2f(adder::f2u(root.dt) - r.od); }
    *best_dt = r.best_dt; *w = r.w;
    return;
  }
  const adder::Node& n = v->nodes[(size_t)k * v->P + i];
  *integ = n.integ; *dt = n.dt; *best_dt = n.best_dt; *w = n.w;
}
}
