/*
 * px_offset.cuh — the OFFSET FORM of a pixel's node stack, and the state machine that works on it.
 *
 * The reference integrates EVERY live node of a pixel every frame (PixelArena::integrate,
 * adder-codec-rs/src/transcoder/event_pixel_tree.rs:340-390): `integration += intensity; delta_t += time` on each
 * level until the first level that fires.  On a scene that changes rarely the stacks are deep (2..11 live nodes after a
 * few hundred frames of a static pixel) and that walk IS the cost of the frame: 16 bytes read and written and ~80
 * instructions per level per pixel, the warp running as long as its deepest stack.
 *
 * All levels of one pixel receive the same increments.  So a level below the root need not be touched while it does
 * not fire: its values are the ROOT's values minus what the root had accumulated when the level was born,
 *
 *     integration_k = X - oi_k        delta_t_k = DT - od_k        (X, DT: the root's integration / delta_t)
 *
 * and the offsets (oi_k, od_k) are constants of the level.  Level k fires in the frame in which the root's new
 * integration X reaches  thr_k = oi_k + 2^d_k  (integrate_main's test `integration + intensity >= D_SHIFT[d]`,
 * :423, on exact integers).  The shallowest such level is what the reference's walk finds (:344-366): it gets its best
 * event and a fresh child, everything deeper is dropped.  Every level record carries
 *
 *     pmin_k, pk_k      the smallest threshold among the levels ABOVE level k and the shallowest level that has it,
 *
 * so the shallowest firing level is found from the deepest stored level (the TOP, level length-2) by hopping to pk while
 * pmin <= X, without looking at the levels that do not fire.
 *
 * The top lives in the second half of the pixel's record 0, next to the root, packed into 16 bytes (a fired level's d
 * determines its best_event.d, so `w` need not be stored).  The stack of a pixel that does not change behaves like a
 * binary counter: every frame exactly one level fires — the tail (a push: the stack grows by one), the top (in place),
 * or a level above it (the stack is cut back to that level, which becomes the top).  On aged 8K stacks these are 39 %,
 * 36 % and 19 % of the pixel-frames.  Only a push writes a level record (the previous top is spilled) and only a cut
 * reads one.  These scattered 32-byte accesses are what DRAM serves worst (a record load brings a whole 128-byte line,
 * profiles/r02r_*): the first version of this form, which rewrote the firing level's record every frame, was bound by
 * them and no faster than the eager form although it executed half the instructions.
 *
 * What makes this exact (bit-identical to the reference's f32 arithmetic):
 *  - a u8 source gives integer intensities, `time` is an integer number of ticks, and every integration / delta_t is a
 *    sum of those starting from 0 (PixelNode::new :502-514; integrate_main :418-479 never scales them): integers, exact
 *    in f32 below 2^24, so sums, differences and comparisons agree with integer arithmetic;
 *  - under PixelMultiMode::Collapse a root is popped when its delta_t reaches delta_t_max (:394-396), after which only
 *    the root integrates (:360-362): while levels below the root are live, DT <= dtm + time and X <= 255 * DT / time.
 *    The host selects this form only when those bounds are below 2^24 (offset_form_eligible, adder_b200.cu); everything
 *    else (PixelMultiMode::Normal, fractional time_spanned, huge delta_t_max / ref ratios) runs the eager form and px_step.
 *  - After a Δt_max pop under Collapse the levels below the root are never integrated again; they are stored with their
 *    actual values ("frozen": the header's popped_dtm bit says which reading applies).
 *
 * The tail.  For length >= 2 the last node of the stack is always a node that has never integrated (all zero): every
 * way the machine produces a stack of two or more levels ends with PixelNode::new (:344-355), and a fresh tail fires on
 * its first integration (its d is set from the intensity just before, :332-335, and 2^floor(log2 v) <= v; v = 0 gives
 * D_ZERO_INTEGRATION whose threshold is 0).  It is therefore not stored at all: level length-1 is implicit.
 *
 * What the machine relies on.  Starting from Video::new and running only this machine (or px_step on eligible parameters,
 * which is the same function of the state), a root that has integrated at least once always holds a best event (a fresh
 * node fires on its first integration, and every later root is either a fresh node or a level that has fired).  The
 * reference's branches for a root WITHOUT a best event at a Δt_max pop (zero event / synthesised event, :155-193), for a
 * popped root with children that has nothing to give (:249-265 looking further down) and for a root at D_MAX can
 * therefore not be reached; px_offset raises ADDER_DEVERR_INTERNAL there instead of carrying a copy of px_step (inlined next to the short
 * path, px_step cost 124 bytes of spills per thread and 25 % on every workload; as a __noinline__ function it crashes
 * ptxas 12.9).  The host simulation runs every shared case and the long runs of tests/test_px_offset_host.py through
 * this very function and checks that the flag stays clear.
 *
 * Layout (state_layout.h): record 0 of a pixel = root node (16 B, eager, as in the eager form) + the top level (16 B:
 * oi | d << 24, od | pk << 24, best_event.delta_t, pmin; meaningful when length >= 3 and the pixel is not frozen);
 * record k >= 1 = level k: { oi, od, best_event.delta_t, w, pmin, pk, -, - } (32 B = one DRAM sector, one 256-bit access);
 * the record of the top level itself is stale while the level is the top.
 */
#pragma once
#include "px_machine.cuh"

namespace adder {

constexpr uint32_t kThrNever = 0xFFFFFFFFu;

struct OffRec {
  uint32_t oi, od; /* offsets (or, frozen: the f32 bits of integration and delta_t) */
  float best_dt;
  uint32_t w;
  uint32_t pmin, pk;
};

/* the top level as it lies in record 0: a = oi | d << 24, b = od | pk << 24 */
struct OffTop {
  uint32_t a, b;
  float best_dt;
  uint32_t pmin;
};
/* `w` of a level that has fired: integrate_main leaves d = best_event.d + 1 (:449-461), or d = best_event.d = 128 for a
 * zero-integration node (d = D_MAX = 127 would need an integration of 2^126) */
ADDER_HD uint32_t w_of_d(uint32_t d) { return d >= 128u ? NODE_PACK(d, d, 1) : NODE_PACK(d, d - 1u, 1); }
ADDER_HD OffRec top_unpack(const OffTop& t) {
  OffRec q;
  q.oi = t.a & 0xFFFFFFu;
  q.od = t.b & 0xFFFFFFu;
  q.best_dt = t.best_dt;
  q.w = w_of_d(t.a >> 24);
  q.pmin = t.pmin;
  q.pk = t.b >> 24;
  return q;
}
ADDER_HD OffTop top_pack(const OffRec& q, uint32_t& errbits) {
  if (((q.oi | q.od) >> 24) || q.w != w_of_d(NODE_D(q.w))) errbits |= ADDER_DEVERR_INTERNAL; /* outside what the form admits */
  OffTop t;
  t.a = q.oi | (NODE_D(q.w) << 24);
  t.b = q.od | (q.pk << 24);
  t.best_dt = q.best_dt;
  t.pmin = q.pmin;
  return t;
}

/* the root integration at which a level with offset oi and decimation d (in w) fires */
ADDER_HD uint32_t off_thr(uint32_t oi, uint32_t w) {
  const uint32_t d = NODE_D(w);
  if (d >= 128u) return 0u;        /* D_SHIFT[128..] = 0: fires whenever it is reached */
  if (d >= 31u) return kThrNever;  /* 2^d beyond any integration this form admits */
  return oi + (1u << d);
}

/* Host side: may a launch with these parameters keep the node stacks in offset form?  Every integration and delta_t
 * that the form subtracts must be an integer below 2^24 (see above); the margins are a factor of two, so that calls with
 * different (eligible) time_spanned values may follow each other. */
inline bool offset_form_eligible(bool collapse, float time, uint32_t dtm) {
  if (!collapse) return false;
  if (!(time >= 1.0f) || !(time < 8388608.0f) || time != (float)(uint32_t)time) return false;
  const uint64_t t = (uint64_t)time;
  const uint64_t dt_max = (uint64_t)dtm + 2ull * t;  /* a live root's delta_t never exceeds dtm + time (:394-396) */
  const uint64_t x_max = 255ull * (dt_max / t + 2ull); /* and it has integrated at most that many samples */
  return dt_max < (1ull << 23) && x_max < (1ull << 23);
}

/* a stored level as the reference's node, given the root's integration / delta_t the offsets refer to */
ADDER_HD Node off_node(const OffRec& q, uint32_t x, uint32_t dt) {
  Node n;
  n.integ = u2f(x - q.oi);
  n.dt = u2f(dt - q.od);
  n.best_dt = q.best_dt;
  n.w = q.w;
  return n;
}

/*
 * One pixel, one frame, offset form (PixelMultiMode::Collapse: offset_form_eligible).  n0 = the root (in: as loaded, out:
 * as to be stored), top likewise; the caller stores header and record 0 afterwards.  Same contract as px_step otherwise.
 */
template <bool kPlain = false, class Mem, class Sink>
ADDER_HD bool px_offset(const PxParams& a, uint32_t v, PxHeader& h, Node& n0, OffTop& top, Mem& mem, Sink& sink, uint32_t& errbits,
                        uint8_t* disp) {
  const float intensity = (float)v;
  const float time = a.time;
  float lf = h.lf;
  uint32_t base = HDR_BASE(h.y), cth = HDR_CTHRESH(h.y), cnt = HDR_COUNTER(h.y);
  uint32_t len = HDR_LENGTH(h.y);
  uint32_t popped = HDR_POPPED(h.y);
  const uint32_t lo = base > cth ? base - cth : 0u;
  const uint32_t hi = base + cth > 255u ? 255u : base + cth;
  bool root_new = false;
  Node r = n0;

  if (v < lo || v > hi) { /* video.rs:1338-1358 -> pop_best_events (:213-287) */
    const uint32_t x0 = f2u(r.integ), dt0 = f2u(r.dt); /* before pop_node: a zero event clears the root's delta_t */
    const bool any = pop_node<kPlain>(a, sink, lf, r);
    bool fresh_root = len > 1u; /* the fresh tail becomes the root (:267-270); a stack of one node keeps its root as it is */
    if (popped) { /* Collapse after a Δt_max pop: the first event only, then D_EMPTY and PixelNode::new (:249-265) */
      if (any) {
        lf = a.running_t_prev; /* running_t before this frame's `+= time` (:337 runs later) */
        sink.push(ADDER_D_EMPTY, f2u(a.running_t_prev));
        fresh_root = true;
      } else if (len > 1u) {
        errbits |= ADDER_DEVERR_INTERNAL; /* (see the header: only a root that has never integrated has nothing to give,
                                           * and such a root — put there by a Δt_max pop in the frame it fired — has no child) */
      }
    } else if (len > 2u) { /* the stored levels in order, the top last; the tail is fresh and has nothing to give (:223-247) */
      for (uint32_t k = 1; k + 2u < len; k++) {
        Node nk = off_node(mem.load_rec(k), x0, dt0);
        pop_node<kPlain>(a, sink, lf, nk);
      }
      Node nk = off_node(top_unpack(top), x0, dt0);
      pop_node<kPlain>(a, sink, lf, nk);
    }
    if (fresh_root) r.integ = 0.0f, r.dt = 0.0f, r.best_dt = 0.0f, r.w = 0u;
    len = 1;
    popped = 0;
    base = v;
    root_new = true;
  }

  /* ---- integrate (:317-413) -------------------------------------------------------------------------------------------
   * Exactly one node of a pixel fires in a frame (the root, or the shallowest level below it that reaches its threshold —
   * the tail if no other does), or none (a root that integrates alone).  The firing arm of integrate_main, with its
   * division, is therefore written ONCE, after the node has been chosen: the pixels of a warp's row choose differently,
   * and a warp runs every arm one of its lanes takes. */
  if (len == 1u && r.dt == 0.0f && r.integ == 0.0f) r.w = (r.w & ~0xFFu) | get_d_from_intensity(intensity); /* :332-335 */
  const float sum0 = rn_add(r.integ, intensity);
  const bool fired0 = sum0 >= d_shift_f32(NODE_D(r.w)); /* :423 */
  uint32_t new_len = len;
  uint32_t x = 0, dt = 0;          /* the root's integration / delta_t after this frame, when a level below it fires */
  uint32_t k = 0;                  /* that level */
  OffRec q{0u, 0u, 0.0f, 0u, kThrNever, 0u};
  Node fn = r;                     /* the node that fires, as it is before this frame */
  bool level_fires = false;
  if (!fired0) {
    r.integ = sum0; /* :468-470 */
    r.dt = rn_add(r.dt, time);
    if (!popped && len > 1u && !(r.dt >= a.dtm_f)) {
      /* the walk below the root (:340-390) in one step: the shallowest level whose threshold the root's integration has
       * reached, else the fresh tail */
      x = f2u(r.integ), dt = f2u(r.dt);
      if ((x | dt) >> 24) errbits |= ADDER_DEVERR_INTERNAL; /* beyond what offset_form_eligible admits */
      bool from_tail = true;
      if (len > 2u) {
        const OffRec t = top_unpack(top);
        const uint32_t thr_t = off_thr(t.oi, t.w);
        if (t.pmin <= x) { /* a level above the top has reached its threshold: the walk stops at the shallowest (:344-366),
                            * which becomes the top (everything below it is dropped, its own record goes stale) */
          k = t.pk;
          q = mem.load_rec(k);
          while (q.pmin <= x) {
            k = q.pk;
            q = mem.load_rec(k);
          }
          from_tail = false;
        } else if (thr_t <= x) { /* the top fires where it is */
          k = len - 2u;
          q = t;
          from_tail = false;
        } else { /* nothing stored fires: the tail does, and the top is spilled to its record */
          k = len - 1u;
          mem.store_rec(len - 2u, t);
          if (t.pmin <= thr_t) { /* ties go to the shallower level, like the walk */
            q.pmin = t.pmin;
            q.pk = t.pk;
          } else {
            q.pmin = thr_t;
            q.pk = len - 2u;
          }
        }
      } else { /* root and tail only */
        k = 1u;
      }
      if (from_tail) { /* PixelNode::new with its d from this intensity (:332-335) */
        fn.integ = 0.0f, fn.dt = 0.0f, fn.best_dt = 0.0f;
        fn.w = get_d_from_intensity(intensity);
      } else { /* the root did not fire: before this frame it held x - v and dt - time */
        fn = off_node(q, x - v, dt - f2u(time));
      }
      if (!(rn_add(fn.integ, intensity) >= d_shift_f32(NODE_D(fn.w)))) errbits |= ADDER_DEVERR_INTERNAL;
      level_fires = true;
    }
  }
  if (fired0 || level_fires) integrate_fire(fn, intensity, time); /* :424-466 */
  if (fired0) {
    r = fn;
    if (NODE_D(r.w) == ADDER_D_MAX) errbits |= ADDER_DEVERR_INTERNAL; /* needs an integration of 2^126 */
  }
  const uint32_t dtm_reached = r.dt >= a.dtm_f ? 1u : 0u; /* :394 */

  if (dtm_reached && !popped) {
    /* ---- pop_top_event (video.rs:1371-1374, :139-210): the root's best event leaves ------------------------------- */
    if (!NODE_HAS_BEST(r.w)) errbits |= ADDER_DEVERR_INTERNAL; /* (see the header) */
    emit_abs<kPlain>(a, sink, lf, NODE_BEST_D(r.w), r.best_dt);
    popped = 1;
    root_new = true;
    if (fired0 || len < 2u) { /* the root fired this very frame: it is replaced by a fresh node (:164-193 -> :195-199) */
      if (!fired0) errbits |= ADDER_DEVERR_INTERNAL; /* a root with a best event that did not fire has a child */
      r.integ = 0.0f, r.dt = 0.0f, r.best_dt = 0.0f;
      r.w = get_d_from_intensity(intensity);
      new_len = 1;
    } else {
      /* the stack moves up a level (:201-204) while every level integrates this frame (:340-390) up to the first one
       * that fires.  What is stored below the new root is never integrated again under Collapse: frozen values, all of
       * them in level records (a frozen pixel keeps no top in record 0). */
      const uint32_t last = len - 1u;
      const uint32_t x_in = f2u(r.integ) - v, dt_in = f2u(r.dt) - f2u(time); /* the root did not fire: it accumulated (:468-470) */
      for (uint32_t kk = 1;; kk++) {
        Node nk;
        if (kk == last) { /* the fresh tail */
          nk.integ = 0.0f, nk.dt = 0.0f, nk.best_dt = 0.0f;
          nk.w = get_d_from_intensity(intensity);
        } else if (kk + 1u == last) {
          nk = off_node(top_unpack(top), x_in, dt_in);
        } else {
          nk = off_node(mem.load_rec(kk), x_in, dt_in);
        }
        const bool fired = integrate_main(nk, intensity, time);
        if (kk == 1u) {
          r = nk;
        } else {
          OffRec f;
          f.oi = f_bits(nk.integ), f.od = f_bits(nk.dt), f.best_dt = nk.best_dt, f.w = nk.w, f.pmin = kThrNever, f.pk = 0u;
          mem.store_rec(kk - 1u, f);
        }
        if (fired || kk == last) {
          if (!fired) errbits |= ADDER_DEVERR_INTERNAL;
          new_len = kk + 1u; /* levels 0 .. kk-1 and a fresh tail */
          break;
        }
      }
    }
  } else if (fired0) { /* :344-355: a fresh child, deeper nodes dropped */
    if (a.depth > 1u) new_len = 2; else errbits |= ADDER_DEVERR_DEPTH;
    root_new = true;
  } else if (level_fires) { /* the level that fired is the top now, with a fresh tail below it */
    q.oi = x - f2u(fn.integ);
    q.od = dt - f2u(fn.dt);
    q.best_dt = fn.best_dt;
    q.w = fn.w;
    top = top_pack(q, errbits);
    if (k + 1u >= a.depth) errbits |= ADDER_DEVERR_DEPTH;
    new_len = k + 2u > a.depth ? a.depth : k + 2u;
  }
  /* (popped: Collapse integrates the root only, :360-362; a stack of one node has nothing below the root) */

  if (cth < a.c_max) { /* :402-412 */
    if (cnt >= a.vel_m1) {
      cth = cth + 1u > 255u ? 255u : cth + 1u;
      cnt = 0;
    } else {
      cnt = cnt + a.cnt_inc > 255u ? 255u : cnt + a.cnt_inc;
    }
  }
  n0 = r;
  h.lf = lf;
  h.y = HDR_PACK(base, cth, cnt, new_len, dtm_reached, popped);
  /* video.rs:713-730.  An unchanged root best event gives the byte already in running_intensities. */
  if (NODE_HAS_BEST(r.w) && a.display && (root_new || a.display == 2u || ADDER_VIEW_OF(a) == 3u)) {
    *disp = frame_value_u8<kPlain>(a, NODE_BEST_D(r.w), f2u(r.best_dt), lf);
    return true;
  }
  return false;
}

}  // namespace adder
