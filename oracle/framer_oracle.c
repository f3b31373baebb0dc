/*
 * framer_oracle.c — CPU restatement of the reference's INSTANTANEOUS framer (events -> u8 frames).
 *
 * TEST INFRASTRUCTURE ONLY (see adder_oracle.h).  Follows, for T = u8 and FramerMode::INSTANTANEOUS:
 *   FrameSequence::new               adder-codec-rs/src/framer/driver.rs:300-399
 *   Framer::ingest_event             :437-562 (feature detection branch left out)
 *   Framer::ingest_events_events     :564-626
 *   Framer::flush_frame_buffer       :633-680
 *   is_frame_filled / pop_next_frame_for_chunk / write_frame_bytes / write_multi_frame_bytes   :820-982
 *   ingest_event_for_chunk           :984-1133
 * PINNED by the reference's own golden pairs (tests/integration_tests.rs:818-962 test_sample_{un,}ordered:
 * sample_3_*.adder -> sample_3.gray, 405 frames; and the `dark` test's lake_scaled_hd_out.adder -> lake_scaled_out),
 * committed under tests/golden/ — see tests/test_framer_oracle.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "adder_oracle.h"

typedef struct fr_frame {
  uint8_t* val;  /* Option<u8>: value */
  uint8_t* some; /* Option<u8>: is_some */
  size_t filled_count;
} fr_frame;

typedef struct fr_deque { /* VecDeque<Frame<Option<u8>>> */
  fr_frame* f;
  size_t head, len, cap;
  size_t px; /* array.len() of every frame of this chunk */
} fr_deque;

struct oracle_framer {
  uint16_t w, h;
  uint8_t c;
  uint32_t chunk_rows, n_chunks;
  /* FrameSequenceState :230-249 */
  int64_t frames_written;
  uint32_t tpf, tps, ref_interval, source_dtm;
  uint8_t codec_version;
  uint32_t source_camera;
  int view_mode, time_mode;
  fr_deque* frames;
  int64_t* frame_idx_offsets;
  uint64_t* pixel_ts;      /* pixel_ts_tracker, (H,W,C) */
  int64_t* last_filled;    /* last_filled_tracker, init -1 */
  uint8_t* last_intensity; /* last_frame_intensity_tracker */
  uint8_t* chunk_filled;
  int64_t buffer_limit; /* < 0 = None */
  int bad;              /* an index the reference would have panicked on */
};

static fr_frame frame_new(size_t px) {
  fr_frame fr;
  fr.val = (uint8_t*)calloc(px ? px : 1, 1);
  fr.some = (uint8_t*)calloc(px ? px : 1, 1);
  fr.filled_count = 0;
  return fr;
}
static void frame_free(fr_frame* fr) {
  free(fr->val);
  free(fr->some);
}
static fr_frame* dq_at(fr_deque* d, size_t i) { return &d->f[d->head + i]; }
static void dq_push_back(fr_deque* d, fr_frame fr) {
  if (d->head + d->len == d->cap) {
    if (d->head > 0) { /* compact */
      memmove(d->f, d->f + d->head, d->len * sizeof(fr_frame));
      d->head = 0;
    }
    if (d->len == d->cap) {
      d->cap = d->cap ? d->cap * 2 : 4;
      d->f = (fr_frame*)realloc(d->f, d->cap * sizeof(fr_frame));
    }
  }
  d->f[d->head + d->len++] = fr;
}
static fr_frame dq_pop_front(fr_deque* d) {
  fr_frame fr = d->f[d->head];
  d->head++;
  d->len--;
  return fr;
}

oracle_framer* oracle_framer_new(uint16_t w, uint16_t h, uint8_t c, uint32_t chunk_rows, uint8_t codec_version, int time_mode,
                                 uint32_t tps, uint32_t ref_interval, uint32_t delta_t_max, float output_fps /* <= 0 = None */,
                                 int view_mode, uint32_t source_camera, int64_t buffer_limit) {
  if (!w || !h || !c || !chunk_rows) return NULL;
  oracle_framer* f = (oracle_framer*)calloc(1, sizeof(*f));
  f->w = w;
  f->h = h;
  f->c = c;
  f->chunk_rows = chunk_rows;
  f->n_chunks = (h + chunk_rows - 1) / chunk_rows; /* ceil(h / chunk_rows) :304 */
  const uint32_t last_chunk_rows = h - (f->n_chunks - 1) * chunk_rows;
  f->frames = (fr_deque*)calloc(f->n_chunks, sizeof(fr_deque));
  for (uint32_t k = 0; k < f->n_chunks; k++) {
    f->frames[k].px = (size_t)(k + 1 == f->n_chunks ? last_chunk_rows : chunk_rows) * w * c;
    dq_push_back(&f->frames[k], frame_new(f->frames[k].px));
  }
  const size_t n = (size_t)w * h * c;
  f->frame_idx_offsets = (int64_t*)calloc(f->n_chunks, sizeof(int64_t));
  f->pixel_ts = (uint64_t*)calloc(n, sizeof(uint64_t));
  f->last_filled = (int64_t*)malloc(n * sizeof(int64_t));
  for (size_t i = 0; i < n; i++) f->last_filled[i] = -1; /* :349-353 */
  f->last_intensity = (uint8_t*)calloc(n, 1);
  f->chunk_filled = (uint8_t*)calloc(f->n_chunks, 1);
  f->tpf = output_fps > 0.0f ? (uint32_t)((float)tps / output_fps) : ref_interval; /* :355-359 */
  f->tps = tps;
  f->ref_interval = ref_interval;
  f->source_dtm = delta_t_max;
  f->codec_version = codec_version;
  f->source_camera = source_camera;
  f->view_mode = view_mode;
  f->time_mode = time_mode;
  f->buffer_limit = buffer_limit;
  return f;
}

void oracle_framer_delete(oracle_framer* f) {
  if (!f) return;
  for (uint32_t k = 0; k < f->n_chunks; k++) {
    for (size_t i = 0; i < f->frames[k].len; i++) frame_free(dq_at(&f->frames[k], i));
    free(f->frames[k].f);
  }
  free(f->frames);
  free(f->frame_idx_offsets);
  free(f->pixel_ts);
  free(f->last_filled);
  free(f->last_intensity);
  free(f->chunk_filled);
  free(f);
}

static int is_framed_camera(uint32_t cam) { return cam <= 5; } /* FramedU8..FramedF64, lib.rs:35-47 */

/* ingest_event_for_chunk, :984-1133.  (y is chunk-relative in the reference; the trackers here are whole-plane
 * arrays, so `gi` is the global pixel index and `li` the index inside the chunk's frame arrays.) */
static int ingest_event_for_chunk(oracle_framer* f, uint32_t chunk, adder_event_t e, size_t gi, size_t li) {
  fr_deque* fc = &f->frames[chunk];
  int64_t* last_filled_frame_ref = &f->last_filled[gi];
  uint64_t* running_ts_ref = &f->pixel_ts[gi];
  const int64_t prev_last_filled_frame = *last_filled_frame_ref;
  const uint64_t prev_running_ts = *running_ts_ref;

  if (f->codec_version >= 2 && f->time_mode == ADDER_TIME_ABSOLUTE_T) { /* :1002-1012 */
    if (prev_running_ts >= (uint64_t)e.t) return dq_at(fc, 0)->filled_count == fc->px;
    *running_ts_ref = e.t;
  } else {
    *running_ts_ref += e.t;
  }

  const uint64_t ts_m1 = *running_ts_ref ? *running_ts_ref - 1 : 0; /* saturating_sub(1) */
  if ((int64_t)ts_m1 / (int64_t)f->tpf > *last_filled_frame_ref) { /* :1014 */
    if (e.d != ADDER_D_EMPTY) {
      const float practical_d_max = oracle_log2_raw(255.0f * (float)(f->source_dtm / f->ref_interval)); /* u8::max_f32() = 255 */
      uint32_t t = e.t;
      if (f->codec_version >= 2 && f->time_mode == ADDER_TIME_ABSOLUTE_T && f->view_mode != ADDER_VIEW_SAE)
        t = (uint32_t)prev_running_ts > t ? 0 : t - (uint32_t)prev_running_ts; /* saturating_sub(prev_running_ts as u32) :1027 */
      f->last_intensity[gi] = oracle_get_frame_value_u8(e.d, t, (double)f->ref_interval, practical_d_max, f->source_dtm, f->view_mode,
                                                       (uint32_t)*running_ts_ref, (uint32_t)prev_running_ts); /* :1031-1042 */
    }
    *last_filled_frame_ref = (int64_t)ts_m1 / (int64_t)f->tpf; /* :1045 */

    const int64_t a = *last_filled_frame_ref - f->frame_idx_offsets[chunk]; /* :1048-1075 */
    if (a > 0) {
      for (int64_t k = 0; k < a; k++) dq_push_back(fc, frame_new(fc->px));
      f->frame_idx_offsets[chunk] += a;
    }
    for (int64_t i = prev_last_filled_frame; i < *last_filled_frame_ref; i++) { /* :1078-1091 */
      const int64_t idx = i - f->frames_written + 1;
      if (idx >= 0) {
        if ((size_t)idx >= fc->len) { /* the reference indexes the VecDeque directly and would panic */
          f->bad = 1;
          break;
        }
        fr_frame* fr = dq_at(fc, (size_t)idx);
        if (!fr->some[li]) {
          fr->some[li] = 1;
          fr->val[li] = f->last_intensity[gi];
          fr->filled_count++;
        }
      }
    }
  }

  /* framed sources: the next event of the pixel starts on a frame boundary, :1094-1113 */
  if (f->codec_version >= 1 && is_framed_camera(f->source_camera) && *running_ts_ref % f->ref_interval > 0)
    *running_ts_ref = (*running_ts_ref / f->ref_interval + 1) * (uint64_t)f->ref_interval;

  if (f->buffer_limit >= 0 && *last_filled_frame_ref > f->frames_written + f->buffer_limit) /* :1115-1121 */
    dq_at(fc, 0)->filled_count = fc->px;
  if (dq_at(fc, 0)->filled_count > fc->px) dq_at(fc, 0)->filled_count = fc->px; /* :1124-1126 */
  return dq_at(fc, 0)->filled_count == fc->px;
}

static int all_chunks_filled(const oracle_framer* f) {
  for (uint32_t k = 0; k < f->n_chunks; k++)
    if (!f->chunk_filled[k]) return 0;
  return 1;
}

/* is_frame_0_filled, :851-866 */
static int is_frame_0_filled(const oracle_framer* f) {
  if (f->buffer_limit >= 0)
    for (uint32_t k = 0; k < f->n_chunks; k++)
      if (f->frames[k].len > (size_t)f->buffer_limit) return 1;
  return all_chunks_filled(f);
}

/* Framer::ingest_event, :437-562 */
int oracle_framer_ingest_event(oracle_framer* f, adder_event_t e) {
  const uint32_t channel = e.c == ADDER_C_NONE ? 0 : e.c;
  const uint32_t chunk = e.y / f->chunk_rows;
  if (chunk >= f->n_chunks) return 0; /* silently handle malformed event :441-444 */
  if (e.x >= f->w || channel >= f->c) { /* ndarray would panic */
    f->bad = 1;
    return 0;
  }
  const size_t gi = ((size_t)e.y * f->w + e.x) * f->c + channel;
  const size_t li = gi - (size_t)chunk * f->chunk_rows * f->w * f->c;
  f->chunk_filled[chunk] = (uint8_t)ingest_event_for_chunk(f, chunk, e, gi, li);
  return all_chunks_filled(f); /* :556-562 */
}

/* Framer::ingest_events_events, :564-626: events[k] belongs to chunk k (asserted equal lengths :566) */
int oracle_framer_ingest_events_events(oracle_framer* f, const adder_event_t* ev, const uint32_t* chunk_counts, uint32_t n_counts) {
  if (n_counts != f->n_chunks) {
    f->bad = 1;
    return 0;
  }
  size_t pos = 0;
  for (uint32_t k = 0; k < f->n_chunks; k++) {
    for (uint32_t j = 0; j < chunk_counts[k]; j++, pos++) {
      const adder_event_t e = ev[pos];
      const uint32_t channel = e.c == ADDER_C_NONE ? 0 : e.c;
      const uint32_t chunk_num = e.y / f->chunk_rows; /* the event's own chunk decides the row, the loop's chunk the frames :601-606 */
      const size_t gi = ((size_t)e.y * f->w + e.x) * f->c + channel;
      const size_t li = gi - (size_t)chunk_num * f->chunk_rows * f->w * f->c;
      if (chunk_num != k || e.x >= f->w || channel >= f->c) { /* mis-chunked input indexes foreign arrays in the reference */
        f->bad = 1;
        continue;
      }
      f->chunk_filled[k] = (uint8_t)ingest_event_for_chunk(f, k, e, gi, li);
    }
  }
  return is_frame_0_filled(f);
}

/* is_frame_filled(0), :820-848: 1 filled, 0 not, -1 error */
static int is_frame_filled0(const oracle_framer* f) {
  for (uint32_t k = 0; k < f->n_chunks; k++) {
    if (f->frames[k].len == 0) return -1;
    const size_t fcnt = f->frames[k].f[f->frames[k].head].filled_count;
    if (fcnt == f->frames[k].px) continue;
    if (fcnt > f->frames[k].px) return -1;
    return 0;
  }
  return 1;
}

/* pop_next_frame_for_chunk, :905-927 */
static fr_frame pop_next_frame_for_chunk(oracle_framer* f, uint32_t k) {
  fr_deque* d = &f->frames[k];
  fr_frame a = dq_pop_front(d); /* rotate_left(1) + pop_back */
  if (d->len == 0) {
    dq_push_back(d, frame_new(d->px));
    f->frame_idx_offsets[k] += 1;
  }
  f->chunk_filled[k] = dq_at(d, 0)->filled_count == d->px;
  return a;
}

/* write_multi_frame_bytes, :971-982 (write_frame_bytes :936-962 inlined): frames appended to out, None -> 0.
 * Returns the number of frames written, or -1 on the reference's error paths / when `cap` is too small. */
int oracle_framer_write_multi_frame_bytes(oracle_framer* f, uint8_t* out, size_t cap, size_t* n_bytes) {
  int frame_count = 0;
  size_t pos = 0;
  const size_t frame_bytes = (size_t)f->w * f->h * f->c;
  for (;;) {
    const int filled = is_frame_filled0(f);
    if (filled < 0) return -1;
    if (!filled) break;
    if (pos + frame_bytes > cap) return -1;
    for (uint32_t k = 0; k < f->n_chunks; k++) {
      fr_frame a = pop_next_frame_for_chunk(f, k);
      for (size_t i = 0; i < f->frames[k].px; i++) out[pos++] = a.some[i] ? a.val[i] : 0; /* None -> T::default() */
      frame_free(&a);
    }
    f->frames_written += 1;
    frame_count++;
  }
  if (n_bytes) *n_bytes = pos;
  return frame_count;
}

/* flush_frame_buffer, :633-680 */
int oracle_framer_flush_frame_buffer(oracle_framer* f) {
  int any_nonempty = 0;
  for (uint32_t k = 0; k < f->n_chunks; k++)
    if (f->frames[k].len > 1) any_nonempty = 1;
  if (any_nonempty) {
    for (uint32_t k = 0; k < f->n_chunks; k++) {
      fr_frame* fr = dq_at(&f->frames[k], 0);
      const size_t base = (size_t)k * f->chunk_rows * f->w * f->c;
      for (size_t i = 0; i < f->frames[k].px; i++) {
        if (!fr->some[i]) {
          fr->some[i] = 1;
          fr->val[i] = f->last_intensity[base + i];
          fr->filled_count++;
          f->last_filled[base + i] += 1;
        }
      }
      f->chunk_filled[k] = 1;
    }
  } else {
    f->chunk_filled[0] = 0;
  }
  return is_frame_0_filled(f);
}

int oracle_framer_bad(const oracle_framer* f) { return f->bad; }
int64_t oracle_framer_frames_written(const oracle_framer* f) { return f->frames_written; }
uint32_t oracle_framer_tpf(const oracle_framer* f) { return f->tpf; }
