/*
 * consumer.c — a plain C consumer of include/adder_b200.h (no Python, no C++): what a Rust `extern "C"` block sees.
 *
 * It does what the replacement body of Video::integrate_matrix in INTEGRATION.md §3 does for every frame:
 *   1. adder_b200_video_integrate_matrix   one frame in (host memory), all events of the frame out (host memory)
 *   2. builds the reference's return value, one vector of events per chunk (Vec<Vec<Event>>, video.rs:677-734;
 *      the framer asserts one vector per chunk, framer/driver.rs:566), mapping the 12-byte record to the
 *      reference's Event{coord{x,y,c:Option<u8>},d,t} layout
 * and times the two parts separately, so that the cost of the drop-in call behind consume() is known with the
 * vector build included (VERDICT r1 task 6).
 *
 * The vector build is a map over independent chunks; the reference runs that loop on its rayon pool (video.rs:677-734).  It is
 * timed twice per frame: on one thread, and on a persistent pool of `threads` workers (chunk ci goes to worker ci % threads),
 * which is what the Rust shim would do with the pool the reference already has.
 *
 *   usage: consumer <cfg1|cfg2|WxHxC:kind> <frames> <out.bin> [pageable|pinned] [threads]
 *     cfg1 = 640x480 gray gradient, Video::new defaults (c_thresh 10), dtm = ref = 255        (BASELINE configs[0])
 *     cfg2 = 1920x1080 RGB noise, crf 3, ref 255, dtm 7650                                     (BASELINE configs[1])
 *   out.bin: per frame  u32 n_chunks | u32 counts[n_chunks] | records[sum] (12 bytes each)   — compared with the oracle by
 *   tests/test_gpu_c_consumer.py.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "adder_b200.h"

/* layout checks a C compiler can make for the Rust side */
_Static_assert(sizeof(adder_event_t) == 12, "adder_event_t is a 12-byte record");
_Static_assert(sizeof(adder_crf_parameters_t) == 8, "adder_crf_parameters_t is 8 bytes");

/* the reference's Event as rustc lays it out today (adder-codec-core/src/lib.rs:371-377, repr(packed)):
 * Coord{x:u16,y:u16,c:Option<u8>} = 2+2+(tag,payload), d:u8, t:u32 = 11 bytes */
#pragma pack(push, 1)
typedef struct ref_event {
  uint16_t x, y;
  uint8_t c_is_some, c;
  uint8_t d;
  uint32_t t;
} ref_event;
#pragma pack(pop)
_Static_assert(sizeof(ref_event) == 11, "packed Event is 11 bytes");

typedef struct chunk_vec { ref_event* data; size_t len, cap; } chunk_vec;

/* chunk ci of the frame: a fresh vector with_capacity(count) (the reference allocates one per chunk per frame, video.rs:693),
 * records [first, first + count) mapped to the reference's Event */
static void build_chunk(chunk_vec* cv, const adder_event_t* ev, size_t first, uint32_t count) {
  free(cv->data);
  cv->len = cv->cap = count;
  cv->data = (ref_event*)malloc((cv->cap ? cv->cap : 1) * sizeof(ref_event));
  for (size_t j = 0; j < cv->len; j++) {
    const adder_event_t* r = &ev[first + j];
    ref_event* e = &cv->data[j];
    e->x = r->x;
    e->y = r->y;
    e->c_is_some = r->c != ADDER_C_NONE;
    e->c = e->c_is_some ? r->c : 0;
    e->d = r->d;
    e->t = r->t;
  }
}

/* a persistent pool (the stand-in for rayon's): the workers wait at a barrier for a frame, build their chunks, meet again */
typedef struct pool {
  pthread_barrier_t bar;
  uint32_t n_threads, n_chunks;
  chunk_vec* chunks;
  const adder_event_t* ev;
  const uint32_t* counts;
  const size_t* first; /* exclusive prefix of counts */
  int stop;
} pool;
typedef struct worker { pool* p; uint32_t id; } worker;
static void* worker_main(void* arg) {
  worker* w = (worker*)arg;
  pool* p = w->p;
  for (;;) {
    pthread_barrier_wait(&p->bar); /* a frame is ready (or stop) */
    if (p->stop) return NULL;
    for (uint32_t ci = w->id; ci < p->n_chunks; ci += p->n_threads) build_chunk(&p->chunks[ci], p->ev, p->first[ci], p->counts[ci]);
    pthread_barrier_wait(&p->bar); /* the frame's vectors are built */
  }
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
/* the synthetic frames of SURVEY.md 8(d) (tests/synth.py): kind 0 gradient, 1 noise */
static void make_frame(uint8_t* out, int kind, uint32_t f, uint32_t w, uint32_t h, uint32_t c) {
  const uint64_t seed = 0xADDE5ull;
  size_t i = 0;
  for (uint32_t y = 0; y < h; y++)
    for (uint32_t x = 0; x < w; x++)
      for (uint32_t ch = 0; ch < c; ch++, i++)
        out[i] = kind == 0 ? (uint8_t)((x + 2u * y + 3u * f) & 255u) : (uint8_t)(splitmix64(seed ^ ((uint64_t)f << 40) ^ (uint64_t)i) & 0xFFu);
}
#define CHECK(call)                                                                      \
  do {                                                                                   \
    int _rc = (call);                                                                    \
    if (_rc != ADDER_OK) {                                                               \
      fprintf(stderr, "%s -> %d: %s\n", #call, _rc, adder_b200_last_error());           \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s <cfg1|cfg2|WxHxC:kind> <frames> <out.bin> [pageable]\n", argv[0]);
    return 1;
  }
  uint32_t w = 640, h = 480, c = 1, ref = 255, dtm = 255;
  int kind = 0, crf = -1;
  if (!strcmp(argv[1], "cfg2")) {
    w = 1920, h = 1080, c = 3, dtm = 7650, kind = 1, crf = 3;
  } else if (strcmp(argv[1], "cfg1")) {
    if (sscanf(argv[1], "%ux%ux%u:%d", &w, &h, &c, &kind) != 4) return 1;
    dtm = 7650, crf = 3;
  }
  const uint32_t n_frames = (uint32_t)atoi(argv[2]);
  const int pageable = argc > 4 && !strcmp(argv[4], "pageable");
  const uint32_t n_threads = argc > 5 ? (uint32_t)atoi(argv[5]) : 0u; /* 0: no pool, the single-threaded build only */
  FILE* out = fopen(argv[3], "wb");
  if (!out) return 1;
  if (adder_b200_abi_version() != ADDER_B200_ABI_VERSION) {
    fprintf(stderr, "header / library ABI mismatch\n");
    return 2;
  }

  adder_b200_video* v = NULL;
  CHECK(adder_b200_video_create((uint16_t)w, (uint16_t)h, (uint8_t)c, ADDER_MODE_FRAME_PERFECT, 0, 0, &v)); /* Video::new */
  int applied = 0;
  CHECK(adder_b200_video_time_parameters(v, ref * 30u, ref, dtm, -1, &applied)); /* auto_time_parameters, framed.rs:94-111 */
  if (!applied) return 2;
  if (crf >= 0) CHECK(adder_b200_video_update_crf(v, (uint8_t)crf));
  adder_b200_video_info_t info;
  CHECK(adder_b200_video_get_info(v, &info));
  const uint32_t n_chunks = info.n_chunks;
  const size_t P = (size_t)w * h * c, cap = P * 4;

  uint8_t* frame = NULL;
  adder_event_t* ev = NULL;
  if (pageable) {
    frame = (uint8_t*)malloc(P);
    ev = (adder_event_t*)malloc(cap * sizeof(adder_event_t));
  } else { /* page-locked, as INTEGRATION.md §4 recommends */
    CHECK(adder_b200_host_alloc(P, (void**)&frame));
    CHECK(adder_b200_host_alloc(cap * sizeof(adder_event_t), (void**)&ev));
  }
  uint32_t* counts = (uint32_t*)malloc(n_chunks * sizeof(uint32_t));
  chunk_vec* chunks = (chunk_vec*)calloc(n_chunks, sizeof(chunk_vec));      /* built by the main thread */
  chunk_vec* chunks_pool = (chunk_vec*)calloc(n_chunks, sizeof(chunk_vec)); /* built by the pool: every vector is allocated and freed by the same worker */
  size_t* first = (size_t*)malloc(n_chunks * sizeof(size_t));
  if (!frame || !ev || !counts || !chunks || !chunks_pool || !first) return 2;
  pool pl;
  pthread_t* threads = NULL;
  worker* workers = NULL;
  if (n_threads) {
    memset(&pl, 0, sizeof(pl));
    pl.n_threads = n_threads, pl.n_chunks = n_chunks, pl.chunks = chunks_pool, pl.ev = ev, pl.counts = counts, pl.first = first;
    if (pthread_barrier_init(&pl.bar, NULL, n_threads + 1u)) return 2;
    threads = (pthread_t*)malloc(n_threads * sizeof(pthread_t));
    workers = (worker*)malloc(n_threads * sizeof(worker));
    if (!threads || !workers) return 2;
    for (uint32_t t = 0; t < n_threads; t++) {
      workers[t].p = &pl, workers[t].id = t;
      if (pthread_create(&threads[t], NULL, worker_main, &workers[t])) return 2;
    }
  }

  double t_call = 0.0, t_vec = 0.0, t_vec_mt = 0.0;
  uint64_t total = 0;
  for (uint32_t f = 0; f < n_frames; f++) {
    make_frame(frame, kind, f, w, h, c); /* stands for the decoder; not timed */
    uint64_t n = 0;
    const double t0 = now_s();
    CHECK(adder_b200_video_integrate_matrix(v, frame, 0, (float)ref, ev, cap, counts, &n));
    const double t1 = now_s();
    /* Vec<Vec<Event>>: one vector per chunk, with_capacity(count), records mapped to the reference's Event */
    size_t k = 0;
    for (uint32_t ci = 0; ci < n_chunks; ci++) {
      first[ci] = k;
      k += counts[ci];
    }
    for (uint32_t ci = 0; ci < n_chunks; ci++) build_chunk(&chunks[ci], ev, first[ci], counts[ci]);
    const double t2 = now_s();
    if (n_threads) { /* the same build on the pool, into its own set of vectors (the ones the test reads back) */
      pthread_barrier_wait(&pl.bar);
      pthread_barrier_wait(&pl.bar);
    }
    const double t3 = now_s();
    t_vec_mt += t3 - t2;
    if (k != n) return 3;
    t_call += t1 - t0;
    t_vec += t2 - t1;
    total += n;
    /* what the test compares: read back from the per-chunk vectors, not from the library's buffer */
    fwrite(&n_chunks, 4, 1, out);
    fwrite(counts, 4, n_chunks, out);
    const chunk_vec* built = n_threads ? chunks_pool : chunks;
    for (uint32_t ci = 0; ci < n_chunks; ci++)
      for (size_t j = 0; j < built[ci].len; j++) {
        const ref_event* e = &built[ci].data[j];
        adder_event_t r = {e->x, e->y, e->c_is_some ? e->c : (uint8_t)ADDER_C_NONE, e->d, 0, e->t};
        fwrite(&r, sizeof(r), 1, out);
      }
  }
  fclose(out);
  printf("{\"workload\": \"%s\", \"plane\": \"%ux%ux%u\", \"frames\": %u, \"events\": %llu, \"buffers\": \"%s\", "
         "\"ms_per_call_integrate_matrix\": %.4f, \"ms_per_frame_vec_vec_event\": %.4f, \"ms_per_consume\": %.4f, \"mpx_per_s\": %.1f, "
         "\"pool_threads\": %u, \"ms_per_frame_vec_vec_event_pool\": %.4f, \"ms_per_consume_pool\": %.4f, \"mpx_per_s_pool\": %.1f}\n",
         argv[1], w, h, c, n_frames, (unsigned long long)total, pageable ? "pageable" : "page-locked", 1e3 * t_call / n_frames,
         1e3 * t_vec / n_frames, 1e3 * (t_call + t_vec) / n_frames, (double)P * n_frames / (t_call + t_vec) / 1e6, n_threads,
         1e3 * t_vec_mt / n_frames, 1e3 * (t_call + t_vec_mt) / n_frames, n_threads ? (double)P * n_frames / (t_call + t_vec_mt) / 1e6 : 0.0);
  if (n_threads) {
    pl.stop = 1;
    pthread_barrier_wait(&pl.bar);
    for (uint32_t t = 0; t < n_threads; t++) pthread_join(threads[t], NULL);
    pthread_barrier_destroy(&pl.bar);
    free(threads);
    free(workers);
  }
  free(first);
  for (uint32_t ci = 0; ci < n_chunks; ci++) free(chunks[ci].data), free(chunks_pool[ci].data);
  free(chunks);
  free(chunks_pool);
  free(counts);
  if (pageable) {
    free(frame);
    free(ev);
  } else {
    adder_b200_host_free(frame);
    adder_b200_host_free(ev);
  }
  adder_b200_video_destroy(v);
  return 0;
}
