/*
 * state_layout.h — how the per-pixel ADΔER state lives in HBM (shared by host and device code).
 *
 * The reference keeps an AoS `Array3<PixelArena>` (video.rs:325), each arena a header plus a
 * SmallVec of 22-byte packed nodes (event_pixel_tree.rs:41-66).  Here the same information is a
 * structure of arrays indexed by the flat raster index i = (y*W + x)*C + c, so that a warp touching
 * 32 consecutive pixels reads 32 consecutive 8-byte headers / 16-byte nodes:
 *
 *   hdr   : uint2[P]        .x = last_fired_t (f32 bits)                       event_pixel_tree.rs:56
 *                           .y = base_val | c_thresh<<8 | c_increase_counter<<16 | length<<24
 *                                | dtm_reached<<29 | popped_dtm<<30              :58-65
 *   nodes : uint4[ceil(K/2)][Ppad][2]   the node stacks, two levels to a 32-byte record        :41-49
 *                           levels 2j and 2j+1 of pixel i are the two halves of record (j, i): a pixel's root and its
 *                           first child arrive in one 256-bit access, and every record is one DRAM sector that belongs to
 *                           ONE pixel (with level-major 16-byte nodes two neighbouring pixels shared a sector, and every
 *                           level only some pixels reach cost up to twice its bytes: profiles/r02d, DRAM / counted 1.43-1.46)
 *                           .x = integration (f32)  .y = delta_t (f32)  .z = best_event.delta_t (f32)
 *                           .w = d | best_event.d<<8 | best_event.is_some()<<16
 *   running : u8[P]         VideoState.running_intensities                      video.rs:212
 *
 * Not stored per pixel, with the reason:
 *   running_t       identical for every pixel under integrate_matrix (all get `time_spanned` each
 *                   frame, event_pixel_tree.rs:337) -> one f32 on the host, advanced with the same
 *                   f32 add per frame.
 *   time_mode       only ever written for all pixels at once (video.rs:500-502, :632-634) -> scalar.
 *   need_to_pop_top transient inside integrate_for_px: video.rs:1371-1374 pops whenever it is set,
 *                   and pop_top_event clears it (event_pixel_tree.rs:152), so it is never live
 *                   between frames.  The entry check at video.rs:1329 can therefore not fire.
 *   coord           derived from i.
 *   alt             debug-assert only (:45).
 */
#ifndef ADDER_B200_STATE_LAYOUT_H
#define ADDER_B200_STATE_LAYOUT_H

#include <stdint.h>

#define ADDER_TILE_PX 256u        /* pixels per tile = threads per CTA */
/* uint4 index of level k of pixel i (ppad = padded pixel count): record (k/2, i), half k%2 */
#define NODE_SLOT(k, i, ppad) ((((unsigned long long)((k) >> 1)) * (ppad) + (i)) * 2ull + ((k) & 1u))
#define NODE_LEVELS_ALLOC(depth) ((((depth) + 1u) >> 1) << 1) /* levels the allocation holds for a depth */
#define ADDER_MAX_DEPTH 31u       /* reference iteration guard, event_pixel_tree.rs:387 */

#define HDR_BASE(y) ((y) & 0xFFu)
#define HDR_CTHRESH(y) (((y) >> 8) & 0xFFu)
#define HDR_COUNTER(y) (((y) >> 16) & 0xFFu)
#define HDR_LENGTH(y) (((y) >> 24) & 0x1Fu)
#define HDR_DTM_REACHED(y) (((y) >> 29) & 1u)
#define HDR_POPPED(y) (((y) >> 30) & 1u)
#define HDR_PACK(base, c, cnt, len, dtmr, popped) \
  ((uint32_t)(base) | ((uint32_t)(c) << 8) | ((uint32_t)(cnt) << 16) | ((uint32_t)(len) << 24) | \
   ((uint32_t)(dtmr) << 29) | ((uint32_t)(popped) << 30))

#define NODE_D(w) ((w) & 0xFFu)
#define NODE_BEST_D(w) (((w) >> 8) & 0xFFu)
#define NODE_HAS_BEST(w) (((w) >> 16) & 1u)
#define NODE_PACK(d, best_d, has_best) ((uint32_t)(d) | ((uint32_t)(best_d) << 8) | ((uint32_t)(has_best) << 16))

/* error bits raised by kernels (device word, OR-ed) */
#define ADDER_DEVERR_CAPACITY 1u /* event buffer too small: records beyond capacity were not written */
#define ADDER_DEVERR_DEPTH 2u    /* node stack would exceed the allocated depth */
#define ADDER_DEVERR_INTERNAL 4u /* state invariant violated (root popped with no child) */

#endif
