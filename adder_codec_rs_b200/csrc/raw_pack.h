/*
 * raw_pack.h — four event records -> their raw .adder wire bytes, on registers (RawOutput::ingest_event,
 * adder-codec-core/src/codec/raw/stream.rs:100-120: bincode fixint, big-endian; EventSingle 9 bytes, Event with
 * c: Some(u8) 11 bytes, codec/header.rs:77-81).  Shared by raw_encode_kernel and the host tests (tests/host_sim).
 */
#ifndef ADDER_B200_RAW_PACK_H
#define ADDER_B200_RAW_PACK_H

#include <stdint.h>

#if defined(__CUDACC__)
#define ADDER_RAW_HD __host__ __device__ __forceinline__
#else
#define ADDER_RAW_HD inline
#endif

namespace adder {

/* Four records = 48 bytes in, 36 (single channel) or 44 bytes out: whole 32-bit words on both sides, so a thread works on
 * registers only.  Record e of the four starts at byte 9e / 11e of the thread's output, i.e. at a byte shift that is a
 * compile-time constant: its bytes, packed into three words in wire order, are shifted into place. */
template <uint32_t ESIZE>
ADDER_RAW_HD void raw_pack4(const uint32_t* in /* 12 words */, uint32_t* o /* ESIZE words, zeroed */) {
#pragma unroll
  for (uint32_t e = 0; e < 4u; e++) {
    const uint32_t w0 = in[3u * e], w1 = in[3u * e + 1u], t = in[3u * e + 2u];
    const uint32_t c = w1 & 0xFFu, d = (w1 >> 8) & 0xFFu;
    /* memory order of a little-endian word: byte 0 is the low byte.  x and y go out big-endian (bincode fixint, big-endian options) */
    const uint32_t p0 = ((w0 >> 8) & 0x00FF00FFu) | ((w0 << 8) & 0xFF00FF00u); /* xh xl yh yl */
    uint32_t p1, p2;
    if (ESIZE == 11u) {
      p1 = 1u | (c << 8) | (d << 16) | ((t >> 24) << 24);                 /* 1 c d t3 */
      p2 = ((t >> 16) & 0xFFu) | (((t >> 8) & 0xFFu) << 8) | ((t & 0xFFu) << 16); /* t2 t1 t0 - */
    } else {
      p1 = d | ((t >> 24) << 8) | (((t >> 16) & 0xFFu) << 16) | (((t >> 8) & 0xFFu) << 24); /* d t3 t2 t1 */
      p2 = t & 0xFFu;                                                                  /* t0 - - - */
    }
    const uint32_t off = ESIZE * e, wi = off >> 2, sh = (off & 3u) * 8u;
    if (sh == 0u) {
      o[wi] |= p0;
      o[wi + 1u] |= p1;
      o[wi + 2u] |= p2;
    } else {
      o[wi] |= p0 << sh;
      o[wi + 1u] |= (p0 >> (32u - sh)) | (p1 << sh);
      o[wi + 2u] |= (p1 >> (32u - sh)) | (p2 << sh);
      if (wi + 3u < ESIZE) o[wi + 3u] |= p2 >> (32u - sh);
    }
  }
}

}  // namespace adder

#endif
