/* TEST INFRASTRUCTURE ONLY -- bit writer, CABAC encoder (H.265 9.3.4) and residual_coding
 * syntax (7.3.8.11) of the CPU oracle.  See hevc_tables.h for scope and pinning status. */
#ifndef ORACLE_HEVC_CABAC_H_
#define ORACLE_HEVC_CABAC_H_
#include <stddef.h>
#include <stdint.h>
#include "hevc_tables.h"

typedef struct {
  uint8_t *buf;
  size_t cap, pos;       /* bytes written */
  uint32_t cur;          /* bit accumulator */
  int nbits;             /* bits held in cur (0..7 after flush of whole bytes) */
  int overflow;
} orc_bits_t;

void orc_bits_init(orc_bits_t *b, uint8_t *buf, size_t cap);
void orc_bits_put(orc_bits_t *b, uint32_t val, int n);      /* n <= 32, MSB first */
void orc_bits_ue(orc_bits_t *b, uint32_t v);
void orc_bits_se(orc_bits_t *b, int32_t v);
void orc_bits_trailing(orc_bits_t *b);                      /* rbsp_trailing_bits / byte_alignment */
size_t orc_bits_bytes(const orc_bits_t *b);                 /* requires byte alignment */

/* Escape an RBSP into a NAL payload (7.4.2: emulation_prevention_three_byte). Returns bytes written. */
size_t orc_nal_escape(const uint8_t *rbsp, size_t n, uint8_t *out, size_t cap);

typedef struct {
  uint32_t low, range;
  int bits_left, num_buffered, buffered_byte;
  orc_bits_t *out;
  uint8_t ctx[CTX_COUNT];     /* (pStateIdx << 1) | valMps */
  uint64_t bins;              /* statistics: number of bins coded */
} orc_cabac_t;

void orc_cabac_init_contexts(orc_cabac_t *c, int init_type, int slice_qp);
void orc_cabac_start(orc_cabac_t *c, orc_bits_t *out);
void orc_cabac_bin(orc_cabac_t *c, int ctx_idx, int bin);
void orc_cabac_bypass(orc_cabac_t *c, int bin);
void orc_cabac_bypass_bits(orc_cabac_t *c, uint32_t bins, int n);
void orc_cabac_terminate(orc_cabac_t *c, int bin);
void orc_cabac_finish(orc_cabac_t *c);      /* flush + the stop/alignment bit + zero bits to the byte */

/* residual_coding( ) for one transform block.  levels: row-major N x N with the given stride.
 * scan_idx: 0 diagonal, 1 horizontal, 2 vertical (7.4.9.11). */
void orc_code_residual(orc_cabac_t *c, const int16_t *levels, int stride, int log2n, int cidx, int scan_idx);
/* ... with sign_data_hiding_enabled_flag = `sign_hiding` (the levels must already obey the parity rule) */
void orc_code_residual2(orc_cabac_t *c, const int16_t *levels, int stride, int log2n, int cidx, int scan_idx, int sign_hiding);

#endif
