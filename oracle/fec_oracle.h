/* oracle/fec_oracle.h -- CPU restatement of the reference's FEC hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library, and
 * only as the checker.  The product (viterbi.dll_b200/csrc) never links it.
 *
 * Parity status: PINNED -- checked against the reference's own implementation
 * compiled unmodified (oracle/_ref, see oracle/Makefile) and against the
 * known-answer vectors recorded from it (tests/golden/kat.json).  The
 * reference repository itself holds no golden vectors.
 */
#ifndef FEC_ORACLE_H
#define FEC_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* deconvolve() of deconvolve.cpp:551-554 (decon_avx2 flavour, 514-526).
 * piData: 4*(framebits+6) words, low byte = soft symbol.  Returns 0, or -2 when
 * framebits is odd or > 9216 (the reference reads uninitialised stack / overruns
 * its decision array there, deconvolve.cpp:126-127 -- not defined behaviour). */
int oracle_deconvolve(unsigned framebits, const uint32_t* piData, int inputLength, uint8_t* output);

/* same decoder, one byte per symbol */
int oracle_deconvolve_u8(unsigned framebits, const uint8_t* syms, uint8_t* output);

/* n frames [n][4*(F+6)] u8 -> [n][ceil(F/8)], sharded over nthreads pthreads */
int oracle_deconvolve_batch_u8(unsigned framebits, const uint8_t* syms, size_t n, uint8_t* out, int nthreads);

/* RScheckSuperframe() of rschecksf.cpp:65-93 incl. the partial-write rule. */
int oracle_rs_check_superframe(const uint8_t* p, int startIx, unsigned rsdims, uint8_t* outVector);

/* n superframes [n][120*s] -> out [n][110*s] (caller pre-fills), ret[n] */
int oracle_rs_check_superframe_batch(const uint8_t* in, unsigned s, size_t n, uint8_t* out, int32_t* ret,
                                     int nthreads);

/* DECODE_RS() of rschecksf.cpp:199-377 on one 120-byte codeword, in place. */
int oracle_rs_decode_codeword(uint8_t cw[120]);

/* GF(256) tables as built by dllmain.cpp:124-150: ato[768], iof[256]. */
void oracle_rs_tables(uint8_t ato_mod[768], uint8_t index_of[256]);

#ifdef __cplusplus
}
#endif
#endif
