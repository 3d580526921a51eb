/*
 * sc_oracle.h -- CPU restatement of the reference SC polar decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this library, and only as the checker / reported CPU baseline.
 *
 * Parity status: PINNED.  Checked (tests/test_oracle*.py) against
 *   (i)  the 9 golden codewords of src/testbench/sc_encoder/sc_encoder.h:74-89,
 *   (ii) the reference's own headers (shared/src/{scalar,vector,functions,library}.h
 *        and src/module/my_module.h) compiled unmodified on the local systemc.h
 *        shim (oracle/shim/, built into oracle/_ref/ by oracle/Makefile), on
 *        random and channel LLRs for every supported (N, PAR, Q, format, EXTENDED).
 *
 * All citations are relative to /root/reference.
 */
#ifndef SC_ORACLE_H
#define SC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCO_CA2 0    /* config.h:11  "#define CA2"    two's complement          */
#define SCO_SIGMAG 1 /* config.h:11  "#define SIGMAG" sign + magnitude, has -0  */

typedef struct {
    int32_t n;        /* _NBITS   (polar_parameters.h:4), power of two            */
    int32_t par;      /* PAR      (polar_parameters.h:8), power of two, 2*par<=n  */
    int32_t llr_bits; /* LLR_BITS (config.h:2), 3..12                             */
    int32_t format;   /* SCO_CA2 / SCO_SIGMAG                                      */
    int32_t extended; /* EXTENDED (config.h:14)                                    */
} sco_config;

/* --- element primitives on raw width-w bit patterns (returned zero-extended) ---
 * A pattern is the low w bits of the sc_bigint<w>/sc_biguint<w> the reference holds. */
uint32_t sco_input(int format, int q, int llr);                       /* wrapper_in.h:33-34 */
uint32_t sco_f(int format, int w, uint32_t a, uint32_t b);             /* functions.h:48-61,124-145 */
uint32_t sco_g(int format, int w, uint32_t a, uint32_t b, int s);      /* functions.h:63-75,147-195 */
uint32_t sco_g_ext(int format, int w, uint32_t a, uint32_t b, int s);  /* functions.h:77-88,197-239 (w+1 bits) */
int sco_f_simp(int format, int w, uint32_t a, uint32_t b, int fb);     /* functions.h:90-101,241-255 */
int sco_g_simp(int format, int w, uint32_t a, uint32_t b, int s, int fb); /* functions.h:103-118,257-281 */
int sco_sign(int w, uint32_t a);                                       /* scalar.h:23-28,101-106 */
int sco_value(int format, int w, uint32_t a); /* numeric value of a pattern (for printing/tests) */

/* Leaf decoder Spec_Polar_Decoder<P,Q> (library.h:149-172, functions.h:354-866):
 * p patterns of width w in, p partial-sum bits out (one byte each). */
void sco_leaf(int format, int extended, int p, int w, const uint32_t* llr,
              const uint8_t* info_flags, uint8_t* bits);

/* Decode nframes frames.  llr: [nframes][n] int8, natural order, as delivered by
 * sc_quantizer (sc_quantizer.h:77-80) before wrapper_in.  info_flags: n bytes, 1 = information
 * bit (Writer.h:86-93).  xhat: [nframes][n] bytes 0/1 = estimated CODEWORD bits, i.e. the
 * stream wrapper_out emits (my_module.h:1859-1866, wrapper_out.h:31-33).
 * Returns 0, or a negative code for an invalid configuration. */
int sco_decode(const sco_config* cfg, const uint8_t* info_flags, const int8_t* llr,
               size_t nframes, uint8_t* xhat);
/* Same, output packed LSB-first: word i/32 bit i%32 of frame f at xhat32[f*n/32 + i/32]. */
int sco_decode_packed(const sco_config* cfg, const uint8_t* info_flags, const int8_t* llr,
                      size_t nframes, uint32_t* xhat32);
/* Multi-threaded variant used for the CPU baseline (frames split statically). */
int sco_decode_packed_mt(const sco_config* cfg, const uint8_t* info_flags, const int8_t* llr,
                         size_t nframes, uint32_t* xhat32, int nthreads);

/* The reference built with PRUNING_LEVEL 2 (REP / SPC / R1 / R0 shortcuts above the leaf word; config.h flags
 * ELAG_R1 = ELAG_REP = ELAG_SPC = ELAG_H0 = 1, the rest 0).  Not plain SC: see sc_oracle.c. */
int sco_decode_l2(const sco_config* cfg, const uint8_t* info_flags, const int8_t* llr,
                  size_t nframes, uint8_t* xhat);
int sco_decode_l2_packed(const sco_config* cfg, const uint8_t* info_flags, const int8_t* llr,
                         size_t nframes, uint32_t* xhat32);

/* Instrumentation for design studies (tests only): for each maximal all-information subtree,
 * counts visits and visits whose input LLRs contain a zero.  sizes indexed by log2(node size). */
typedef struct {
    uint64_t r1_visits[24];
    uint64_t r1_with_zero[24];
} sco_stats;
int sco_decode_stats(const sco_config* cfg, const uint8_t* info_flags, const int8_t* llr,
                     size_t nframes, uint8_t* xhat, sco_stats* st);

/* --- u = x * F^{(x)n}: natural-order polar transform (its own inverse).  Used to check the
 * golden codewords against the frozen sets and for the encoder extra. --- */
void sco_polar_transform(uint8_t* bits, int n);

/* --- channel chain of src/testbench (App. C of SURVEY.md) --- */
typedef struct {
    uint32_t x, y, z, w;
} sco_xs128;
void sco_xs128_seed(sco_xs128* a, sco_xs128* b, uint8_t seed); /* sc_xorshift128.h:61-66,99-104 */
uint32_t sco_xs128_next(sco_xs128* s);                          /* sc_xorshift128.h:78-85 */
float sco_xs128_uniform(uint32_t w);                            /* sc_xorshift128.h:86 */
float sco_sigma(float ebn0_db, float rate);                     /* main.cpp:91-98 */
/* Generate quantised LLRs for nframes frames starting at frame index frame0 of the stream.
 * Frame f uses draws [f*n/2, (f+1)*n/2) of both xorshift streams (two Gaussian samples per
 * draw: Ph then Qu, sc_awgn.h:66-88).  codeword: n bytes 0/1 per frame if per_frame_cw, else
 * one shared codeword, or NULL for all-zero.  llr out: [nframes][n] int8 in [-31,31]. */
void sco_channel(int n, size_t frame0, size_t nframes, uint8_t seed, float sigma,
                 const uint8_t* codeword, int per_frame_cw, int8_t* llr);
/* Quantiser alone (sc_quantizer.h:77-80). */
int sco_quantize(float y);

/* Error counter (sc_error_counter.h:68-125).  counters: [0]=bit errors, [1]=frame errors,
 * [2]=bits, [3]=frames, [4]=bit errors with the reference's 10-bit wrap (G9),
 * [5]=frame errors with the wrap. */
void sco_count_errors(int n, size_t nframes, const uint8_t* xhat, const uint8_t* ref,
                      int per_frame_ref, uint64_t counters[6]);

#ifdef __cplusplus
}
#endif
#endif
