/*
 * scpd.h -- C ABI of the B200-native batched successive-cancellation polar decoder.
 *
 * Drop-in boundary for the decode path of ydelomier/SC_Polar_decoder_HLS.  The reference has
 * no function-call API: its decoder is SC_MODULE(my_module) (src/module/my_module.h:15-46)
 * wired between wrapper_in / wrapper_out by sc_top_module (src/testbench/sc_top_module.h:146-160),
 * and configured by macros (src/module/config.h, src/module/polar_parameters.h).  Each entry
 * point below names the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions kept from the reference:
 *   - LLRs: one signed 8-bit value per code bit, natural index order, positive <=> bit 0
 *     (sc_bpsk.h:53), produced by the quantiser clamp(trunc(4y), -31, 31) (sc_quantizer.h:77-80);
 *     any value in [-(2^(Q-1)-1), 2^(Q-1)-1] is accepted (Q = llr_bits).  With llr_bits = 5 the
 *     alphabet does not fit and the reference truncates it to 5 bits on the sc_fifo<LLR> write
 *     (wrapper_in.h:33-34); that configuration reproduces the truncation for any int8 input.
 *   - frozen table: N flags, 1 = information bit, 0 = frozen (Writer.h:86-93; my_module.h:92-95).
 *   - output: the estimated CODEWORD x^ (what my_module streams out of bit_mem_1,
 *     my_module.h:1859-1866), packed LSB-first: bit i of a frame is bit (i % 32) of word i / 32
 *     (the PAR-bit words of wrapper_out.h:31-33 laid end to end).
 *   - arithmetic: bit-exact with the reference for (format, llr_bits, par, extended):
 *     shared/src/functions.h:48-347 (f, g, h), :354-866 (PAR-wide leaf decoders).
 *
 * All pointers named d_* are DEVICE pointers; h_* are host pointers.  Calls taking a stream are
 * asynchronous on that stream.  A handle is bound to one device and is not thread-safe.
 * There is no CPU fallback: every compute entry point fails with SCPD_E_CUDA if no device works.
 */
#ifndef SCPD_H
#define SCPD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCPD_VERSION 100

typedef enum {
    SCPD_OK = 0,
    SCPD_E_ARG = 1,         /* null pointer / bad size */
    SCPD_E_CONFIG = 2,      /* N not a power of two, 2*PAR > N, K != popcount(flags), ... */
    SCPD_E_UNSUPPORTED = 3, /* beyond the kernels: n above 2^20 or par above 512 */
    SCPD_E_IO = 4,          /* file missing / unparsable */
    SCPD_E_CUDA = 5,        /* CUDA runtime error (scpd_last_error() has the text) */
    SCPD_E_NOMEM = 6
} scpd_status;

typedef enum { SCPD_FMT_CA2 = 0, SCPD_FMT_SIGMAG = 1 } scpd_format; /* config.h:11 CA2 / SIGMAG */

/* Pruning.  Modes 0..2: which proven-bit-identical node shortcuts the schedule may use.  None of them changes
 * a single output bit w.r.t. plain SC (reference PRUNING_LEVEL 0); they only skip work.
 * Mode 3 is a DIFFERENT DECODER, opt-in: the reference as it is checked in (config.h:16-30: PRUNING_LEVEL 2 with
 * ELAG_R1 / ELAG_REP / ELAG_SPC / ELAG_H0).  Above the PAR-wide leaf a child whose words are all-frozen but a last
 * repetition word is decided by the sign of a saturating sum (my_module.h:1292-1390), an all-information child by the
 * sign of g (:1571-1642), a child that is one single-parity-check word followed by all-information words by the
 * signs of g with the least reliable one flipped on odd parity (:1737-1842).  Its output differs from plain SC on
 * noisy frames (as the reference's own does, tests/test_oracle.py); it is bit-exact with the reference built that way. */
typedef enum {
    SCPD_PRUNE_NONE = 0,  /* visit every node, as PRUNING_LEVEL 0 (config.h:16) */
    SCPD_PRUNE_R0 = 1,    /* skip all-frozen subtrees */
    SCPD_PRUNE_R0_R1 = 2, /* + all-information subtrees by hard decision when no input LLR is a
                             CA2 zero (falls back to plain SC on that node otherwise) */
    SCPD_PRUNE_REF_LEVEL2 = 3 /* the reference's PRUNING_LEVEL 2 decoder (R0 / R1 / REP / SPC / H0); par 2..256 */
} scpd_pruning;

/* Replaces the compile-time macros of config.h:2-16 and polar_parameters.h:4-11. */
typedef struct {
    uint32_t n;        /* _NBITS                                  */
    uint32_t k;        /* number of information bits (checked)    */
    uint32_t par;      /* PAR: width of the leaf decoder          */
    uint32_t llr_bits; /* LLR_BITS                                */
    uint32_t format;   /* scpd_format                             */
    uint32_t extended; /* EXTENDED                                */
    uint32_t pruning;  /* scpd_pruning                            */
    uint32_t reserved;
} scpd_config;

typedef struct scpd_decoder scpd_decoder; /* opaque; one per (device, configuration, frozen set) */

/* ---- frozen-bit tables: Frozen_Bit_Generator/src/Writer.h:21-171 ---- */
/* Frozen_Bit_Tab/FB_N*_K*.txt: "N\r\n0\r\n0\r\n" then channel indices, most reliable first
 * (Writer.h:35-58).  Indices >= n are dropped, the first k survivors are information bits
 * (Writer.h:64-93).  flags_out: n bytes. */
int scpd_frozen_load_order(const char* path, uint32_t n, uint32_t k, uint8_t* flags_out);
/* Generated_Frozen_Bit/frozen_n_*_k_*.txt: one line of n tokens 0|1 (Writer.h:95-105).
 * k_out (optional) receives the number of ones. */
int scpd_frozen_load_flags(const char* path, uint32_t n, uint8_t* flags_out, uint32_t* k_out);
/* The "affect" file Writer.h:75-80 writes back: order restricted to indices < n. */
int scpd_frozen_write_order(const char* path, uint32_t n, const uint32_t* order);
/* Single-line flag file in the Generated_Frozen_Bit format. */
int scpd_frozen_write_flags(const char* path, uint32_t n, const uint8_t* flags);
/* polar_parameters.h exactly as Writer.h:110-162 emits it (en = "En" argument, main.cpp:29),
 * so the HLS flow can still consume tables loaded here. */
int scpd_write_polar_parameters(const char* path, uint32_t n, uint32_t par, int en,
                                const uint8_t* flags);

/* ---- decoder: SC_MODULE(my_module), my_module.h:15-46 ---- */
/* FB port + do_prunning (my_module.h:61-166): ingest the table, build the device schedule. */
int scpd_create(const scpd_config* cfg, const uint8_t* h_info_flags, int device,
                scpd_decoder** out);
void scpd_destroy(scpd_decoder* dec);
/* e -> s ports (my_module.h:34-35): decode nframes frames.
 * d_llr  : [nframes][n] int8, device.      d_xhat : [nframes][n/32] uint32, device.
 * Large batches run on the bit-sliced kernel (32 frames per register bit; the handle then keeps a
 * bit-plane copy of the batch, nframes * n bytes, plus a workspace for the resident frame groups); batches
 * too small to fill the GPU that way run on the int16x2 kernel.  Results are identical either way.
 * Configurations outside those two datapaths (llr_bits 5, SIGMAG combinations without a bit-sliced
 * instantiation, n < 128 in SIGMAG, EXTENDED leaves wider than 16 bits) run on the raw-pattern kernel,
 * which models the reference's W-bit wrap-around exactly and is slower. */
int scpd_decode(scpd_decoder* dec, const int8_t* d_llr, size_t nframes, uint32_t* d_xhat,
                void* cuda_stream);
/* Same through host buffers (pinned or pageable): the batch is cut into chunks that flow through an
 * H2D copy / decode / D2H copy pipeline on three streams; returns when h_xhat is complete.
 * This is the call a host-only caller such as the reference testbench would make. */
int scpd_decode_host(scpd_decoder* dec, const int8_t* h_llr, size_t nframes, uint32_t* h_xhat);
/* Input contract of scpd_decode / scpd_decode_host: |llr| <= 2^(llr_bits-1) - 1.  The reference's quantiser alphabet is
 * +-31 whatever LLR_BITS is (main.cpp:16-18); wider values up to the internal saturation are accepted.  Values outside
 * that range (and -128) are NOT checked by the decode calls and give unspecified bits (the reference wraps them modulo
 * 2^LLR_BITS on the sc_fifo<LLR> write, wrapper_in.h:33-34; only llr_bits = 5, where the +-31 alphabet itself wraps, is
 * reproduced -- by the raw-pattern kernel).  scpd_validate_llr counts the offending values of a device batch
 * (synchronous on `cuda_stream`). */
int scpd_validate_llr(scpd_decoder* d, const int8_t* d_llr, size_t nframes, uint64_t* h_out_of_range, void* cuda_stream);
/* Information-bit estimate u^ = x^ * F^(x)n (extra; the reference outputs x^ only). */
int scpd_extract_info(scpd_decoder* dec, const uint32_t* d_xhat, size_t nframes, uint32_t* d_uhat,
                      void* cuda_stream);
/* Introspection. */
int scpd_get_config(const scpd_decoder* dec, scpd_config* out);
/* Number of schedule operations and of f/g element updates per frame after pruning. */
int scpd_schedule_stats(const scpd_decoder* dec, uint64_t* n_ops, uint64_t* n_fg_updates);
/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
uint64_t scpd_launch_count(const scpd_decoder* dec);
/* Measurement aids (the reference's sc_monitor.h:50-441 counts cycles per FSM state; here the unit is
 * the decode kernel): with timing enabled every scpd_decode brackets its tree-walk kernel with CUDA
 * events on the caller's stream; scpd_last_kernel_ms waits for the last one and returns its duration. */
int scpd_kernel_timing(scpd_decoder* dec, int enable);
int scpd_last_kernel_ms(scpd_decoder* dec, float* ms);
/* Which kernel family and layout this handle launches (static text, valid until the next call). */
const char* scpd_kernel_name(const scpd_decoder* dec);
/* The kernel the last scpd_decode of this handle launched (scpd_decode picks by batch size: bit-sliced for large
 * batches, int16x2 with 8 / 16 / 32 lanes or a whole CTA per frame pair below); "" before the first decode. */
const char* scpd_last_kernel_name(const scpd_decoder* dec);
/* The function x level matrix of sc_monitor (src/rtl_simu_testbench/sc_monitor/sc_monitor.h:50-441), from
 * the frozen table alone (host only, no device needed): per FSM function and node size 2^l, the node visits
 * per frame and the trip counts of the reference's pipelined loops (f_loop / g_loop / h_loop,
 * my_module.h:373,704,903: half the node's PAR-wide words; one per leaf decode R, :592-595).  With
 * SCPD_PRUNE_NONE this is the reference at PRUNING_LEVEL 0 and total_iterations its loop cycles per frame;
 * with pruning it is the walk this library executes (R0 / R1: nodes skipped / decided by hard decision); with
 * SCPD_PRUNE_REF_LEVEL2 the reference FSM at PRUNING_LEVEL 2: F / G / H / R as above, R0 = children skipped (H0 path),
 * R1 = G_R1_STATE, REP = F_REP_STATE, SPC = G_SPC_STATE (their loops run over the child's words). */
enum { SCPD_STAGE_F = 0, SCPD_STAGE_G = 1, SCPD_STAGE_H = 2, SCPD_STAGE_R = 3, SCPD_STAGE_R0 = 4,
       SCPD_STAGE_R1 = 5, SCPD_STAGE_REP = 6, SCPD_STAGE_SPC = 7, SCPD_STAGE_FUNCS = 8 };
typedef struct {
    uint64_t visits[SCPD_STAGE_FUNCS][32];     /* [function][l]: node visits per frame            */
    uint64_t iterations[SCPD_STAGE_FUNCS][32]; /* [function][l]: loop iterations (PAR-wide words) */
    uint64_t total_iterations;
} scpd_stage_matrix;
int scpd_stage_profile(const scpd_config* cfg, const uint8_t* h_info_flags, scpd_stage_matrix* out);
/* The measured counterpart (sc_monitor.h:50-441 histograms cycles, not trips): scpd_stage_timing(d, 1) makes every later
 * scpd_decode on the slot-sliced kernel add, for one warp (32 frames) per CTA, the SM-clock cycles between the fetch of a
 * schedule op and the fetch of the next to cycles[function][level] (rows as SCPD_STAGE_*: F, G, H, R = a whole 64-LLR
 * node incl. the level-7 f / g fused in front of it, R0, R1 = hard decision) and 1 to visits[function][level];
 * scpd_stage_time reads and clears the histogram.  SCPD_E_UNSUPPORTED for handles without that kernel. */
int scpd_stage_timing(scpd_decoder* d, int enable);
int scpd_stage_time(scpd_decoder* d, uint64_t cycles[6][32], uint64_t visits[6][32]);
/* With the same switch on: how often the profiled warps reached an all-information (R1) node of 2^l LLRs and how often a
 * zero LLR in one of the warp's 32 frames sent the node to the full walk instead of the hard-decision shortcut (CA2:
 * hd(0) = 0 whatever the sign bit says, functions.h:48-70 / SURVEY G3).  What scpd_schedule_stats reports as pruned is
 * the best case; this is the measured share.  Reads and clears. */
int scpd_r1_votes(scpd_decoder* d, uint64_t votes[32], uint64_t fallbacks[32]);

/* ---- testbench harness on the device: src/testbench/ ---- */
/* sigma = 1/sqrt(2 R 10^(EbN0/10)), main.cpp:91-98 (the reference hard-codes R = 0.5). */
float scpd_sigma(float ebn0_db, float rate);
/* sc_xorshift128 (two streams, seed byte) -> sc_awgn (Box-Muller, 2 samples per draw) ->
 * sc_bpsk + sc_adder -> sc_quantizer.  Frame f of the stream uses draws [f*n/2, (f+1)*n/2) of
 * each generator, reached by GF(2) jump-ahead.  d_codeword: n bytes (0/1) shared by all frames
 * if per_frame == 0, [nframes][n] if 1, packed rows [nframes][n/32] (LSB first) if 2, or NULL for the all-zero
 * codeword (sc_encoder.h:105-110).
 * d_llr: [nframes][n] int8. */
int scpd_channel_generate(uint32_t n, uint64_t first_frame, size_t nframes, uint8_t seed,
                          float sigma, const uint8_t* d_codeword, int per_frame, int8_t* d_llr,
                          void* cuda_stream);
/* How the Box-Muller step of the channel (sc_awgn.h:61-77) is evaluated, process-wide; returns the previous mode.
 * 0: logf / sqrtf / sincosf as libm-grade code.  2 (default): SFU approximations, with the libm-grade code re-run for
 * every sample that lands within a guard band of a quantiser bin edge -- the same LLRs at about half the instructions
 * (the band is 16x the largest deviation found over 8e10 draws, tools/probe/chan_err.cu).  1: approximations only (about
 * 1e-5 of the LLRs differ by one quantiser step).  mode < 0 only reads the current mode. */
int scpd_channel_mode(int mode);
/* sc_error_counter.h:68-125.  d_ref_words: packed reference codeword(s) ([n/32] shared or
 * [nframes][n/32]) or NULL for all-zero.  d_counters (6 x uint64, accumulated, caller zeroes):
 * [0] bit errors [1] frame errors [2] bits [3] frames [4],[5] as [0],[1] with the reference's
 * 10-bit per-frame wrap (sc_uint<10> err, :70-71). */
int scpd_count_errors(uint32_t n, size_t nframes, const uint32_t* d_xhat,
                      const uint32_t* d_ref_words, int per_frame, uint64_t* d_counters,
                      void* cuda_stream);
/* Whole Monte-Carlo loop of src/testbench/main.cpp on the device: generate -> decode -> count,
 * nframes frames starting at stream position first_frame, in batches; h_counters as above. */
int scpd_run_ber(scpd_decoder* dec, float ebn0_db, float rate, uint64_t first_frame,
                 uint64_t nframes, uint8_t seed, const uint8_t* h_codeword, uint64_t h_counters[6]);
/* The same loop with the other codeword sources and information-bit counters.
 *   SCPD_SRC_CODEWORDS: h_codewords = ncw codewords of n bytes (0/1); frame g of the stream sends codeword g % ncw --
 *     ncw = 3 with cw512x256 / cw1024x512 / cw8x4 is the reference's sc_encoder (sc_encoder.h:91-113); NULL / 0 = all-zero.
 *   SCPD_SRC_RANDOM: K random information bits per frame (counter-based generator keyed by payload_seed and the frame
 *     index; the reference has no payload source), encoded on the device: x = u F^(x)n, natural order.
 * counters[0..5] as scpd_count_errors (codeword bits: what sc_error_counter.h:68-125 counts); counters[6..9] on the
 * information positions of u^ = x^ F^(x)n: bit errors, frames in error, bits (k per frame), frames. */
enum { SCPD_SRC_CODEWORDS = 0, SCPD_SRC_RANDOM = 1 };
int scpd_run_ber_ex(scpd_decoder* d, float ebn0_db, float rate, uint64_t first_frame, uint64_t nframes, uint8_t seed,
                    int src_mode, const uint8_t* h_codewords, uint32_t ncw, uint64_t payload_seed, uint64_t h_counters[10]);

const char* scpd_last_error(void);
const char* scpd_status_string(int status);

#ifdef __cplusplus
}
#endif
#endif /* SCPD_H */
