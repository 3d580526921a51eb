// decode_raw.cuh -- the "every configuration" SC decoder: raw W-bit patterns, one frame per lane group.
//
// The bit-sliced and int16x2 kernels cover the configurations that matter for throughput and assume
// in-range values.  This kernel covers the rest of what the reference can be configured to
// (config.h:2-30): SIGMAG for any (LLR_BITS, PAR, EXTENDED), N below 128, LLR_BITS = 5 where the
// +-31 quantiser alphabet wraps modulo 2^5 on the sc_fifo<LLR> write (wrapper_in.h:33-34), and
// EXTENDED leaves wider than 16 bits.  Every value is kept as the W-bit pattern an
// sc_bigint<W> / sc_biguint<W> would hold, and every operation truncates the way the SystemC
// assignment does, so wrap-around is reproduced, not only the in-range behaviour.
//
// G lanes decode ONE frame (32/G frames per warp, all in lock step on the shared schedule of
// schedule.h).  Only all-frozen nodes are pruned here (always identical: every decision of the
// subtree is "& 0"); the rate-1 shortcut is not used because its proof needs in-range inputs.
//
// It is also the kernel of the reference-pruning mode (SCPD_PRUNE_REF_LEVEL2): the decoder the reference
// is when built with its checked-in PRUNING_LEVEL 2 -- repetition children decided by the sign of a
// saturating sum (OP_REP), all-information children by the sign of g (OP_GR1), single-parity-check
// children by the signs of g with the least reliable one flipped on odd parity (OP_GSPC).  That decoder
// is not plain SC; tests pin it on the reference's own sources built that way (oracle/_ref/*_pl2.so).
//
// Arithmetic contract (reference):
//   CA2     f  F_function_C2  functions.h:48-61     qabs/qmin/qsign  scalar.h:9-36
//           g  G_function_C2  functions.h:63-75     qsat             scalar.h:15-21
//              G_extended_C2  functions.h:77-88
//   SIGMAG  f  F_function_SM  functions.h:124-145   qabs_sm/qmin_sm  scalar.h:88-130
//           g  G_function_SM  functions.h:147-195   qfull_add_sub_sm scalar.h:196-225, qsat_sm :94-99
//              G_extended_SM  functions.h:197-239
//   input   Adapt_format library.h:18-28, qconv_format scalar.h:229-239
//   terminals Spec_P2 / Spec_P1 functions.h:354-384 with F/G_simplified functions.h:90-118,241-281
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "schedule.h"

namespace scpd {

struct RawParams {
    const uint32_t* sched;  // device copy of the schedule (pruning NONE or R0 only)
    const int8_t* llr;      // [nframes][n]
    uint32_t* xhat;         // [nframes][wpf]
    unsigned long long nframes;
    uint32_t n, log2n, wpf;
    uint32_t q;        // LLR_BITS
    uint32_t sigmag;   // 0 = CA2, 1 = SIGMAG
    uint32_t log2par;  // PAR-wide words (reference-pruning mode: word sums, minimum search order)
    uint32_t ls;             // alpha levels 0..ls in shared memory, ls+1..log2n-1 in the workspace
    uint32_t beta_in_smem;   // partial sums in shared memory (else workspace)
    uint32_t sm_words_per_frame;
    uint32_t* ws;            // global workspace, ws_words_per_frame per resident frame
    unsigned long long ws_words_per_frame;
};

namespace raw {
__device__ __forceinline__ uint32_t mk(int w) { return (1u << w) - 1u; }
__device__ __forceinline__ uint32_t sgn(uint32_t p, int w) { return (p >> (w - 1)) & 1u; }
// two's-complement value of a w-bit pattern
__device__ __forceinline__ int val(uint32_t p, int w) {
    const uint32_t m = 1u << (w - 1);
    return (int)(((p & mk(w)) ^ m) - m);
}
// ---- CA2
__device__ __forceinline__ uint32_t abs_c2(uint32_t a, int w) {  // qabs: -(-2^(w-1)) wraps back onto itself
    const int v = val(a, w);
    return (uint32_t)(v < 0 ? -v : v) & mk(w);
}
__device__ __forceinline__ uint32_t f_c2(uint32_t a, uint32_t b, int w) {
    const uint32_t aa = abs_c2(a, w), ab = abs_c2(b, w);
    const uint32_t mn = val(aa, w) < val(ab, w) ? aa : ab;  // qmin compares as signed
    return (sgn(a, w) ^ sgn(b, w)) ? ((uint32_t)(-val(mn, w)) & mk(w)) : mn;
}
__device__ __forceinline__ int addsub_c2(uint32_t a, uint32_t b, uint32_t u, int w) {
    return u ? val(b, w) - val(a, w) : val(b, w) + val(a, w);
}
__device__ __forceinline__ uint32_t g_c2(uint32_t a, uint32_t b, uint32_t u, int w, bool nosat) {
    int v = addsub_c2(a, b, u, w);
    if (nosat) return (uint32_t)v & mk(w + 1);
    const int m = (1 << (w - 1)) - 1;
    v = v > m ? m : (v < -m ? -m : v);
    return (uint32_t)v & mk(w);
}
// ---- SIGMAG: pattern = (sign, w-1 magnitude bits); -0 exists
__device__ __forceinline__ uint32_t f_sm(uint32_t a, uint32_t b, int w) {
    const uint32_t ma = a & mk(w - 1), mb = b & mk(w - 1);
    return ((sgn(a, w) ^ sgn(b, w)) << (w - 1)) | (ma < mb ? ma : mb);
}
// (w+1)-bit result: sign at bit w, w-bit end-around sum below it
__device__ __forceinline__ uint32_t addsub_sm(uint32_t a, uint32_t b, uint32_t u, int w) {
    const uint32_t sa = sgn(a, w) ^ (u & 1u), sb = sgn(b, w);
    const uint32_t x = sa ^ sb;
    const uint32_t ma = a & mk(w - 1), mb = b & mk(w - 1);
    const uint32_t a_smaller = ma < mb ? 1u : 0u;
    const uint32_t ta = (x & a_smaller) ? (~ma & mk(w)) : ma;
    const uint32_t tb = (x & (a_smaller ^ 1u)) ? (~mb & mk(w)) : mb;
    const uint32_t sum = (ta + tb + x) & mk(w);
    return ((a_smaller ? sb : sa) << w) | sum;
}
__device__ __forceinline__ uint32_t g_sm(uint32_t a, uint32_t b, uint32_t u, int w, bool nosat) {
    const uint32_t r = addsub_sm(a, b, u, w);
    if (nosat) return r;
    const uint32_t mag = r & mk(w), top = mk(w - 2);  // qsat_sm<w-1>: the magnitude saturates at 2^(w-2) - 1
    return (((r >> w) & 1u) << (w - 1)) | (mag > top ? top : (mag & mk(w - 1)));
}
__device__ __forceinline__ uint32_t input(int llr, int q, bool sm) {
    const uint32_t p = (uint32_t)llr & mk(q);
    if (!sm || !sgn(p, q)) return p;
    return ((~(p & mk(q - 1)) & mk(q)) + 1u) & mk(q);
}
}  // namespace raw

template <int G>
struct RawDecoder {
    const RawParams& p;
    uint32_t* sm_alpha;  // index (1 << l) + i
    uint32_t* gl_alpha;
    uint32_t* beta;
    const int8_t* llr;
    int gl;
    bool sm;
    unsigned long long ew;  // 4 bits per level l < 16: width of alpha[l] beyond LLR_BITS (inside EXTENDED leaves)

    __device__ RawDecoder(const RawParams& p_) : p(p_) {}

    __device__ __forceinline__ int extra(int l) const { return l < 16 ? (int)((ew >> (4 * l)) & 15ull) : 0; }
    __device__ __forceinline__ int width(int l) const { return (int)p.q + extra(l); }
    __device__ __forceinline__ void set_extra(int l, int e) {
        if (l < 16) ew = (ew & ~(15ull << (4 * l))) | ((unsigned long long)e << (4 * l));
    }
    __device__ __forceinline__ uint32_t* alpha(int l) const {
        return ((uint32_t)l <= p.ls ? sm_alpha : gl_alpha) + (1u << l);
    }
    __device__ __forceinline__ uint32_t ld(int l, uint32_t i) const {
        if ((uint32_t)l == p.log2n) return raw::input(llr[i], (int)p.q, sm);
        return alpha(l)[i];
    }
    __device__ void op_f(int l) {
        const uint32_t h = 1u << (l - 1);
        const int w = width(l);
        uint32_t* dst = alpha(l - 1);
        for (uint32_t i = gl; i < h; i += G) {
            const uint32_t a = ld(l, i), b = ld(l, i + h);
            dst[i] = sm ? raw::f_sm(a, b, w) : raw::f_c2(a, b, w);
        }
        set_extra(l - 1, extra(l));
        __syncwarp();
    }
    __device__ void op_g(int l, uint32_t o, bool nosat, bool zero_beta) {
        const uint32_t h = 1u << (l - 1);
        const int w = width(l);
        uint32_t* dst = alpha(l - 1);
        for (uint32_t i = gl; i < h; i += G) {
            const uint32_t pos = o + i;
            const uint32_t u = zero_beta ? 0u : (beta[pos >> 5] >> (pos & 31u)) & 1u;
            const uint32_t a = ld(l, i), b = ld(l, i + h);
            dst[i] = sm ? raw::g_sm(a, b, u, w, nosat) : raw::g_c2(a, b, u, w, nosat);
        }
        set_extra(l - 1, extra(l) + (nosat ? 1 : 0));
        __syncwarp();
    }
    // beta[o..o+h) (op)= beta[o+h..o+2h)
    __device__ void op_h(int l, uint32_t o, bool copy) {
        const uint32_t h = 1u << (l - 1);
        if (h >= 32) {
            const uint32_t w0 = o >> 5, w1 = (o + h) >> 5, nw = h >> 5;
            for (uint32_t w = gl; w < nw; w += G) beta[w0 + w] = copy ? beta[w1 + w] : (beta[w0 + w] ^ beta[w1 + w]);
        } else if (gl == 0) {
            const uint32_t w = o >> 5, lm = ((1u << h) - 1u) << (o & 31u);
            const uint32_t a = beta[w];
            beta[w] = copy ? ((a & ~lm) | ((a >> h) & lm)) : (a ^ ((a >> h) & lm));
        }
        __syncwarp();
    }
    __device__ void op_r0(int l, uint32_t o) {
        const uint32_t n = 1u << l;
        if (n >= 32) {
            for (uint32_t w = gl; w < (n >> 5); w += G) beta[(o >> 5) + w] = 0u;
        } else if (gl == 0) {
            beta[o >> 5] &= ~(((1u << n) - 1u) << (o & 31u));
        }
        __syncwarp();
    }
    // ---- reference-pruning mode.  All three ops run on a node of 2^l LLRs above the leaf (l > log2par, width q) and use
    // alpha[l-1] as scratch, as the f / g loops they replace would.
    // beta[o..o+h) = sign bits of scratch[0..h)
    __device__ void pack_signs(const uint32_t* v, uint32_t o, uint32_t h, int w) {
        if (h >= 32) {
            for (uint32_t wd = gl; wd < (h >> 5); wd += G) {
                uint32_t bits = 0;
                for (uint32_t b = 0; b < 32; b++) bits |= raw::sgn(v[32 * wd + b], w) << b;
                beta[(o >> 5) + wd] = bits;
            }
        } else if (gl == 0) {
            uint32_t bits = 0;
            for (uint32_t b = 0; b < h; b++) bits |= raw::sgn(v[b], w) << b;
            const uint32_t wd = o >> 5, sh = o & 31u;
            beta[wd] = (beta[wd] & ~(((1u << h) - 1u) << sh)) | (bits << sh);
        }
    }
    // F_REP_STATE my_module.h:1292-1390: sum = ADD_TREE_FUNCTION(f word, sum) over the words of the left child, all its
    // bits = sign(sum).  ADDER_TREE_<PAR> (functions.h:3141-3260): the word folded by halves (exact; in SIGMAG with
    // qfull_adder_sm, whose zero keeps a sign), then one saturating addition at width q + log2par + 1.
    __device__ void op_rep(int l, uint32_t o) {
        const uint32_t h = 1u << (l - 1);
        const int w = width(l), lp = (int)p.log2par;
        uint32_t* v = alpha(l - 1);
        for (uint32_t i = gl; i < h; i += G) {
            const uint32_t a = ld(l, i), b = ld(l, i + h);
            v[i] = sm ? raw::f_sm(a, b, w) : (uint32_t)raw::val(raw::f_c2(a, b, w), w);
        }
        __syncwarp();
        const uint32_t nwords = h >> lp;
        for (int s = 0; s < lp; s++) {  // stage s: element i += element i + hh of every word, one bit wider
            const uint32_t hh = (1u << lp) >> (s + 1);
            for (uint32_t t = gl; t < nwords * hh; t += G) {
                const uint32_t at = ((t / hh) << lp) + (t % hh);
                v[at] = sm ? raw::addsub_sm(v[at], v[at + hh], 0u, w + s) : (uint32_t)((int)v[at] + (int)v[at + hh]);
            }
            __syncwarp();
        }
        const int wt = w + lp, ws = w + lp + 1;  // word sum: wt bits; running sum: ws bits
        uint32_t bit = 0;
        if (gl == 0) {
            uint32_t sum = 0;
            for (uint32_t k = 0; k < nwords; k++) {
                const uint32_t t = v[k << lp];
                if (sm) {  // qfull_adder_sat_sm<ws>(word sum extended by one magnitude bit, sum)   scalar.h:165-194
                    const uint32_t ext = (raw::sgn(t, wt) << (ws - 1)) | (t & raw::mk(wt - 1));
                    const uint32_t r = raw::addsub_sm(ext, sum, 0u, ws);
                    const uint32_t mag = r & raw::mk(ws), top = raw::mk(ws - 2);
                    sum = (((r >> ws) & 1u) << (ws - 1)) | (mag > top ? top : (mag & raw::mk(ws - 1)));
                } else {  // qadd<ws>   scalar.h:49-54
                    const int m = (1 << (ws - 1)) - 1;
                    int x = raw::val(sum, ws) + (int)t;
                    x = x > m ? m : (x < -m ? -m : x);
                    sum = (uint32_t)x & raw::mk(ws);
                }
            }
            bit = raw::sgn(sum, ws);
        }
        bit = __shfl_sync(0xFFFFFFFFu, bit, 0, G);
        const uint32_t fill = bit ? 0xFFFFFFFFu : 0u;
        if (h >= 32) {
            for (uint32_t wd = gl; wd < (h >> 5); wd += G) beta[(o >> 5) + wd] = fill;
        } else if (gl == 0) {
            const uint32_t wd = o >> 5, lm = ((1u << h) - 1u) << (o & 31u);
            beta[wd] = (beta[wd] & ~lm) | (fill & lm);
        }
        __syncwarp();
    }
    // G_R1_STATE :1571-1642 (spc = false) and G_SPC_STATE :1737-1842: right child = sign(g).  SPC: parity over the
    // child; Min_Mask_TREE_<PAR> (functions.h:3450-3973) is a tournament over halves in which the upper element wins only
    // when strictly smaller, and a later word replaces the minimum only when strictly smaller -- i.e. the minimum of
    // (|g|, word, bit-reversed position in the word); that bit is flipped when the parity is odd.
    __device__ void op_gsign(int l, uint32_t o, bool zero_beta, bool spc) {
        const uint32_t h = 1u << (l - 1);
        const int w = width(l), lp = (int)p.log2par;
        uint32_t* v = alpha(l - 1);
        uint32_t parity = 0, best = 0xFFFFFFFFu;
        for (uint32_t i = gl; i < h; i += G) {
            const uint32_t pos = o + i;
            const uint32_t u = zero_beta ? 0u : (beta[pos >> 5] >> (pos & 31u)) & 1u;
            const uint32_t a = ld(l, i), b = ld(l, i + h);
            const uint32_t r = sm ? raw::g_sm(a, b, u, w, false) : raw::g_c2(a, b, u, w, false);
            v[i] = r;
            if (spc) {
                parity ^= raw::sgn(r, w);
                const uint32_t mag = sm ? (r & raw::mk(w - 1)) : raw::abs_c2(r, w);
                const uint32_t j = i & ((1u << lp) - 1u);
                const uint32_t key = (mag << 20) | ((i >> lp) << lp) | (__brev(j) >> (32 - lp));
                best = key < best ? key : best;
            }
        }
        __syncwarp();
        pack_signs(v, o + h, h, w);
        if (spc) {
            for (int m = G / 2; m >= 1; m >>= 1) {
                parity ^= __shfl_xor_sync(0xFFFFFFFFu, parity, m, G);
                const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, best, m, G);
                best = other < best ? other : best;
            }
            __syncwarp();
            if (gl == 0 && parity) {
                const uint32_t kj = best & 0xFFFFFu, j = __brev(kj & ((1u << lp) - 1u)) >> (32 - lp);
                const uint32_t pos = o + h + ((kj >> lp) << lp) + j;
                beta[pos >> 5] ^= 1u << (pos & 31u);
            }
        }
        __syncwarp();
    }
    __device__ void op_p2(uint32_t o, uint32_t lf) {
        if (gl == 0) {
            const int w = width(1);
            const uint32_t a = ld(1, 0), b = ld(1, 1);
            const uint32_t u0 = (raw::sgn(a, w) ^ raw::sgn(b, w)) & lf & 1u;
            uint32_t u1;
            if (sm) {  // G_simplified_SM: sign of the larger magnitude, tie -> the a term; each sign gated by the flag
                const uint32_t f1 = (lf >> 1) & 1u;
                const uint32_t sa = (raw::sgn(a, w) ^ u0) & f1, sb = raw::sgn(b, w) & f1;
                u1 = (a & raw::mk(w - 1)) < (b & raw::mk(w - 1)) ? sb : sa;
            } else {  // G_simplified_C2: sign of the exact (w+1)-bit sum
                u1 = raw::sgn((uint32_t)raw::addsub_c2(a, b, u0, w) & raw::mk(w + 1), w + 1) & (lf >> 1) & 1u;
            }
            const uint32_t wd = o >> 5, sh = o & 31u;
            beta[wd] = (beta[wd] & ~(3u << sh)) | (((u0 ^ u1) | (u1 << 1)) << sh);
        }
        __syncwarp();
    }
    __device__ void op_p1(uint32_t o, uint32_t lf) {
        if (gl == 0) {
            const uint32_t a = ld(0, 0);
            const uint32_t wd = o >> 5, sh = o & 31u;
            beta[wd] = (beta[wd] & ~(1u << sh)) | ((raw::sgn(a, width(0)) & lf & 1u) << sh);
        }
        __syncwarp();
    }

    __device__ void run() {
        ew = 0ull;
        for (uint32_t pc = 0;; pc++) {
            const uint32_t w = __ldg(p.sched + pc);
            const uint32_t opc = op_code(w);
            const int l = (int)op_level(w);
            const uint32_t o = op_offset(w);
            if (opc == OP_END) break;
            switch (opc) {
                case OP_F: op_f(l); break;
                case OP_G: op_g(l, o, op_nosat(w), false); break;
                case OP_G0: op_g(l, o, op_nosat(w), true); break;
                case OP_H: op_h(l, o, false); break;
                case OP_HCOPY: op_h(l, o, true); break;
                case OP_R0: op_r0(l, o); break;
                case OP_P2: op_p2(o, op_lf(w)); break;
                case OP_P1: op_p1(o, op_lf(w)); break;
                case OP_REP: op_rep(l, o); break;
                case OP_GR1: op_gsign(l, o, op_nosat(w), false); break;
                case OP_GSPC: op_gsign(l, o, op_nosat(w), true); break;
                default: break;
            }
        }
    }
};

template <int G>
__global__ void __launch_bounds__(128) sc_decode_raw_kernel(const RawParams p) {
    extern __shared__ __align__(16) uint8_t smem_fast[];
    uint32_t* smem = reinterpret_cast<uint32_t*>(smem_fast);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int FPW = 32 / G;  // frames per warp
    const int nwarps = blockDim.x >> 5;
    const unsigned long long f_per_cta = (unsigned long long)nwarps * FPW;
    const unsigned long long slot = (unsigned long long)warp * FPW + lane / G;

    RawDecoder<G> d(p);
    d.gl = lane % G;
    d.sm = p.sigmag != 0u;
    uint32_t* sm = smem + slot * p.sm_words_per_frame;
    uint32_t* ws = p.ws + ((unsigned long long)blockIdx.x * f_per_cta + slot) * p.ws_words_per_frame;
    d.sm_alpha = sm;
    d.gl_alpha = ws;
    d.beta = p.beta_in_smem ? sm + (2u << p.ls) : ws + p.n;

    for (unsigned long long base = (unsigned long long)blockIdx.x * f_per_cta; base < p.nframes;
         base += (unsigned long long)gridDim.x * f_per_cta) {
        if (base + (unsigned long long)warp * FPW >= p.nframes) break;  // warp-uniform
        unsigned long long f = base + slot;
        const bool valid = f < p.nframes;
        if (!valid) f = p.nframes - 1;  // idle group shadows the last frame, stores nothing
        d.llr = p.llr + f * p.n;
        d.run();
        if (valid) {
            const uint32_t tail = p.n < 32 ? ((1u << p.n) - 1u) : 0xFFFFFFFFu;  // n < 32: one partial word
            for (uint32_t w = d.gl; w < p.wpf; w += G) p.xhat[f * p.wpf + w] = d.beta[w] & tail;
        }
        __syncwarp();
    }
}

}  // namespace scpd
