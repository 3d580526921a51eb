// decode_generic.cuh -- schedule-interpreting SC decoder, any N / PAR / Q / EXTENDED (CA2).
//
// One group of G lanes decodes one FRAME PAIR: the two frames ride in the two 16-bit halves of
// every register (int16x2 SIMD: VIMNMX.S16x2 / VIADD.16x2 / VIADDMNMX are single instructions on
// sm_100a), which is legal because all frames of a batch share the frozen set and therefore the
// control flow.  A warp holds 32/G groups.  The LLR stack alpha[l] (l = log2 of the node size)
// lives in shared memory for the small levels and in an L2-resident global workspace above
// that; partial sums beta are bit arrays.  The walk itself is the host-built schedule
// (schedule.h); the only data-dependent control flow is the rate-1 fallback.
//
// Arithmetic contract (reference, CA2):  f  F_function_C2  functions.h:48-61
//                                        g  G_function_C2  functions.h:63-75 (saturating)
//                                           G_extended_C2  functions.h:77-88 (inside the leaf)
//                                        terminals Spec_P2 / Spec_P1 functions.h:354-384
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "schedule.h"

namespace scpd {

struct DecodeParams {
    const uint32_t* sched;  // device copy of the schedule
    const int8_t* llr;      // [nframes][n]
    uint32_t* xhat;         // [nframes][wpf]
    unsigned long long nframes;
    unsigned long long num_fp;  // frame pairs = ceil(nframes / 2)
    uint32_t n, log2n, wpf;
    uint32_t satv;     // 2^(Q-1) - 1
    uint32_t log2par;  // log2(PAR)
    uint32_t extended;
    uint32_t ls;             // alpha levels 0..ls are in shared memory, ls+1..log2n-1 in the workspace
    uint32_t beta_in_smem;   // partial sums in shared memory (else workspace)
    uint32_t sm_words_per_fp;
    uint32_t* ws;            // global workspace, ws_words_per_fp per resident frame pair
    unsigned long long ws_words_per_fp;
};

// ---------------------------------------------------------------- packed int16x2 primitives
__device__ __forceinline__ uint32_t pack2(int a, int b) {
    return ((uint32_t)a & 0xFFFFu) | ((uint32_t)b << 16);
}
// f(a,b) = sign(a) sign(b) min(|a|,|b|) = max(a+b, 0) - max(a, b)   (exact identity on integers;
// yields +0 whenever either input is 0, as qsign/qmin do in F_function_C2)
__device__ __forceinline__ uint32_t f_c2(uint32_t a, uint32_t b) {
    uint32_t s = __viaddmax_s16x2(a, b, 0u);
    uint32_t m = __vmaxs2(a, b);
    return __vsub2(s, m);
}
// mask: 0xFFFF in a half where the partial sum is 1 (then b - a), 0 where it is 0 (b + a)
__device__ __forceinline__ uint32_t g_c2_nosat(uint32_t a, uint32_t b, uint32_t mask) {
    uint32_t an = __vsub2(a ^ mask, mask);  // conditional negate
    return __vadd2(b, an);
}
__device__ __forceinline__ uint32_t g_c2_sat(uint32_t a, uint32_t b, uint32_t mask, uint32_t satp, uint32_t satn) {
    uint32_t an = __vsub2(a ^ mask, mask);
    uint32_t r = __viaddmin_s16x2(b, an, satp);  // min(b + a', +M)
    return __vmaxs2(r, satn);                    // max(., -M)           qsat scalar.h:15-21
}
__device__ __forceinline__ uint32_t bits_to_mask(uint32_t ba, uint32_t bb) {
    return (0u - ba) & 0xFFFFu | ((0u - bb) << 16);
}

template <int G>
struct GenericDecoder {
    static constexpr unsigned FULL = 0xFFFFFFFFu;
    const DecodeParams& p;
    uint32_t* sm_alpha;  // index (1 << l) + i
    uint32_t* gl_alpha;  // same indexing, levels > ls
    uint32_t* betaA;
    uint32_t* betaB;
    const int8_t* llrA;
    const int8_t* llrB;
    int gl, grp;
    uint32_t satp, satn;

    __device__ GenericDecoder(const DecodeParams& p_) : p(p_) {}

    __device__ __forceinline__ uint32_t* alpha(int l) const {
        return ((uint32_t)l <= p.ls ? sm_alpha : gl_alpha) + (1u << l);
    }
    __device__ __forceinline__ uint32_t ld(int l, uint32_t i) const {
        if ((uint32_t)l == p.log2n) return pack2(llrA[i], llrB[i]);
        return alpha(l)[i];
    }
    __device__ __forceinline__ uint32_t beta_mask(uint32_t pos) const {
        uint32_t w = pos >> 5, b = pos & 31u;
        return bits_to_mask((betaA[w] >> b) & 1u, (betaB[w] >> b) & 1u);
    }

    __device__ void op_f(int l) {
        const uint32_t h = 1u << (l - 1);
        uint32_t* dst = alpha(l - 1);
        for (uint32_t i = gl; i < h; i += G) dst[i] = f_c2(ld(l, i), ld(l, i + h));
        __syncwarp();
    }
    __device__ void op_g(int l, uint32_t o, bool nosat, bool zero_beta) {
        const uint32_t h = 1u << (l - 1);
        uint32_t* dst = alpha(l - 1);
        for (uint32_t i = gl; i < h; i += G) {
            uint32_t m = zero_beta ? 0u : beta_mask(o + i);
            uint32_t a = ld(l, i), b = ld(l, i + h);
            dst[i] = nosat ? g_c2_nosat(a, b, m) : g_c2_sat(a, b, m, satp, satn);
        }
        __syncwarp();
    }
    // beta[o..o+h) (op)= beta[o+h..o+2h)
    __device__ void op_h(int l, uint32_t o, bool copy) {
        const uint32_t h = 1u << (l - 1);
        if (h >= 32) {
            const uint32_t w0 = o >> 5, w1 = (o + h) >> 5, nw = h >> 5;
            for (uint32_t w = gl; w < nw; w += G) {
                uint32_t a = betaA[w1 + w], b = betaB[w1 + w];
                if (!copy) {
                    a ^= betaA[w0 + w];
                    b ^= betaB[w0 + w];
                }
                betaA[w0 + w] = a;
                betaB[w0 + w] = b;
            }
        } else if (gl == 0) {
            const uint32_t w = o >> 5, sh = o & 31u;
            const uint32_t lm = ((1u << h) - 1u) << sh;
            uint32_t a = betaA[w], b = betaB[w];
            if (copy) {
                a = (a & ~lm) | ((a >> h) & lm);
                b = (b & ~lm) | ((b >> h) & lm);
            } else {
                a ^= (a >> h) & lm;
                b ^= (b >> h) & lm;
            }
            betaA[w] = a;
            betaB[w] = b;
        }
        __syncwarp();
    }
    __device__ void op_r0(int l, uint32_t o) {
        const uint32_t n = 1u << l;
        if (n >= 32) {
            for (uint32_t w = gl; w < (n >> 5); w += G) {
                betaA[(o >> 5) + w] = 0u;
                betaB[(o >> 5) + w] = 0u;
            }
        } else if (gl == 0) {
            const uint32_t m = ~(((1u << n) - 1u) << (o & 31u));
            betaA[o >> 5] &= m;
            betaB[o >> 5] &= m;
        }
        __syncwarp();
    }
    // hard decision of a whole node; returns (warp-uniform) whether any LLR of any group was 0
    __device__ bool op_hd(int l, uint32_t o) {
        const uint32_t n = 1u << l;
        bool zero = false;
        uint32_t accA = 0, accB = 0;
        const uint32_t gmask = (G == 32) ? 0xFFFFFFFFu : ((1u << G) - 1u);
        const uint32_t iters = (n + G - 1) / G;
        for (uint32_t k = 0; k < iters; k++) {
            const uint32_t i = k * G + gl;
            uint32_t v = 0x00010001u;  // inactive lanes: positive, non-zero
            if (i < n) v = ld(l, i);
            zero |= ((v & 0xFFFFu) == 0u) | ((v >> 16) == 0u);
            uint32_t ba = __ballot_sync(FULL, (v >> 15) & 1u);
            uint32_t bb = __ballot_sync(FULL, v >> 31);
            ba = (ba >> (grp * G)) & gmask;
            bb = (bb >> (grp * G)) & gmask;
            const uint32_t pos = o + k * G;
            accA |= ba << (pos & 31u);
            accB |= bb << (pos & 31u);
            if (n >= 32 && ((pos + G) & 31u) == 0u) {  // a full word is ready
                if (gl == 0) {
                    betaA[pos >> 5] = accA;
                    betaB[pos >> 5] = accB;
                }
                accA = accB = 0;
            }
        }
        if (n < 32 && gl == 0) {
            const uint32_t m = ((1u << n) - 1u) << (o & 31u);
            betaA[o >> 5] = (betaA[o >> 5] & ~m) | accA;
            betaB[o >> 5] = (betaB[o >> 5] & ~m) | accB;
        }
        zero = __any_sync(FULL, zero);
        __syncwarp();
        return zero;
    }
    __device__ __forceinline__ static void p2_scalar(int a, int b, int f0, int f1, uint32_t& x0, uint32_t& x1) {
        int u0 = ((a < 0) ^ (b < 0)) & f0;      // F_simplified_C2 functions.h:90-101
        int s = u0 ? (b - a) : (b + a);         // G_simplified_C2 functions.h:103-118 (exact sum)
        int u1 = (s < 0) & f1;
        x0 = (uint32_t)(u0 ^ u1);               // Spec_P2 functions.h:380-383
        x1 = (uint32_t)u1;
    }
    __device__ void op_p2(uint32_t o, uint32_t lf) {
        if (gl == 0) {
            uint32_t a = ld(1, 0), b = ld(1, 1);
            uint32_t a0, a1, b0, b1;
            p2_scalar((short)(a & 0xFFFFu), (short)(b & 0xFFFFu), lf & 1, (lf >> 1) & 1, a0, a1);
            p2_scalar((short)(a >> 16), (short)(b >> 16), lf & 1, (lf >> 1) & 1, b0, b1);
            const uint32_t w = o >> 5, sh = o & 31u;
            betaA[w] = (betaA[w] & ~(3u << sh)) | ((a0 | (a1 << 1)) << sh);
            betaB[w] = (betaB[w] & ~(3u << sh)) | ((b0 | (b1 << 1)) << sh);
        }
        __syncwarp();
    }
    __device__ void op_p1(uint32_t o, uint32_t lf) {
        if (gl == 0) {
            uint32_t a = ld(0, 0);
            const uint32_t w = o >> 5, sh = o & 31u;
            uint32_t xa = ((a >> 15) & 1u) & lf, xb = (a >> 31) & lf;  // Spec_P1 functions.h:355-364
            betaA[w] = (betaA[w] & ~(1u << sh)) | (xa << sh);
            betaB[w] = (betaB[w] & ~(1u << sh)) | (xb << sh);
        }
        __syncwarp();
    }
    __device__ __forceinline__ bool nosat(int l) const { return p.extended && (uint32_t)l <= p.log2par; }

    // Plain SC of an all-information node (no pruning): the rate-1 fallback.  Iterative DFS over
    // the terminals t of the node, equivalent to the F/R/G/H state sequence of my_module.h.
    __device__ void generic_sc(int l, uint32_t o) {
        const int tl = p.log2par >= 1 ? 1 : 0;
        const uint32_t nt = 1u << (l - tl);
        for (uint32_t t = 0; t < nt; t++) {
            const uint32_t to = o + (t << tl);
            if (t == 0) {
                for (int lv = l; lv > tl; lv--) op_f(lv);
            } else {
                const int lv0 = (__ffs(t) - 1) + tl + 1;
                op_g(lv0, to & ~((1u << lv0) - 1u), nosat(lv0), false);
                for (int lv = lv0 - 1; lv > tl; lv--) op_f(lv);
            }
            if (tl)
                op_p2(to, 3u);
            else
                op_p1(to, 1u);
            const int ones = __ffs(~t) - 1;  // trailing ones of t: nodes completed by this terminal
            for (int j = 1; j <= ones && tl + j <= l; j++) {
                const int lv = tl + j;
                op_h(lv, to + (1u << tl) - (1u << lv), false);
            }
        }
    }

    __device__ void run() {
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
                case OP_R1:
                    if (op_hd(l, o)) generic_sc(l, o);
                    break;
                case OP_P2: op_p2(o, op_lf(w)); break;
                case OP_P1: op_p1(o, op_lf(w)); break;
                default: break;
            }
        }
    }
};

template <int G>
__global__ void __launch_bounds__(128) sc_decode_generic_kernel(const DecodeParams p) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int GPW = 32 / G;  // groups (frame pairs) per warp
    const int nwarps = blockDim.x >> 5;
    const unsigned long long fp_per_cta = (unsigned long long)nwarps * GPW;
    const int grp = lane / G;
    const unsigned long long slot = (unsigned long long)warp * GPW + grp;

    GenericDecoder<G> d(p);
    d.gl = lane % G;
    d.grp = grp;
    d.satp = p.satv * 0x00010001u;
    d.satn = ((0u - p.satv) & 0xFFFFu) * 0x00010001u;
    uint32_t* sm = smem + slot * p.sm_words_per_fp;
    uint32_t* ws = p.ws + ((unsigned long long)blockIdx.x * fp_per_cta + slot) * p.ws_words_per_fp;
    d.sm_alpha = sm;
    d.gl_alpha = ws;
    if (p.beta_in_smem) {
        d.betaA = sm + (2u << p.ls);
        d.betaB = d.betaA + p.wpf;
    } else {
        d.betaA = ws + p.n;
        d.betaB = d.betaA + p.wpf;
    }

    for (unsigned long long base = (unsigned long long)blockIdx.x * fp_per_cta; base < p.num_fp;
         base += (unsigned long long)gridDim.x * fp_per_cta) {
        if (base + (unsigned long long)warp * GPW >= p.num_fp) break;  // warp-uniform
        unsigned long long fp = base + slot;
        const bool valid = fp < p.num_fp;
        if (!valid) fp = p.num_fp - 1;  // idle group shadows the last pair, stores nothing
        const unsigned long long fa = 2 * fp;
        const bool has_b = fa + 1 < p.nframes;
        const unsigned long long fb = has_b ? fa + 1 : fa;
        d.llrA = p.llr + fa * p.n;
        d.llrB = p.llr + fb * p.n;
        d.run();
        if (valid) {
            const uint32_t tail = p.n < 32 ? ((1u << p.n) - 1u) : 0xFFFFFFFFu;  // n < 32: one partial word
            for (uint32_t w = d.gl; w < p.wpf; w += G) {
                p.xhat[fa * p.wpf + w] = d.betaA[w] & tail;
                if (has_b) p.xhat[fb * p.wpf + w] = d.betaB[w] & tail;
            }
        }
        __syncwarp();
    }
}

}  // namespace scpd
