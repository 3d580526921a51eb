// decode_ss.cuh -- the slot-sliced SC decode kernel: every LANE decodes one frame, bit-sliced over 32
// consecutive code positions ("a chunk"); the bottom of the tree runs as fp16x2 arithmetic on the FMA pipe.
//
// Why (profiles/ncu_r1_bs_c2_summary.txt): the frame-sliced kernel (decode_bs.cuh: one register = one bit of
// one LLR for 32 FRAMES, the lanes of a warp spread over the LLRs of a node) keeps only n/2 lanes busy on a
// node of n <= 32 LLRs and walks those nodes through shared memory, which cost 5.4x the instructions of the
// plane arithmetic.  Turning the slicing by 90 degrees removes the problem:
//   * a 32-bit register holds one bit plane of 32 CONSECUTIVE LLRs of ONE frame; lane = frame, a warp = 32 frames
//     that share nothing but the schedule.  f / g of a node of n >= 64 LLRs are the same LOP3 sequences as before
//     (bs_arith.cuh: f 2P+1, saturating g 5P+2 for P magnitude planes) on n/64 independent chunk pairs per lane:
//     every lane is busy at every level, there is no cross-lane traffic, no __syncwarp, and the partial sums
//     (one bit per code position) are already the packed output words -- no transposition on the way out.
//   * a node of 32 LLRs is one chunk.  Its sign-magnitude planes are transposed IN the lane (three delta-swap
//     stages on 8 registers) into sign-magnitude bytes, and a sign-magnitude byte spread to 16 bits IS an fp16
//     denormal of the same value: 16 registers of fp16x2 hold the node.  Below that, f(a,b) = (|a+b| - |a-b|)/2
//     is three HADD2 (the factor 2 is carried as a static scale: f, g and the hard decisions are homogeneous),
//     g(a,b,u) = b + (1-2u) a is one HFMA2 with the partial sum kept as +-1.0, h is one HMUL2, and the hard
//     decision is the sign bit.  All values are integers below 2^11 times a power of two, so fp16 is exact;
//     x - x = +0 in round-to-nearest, so a CA2 zero never carries a sign (hd(0) = 0).  These run on
//     the FMA pipe, which the LOP3 stream of the upper levels (ALU pipe) leaves idle.
//   * alpha[l] (the 2^l LLRs a node receives) for l = 6 .. lsa lives in shared memory, above in a per-warp
//     workspace, the channel level in the plane buffer written by ss_planes_kernel.  Quad q of chunk c sits at
//     uint4 index (2c + q) * 32 + lane: conflict-free / coalesced 512-byte rows.  alpha[5] and below never
//     leave registers.
//
// Bit-exactness: CA2 only (SIGMAG stays on decode_bs.cuh: its -0 and tie rules do not map onto IEEE zeros).
// f / g at n >= 64: bs_arith.cuh.  Inside a 32-LLR node: g saturates at +-(2^(Q-1)-1) (times the static scale)
// unless the node lies inside the PAR-wide un-saturated leaf decoder (Spec_P*_ext, functions.h:413-546);
// Spec_P2 (functions.h:367-384) = the leaf() cases below.  All-information nodes are replaced by the hard
// decision only at n >= 32 and only when no LLR of any frame of the warp is zero (DESIGN.md "why pruning is
// bit-identical"); inside the fp16x2 walker they are simply decoded.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bs_arith.cuh"
#include "ss_plan.h"

#ifndef SS_SIDE_UNROLL
#define SS_SIDE_UNROLL 3  // 1: the (left, right) loops of the fp16x2 walker stay loops (one copy of the 8-LLR routines)
#endif
// 3: unrolled, with the 16- and 8-LLR routines as out-of-line functions (arguments and results in registers)
#if SS_SIDE_UNROLL == 1
#define SS_SIDE_PRAGMA _Pragma("unroll 1")
#else
#define SS_SIDE_PRAGMA _Pragma("unroll")
#endif
#if defined(__CUDACC__)
#define SS_NOINLINE __device__ __noinline__
#define SS_DEV __device__ __forceinline__
#define SS_ANY(x) __any_sync(0xFFFFFFFFu, (x))
#else
#define SS_NOINLINE inline
#define SS_DEV inline
#define SS_ANY(x) (x)
#endif

namespace scpd {

struct SsParams {
    const uint32_t* sched;
    uint32_t sched_words;  // > 0: every CTA keeps a copy of the schedule in shared memory
    const uint4* planes;   // channel planes of every 32-frame task (ss_planes_kernel)
    unsigned long long planes_stride;  // uint4 per task
    uint32_t* xhat;
    unsigned long long nframes, ntasks;
    uint32_t n, log2n, wpf;
    uint32_t lsa, lwin, win_words;
    uint32_t ltm;          // alpha level kept in tensor memory (0 = none); tm_cols: columns per warp
    uint32_t tm_cols;
    uint32_t sm_stride, sm_beta_off;  // uint4, per warp
    uint4* ws;
    unsigned long long ws_stride;  // uint4 per warp
    uint32_t ws_beta_off;
    uint32_t aoff[24];
    // the leading f chain F(log2n), F(log2n - 1), ... of the schedule is computed by ss_planes_kernel (it depends on the
    // channel alone): alpha[l] of the leftmost node, l = lpre .. log2n, sits at planes + poff[l] (poff[log2n] = 0,
    // lpre = log2n: nothing prefused) and the walk starts below it
    uint32_t lpre;
    uint32_t poff[24];
    // alpha[lhot] (the smallest workspace level) in its own dense array: hot + slot * hot_stride; lhot = 0: none
    uint4* hot;
    uint32_t lhot, hot_stride;
    // != nullptr: per (function, level) SM-clock cycles [6][32] and visits [6][32] of warp 0 of every CTA (scpd_stage_time)
    unsigned long long* prof;
};

namespace ss {

// ------------------------------------------------------------------------------------------------ fp16x2
#if defined(__CUDA_ARCH__)
struct H2 {
    uint32_t u;
};
SS_DEV H2 h2_add(H2 a, H2 b) {
    H2 r;
    asm("add.f16x2 %0, %1, %2;" : "=r"(r.u) : "r"(a.u), "r"(b.u));
    return r;
}
SS_DEV H2 h2_sub(H2 a, H2 b) {
    H2 r;
    asm("sub.f16x2 %0, %1, %2;" : "=r"(r.u) : "r"(a.u), "r"(b.u));
    return r;
}
SS_DEV H2 h2_abssub(H2 s, H2 d) {  // |s| - |d|: one HADD2 with operand modifiers
    uint32_t as, ad;
    asm("abs.f16x2 %0, %1;" : "=r"(as) : "r"(s.u));
    asm("abs.f16x2 %0, %1;" : "=r"(ad) : "r"(d.u));
    H2 r;
    asm("sub.f16x2 %0, %1, %2;" : "=r"(r.u) : "r"(as), "r"(ad));
    return r;
}
SS_DEV H2 h2_fma(H2 a, H2 b, H2 c) {
    H2 r;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r.u) : "r"(a.u), "r"(b.u), "r"(c.u));
    return r;
}
SS_DEV H2 h2_mul(H2 a, H2 b) {
    H2 r;
    asm("mul.f16x2 %0, %1, %2;" : "=r"(r.u) : "r"(a.u), "r"(b.u));
    return r;
}
// half bits of v * 2^-24 (v < 2^11 << k, at most 11 significant bits)
__host__ __device__ constexpr uint32_t h_bits(uint32_t v) {
    if (v < 1024u) return v;  // denormal
    int e = 0;
    for (uint32_t t = v; t > 1u; t >>= 1) e++;
    const uint32_t mant = (e >= 10 ? (v >> (e - 10)) : (v << (10 - e))) & 0x3FFu;
    return ((uint32_t)(e - 9) << 10) | mant;
}
template <uint32_t V>
SS_DEV H2 h2_clamp(H2 x) {  // to +-V (units of 2^-24)
    constexpr uint32_t hb = h_bits(V), pos = hb | (hb << 16), neg = pos | 0x80008000u;
    H2 r;
    asm("max.f16x2 %0, %1, %2;" : "=r"(r.u) : "r"(x.u), "r"(neg));
    asm("min.f16x2 %0, %1, %2;" : "=r"(r.u) : "r"(r.u), "r"(pos));
    return r;
}
SS_DEV H2 h2_one() { return H2{0x3C003C00u}; }
SS_DEV H2 h2_half() { return H2{0x38003800u}; }
SS_DEV H2 h2_sumboth(H2 a) {  // (lo + hi) in both halves
    H2 sw{__byte_perm(a.u, a.u, 0x1032)};
    return h2_add(a, sw);
}
SS_DEV H2 h2_sgn(H2 a) { return H2{bs::lop3<((bs::LA & bs::LB) | bs::LC) & 0xFFu>(a.u, 0x80008000u, 0x3C003C00u)}; }
SS_DEV H2 h2_xorsgn(H2 a) {  // (sign(lo) ^ sign(hi) as +-1, +1)
    const uint32_t t = a.u ^ (a.u >> 16);
    return H2{(t & 0x8000u) | 0x3C003C00u};
}
// sign-magnitude bytes (s << 7 | mag): byte c of wa -> low half, byte c of wb -> high half
SS_DEV H2 h2_from_sm(uint32_t wa, uint32_t wb, int c) {  // c folds to an immediate selector after unrolling
    const uint32_t sel = (uint32_t)c | ((uint32_t)c << 4) | ((uint32_t)(4 + c) << 8) | ((uint32_t)(4 + c) << 12);
    return H2{__byte_perm(wa, wb, sel) & 0x807F807Fu};
}
// sign bits of 16 registers of +-1.0 pairs -> one word, bit 2j = low half of B[j]
SS_DEV uint32_t h2_pack16(const H2 (&B)[16]) {
    // One IDP.4A per register: the bytes of +-1.0 pairs are (0, 0x3C | s_lo << 7, 0, 0x3C | s_hi << 7), so with the byte
    // weights (0, 4^r, 0, 2 4^r) four registers add up to 0x3C * 3 * 85 + 128 * (their eight sign bits in order); the
    // constant is folded into the start value.
    uint32_t v[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        uint32_t acc = 0u - 0x3Cu * 3u * 85u;
#pragma unroll
        for (int r = 0; r < 4; r++) acc = __dp4a(B[4 * g + r].u, (uint32_t)((1u << (2 * r + 8)) | (2u << (2 * r + 24))), acc);
        v[g] = acc;  // 128 * (8 sign bits)
    }
    const uint32_t lo = v[0] + (v[1] << 8), hi = v[2] + (v[3] << 8);  // 128 * (16 sign bits)
    return (lo >> 7) | (hi << 9);
}
#else
struct H2 {
    float lo, hi;
};
SS_DEV H2 h2_add(H2 a, H2 b) { return H2{a.lo + b.lo, a.hi + b.hi}; }
SS_DEV H2 h2_sub(H2 a, H2 b) { return H2{a.lo - b.lo, a.hi - b.hi}; }
SS_DEV float h2_fabs(float x) { return x < 0 ? -x : (x == 0 ? 0.f : x); }
SS_DEV H2 h2_abssub(H2 s, H2 d) { return H2{h2_fabs(s.lo) - h2_fabs(d.lo), h2_fabs(s.hi) - h2_fabs(d.hi)}; }
SS_DEV H2 h2_fma(H2 a, H2 b, H2 c) { return H2{a.lo * b.lo + c.lo, a.hi * b.hi + c.hi}; }
SS_DEV H2 h2_mul(H2 a, H2 b) { return H2{a.lo * b.lo, a.hi * b.hi}; }
template <uint32_t V>
SS_DEV H2 h2_clamp(H2 x) {
    const float m = (float)V;
    auto c = [m](float v) { return v > m ? m : (v < -m ? -m : v); };
    return H2{c(x.lo), c(x.hi)};
}
SS_DEV H2 h2_one() { return H2{1.f, 1.f}; }
SS_DEV H2 h2_half() { return H2{0.5f, 0.5f}; }
SS_DEV H2 h2_sumboth(H2 a) { return H2{a.lo + a.hi, a.lo + a.hi}; }
SS_DEV bool h2_neg(float x) { return __builtin_signbit(x) != 0; }
SS_DEV H2 h2_sgn(H2 a) { return H2{h2_neg(a.lo) ? -1.f : 1.f, h2_neg(a.hi) ? -1.f : 1.f}; }
SS_DEV H2 h2_xorsgn(H2 a) { return H2{h2_neg(a.lo) != h2_neg(a.hi) ? -1.f : 1.f, 1.f}; }
SS_DEV H2 h2_from_sm(uint32_t wa, uint32_t wb, int c) {
    const uint32_t a = (wa >> (8 * c)) & 0xFFu, b = (wb >> (8 * c)) & 0xFFu;
    auto v = [](uint32_t x) { const float m = (float)(x & 0x7Fu); return (x & 0x80u) ? -m : m; };
    return H2{v(a), v(b)};
}
SS_DEV uint32_t h2_pack16(const H2 (&B)[16]) {
    uint32_t acc = 0u;
    for (int j = 0; j < 16; j++) acc |= (h2_neg(B[j].lo) ? 1u : 0u) << (2 * j) | (h2_neg(B[j].hi) ? 1u : 0u) << (2 * j + 1);
    return acc;
}
#endif

struct R4 {
    H2 v[4];
};
struct R8 {
    H2 v[8];
};

// 8 rows x 32 columns of bits, seen as four 8 x 8 matrices (one per byte column): transpose each of them.
// in: row p bit (8c + k);  out: row k byte c bit p.  Three delta-swap stages.
SS_DEV void transpose8x8x4(uint32_t (&w)[8]) {
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const uint32_t t = ((w[p] >> 4) ^ w[p + 4]) & 0x0F0F0F0Fu;
        w[p + 4] ^= t;
        w[p] ^= t << 4;
    }
#pragma unroll
    for (int p = 0; p < 8; p++) {
        if (p & 2) continue;
        const uint32_t t = ((w[p] >> 2) ^ w[p + 2]) & 0x33333333u;
        w[p + 2] ^= t;
        w[p] ^= t << 2;
    }
#pragma unroll
    for (int p = 0; p < 8; p += 2) {
        const uint32_t t = ((w[p] >> 1) ^ w[p + 1]) & 0x55555555u;
        w[p + 1] ^= t;
        w[p] ^= t << 1;
    }
}

// 32 int8 LLRs of one frame (v[j] = LLRs 4j .. 4j+3, as they lie in memory) -> one chunk of sign-magnitude planes,
// bit i of every plane = LLR i.                                      wrapper_in.h:30-42, qconv_format scalar.h:229-239
template <int P>
SS_DEV void chunk_planes(const uint32_t (&v)[8], bs::Val<P>& x) {
    // row j of the bit matrix = bytes of the LLRs (j, 8 + j, 16 + j, 24 + j): LLR s is byte s % 4 of word s / 4
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t b = j & 3, sel = b | ((4u + b) << 4);
        const uint32_t t0 = bs::bperm(v[j / 4], v[2 + j / 4], sel), t1 = bs::bperm(v[4 + j / 4], v[6 + j / 4], sel);
        w[j] = bs::bperm(t0, t1, 0x5410);
    }
    transpose8x8x4(w);  // w[b] bit i = bit b of LLR i (two's complement)
    bs::from_int8_planes<P>(w, x);
}

// The fp16x2 walker of a node of 32 LLRs and everything below it.  Register j of an array holds the LLRs
// (2j, 2j+1) of the node, so the pair (i, i + n/2) of a node of n >= 4 is (register i/2, register i/2 + n/4):
// every op is a full-width fp16x2 op; only the 2-LLR terminals look inside a register.
//
// Nodes of 8 LLRs are decoded by straight-line routines specialised at compile time for the information-flag
// patterns that polar constructions produce (nine 8-bit patterns cover all 47 frozen tables of the reference:
// ss_plan.h, SS_KNOWN8), selected through a jump table; any other pattern (arbitrary flag tables are legal
// input) runs the same template with the flags read at run time.  The two levels above (16, 32) are loops over
// (left, right) so that the routines exist once.

// node types of the 7 nodes of an 8-LLR node from its flag byte, local heap order (0: the node, 1-2: the
// 4-LLR children, 3-6: the 2-LLR terminals)
template <uint32_t PAT>
struct CtDesc8 {
    template <int HEAP>
    SS_HD static constexpr uint32_t tp() {
        if (HEAP == 0) return PAT == 0u ? SS_T_R0 : SS_T_MIX;
        if (HEAP <= 2) return ((PAT >> (4 * (HEAP >= 1 ? HEAP - 1 : 0))) & 15u) == 0u ? SS_T_R0 : SS_T_MIX;
        const uint32_t f = (PAT >> (2 * (HEAP >= 3 ? HEAP - 3 : 0))) & 3u;  // f0 | f1 << 1
        return f == 2u ? 0u : f == 0u ? 1u : f == 3u ? 2u : 3u;  // (0,1) (0,0) (1,1) (1,0)
    }
};
struct RtDesc8 {
    uint32_t fb;     // flag byte
    uint32_t prune;  // 0: all-frozen nodes are decoded like any other (pruning mode NONE)
    template <int HEAP>
    SS_DEV uint32_t tp() const {
        if (HEAP == 0) return (prune && fb == 0u) ? SS_T_R0 : SS_T_MIX;
        if (HEAP <= 2) return (prune && ((fb >> (4 * (HEAP >= 1 ? HEAP - 1 : 0))) & 15u) == 0u) ? SS_T_R0 : SS_T_MIX;
        const uint32_t f = (fb >> (2 * (HEAP >= 3 ? HEAP - 3 : 0))) & 3u;
        return f == 2u ? 0u : f == 0u ? 1u : f == 3u ? 2u : 3u;
    }
};

template <int Q, int LOG2PAR, bool EXT>
struct Walk {
    static constexpr int P = Q - 1;
    static constexpr uint32_t MAXV = (1u << (Q - 1)) - 1u;
    // g of a node of n LLRs saturates unless the node lies inside the un-saturated PAR-wide leaf decoder
    SS_HD static constexpr bool sat(int n) { return !(EXT && n <= (1 << LOG2PAR)); }
    // f doubles the scale (2 f = |a+b| - |a-b|).  Where a saturating g follows further down the scale must be
    // known, so there f is halved again (one more HMUL2); elsewhere the factor rides along (f, g, hd are homogeneous)
    SS_HD static constexpr bool halve(int n) { return n >= 8 && sat(n / 2); }

    static SS_DEV H2 f2(H2 a, H2 b) { return h2_abssub(h2_add(a, b), h2_sub(a, b)); }  // F_function_C2 functions.h:48-61

    // N <= 8 LLRs in A (scale 2^SC), node HEAP of the descriptor D; B = partial sums as +-1.0 (-1 = bit 1)
    template <int N, int HEAP, int SC, class D>
    static SS_DEV void node(const H2 (&A)[N / 2], const D& d, H2 (&B)[N / 2]) {
        const uint32_t t = d.template tp<HEAP>();
        if constexpr (N == 2) {  // Spec_P2, functions.h:367-384 with F_simplified / G_simplified :90-118
            if (t == 1u)
                B[0] = h2_one();  // (0,0)
            else if (t == 2u)
                B[0] = h2_sgn(A[0]);  // (1,1): (u0 ^ u1, u1) = (hd(a), hd(b)) for every input
            else if (t == 0u)
                B[0] = h2_sgn(h2_sumboth(A[0]));  // (0,1): u1 = sign of the exact sum, sum 0 -> 0
            else
                B[0] = h2_xorsgn(A[0]);  // (1,0)
        } else {
            constexpr int H = N / 4;  // registers per half
            if (t == SS_T_R0) {
#pragma unroll
                for (int i = 0; i < 2 * H; i++) B[i] = h2_one();
                return;
            }
            const uint32_t tl = d.template tp<2 * HEAP + 1>(), tr = d.template tp<2 * HEAP + 2>();
            H2 BL[H], BR[H];
            if (tl != SS_T_R0) {
                H2 AL[H];
#pragma unroll
                for (int i = 0; i < H; i++) AL[i] = f2(A[i], A[i + H]);
                node<N / 2, 2 * HEAP + 1, SC + 1>(AL, d, BL);
            }
            if (tr != SS_T_R0) {
                H2 AR[H];
#pragma unroll
                for (int i = 0; i < H; i++) {  // g = b + (1 - 2u) a                    G_function_C2 functions.h:63-88
                    AR[i] = tl != SS_T_R0 ? h2_fma(A[i], BL[i], A[i + H]) : h2_add(A[i], A[i + H]);
                    if constexpr (sat(N)) AR[i] = h2_clamp<(MAXV << SC)>(AR[i]);  // qsat scalar.h:15-21
                }
                node<N / 2, 2 * HEAP + 2, SC>(AR, d, BR);
            }
#pragma unroll
            for (int i = 0; i < H; i++) {  // H_STATE my_module.h:903-932: (left ^ right, right)
                if (tl == SS_T_R0) {
                    B[i] = tr != SS_T_R0 ? BR[i] : h2_one();
                } else {
                    B[i] = tr != SS_T_R0 ? h2_mul(BL[i], BR[i]) : BL[i];
                }
                B[i + H] = tr != SS_T_R0 ? BR[i] : h2_one();
            }
        }
    }
    template <uint32_t PAT>
    static SS_DEV void node8c(const H2 (&A)[4], H2 (&B)[4]) {
        node<8, 0, 0>(A, CtDesc8<PAT>(), B);
    }
    // id: index into SS_KNOWN8 (ss_plan.h), anything else: the flags at run time
    static SS_DEV void dispatch8_body(uint32_t id, uint32_t fb, uint32_t prune, const H2 (&A)[4], H2 (&B)[4]) {
        switch (id) {
            case 1: node8c<0xFFu>(A, B); break;
            case 2: node8c<0xFEu>(A, B); break;
            case 3: node8c<0xE8u>(A, B); break;
            case 4: node8c<0x80u>(A, B); break;
            case 5: node8c<0xE0u>(A, B); break;
            case 6: node8c<0xFCu>(A, B); break;
            case 7: node8c<0xF8u>(A, B); break;
            case 8: node8c<0xC0u>(A, B); break;
            default: node<8, 0, 0>(A, RtDesc8{fb, prune}, B); break;
        }
    }
    // out of line: the routines exist once (instruction cache); four registers in, four out, all in registers
    static SS_NOINLINE R4 dispatch8_fn(H2 a0, H2 a1, H2 a2, H2 a3, uint32_t id, uint32_t fb, uint32_t prune) {
        const H2 A[4] = {a0, a1, a2, a3};
        R4 r;
        dispatch8_body(id, fb, prune, A, r.v);
        return r;
    }
    static SS_DEV void dispatch8(uint32_t id, uint32_t fb, uint32_t prune, const H2 (&A)[4], H2 (&B)[4]) {
#if SS_SIDE_UNROLL == 3
        const R4 r = dispatch8_fn(A[0], A[1], A[2], A[3], id, fb, prune);
#pragma unroll
        for (int i = 0; i < 4; i++) B[i] = r.v[i];
#else
        dispatch8_body(id, fb, prune, A, B);
#endif
    }
    // 16 LLRs; ids / fl: pattern ids (4 bits each) and flag bytes of the two 8-LLR children
    static SS_DEV void walk16_body(const H2 (&A)[8], uint32_t ids, uint32_t fl, uint32_t prune, H2 (&B)[8]) {
        H2 BL[4], Bc[4];
SS_SIDE_PRAGMA
        for (int side = 0; side < 2; side++) {
            const uint32_t id = (ids >> (4 * side)) & 15u, fb = (fl >> (8 * side)) & 255u;
            if (id != 0u) {
                H2 A8[4];
                if (side == 0) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        A8[i] = f2(A[i], A[i + 4]);
                        if constexpr (halve(16)) A8[i] = h2_mul(A8[i], h2_half());
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) {  // BL = +1 where the left child is all-frozen
                        A8[i] = h2_fma(A[i], BL[i], A[i + 4]);
                        if constexpr (sat(16)) A8[i] = h2_clamp<MAXV>(A8[i]);
                    }
                }
                dispatch8(id, fb, prune, A8, Bc);
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) Bc[i] = h2_one();
            }
            if (side == 0) {
#pragma unroll
                for (int i = 0; i < 4; i++) BL[i] = Bc[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            B[i] = h2_mul(BL[i], Bc[i]);
            B[i + 4] = Bc[i];
        }
    }
    // inlined into its two call sites in walk32: the routine itself is ~50 instructions, the call cost 19 register moves
    // (c1 385 -> 393 Gb/s, 112 -> 104 registers); the 8-LLR routines behind dispatch8 stay out of line
    static SS_DEV void walk16(const H2 (&A)[8], uint32_t ids, uint32_t fl, uint32_t prune, H2 (&B)[8]) {
        walk16_body(A, ids, fl, prune, B);
    }
    // 32 LLRs (scale 1); ids: four pattern ids, fl: the 32 information flags
    static SS_DEV void walk32(const H2 (&A)[16], uint32_t ids, uint32_t fl, uint32_t prune, H2 (&B)[16]) {
        H2 BL[8], Bc[8];
SS_SIDE_PRAGMA
        for (int side = 0; side < 2; side++) {
            const uint32_t id2 = (ids >> (8 * side)) & 255u, fl2 = (fl >> (16 * side)) & 0xFFFFu;
            if (id2 != 0u) {
                H2 A16[8];
                if (side == 0) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        A16[i] = f2(A[i], A[i + 8]);
                        if constexpr (halve(32)) A16[i] = h2_mul(A16[i], h2_half());
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        A16[i] = h2_fma(A[i], BL[i], A[i + 8]);
                        if constexpr (sat(32)) A16[i] = h2_clamp<MAXV>(A16[i]);
                    }
                }
                walk16(A16, id2, fl2, prune, Bc);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) Bc[i] = h2_one();
            }
            if (side == 0) {
#pragma unroll
                for (int i = 0; i < 8; i++) BL[i] = Bc[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            B[i] = h2_mul(BL[i], Bc[i]);
            B[i + 8] = Bc[i];
        }
    }

    // one chunk of sign-magnitude planes -> 16 fp16x2 registers
    static SS_DEV void to_h2(const bs::Val<P>& r, H2 (&A)[16]) {
        uint32_t w[8];
#pragma unroll
        for (int p = 0; p < 7; p++) w[p] = p < P ? r.m[p] : 0u;
        w[7] = r.s & bs::nonzero<P>(r);  // a CA2 zero carries no sign
        transpose8x8x4(w);  // w[k] byte c = LLR 8c + k as (s << 7 | mag)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            A[4 * c + 0] = h2_from_sm(w[0], w[1], c);
            A[4 * c + 1] = h2_from_sm(w[2], w[3], c);
            A[4 * c + 2] = h2_from_sm(w[4], w[5], c);
            A[4 * c + 3] = h2_from_sm(w[6], w[7], c);
        }
    }
};

}  // namespace ss

// One lane's view of the decoder state.  All pointers already include the lane.
template <int Q, int LOG2PAR, bool EXT, bool PROF = false, bool XF = false>
struct SsThread {
    static constexpr int P = Q - 1;
    static constexpr int FMT = bs::FMT_CA2;
    using V = bs::Val<P>;
    using W = ss::Walk<Q, LOG2PAR, EXT>;
    static_assert(Q >= 5 && Q <= 8, "two quads per chunk");
    static_assert(LOG2PAR <= 5, "the un-saturated leaf decoder must sit inside a 32-LLR node");

    const SsParams& p;
    uint4* sm;          // the warp's shared region + lane
    uint4* wsl;         // the warp's workspace + lane
    const uint4* pl;    // channel planes of the current task + lane
    const uint32_t* sched;  // shared copy or global

    SS_DEV SsThread(const SsParams& p_) : p(p_) {}

    SS_DEV uint4* aptr(uint32_t l) const {  // not for l == p.ltm (tensor memory)
        return (l <= p.lsa ? sm : l == p.lhot ? wsh : wsl) + p.aoff[l];
    }
    // alpha[l] of the node whose first partial-sum word is wd: the root and the prefused leftmost nodes come from the
    // task's plane buffer
    SS_DEV bool in_planes(uint32_t l, uint32_t wd) const { return wd == 0u && l >= p.lpre; }
    SS_DEV const uint4* asrc(uint32_t l, uint32_t wd) const { return in_planes(l, wd) ? pl + p.poff[l] : aptr(l); }
    // partial sums of node level l, word w: component w & 3 of a quad.  The two bases are kept in registers and the
    // choice is a select: as a branch with the base recomputed from the kernel parameters this helper alone was 15 % of
    // the kernel's text and 5 % of its executed instructions
    uint4* bsm;  // sm + p.sm_beta_off
    uint4* bws;  // wsl + p.ws_beta_off
    uint4* a6;   // alpha[6]: always in shared memory, touched by every 64-LLR node
    uint4* wsh;  // the warp's block of the hot-level array + lane
    SS_DEV void bind(uint4* sm_, uint4* wsl_, uint4* wsh_) {
        sm = sm_;
        wsl = wsl_;
        wsh = wsh_;
        a6 = sm_ + p.aoff[6];
        bsm = sm_ + p.sm_beta_off;
        bws = wsl_ + p.ws_beta_off;
    }
    SS_DEV uint4* bquad(uint32_t l, uint32_t w) const {
        const bool win = l < p.lwin;
        uint4* base = win ? bsm : bws;
        const uint32_t ww = win ? (w & (p.win_words - 1u)) : w;
        return base + ((ww >> 2) * 32u);
    }
    SS_DEV uint32_t* bword(uint32_t l, uint32_t w) const { return reinterpret_cast<uint32_t*>(bquad(l, w)) + (w & 3u); }

    static SS_DEV void load(const uint4* q, V& x) {
        const uint4 a = q[0], b = q[32];
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        x.s = w[0];
#pragma unroll
        for (int k = 0; k < P; k++) x.m[k] = w[1 + k];
    }
    static SS_DEV void store(uint4* q, const V& x) {
        uint32_t w[8];
        w[0] = x.s;
#pragma unroll
        for (int k = 1; k < 8; k++) w[k] = k <= P ? x.m[k - 1] : 0u;
        q[0] = make_uint4(w[0], w[1], w[2], w[3]);
        q[32] = make_uint4(w[4], w[5], w[6], w[7]);
    }

    // ---------------------------------------------------------------- tensor memory (one alpha level, optional)
    // The level p.ltm lives in the SM's tensor memory: lane = TMEM lane, the 8 words of chunk c = columns 8c .. 8c+7
    // of the warp's column range, so one tcgen05.ld / tcgen05.st (32x32b.x8) moves a chunk for the whole warp.
#if defined(__CUDA_ARCH__)
    uint32_t tm;  // TMEM address of the warp's column range (lane quarter in bits 31:16)
    SS_DEV void tm_load(uint32_t c, V& x) const {
        uint32_t w[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "r"(tm + 8u * c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        x.s = w[0];
#pragma unroll
        for (int k = 0; k < P; k++) x.m[k] = w[1 + k];
    }
    SS_DEV void tm_store(uint32_t c, const V& x) const {
        uint32_t w[8];
        w[0] = x.s;
#pragma unroll
        for (int k = 1; k < 8; k++) w[k] = k <= P ? x.m[k - 1] : 0u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tm + 8u * c),
                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                     : "memory");
    }
    SS_DEV void tm_fence() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#else
    uint32_t* tm;  // emulation: 8 words per chunk of the lane
    SS_DEV void tm_load(uint32_t c, V& x) const {
        x.s = tm[8 * c];
        for (int k = 0; k < P; k++) x.m[k] = tm[8 * c + 1 + k];
    }
    SS_DEV void tm_store(uint32_t c, const V& x) const {
        tm[8 * c] = x.s;
        for (int k = 1; k < 8; k++) tm[8 * c + k] = k <= P ? x.m[k - 1] : 0u;
    }
    SS_DEV void tm_fence() const {}
#endif

    // ---------------------------------------------------------------- nodes of 128 LLRs and more
    // One pair of input chunks (c, c + half) of alpha[l]
    template <bool SRC_TM>
    SS_DEV void load_pair(const uint4* src, uint32_t c, uint32_t half, V& a, V& b) const {
        if constexpr (SRC_TM) {
            tm_load(c, a);
            tm_load(c + half, b);
        } else {
            load(src + c * 64u, a);
            load(src + (c + half) * 64u, b);
        }
    }
    template <bool DST_TM>
    SS_DEV void store_chunk(uint4* dst, uint32_t c, const V& r) const {
        if constexpr (DST_TM)
            tm_store(c, r);
        else
            store(dst + c * 64u, r);
    }
    // alpha[l-1] chunk c = f(alpha[l] chunk c, chunk c + half)                   F_STATE my_module.h:373-445
    // alpha[l-1] chunk c = g(..., beta word c of the left child), beta = 0 when it is all-frozen   G_STATE :704-781
    // Two chunks per trip, software-pipelined: the loads of the next chunk are in flight while this one is computed
    // (alpha levels in the workspace / plane buffer answer from L2 or DRAM).
    template <bool G, bool SRC_TM, bool DST_TM>
    SS_DEV void op_x(uint32_t l, uint32_t wd, bool zero) {
        const uint4* src = SRC_TM ? nullptr : asrc(l, wd);
        uint4* dst = DST_TM ? nullptr : aptr(l - 1);
        const uint32_t half = 1u << (l - 6);
        const uint32_t* ub = (G && !zero) ? bword(l - 1, wd) : nullptr;  // words wd .. wd + half - 1 lie in one storage
        V a0, b0, a1, b1, r;
        if constexpr (!SRC_TM && !DST_TM) {
            // source and destination in global / shared memory (the levels that stream through DRAM at c2 and above): four
            // chunk pairs per trip, three of them in flight while one is computed (half >= 4: f / g of a 128-LLR node is
            // always the fused SS_XS, never this op).  c2 456 -> 471 Gb/s at 128 registers.  The same depth in the
            // tensor-memory variants costs more in code than it hides in latency (c1 396 -> 388, c2 471 -> 447), and so
            // does a second, shallower loop inlined beside this one (c1 379, c2 438).
            V a2, b2, a3, b3;
            load_pair<false>(src, 0, half, a0, b0);
            load_pair<false>(src, 1, half, a1, b1);
            load_pair<false>(src, 2, half, a2, b2);
            for (uint32_t c = 0; c < half; c += 4) {
                load_pair<false>(src, c + 3u, half, a3, b3);
                uint4 u = make_uint4(0u, 0u, 0u, 0u);
                if (G && !zero) u = *reinterpret_cast<const uint4*>(ub + (c >> 2) * 128u);
                const bool more = c + 4u < half;
                if constexpr (G) bs::g_sat_ca2<P>(a0, b0, u.x, r); else bs::f_op<P>(a0, b0, r);
                store(dst + c * 64u, r);
                if (more) load_pair<false>(src, c + 4u, half, a0, b0);
                if constexpr (G) bs::g_sat_ca2<P>(a1, b1, u.y, r); else bs::f_op<P>(a1, b1, r);
                store(dst + (c + 1u) * 64u, r);
                if (more) load_pair<false>(src, c + 5u, half, a1, b1);
                if constexpr (G) bs::g_sat_ca2<P>(a2, b2, u.z, r); else bs::f_op<P>(a2, b2, r);
                store(dst + (c + 2u) * 64u, r);
                if (more) load_pair<false>(src, c + 6u, half, a2, b2);
                if constexpr (G) bs::g_sat_ca2<P>(a3, b3, u.w, r); else bs::f_op<P>(a3, b3, r);
                store(dst + (c + 3u) * 64u, r);
            }
            return;
        }
        load_pair<SRC_TM>(src, 0, half, a0, b0);
        for (uint32_t c = 0; c < half; c += 2) {
            load_pair<SRC_TM>(src, c + 1u, half, a1, b1);
            uint2 u = make_uint2(0u, 0u);
            if (G && !zero) u = *reinterpret_cast<const uint2*>(ub + (c >> 2) * 128u + (c & 3u));
            if constexpr (G)
                bs::g_sat_ca2<P>(a0, b0, u.x, r);
            else
                bs::f_op<P>(a0, b0, r);
            store_chunk<DST_TM>(dst, c, r);
            if (c + 2u < half) load_pair<SRC_TM>(src, c + 2u, half, a0, b0);
            if constexpr (G)
                bs::g_sat_ca2<P>(a1, b1, u.y, r);
            else
                bs::f_op<P>(a1, b1, r);
            store_chunk<DST_TM>(dst, c + 1u, r);
        }
        if constexpr (DST_TM) tm_fence();
    }
    template <bool G>
    SS_DEV void op_fg(uint32_t l, uint32_t wd, bool zero) {
        if (l == p.ltm && !in_planes(l, wd))
            op_x<G, true, false>(l, wd, zero);
        else if (l == p.ltm + 1u)
            op_x<G, false, true>(l, wd, zero);
        else
            op_x<G, false, false>(l, wd, zero);
    }
    // SS_XF_*: alpha[l-1] of the child = f / g of alpha[l], and alpha[l-2] = f of that child, in one pass: output chunk c of
    // level l-2 needs chunks c and c + q of level l-1, i.e. chunks c, c + q, c + 2q, c + 3q of level l.  Both levels
    // are stored (their g ops come later); alpha[l-1] is not read back.  All three levels are in global memory.
    template <bool G>
    SS_DEV void op_xf(uint32_t l, uint32_t wd, bool zero) {
        const uint4* src = asrc(l, wd);
        uint4* d1 = aptr(l - 1);
        uint4* d2 = aptr(l - 2);
        const uint32_t half = 1u << (l - 6), q = half >> 1;
        const uint32_t* ub = (G && !zero) ? bword(l - 1, wd) : nullptr;
        V a0, b0, a1, b1;
        load_pair<false>(src, 0, half, a0, b0);
        load_pair<false>(src, q, half, a1, b1);
        for (uint32_t c = 0; c < q; c++) {
            V r0, r1, r2;
            uint32_t u0 = 0u, u1 = 0u;
            if (G && !zero) {
                u0 = ub[(c >> 2) * 128u + (c & 3u)];
                u1 = ub[((c + q) >> 2) * 128u + ((c + q) & 3u)];
            }
            if constexpr (G) {
                bs::g_sat_ca2<P>(a0, b0, u0, r0);
                bs::g_sat_ca2<P>(a1, b1, u1, r1);
            } else {
                bs::f_op<P>(a0, b0, r0);
                bs::f_op<P>(a1, b1, r1);
            }
            if (c + 1u < q) {  // the next unit's loads are in flight during the stores and the second-level f
                load_pair<false>(src, c + 1u, half, a0, b0);
                load_pair<false>(src, c + 1u + q, half, a1, b1);
            }
            store(d1 + c * 64u, r0);
            store(d1 + (c + q) * 64u, r1);
            bs::f_op<P>(r0, r1, r2);
            store(d2 + c * 64u, r2);
        }
    }
    // node (l, wd) := (left ^ right, right); copy: the left child is all-frozen      H_STATE my_module.h:903-932
    SS_DEV void op_h(uint32_t l, uint32_t wd, bool copy) {
        const uint32_t nw = 1u << (l - 6);  // words per half
        if (nw == 2u) {  // l = 7: the node is one quad, always inside the window
            uint4* q = bquad(7, wd);
            uint4 x = *q;
            if (copy) {
                x.x = x.z;
                x.y = x.w;
            } else {
                x.x ^= x.z;
                x.y ^= x.w;
            }
            *q = x;
            return;
        }
        // children in the window and the node in the workspace: the right half moves along
        const bool move = l == p.lwin;
        const uint4* cl = bquad(l - 1, wd);
        const uint4* cr = bquad(l - 1, wd + nw);
        uint4* d = bquad(l, wd);
        for (uint32_t i = 0; i < nw; i += 4) {
            uint4 x = cr[i * 8u];  // quad index advances by 32 per 4 words
            if (move) d[(nw + i) * 8u] = x;
            if (!copy) {
                const uint4 y = cl[i * 8u];
                x.x ^= y.x;
                x.y ^= y.y;
                x.z ^= y.z;
                x.w ^= y.w;
            }
            d[i * 8u] = x;
        }
    }
    SS_DEV void op_r0(uint32_t l, uint32_t wd) {
        if (l == 6u) {
            *reinterpret_cast<uint2*>(bword(6, wd)) = make_uint2(0u, 0u);
            return;
        }
        const uint32_t nw = 1u << (l - 5);
        for (uint32_t i = 0; i < nw; i += 4) *bquad(l, wd + i) = make_uint4(0u, 0u, 0u, 0u);
    }
    // hard decision of alpha[l]; returns true (warp-uniform) when some LLR of some frame is zero
    template <bool SRC_TM>
    SS_DEV bool op_hd_x(uint32_t l, uint32_t wd) {
        const uint4* src = SRC_TM ? nullptr : asrc(l, wd);
        const uint32_t nc = 1u << (l - 5);
        uint4* d = bquad(l, wd);
        uint32_t z = 0u;
        for (uint32_t c = 0; c < nc; c += 4) {
            uint32_t x[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                V a;
                if constexpr (SRC_TM)
                    tm_load(c + k, a);
                else
                    load(src + (c + k) * 64u, a);
                const uint32_t nz = bs::nonzero<P>(a);
                z |= ~nz;
                x[k] = a.s & nz;
            }
            d[c * 8u] = make_uint4(x[0], x[1], x[2], x[3]);
        }
        return SS_ANY(z != 0u);
    }
    SS_DEV bool op_hd(uint32_t l, uint32_t wd) {
        return (l == p.ltm && !in_planes(l, wd)) ? op_hd_x<true>(l, wd) : op_hd_x<false>(l, wd);
    }

    // PROF build only: how often the R1 shortcut (hard decision = sign) is voted on at nodes of 2^l LLRs and how often a
    // zero LLR somewhere in the warp sends the node to the full walk instead (scpd_r1_votes; profiled warps only)
    SS_DEV void prof_vote(uint32_t l, bool fallback) {
#if defined(__CUDA_ARCH__)
        if (PROF && prof_on && (threadIdx.x & 31u) == 0u) {
            atomicAdd(p.prof + 384u + l, 1ull);
            if (fallback) atomicAdd(p.prof + 416u + l, 1ull);
        }
#else
        (void)l;
        (void)fallback;
#endif
    }
    // ---------------------------------------------------------------- nodes of 64 and 32 LLRs
    // a, b: the two chunks of alpha[6] (also stored at aptr(6) when the node needs them again for g)
    // d0: node types (64, left 32, right 32) + kind + pruning flag; d1: the 64 flags' low word, d2: high word,
    // d3: pattern ids of the eight 8-LLR nodes (4 bits each)
    SS_DEV void sub64(uint32_t wd, const V& a, const V& b, bool stored, uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3) {
        const uint32_t t64 = d0 & 3u, tl = (d0 >> 2) & 3u, tr = (d0 >> 4) & 3u, prune = (d0 >> 10) & 1u;
        uint32_t bl = 0u, br = 0u;
        bool walk = true;
        if (t64 == SS_T_R1) {
            const uint32_t nza = bs::nonzero<P>(a), nzb = bs::nonzero<P>(b);
            if (!SS_ANY((~nza | ~nzb) != 0u)) {
                *reinterpret_cast<uint2*>(bword(6, wd)) = make_uint2(a.s, b.s);
                walk = false;
            }
            prof_vote(6u, walk);
        }
        if (walk) {
            // both children in one loop that is not unrolled: the 32-LLR walker (plane -> fp16x2 conversion, walk32, sign-bit
            // packing) is inlined exactly once, with no call and no argument moves; alpha[6] is re-read from shared memory
            // by either side instead of staying in registers (as an out-of-line function with twelve register arguments
            // the walker cost 112 instead of 94 registers and c2 ran at 449 instead of 455 Gb/s)
            if (!stored) {
                store(a6, a);
                store(a6 + 64, b);
            }
            _Pragma("unroll 1")
            for (uint32_t side = 0; side < 2u; side++) {
                const uint32_t ts = side ? tr : tl;
                if (ts == SS_T_R0) continue;
                V a2, b2, r;
                load(a6, a2);
                load(a6 + 64, b2);
                if (side == 0u)
                    bs::f_op<P>(a2, b2, r);
                else
                    bs::g_sat_ca2<P>(a2, b2, bl, r);  // bl = 0 when the left child is all-frozen
                uint32_t bits;
                bool hard = false;
                if (ts == SS_T_R1) {
                    const uint32_t nz = bs::nonzero<P>(r);
                    hard = !SS_ANY(~nz != 0u);  // no zero anywhere in the warp: hd = sign
                    prof_vote(5u, !hard);
                }
                if (hard) {
                    bits = r.s;
                } else {
                    ss::H2 A[16], B[16];
                    W::to_h2(r, A);
                    W::walk32(A, side ? d3 >> 16 : d3 & 0xFFFFu, side ? d2 : d1, prune, B);
                    bits = ss::h2_pack16(B);
                }
                if (side) br = bits; else bl = bits;
            }
            *reinterpret_cast<uint2*>(bword(6, wd)) = make_uint2(bl ^ br, br);
        }
    }
    // SS_SUB: the 64-LLR node from alpha[6];  SS_XS: f / g / g0 of the 128-LLR node above, straight into the child
    SS_DEV void op_sub(uint32_t wd, uint32_t pc, bool xs) {
        const uint32_t d0 = sched[pc + 1], kind = (d0 >> 8) & 3u;
        V r0, r1;
        if (!xs) {
            load(a6, r0);
            load(a6 + 64, r1);
        } else {
            const uint4* src = asrc(7, wd & ~3u);  // N = 128: the root
            V a0, b0, a1, b1;
            load(src, a0);
            load(src + 128, b0);
            load(src + 64, a1);
            load(src + 192, b1);
            if (kind == SS_F) {
                bs::f_op<P>(a0, b0, r0);
                bs::f_op<P>(a1, b1, r1);
            } else {
                uint2 u = make_uint2(0u, 0u);
                if (kind == SS_G) u = *reinterpret_cast<const uint2*>(bword(6, wd & ~3u));
                bs::g_sat_ca2<P>(a0, b0, u.x, r0);
                bs::g_sat_ca2<P>(a1, b1, u.y, r1);
            }
        }
        sub64(wd, r0, r1, !xs, d0, sched[pc + 2], sched[pc + 3], sched[pc + 4]);
    }

    bool prof_on = false;
    uint32_t prof_fn = 6u, prof_l = 0u;
    long long prof_t0 = 0;
    SS_DEV void run() {
        uint32_t pc = 0;
        prof_fn = 6u;
        for (;;) {
            const uint32_t w = sched[pc];
            const uint32_t code = ss_op_code(w), l = ss_op_level(w), wd = ss_op_word(w);
#if defined(__CUDA_ARCH__)
            // measured counterpart of the reference's function x level monitor (sc_monitor.h:50-441): cycles between
            // the fetch of this op and the fetch of the next, for one warp per CTA
            if (PROF && prof_on) {
                const long long now = clock64();
                if (prof_fn < 6u && (threadIdx.x & 31u) == 0u) {
                    atomicAdd(p.prof + prof_fn * 32u + prof_l, (unsigned long long)(now - prof_t0));
                    atomicAdd(p.prof + 192u + prof_fn * 32u + prof_l, 1ull);
                }
                // rows: 0 F, 1 G (g and g0), 2 H, 3 R (64-LLR node incl. a fused level-7 op), 4 R0, 5 R1 (hard decision)
                prof_fn = (code == SS_F || code == SS_XF_F) ? 0u : (code == SS_G || code == SS_G0 || code == SS_XF_G || code == SS_XF_G0) ? 1u : (code == SS_H || code == SS_HCOPY) ? 2u
                          : (code == SS_SUB || code == SS_XS) ? 3u : code == SS_R0 ? 4u : code == SS_R1 ? 5u : 6u;
                prof_l = (code == SS_SUB || code == SS_XS) ? 6u : l;
                prof_t0 = clock64();
            }
#endif
            switch (code) {
                case SS_END: return;
                case SS_F:
                    op_fg<false>(l, wd, false);
                    pc++;
                    break;
                case SS_G:
                case SS_G0:
                    op_fg<true>(l, wd, code == SS_G0);
                    pc++;
                    break;
                case SS_XF_F:
                case SS_XF_G:
                case SS_XF_G0:
                    if constexpr (XF) {
                        if (code == SS_XF_F)
                            op_xf<false>(l, wd, false);
                        else
                            op_xf<true>(l, wd, code == SS_XF_G0);
                        pc++;
                        break;
                    } else {
                        return;  // a schedule with fused ops is never given to the lean kernel
                    }
                case SS_H:
                    op_h(l, wd, false);
                    pc++;
                    break;
                case SS_HCOPY:
                    op_h(l, wd, true);
                    pc++;
                    break;
                case SS_R0:
                    op_r0(l, wd);
                    pc++;
                    break;
                case SS_R1: {
                    const bool z = op_hd(l, wd);
                    prof_vote(l, z);
                    pc += 2u + (z ? 0u : sched[pc + 1]);
                    break;
                }
                case SS_SUB:
                case SS_XS:
                    op_sub(wd, pc, code == SS_XS);
                    pc += 5;
                    break;
                default: return;
            }
        }
    }
    // the root's partial sums are the packed row of the lane's frame (wrapper_out.h:31-33 laid end to end)
    SS_DEV void write_output(unsigned long long frame) {
        if (frame >= p.nframes) return;
        uint4* out = reinterpret_cast<uint4*>(p.xhat + frame * p.wpf);
        const uint4* b = bquad(p.log2n, 0);
#pragma unroll 1
        for (uint32_t q = 0; q < p.wpf / 4u; q++) out[q] = b[q * 32u];
    }
};

// int8 rows -> sign-magnitude planes in the chunk layout of the decode kernel (wrapper_in.h:30-42), plus the first D
// f levels of the tree, which depend on the channel alone: alpha[log2n - s] of the leftmost node, s = 1 .. D, so that the
// walk kernel neither runs those f ops nor reads the channel planes for them (the conversion is bound by DRAM, its ALUs
// are idle).  One call = one lane (frame) and one unit of 8 chunks: 2^D groups, C >> D chunks apart, of 8 >> D
// consecutive chunks, i.e. every operand of the D-level f tree over them.  Level s pairs entries 8 >> s apart.
struct SsPre {
    uint32_t off[4];  // off[s]: uint4 offset of alpha[log2n - s] of the leftmost node in the task's plane buffer
};
namespace ss {
template <int P, int D>
SS_DEV void planes_unit(const int8_t* row, bool valid, uint32_t nchunks, uint32_t unit, uint4* dst, const SsPre& pre) {
    constexpr uint32_t PER = 8u >> D;
    const uint32_t gstride = nchunks >> D, cb = unit * PER;
    bs::Val<P> x[8];
    auto put = [&](uint4* base, uint32_t c, const bs::Val<P>& r) {
        uint32_t o[8];
        o[0] = r.s;
#pragma unroll
        for (int k = 1; k < 8; k++) o[k] = k <= P ? r.m[k - 1] : 0u;
        base[(2u * c) * 32u] = make_uint4(o[0], o[1], o[2], o[3]);
        base[(2u * c + 1u) * 32u] = make_uint4(o[4], o[5], o[6], o[7]);
    };
#pragma unroll
    for (uint32_t a = 0; a < 8u; a++) {
        const uint32_t c = (a / PER) * gstride + cb + (a % PER);
        uint32_t v[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (valid) {
#if defined(__CUDA_ARCH__)
            const uint4 lo = __ldg(reinterpret_cast<const uint4*>(row + 32u * c));
            const uint4 hi = __ldg(reinterpret_cast<const uint4*>(row + 32u * c) + 1);
            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
            v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
#else
            for (int k = 0; k < 8; k++) {
                uint32_t t = 0;
                for (int b = 0; b < 4; b++) t |= (uint32_t)(uint8_t)row[32u * c + 4 * k + b] << (8 * b);
                v[k] = t;
            }
#endif
        }
        chunk_planes<P>(v, x[a]);
        put(dst, c, x[a]);
    }
#pragma unroll
    for (int s = 1; s <= D; s++) {
        const uint32_t dist = 8u >> s;
#pragma unroll
        for (uint32_t a = 0; a < dist; a++) {
            bs::Val<P> r;
            bs::f_op<P>(x[a], x[a + dist], r);
            x[a] = r;
            put(dst + pre.off[s], (a / PER) * gstride + cb + (a % PER), r);
        }
    }
}
}  // namespace ss

#if defined(__CUDACC__)
// One warp per (task, unit); lane = frame: 32 bytes of the lane's row per chunk.
template <int Q, int D>
__global__ void __launch_bounds__(256) ss_planes_kernel(const int8_t* __restrict__ llr, unsigned long long nframes, uint32_t n,
                                                        uint4* __restrict__ planes, unsigned long long planes_stride,
                                                        const SsPre pre) {
    constexpr int P = Q - 1;
    const int lane = threadIdx.x & 31;
    const unsigned long long ntasks = (nframes + 31) / 32;
    const uint32_t units_per_task = n / 256u;  // n >= 256; n = 128 handled by the tail below (D = 0)
    const unsigned long long nunits = ntasks * (units_per_task ? units_per_task : 1u);
    const unsigned long long wstride = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    for (unsigned long long t = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < nunits; t += wstride) {
        const unsigned long long task = units_per_task ? t / units_per_task : t;
        const unsigned long long f = task * 32ull + lane;
        const bool valid = f < nframes;
        const int8_t* row = llr + (valid ? f : 0ull) * n;
        uint4* dst = planes + task * planes_stride + lane;
        if (units_per_task) {
            ss::planes_unit<P, D>(row, valid, n / 32u, (uint32_t)(t % units_per_task), dst, pre);
        } else {
            for (uint32_t c = 0; c < n / 32u; c++) {
                uint32_t v[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                if (valid) {
                    const uint4 x = __ldg(reinterpret_cast<const uint4*>(row + 32u * c));
                    const uint4 y = __ldg(reinterpret_cast<const uint4*>(row + 32u * c) + 1);
                    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
                    v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
                }
                bs::Val<P> x;
                ss::chunk_planes<P>(v, x);
                uint32_t o[8];
                o[0] = x.s;
#pragma unroll
                for (int k = 1; k < 8; k++) o[k] = k <= P ? x.m[k - 1] : 0u;
                dst[(2u * c) * 32u] = make_uint4(o[0], o[1], o[2], o[3]);
                dst[(2u * c + 1u) * 32u] = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
    }
}

#ifndef SCPD_SS_THREADS
#define SCPD_SS_THREADS 512  // upper bound of the CTA size: 16 warps, one CTA per SM, up to 128 registers
#endif
// PROF: the build with the per-stage clock64() histogram (scpd_stage_timing); the production kernel carries none of it
// XF: the build that also knows the fused SS_XF_* ops (large trees); the lean kernel of the small ones carries none of it
template <int Q, int LOG2PAR, bool EXT, bool PROF = false, bool XF = false>
__global__ void __launch_bounds__(SCPD_SS_THREADS, 1) sc_decode_ss_kernel(const SsParams p) {
    extern __shared__ __align__(16) uint4 ss_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    SsThread<Q, LOG2PAR, EXT, PROF, XF> t(p);
    uint4* const sm_warp = ss_smem + (size_t)warp * p.sm_stride + lane;
    uint32_t* sm_sched = reinterpret_cast<uint32_t*>(ss_smem + (size_t)nwarps * p.sm_stride);
    // tensor memory for the alpha level p.ltm: warp w owns lanes 32 (w % 4) .. +31 (the only ones it can reach) and the
    // column range (w / 4) * tm_cols .. of the CTA's allocation
    uint32_t tm_base = 0u, tm_alloc_cols = 0u;
    if (p.ltm) {
        tm_alloc_cols = 32u;
        while (tm_alloc_cols < p.tm_cols * (uint32_t)((nwarps + 3) / 4)) tm_alloc_cols <<= 1;
        uint32_t* slot = sm_sched + p.sched_words;
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(slot)),
                         "r"(tm_alloc_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tm_base = *reinterpret_cast<volatile uint32_t*>(slot);
        t.tm = tm_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * p.tm_cols;
        __syncthreads();  // the slot is re-used below
    }
    if (p.sched_words) {
        for (uint32_t i = threadIdx.x; i < p.sched_words; i += blockDim.x) sm_sched[i] = __ldg(p.sched + i);
        __syncthreads();
        t.sched = sm_sched;
    } else {
        t.sched = p.sched;
    }
    t.prof_on = PROF && p.prof != nullptr && warp == 0;
    const unsigned long long slot_id = (unsigned long long)blockIdx.x * nwarps + warp;
    t.bind(sm_warp, p.ws + slot_id * p.ws_stride + lane, p.hot + slot_id * p.hot_stride + lane);
    for (unsigned long long task = slot_id; task < p.ntasks; task += (unsigned long long)gridDim.x * nwarps) {
        t.pl = p.planes + task * p.planes_stride + lane;
        t.run();
        t.write_output(task * 32ull + lane);
    }
    if (p.ltm) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "r"(tm_alloc_cols) : "memory");
    }
}
#endif

}  // namespace scpd
