// decode_fast.cuh -- the throughput kernel for the headline settings (CA2, Q <= 8, PAR <= 8G).
//
// Layout of the work
//   * A group of G lanes owns one FRAME PAIR (two frames in the two int16 halves of a register);
//     a warp owns 32/G pairs that run the same host-built schedule in lock step.
//   * Nodes of size >= 16G ("memory levels"): the LLR stack is kept as int8 CELLS
//     (byte 0 = frame A, byte 1 = frame B; values are saturated to +-(2^(Q-1)-1) <= 127 there) in
//     shared memory -- or, for the largest levels, in an L2-resident workspace -- and moved with
//     128-bit accesses, 8 cells per lane; PRMT unpacks a cell to int16x2 and packs it back.
//     Partial sums use the same cell format with 0x00 / 0xFF bytes, so the PRMT that unpacks a cell
//     also expands a partial sum to the 0x0000 / 0xFFFF mask g needs, and h is a plain XOR of
//     packed words.
//   * Nodes of size S = 8G and below: one fully unrolled register routine per subtree.  Lane gl
//     holds elements gl, gl+G, ... so every f/g pair (i, i + n/2) is lane-local down to n = 2G;
//     the last log2(G) levels use __shfl_xor_sync.  The PAR-wide leaf decoder of the reference
//     (un-saturated g when EXTENDED) lives entirely inside this routine.  Node types (all-frozen,
//     all-information, mixed) come from a 2-bit-per-node descriptor in the schedule: branches are
//     warp-uniform.
//
// Bit-exactness: f = F_function_C2 (functions.h:48-61), g = G_function_C2 / G_extended_C2
// (:63-88), size-2 terminal = Spec_P2 (:367-384); all-information nodes are replaced by the hard
// decision only when no input LLR is 0 (checked with a warp vote), all-frozen
// nodes are skipped (their partial sums are 0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "schedule.h"

namespace scpd {

struct FastParams {
    const uint32_t* sched;
    const int8_t* llr;
    uint32_t* xhat;
    unsigned long long nframes, num_fp;
    uint32_t n, log2n, wpf;
    uint32_t satv;
    uint32_t lsa;            // alpha levels log2S .. lsa in shared memory, above in the workspace
    uint32_t lsb;            // partial sums of nodes up to level lsb in shared memory (block of 2^(lsb+1) cells)
    uint32_t sm_alpha_cells; // cells reserved for alpha in shared memory, per frame pair
    uint32_t sm_stride;      // bytes between the shared-memory regions of consecutive frame pairs
    uint8_t* ws;             // workspace: per resident frame pair  [alpha: 2n cells][beta: n cells]
    unsigned long long ws_stride;  // bytes
};

enum : uint32_t { T_MIX = 0, T_R0 = 1, T_R1 = 2, T_X = 3 };

// PRMT with 4-bit selectors: bit 3 of a selector replicates the sign of the selected byte.
// (__byte_perm masks the selectors to 3 bits, so the instruction is issued through PTX.)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
#if defined(__CUDACC__)
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
#else
    return cuda_emu_prmt(a, b, s);  // CPU emulation build (tests/emu)
#endif
}

// ---------------------------------------------------------------- int16x2 arithmetic
__device__ __forceinline__ uint32_t signmask2(uint32_t v) { return prmt(v, 0u, 0xBB99); }  // 0xFFFF where half < 0
// f = max(min(a,b), -max(a,b));  -x = ~x + 1 folded into VIADDMNMX: max(~mx + 1, mn)
__device__ __forceinline__ uint32_t f16x2(uint32_t a, uint32_t b) {
    const uint32_t mn = __vmins2(a, b);
    const uint32_t mx = __vmaxs2(a, b);
    return __viaddmax_s16x2(~mx, 0x00010001u, mn);
}
// g with partial-sum mask m (0xFFFF -> b - a, 0 -> b + a):  b + (a ^ m) + (m & 1)
__device__ __forceinline__ uint32_t g16x2_sat(uint32_t a, uint32_t b, uint32_t m, uint32_t satp, uint32_t satn) {
    const uint32_t t = __vadd2(b, a ^ m);
    const uint32_t r = __viaddmin_s16x2(t, m & 0x00010001u, satp);
    return __vmaxs2(r, satn);
}
__device__ __forceinline__ uint32_t g16x2_nosat(uint32_t a, uint32_t b, uint32_t m) {
    return __vadd2(__vadd2(b, a ^ m), m & 0x00010001u);
}
__device__ __forceinline__ uint32_t g016x2_sat(uint32_t a, uint32_t b, uint32_t satp, uint32_t satn) {
    return __vmaxs2(__viaddmin_s16x2(a, b, satp), satn);
}
// accumulates the "some half is zero" test: bit 15/31 of the result set if a half of v is 0
__device__ __forceinline__ uint32_t zacc2(uint32_t z, uint32_t v) { return z | ((v - 0x00010001u) & ~v); }

// ---------------------------------------------------------------- cells
__device__ __forceinline__ void unpack8(const uint4 w, uint32_t (&v)[8]) {
    v[0] = prmt(w.x, 0u, 0x9180);
    v[1] = prmt(w.x, 0u, 0xB3A2);
    v[2] = prmt(w.y, 0u, 0x9180);
    v[3] = prmt(w.y, 0u, 0xB3A2);
    v[4] = prmt(w.z, 0u, 0x9180);
    v[5] = prmt(w.z, 0u, 0xB3A2);
    v[6] = prmt(w.w, 0u, 0x9180);
    v[7] = prmt(w.w, 0u, 0xB3A2);
}
__device__ __forceinline__ uint4 pack8(const uint32_t (&v)[8]) {
    uint4 w;
    w.x = prmt(v[0], v[1], 0x6420);
    w.y = prmt(v[2], v[3], 0x6420);
    w.z = prmt(v[4], v[5], 0x6420);
    w.w = prmt(v[6], v[7], 0x6420);
    return w;
}
// two int8 rows (frames A and B, 8 consecutive LLRs each) -> 8 cells
__device__ __forceinline__ uint4 rows_to_cells(const uint2 a, const uint2 b) {
    uint4 w;
    w.x = prmt(a.x, b.x, 0x5140);
    w.y = prmt(a.x, b.x, 0x7362);
    w.z = prmt(a.y, b.y, 0x5140);
    w.w = prmt(a.y, b.y, 0x7362);
    return w;
}

template <int G, int LOG2PAR, bool EXT>
struct FastDecoder {
    static constexpr unsigned FULL = 0xFFFFFFFFu;
    static constexpr int S = 8 * G;
    static constexpr int LOG2S = (G == 1) ? 3 : (G == 2) ? 4 : (G == 4) ? 5 : (G == 8) ? 6 : (G == 16) ? 7 : 8;
    static constexpr int DESC_WORDS = (2 * (S - 1) + 31) / 32;

    const FastParams& p;
    uint16_t* sm_alpha;  // level l at cell offset (1 << l)
    uint16_t* sm_beta;   // block of 2^(lsb+1) cells
    uint16_t* gl_alpha;  // workspace, level l at cell offset (1 << l)
    uint16_t* gl_beta;   // workspace, absolute positions
    const int8_t* llrA;
    const int8_t* llrB;
    int gl;
    uint32_t satp, satn;
    uint32_t desc[DESC_WORDS];

    __device__ FastDecoder(const FastParams& p_) : p(p_) {}

    // ------------------------------------------------------------ storage
    __device__ __forceinline__ uint16_t* alpha(int l) const {
        return ((uint32_t)l <= p.lsa ? sm_alpha : gl_alpha) + (1u << l);
    }
    // partial sums of the node (l, o)
    __device__ __forceinline__ uint16_t* beta(int l, uint32_t o) const {
        return (uint32_t)l <= p.lsb ? sm_beta + (o & ((2u << p.lsb) - 1u)) : gl_beta + o;
    }
    __device__ __forceinline__ uint4 ld_cells(int l, uint32_t i) const {
        if ((uint32_t)l == p.log2n) {
            const uint2 a = __ldg(reinterpret_cast<const uint2*>(llrA + i));
            const uint2 b = __ldg(reinterpret_cast<const uint2*>(llrB + i));
            return rows_to_cells(a, b);
        }
        return *reinterpret_cast<const uint4*>(alpha(l) + i);
    }

    // ------------------------------------------------------------ memory-level operations
    __device__ void op_f(int l) {
        const uint32_t h = 1u << (l - 1);
        uint16_t* dst = alpha(l - 1);
        for (uint32_t i = 8u * gl; i < h; i += 8u * G) {
            uint32_t a[8], b[8], r[8];
            unpack8(ld_cells(l, i), a);
            unpack8(ld_cells(l, i + h), b);
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] = f16x2(a[j], b[j]);
            *reinterpret_cast<uint4*>(dst + i) = pack8(r);
        }
        __syncwarp();
    }
    template <bool ZERO>
    __device__ void op_g(int l, uint32_t o) {
        const uint32_t h = 1u << (l - 1);
        uint16_t* dst = alpha(l - 1);
        const uint16_t* bsrc = ZERO ? nullptr : beta(l - 1, o);
        for (uint32_t i = 8u * gl; i < h; i += 8u * G) {
            uint32_t a[8], b[8], m[8], r[8];
            unpack8(ld_cells(l, i), a);
            unpack8(ld_cells(l, i + h), b);
            if (!ZERO) unpack8(*reinterpret_cast<const uint4*>(bsrc + i), m);
#pragma unroll
            for (int j = 0; j < 8; j++)
                r[j] = ZERO ? g016x2_sat(a[j], b[j], satp, satn) : g16x2_sat(a[j], b[j], m[j], satp, satn);
            *reinterpret_cast<uint4*>(dst + i) = pack8(r);
        }
        __syncwarp();
    }
    // node (l,o) := (left ^ right, right) of its children (l-1,o), (l-1,o+h); COPY: left child all-frozen
    template <bool COPY>
    __device__ void op_h(int l, uint32_t o) {
        const uint32_t h = 1u << (l - 1);
        const uint16_t* cl = beta(l - 1, o);
        const uint16_t* cr = beta(l - 1, o + h);
        uint16_t* d = beta(l, o);
        const bool moved = (d != cl);  // crossing from the shared block to the workspace
        for (uint32_t i = 8u * gl; i < h; i += 8u * G) {
            uint4 x = *reinterpret_cast<const uint4*>(cr + i);
            if (moved) *reinterpret_cast<uint4*>(d + h + i) = x;
            if (!COPY) {
                const uint4 y = *reinterpret_cast<const uint4*>(cl + i);
                x.x ^= y.x;
                x.y ^= y.y;
                x.z ^= y.z;
                x.w ^= y.w;
            }
            *reinterpret_cast<uint4*>(d + i) = x;
        }
        __syncwarp();
    }
    __device__ void op_r0(int l, uint32_t o) {
        uint16_t* d = beta(l, o);
        for (uint32_t i = 8u * gl; i < (1u << l); i += 8u * G) *reinterpret_cast<uint4*>(d + i) = make_uint4(0, 0, 0, 0);
        __syncwarp();
    }
    // hard decision of node (l,o); returns true (warp-uniform) if some LLR of some pair was 0
    __device__ bool op_hd(int l, uint32_t o) {
        uint16_t* d = beta(l, o);
        uint32_t z = 0;
        for (uint32_t i = 8u * gl; i < (1u << l); i += 8u * G) {
            uint4 w = ld_cells(l, i);
            z |= ((w.x - 0x01010101u) & ~w.x) | ((w.y - 0x01010101u) & ~w.y) | ((w.z - 0x01010101u) & ~w.z) |
                 ((w.w - 0x01010101u) & ~w.w);
            w.x = prmt(w.x, 0u, 0xBA98);
            w.y = prmt(w.y, 0u, 0xBA98);
            w.z = prmt(w.z, 0u, 0xBA98);
            w.w = prmt(w.w, 0u, 0xBA98);
            *reinterpret_cast<uint4*>(d + i) = w;
        }
        const bool any = __any_sync(FULL, (z & 0x80808080u) != 0u);
        __syncwarp();
        return any;
    }

    // ------------------------------------------------------------ register subtree
    template <int HEAP>
    __device__ __forceinline__ uint32_t ntype() const {
        return (desc[(2 * HEAP) >> 5] >> ((2 * HEAP) & 31)) & 3u;
    }
    template <int NODE>
    __device__ __forceinline__ uint32_t gfun(uint32_t a, uint32_t b, uint32_t m) const {
        if (EXT && NODE <= (1 << LOG2PAR)) return g16x2_nosat(a, b, m);
        return g16x2_sat(a, b, m, satp, satn);
    }
    template <int NODE>
    __device__ __forceinline__ uint32_t g0fun(uint32_t a, uint32_t b) const {
        if (EXT && NODE <= (1 << LOG2PAR)) return __vadd2(a, b);
        return g016x2_sat(a, b, satp, satn);
    }
    // Spec_P2 on (a, b) for both frames at once; returns the two partial sums as masks
    __device__ __forceinline__ void p2(uint32_t t, uint32_t a, uint32_t b, uint32_t& x0, uint32_t& x1) const {
        if (t == T_R0) {  // flags (0,0)
            x0 = 0u;
            x1 = 0u;
        } else if (t == T_R1) {  // flags (1,1): u0^u1 = hd(a), u1 = hd(b) for every input, zeros included
            x0 = signmask2(a);
            x1 = signmask2(b);
        } else if (t == T_MIX) {  // flags (0,1): u0 = 0, u1 = sign(b + a)
            x1 = signmask2(__vadd2(a, b));
            x0 = x1;
        } else {  // flags (1,0): u0 = sign(a) ^ sign(b), u1 = 0
            x0 = signmask2(a ^ b);
            x1 = 0u;
        }
    }

    // node of size M <= G: element gl of the node sits in lane gl (gl < M) of the group
    template <int M, int HEAP>
    __device__ __forceinline__ void cross(uint32_t v, uint32_t& bout) {
        const uint32_t t = ntype<HEAP>();
        if constexpr (M == 2) {
            const uint32_t a = __shfl_sync(FULL, v, 0, G);
            const uint32_t b = __shfl_sync(FULL, v, 1, G);
            uint32_t x0, x1;
            p2(t, a, b, x0, x1);
            bout = (gl & 1) ? x1 : x0;
        } else {
            if (t == T_R0) {
                bout = 0u;
                return;
            }
            if (t == T_R1) {
                bout = signmask2(v);
                const uint32_t z = (gl < M) ? zacc2(0u, v) : 0u;
                if (!__any_sync(FULL, (z & 0x80008000u) != 0u)) return;
            }
            const uint32_t pv = __shfl_xor_sync(FULL, v, M / 2);
            const uint32_t tl = ntype<2 * HEAP + 1>(), tr = ntype<2 * HEAP + 2>();
            uint32_t bl = 0u, br = 0u;
            if (tl != T_R0) cross<M / 2, 2 * HEAP + 1>(f16x2(v, pv), bl);
            if (tr != T_R0) {
                const uint32_t ar = (tl == T_R0) ? g0fun<M>(v, pv) : gfun<M>(v, pv, bl);
                cross<M / 2, 2 * HEAP + 2>(ar, br);
            }
            const uint32_t up = __shfl_xor_sync(FULL, br, M / 2);
            bout = (gl & (M / 2)) ? up : (bl ^ br);
        }
    }

    // node of size R*G: lane holds elements gl + G*r, r < R
    template <int R, int HEAP>
    __device__ __forceinline__ void local(const uint32_t (&a)[R], uint32_t (&bout)[R]) {
        const uint32_t t = ntype<HEAP>();
        if constexpr (R * G == 2) {  // G == 1: terminal pair inside one lane
            p2(t, a[0], a[1], bout[0], bout[1]);
        } else {
            if (t == T_R0) {
#pragma unroll
                for (int r = 0; r < R; r++) bout[r] = 0u;
                return;
            }
            if (t == T_R1) {
                uint32_t z = 0u;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    bout[r] = signmask2(a[r]);
                    z = zacc2(z, a[r]);
                }
                if (!__any_sync(FULL, (z & 0x80008000u) != 0u)) return;
            }
            constexpr int H = R / 2;
            const uint32_t tl = ntype<2 * HEAP + 1>(), tr = ntype<2 * HEAP + 2>();
            uint32_t x[H], bl[H], br[H];
#pragma unroll
            for (int r = 0; r < H; r++) bl[r] = br[r] = 0u;
            if (tl != T_R0) {
#pragma unroll
                for (int r = 0; r < H; r++) x[r] = f16x2(a[r], a[r + H]);
                child<H, 2 * HEAP + 1>(x, bl);
            }
            if (tr != T_R0) {
                if (tl == T_R0) {
#pragma unroll
                    for (int r = 0; r < H; r++) x[r] = g0fun<R * G>(a[r], a[r + H]);
                } else {
#pragma unroll
                    for (int r = 0; r < H; r++) x[r] = gfun<R * G>(a[r], a[r + H], bl[r]);
                }
                child<H, 2 * HEAP + 2>(x, br);
            }
#pragma unroll
            for (int r = 0; r < H; r++) {
                bout[r] = bl[r] ^ br[r];
                bout[r + H] = br[r];
            }
        }
    }
    template <int H, int HEAP>
    __device__ __forceinline__ void child(const uint32_t (&x)[H], uint32_t (&b)[H]) {
        if constexpr (H == 1 && G > 1)
            cross<G, HEAP>(x[0], b[0]);
        else
            local<H, HEAP>(x, b);
    }

    // subtree rooted at (LOG2S, o): alpha[LOG2S] -> partial sums of the node
    __device__ void op_subtree(uint32_t o) {
        const uint16_t* src = alpha(LOG2S);
        uint32_t a[8], b[8];
#pragma unroll
        for (int r = 0; r < 8; r++) a[r] = prmt((uint32_t)src[G * r + gl], 0u, 0x9180);
        local<8, 0>(a, b);
        uint16_t* d = beta(LOG2S, o);
#pragma unroll
        for (int r = 0; r < 8; r++) d[G * r + gl] = (uint16_t)prmt(b[r], 0u, 0x4420);
        __syncwarp();
    }

    // plain SC of an all-information node above the register subtree (rate-1 fallback)
    __device__ void generic_sc(int l, uint32_t o) {
#pragma unroll
        for (int k = 0; k < DESC_WORDS; k++) desc[k] = 0xAAAAAAAAu;  // every node all-information
        const uint32_t nt = 1u << (l - LOG2S);
        for (uint32_t t = 0; t < nt; t++) {
            const uint32_t to = o + (t << LOG2S);
            if (t == 0) {
                for (int lv = l; lv > LOG2S; lv--) op_f(lv);
            } else {
                const int lv0 = (__ffs(t) - 1) + LOG2S + 1;
                op_g<false>(lv0, to & ~((1u << lv0) - 1u));
                for (int lv = lv0 - 1; lv > LOG2S; lv--) op_f(lv);
            }
            op_subtree(to);
            const int ones = __ffs(~t) - 1;
            for (int j = 1; j <= ones && LOG2S + j <= l; j++) {
                const int lv = LOG2S + j;
                op_h<false>(lv, to + (1u << LOG2S) - (1u << lv));
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
                case OP_G: op_g<false>(l, o); break;
                case OP_G0: op_g<true>(l, o); break;
                case OP_H: op_h<false>(l, o); break;
                case OP_HCOPY: op_h<true>(l, o); break;
                case OP_R0: op_r0(l, o); break;
                case OP_R1:
                    if (op_hd(l, o)) {
                        if (l == LOG2S) {
#pragma unroll
                            for (int k = 0; k < DESC_WORDS; k++) desc[k] = 0xAAAAAAAAu;
                            op_subtree(o);
                        } else {
                            generic_sc(l, o);
                        }
                    }
                    break;
                case OP_SUB:
#pragma unroll
                    for (int k = 0; k < DESC_WORDS; k++) desc[k] = __ldg(p.sched + pc + 1 + k);
                    pc += DESC_WORDS;
                    op_subtree(o);
                    break;
                default: break;
            }
        }
    }

    // final partial sums (cells) of the root -> packed bits of both frames (32 cells per word)
    __device__ void write_output(uint32_t* outA, uint32_t* outB) {
        const uint16_t* b = beta((int)p.log2n, 0);
        for (uint32_t w = gl; w < p.wpf; w += G) {
            uint32_t ba = 0u, bb = 0u;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint4 c = *reinterpret_cast<const uint4*>(b + 32u * w + 8u * q);
                const uint32_t a0 = prmt(c.x, c.y, 0x6420), a1 = prmt(c.z, c.w, 0x6420);
                const uint32_t b0 = prmt(c.x, c.y, 0x7531), b1 = prmt(c.z, c.w, 0x7531);
                const uint32_t na = (((a0 & 0x08040201u) * 0x01010101u) >> 24) |
                                    ((((a1 & 0x08040201u) * 0x01010101u) >> 24) << 4);
                const uint32_t nb = (((b0 & 0x08040201u) * 0x01010101u) >> 24) |
                                    ((((b1 & 0x08040201u) * 0x01010101u) >> 24) << 4);
                ba |= na << (8 * q);
                bb |= nb << (8 * q);
            }
            if (outA) outA[w] = ba;
            if (outB) outB[w] = bb;
        }
        __syncwarp();
    }
};

template <int G, int LOG2PAR, bool EXT>
__global__ void __launch_bounds__(128) sc_decode_fast_kernel(const FastParams p) {
    extern __shared__ __align__(16) uint8_t smem_fast[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int GPW = 32 / G;
    const int nwarps = blockDim.x >> 5;
    const unsigned long long fp_per_cta = (unsigned long long)nwarps * GPW;
    const int grp = lane / G;
    const unsigned long long slot = (unsigned long long)warp * GPW + grp;

    FastDecoder<G, LOG2PAR, EXT> d(p);
    d.gl = lane % G;
    d.satp = p.satv * 0x00010001u;
    d.satn = ((0u - p.satv) & 0xFFFFu) * 0x00010001u;
    uint8_t* sm = smem_fast + slot * p.sm_stride;
    d.sm_alpha = reinterpret_cast<uint16_t*>(sm);
    d.sm_beta = d.sm_alpha + p.sm_alpha_cells;
    uint8_t* ws = p.ws + ((unsigned long long)blockIdx.x * fp_per_cta + slot) * p.ws_stride;
    d.gl_alpha = reinterpret_cast<uint16_t*>(ws);
    d.gl_beta = d.gl_alpha + 2ull * p.n;

    for (unsigned long long base = (unsigned long long)blockIdx.x * fp_per_cta; base < p.num_fp;
         base += (unsigned long long)gridDim.x * fp_per_cta) {
        if (base + (unsigned long long)warp * GPW >= p.num_fp) break;
        unsigned long long fp = base + slot;
        const bool valid = fp < p.num_fp;
        if (!valid) fp = p.num_fp - 1;
        const unsigned long long fa = 2 * fp;
        const bool has_b = fa + 1 < p.nframes;
        const unsigned long long fb = has_b ? fa + 1 : fa;
        d.llrA = p.llr + fa * p.n;
        d.llrB = p.llr + fb * p.n;
        d.run();
        d.write_output(valid ? p.xhat + fa * p.wpf : nullptr, (valid && has_b) ? p.xhat + fb * p.wpf : nullptr);
    }
}

}  // namespace scpd
