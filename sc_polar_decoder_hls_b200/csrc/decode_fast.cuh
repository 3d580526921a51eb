// decode_fast.cuh -- the throughput kernel for the headline settings (CA2, Q <= 8, PAR <= 8G).
//
// Layout of the work
//   * A group of G lanes owns one FRAME PAIR (two frames in the two int16 halves of a register);
//     a warp owns 32/G pairs that run the same host-built schedule in lock step.
//   * Nodes of size >= 16G ("memory levels"): the LLR stack is kept as int8 CELLS
//     (byte 0 = frame A, byte 1 = frame B; values are saturated to +-(2^(Q-1)-1) <= 127 there) in
//     shared memory -- or, for the largest levels, in an L2-resident workspace -- and moved with
//     128-bit accesses, 8 cells per lane; PRMT unpacks a cell to int16x2 and packs it back.
//     Partial sums take one byte per element (low nibble = frame A, high nibble = frame B, 0xF
//     when the bit is 1): one shift and one PRMT expand an element to the 0x0000 / 0xFFFF mask g
//     needs, and h is a plain XOR of packed words.
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
    uint32_t lsb;            // partial sums of nodes up to level lsb in shared memory (block of 2^(lsb+1) bytes)
    uint32_t sm_alpha_cells; // cells reserved for alpha in shared memory, per frame pair
    uint32_t sm_stride;      // bytes between the shared-memory regions of consecutive frame pairs
    uint8_t* ws;             // workspace: per resident frame pair  [alpha: 4n bytes][partial sums: n bytes]
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

// Shared-memory accesses by 32-bit shared-window address (LDS.128 / STS.128, no generic addressing).
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, const uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}
#else  // CPU emulation build: the "shared window" is the smem_fast array
extern uint8_t smem_fast[];
static inline uint32_t smem_u32(const void* p) { return (uint32_t)((const uint8_t*)p - smem_fast); }
static inline uint4 lds128(uint32_t a) { return *reinterpret_cast<const uint4*>(smem_fast + a); }
static inline void sts128(uint32_t a, const uint4 v) { *reinterpret_cast<uint4*>(smem_fast + a) = v; }
static inline uint2 lds64(uint32_t a) { return *reinterpret_cast<const uint2*>(smem_fast + a); }
static inline void sts64(uint32_t a, const uint2 v) { *reinterpret_cast<uint2*>(smem_fast + a) = v; }
#endif

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

template <bool SM>
struct MemRef {  // a byte-addressed operand array in shared memory (SM) or in the workspace
    uint32_t a;
    uint8_t* g;
    __device__ __forceinline__ uint4 ld128(uint32_t off) const {
        if (SM) return lds128(a + off);
        return *reinterpret_cast<const uint4*>(g + off);
    }
    __device__ __forceinline__ void st128(uint32_t off, const uint4 v) const {
        if (SM)
            sts128(a + off, v);
        else
            *reinterpret_cast<uint4*>(g + off) = v;
    }
    __device__ __forceinline__ uint2 ld64(uint32_t off) const {
        if (SM) return lds64(a + off);
        return *reinterpret_cast<const uint2*>(g + off);
    }
    __device__ __forceinline__ void st64(uint32_t off, const uint2 v) const {
        if (SM)
            sts64(a + off, v);
        else
            *reinterpret_cast<uint2*>(g + off) = v;
    }
};

// partial sums are stored one byte per element: low nibble 0xF if frame A's bit is 1, high nibble
// 0xF if frame B's is.  4 elements (one word) -> four 0x0000/0xFFFF-per-half masks for g.
__device__ __forceinline__ void expand4(uint32_t w, uint32_t* m) {
    const uint32_t wa = w << 4;
    m[0] = prmt(wa, w, 0xCC88);
    m[1] = prmt(wa, w, 0xDD99);
    m[2] = prmt(wa, w, 0xEEAA);
    m[3] = prmt(wa, w, 0xFFBB);
}
// int16x2 mask (0xFFFF per half) -> partial-sum byte
__device__ __forceinline__ uint32_t mask_to_byte(uint32_t m) {
    const uint32_t x = m & 0x00F0000Fu;
    return (x | (x >> 16)) & 0xFFu;
}
// 8 cells -> 8 partial-sum bytes of their hard decisions
__device__ __forceinline__ uint2 cells_to_hd_bytes(const uint4 w) {
    uint32_t u[4];
    const uint32_t c[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t t = prmt(c[k], 0u, 0xBA98) & 0xF00FF00Fu;  // A bytes -> 0x0F, B bytes -> 0xF0
        u[k] = t | (t >> 8);                                       // bytes 0 and 2 hold the two elements
    }
    return make_uint2(prmt(u[0], u[1], 0x6420), prmt(u[2], u[3], 0x6420));
}

// the mixed size-8 patterns (bit i = information flag of element i) that occur in the reference's
// frozen tables (Frozen_Bit_Tab/, Generated_Frozen_Bit/) plus the all-information one; anything
// else takes the descriptor-driven routine
#define SCPD_PAT8_CASES(X) X(0x80u) X(0xC0u) X(0xE0u) X(0xE8u) X(0xF8u) X(0xFCu) X(0xFEu) X(0xFFu)

// Everything a node of size <= 8 needs, small enough to be passed by value to out-of-line routines.
template <int G, int LOG2PAR, bool EXT>
struct LeafCtx {
    static constexpr unsigned FULL = 0xFFFFFFFFu;
    int gl;
    uint32_t satp, satn;
    uint32_t sub;  // node types of one size-8 subtree, 2 bits per node, local heap order (7 nodes)

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

    // ---- pattern-specialised routines: the information flags PAT are a template parameter, so the
    // whole walk below the node is straight-line code (only the rate-1 zero vote remains a run-time
    // branch).  FORCE: treat an all-information node as mixed (the fallback after a zero was seen).
    template <int M, uint32_t PAT, bool FORCE>
    __device__ __forceinline__ void cross_ct(uint32_t v, uint32_t& bout) const {
        constexpr uint32_t ALL = (M == 32) ? 0xFFFFFFFFu : ((1u << M) - 1u);
        if constexpr (PAT == 0u) {
            bout = 0u;
        } else if constexpr (M == 2) {
            const uint32_t a = __shfl_sync(FULL, v, 0, G);
            const uint32_t b = __shfl_sync(FULL, v, 1, G);
            uint32_t x0, x1;
            p2(PAT == 3u ? T_R1 : PAT == 2u ? T_MIX : T_X, a, b, x0, x1);
            bout = (gl & 1) ? x1 : x0;
        } else {
            if constexpr (PAT == ALL && !FORCE) {
                bout = signmask2(v);
                const uint32_t z = (gl < M) ? zacc2(0u, v) : 0u;
                if (__any_sync(FULL, (z & 0x80008000u) != 0u)) cross_ct<M, PAT, true>(v, bout);
            } else {
                constexpr uint32_t PL = PAT & ((1u << (M / 2)) - 1u), PR = PAT >> (M / 2);
                const uint32_t pv = __shfl_xor_sync(FULL, v, M / 2);
                uint32_t bl = 0u, br = 0u;
                if constexpr (PL != 0u) cross_ct<M / 2, PL, false>(f16x2(v, pv), bl);
                if constexpr (PR != 0u) {
                    const uint32_t ar = (PL == 0u) ? g0fun<M>(v, pv) : gfun<M>(v, pv, bl);
                    cross_ct<M / 2, PR, false>(ar, br);
                    const uint32_t up = __shfl_xor_sync(FULL, br, M / 2);
                    bout = (gl & (M / 2)) ? up : (bl ^ br);
                } else {
                    bout = (gl & (M / 2)) ? 0u : bl;
                }
            }
        }
    }
    template <int R, uint32_t PAT, bool FORCE>
    __device__ __forceinline__ void local_ct(const uint32_t (&a)[R], uint32_t (&bout)[R]) const {
        constexpr int NODE = R * G;
        constexpr uint32_t ALL = (NODE == 32) ? 0xFFFFFFFFu : ((1u << NODE) - 1u);
        if constexpr (PAT == 0u) {
#pragma unroll
            for (int r = 0; r < R; r++) bout[r] = 0u;
        } else if constexpr (NODE == 2) {
            p2(PAT == 3u ? T_R1 : PAT == 2u ? T_MIX : T_X, a[0], a[1], bout[0], bout[1]);
        } else if constexpr (PAT == ALL && !FORCE) {
            uint32_t z = 0u;
#pragma unroll
            for (int r = 0; r < R; r++) {
                bout[r] = signmask2(a[r]);
                z = zacc2(z, a[r]);
            }
            if (__any_sync(FULL, (z & 0x80008000u) != 0u)) local_ct<R, PAT, true>(a, bout);
        } else {
            constexpr int H = R / 2;
            constexpr uint32_t PL = PAT & ((1u << (NODE / 2)) - 1u), PR = PAT >> (NODE / 2);
            uint32_t x[H], bl[H], br[H];
#pragma unroll
            for (int r = 0; r < H; r++) bl[r] = br[r] = 0u;
            if constexpr (PL != 0u) {
#pragma unroll
                for (int r = 0; r < H; r++) x[r] = f16x2(a[r], a[r + H]);
                child_ct<H, PL>(x, bl);
            }
            if constexpr (PR != 0u) {
#pragma unroll
                for (int r = 0; r < H; r++)
                    x[r] = (PL == 0u) ? g0fun<NODE>(a[r], a[r + H]) : gfun<NODE>(a[r], a[r + H], bl[r]);
                child_ct<H, PR>(x, br);
            }
#pragma unroll
            for (int r = 0; r < H; r++) {
                bout[r] = bl[r] ^ br[r];
                bout[r + H] = br[r];
            }
        }
    }
    template <int H, uint32_t PAT>
    __device__ __forceinline__ void child_ct(const uint32_t (&x)[H], uint32_t (&b)[H]) const {
        if constexpr (H == 1 && G > 1)
            cross_ct<G, PAT, false>(x[0], b[0]);
        else
            local_ct<H, PAT, false>(x, b);
    }

    // ---- descriptor-driven routines for a size-8 subtree (node types from `sub`, local heap)
    template <int HEAP>
    __device__ __forceinline__ uint32_t stype() const { return (sub >> (2 * HEAP)) & 3u; }
    template <int M, int HEAP>
    __device__ __forceinline__ void cross_rt(uint32_t v, uint32_t& bout) const {
        const uint32_t t = stype<HEAP>();
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
            const uint32_t tl = stype<2 * HEAP + 1>(), tr = stype<2 * HEAP + 2>();
            uint32_t bl = 0u, br = 0u;
            if (tl != T_R0) cross_rt<M / 2, 2 * HEAP + 1>(f16x2(v, pv), bl);
            if (tr != T_R0) {
                const uint32_t ar = (tl == T_R0) ? g0fun<M>(v, pv) : gfun<M>(v, pv, bl);
                cross_rt<M / 2, 2 * HEAP + 2>(ar, br);
            }
            const uint32_t up = __shfl_xor_sync(FULL, br, M / 2);
            bout = (gl & (M / 2)) ? up : (bl ^ br);
        }
    }
    template <int R, int HEAP>
    __device__ __forceinline__ void local_rt(const uint32_t (&a)[R], uint32_t (&bout)[R]) const {
        const uint32_t t = stype<HEAP>();
        if constexpr (R * G == 2) {
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
            const uint32_t tl = stype<2 * HEAP + 1>(), tr = stype<2 * HEAP + 2>();
            uint32_t x[H], bl[H], br[H];
#pragma unroll
            for (int r = 0; r < H; r++) bl[r] = br[r] = 0u;
            if (tl != T_R0) {
#pragma unroll
                for (int r = 0; r < H; r++) x[r] = f16x2(a[r], a[r + H]);
                child_rt<H, 2 * HEAP + 1>(x, bl);
            }
            if (tr != T_R0) {
#pragma unroll
                for (int r = 0; r < H; r++)
                    x[r] = (tl == T_R0) ? g0fun<R * G>(a[r], a[r + H]) : gfun<R * G>(a[r], a[r + H], bl[r]);
                child_rt<H, 2 * HEAP + 2>(x, br);
            }
#pragma unroll
            for (int r = 0; r < H; r++) {
                bout[r] = bl[r] ^ br[r];
                bout[r + H] = br[r];
            }
        }
    }
    template <int H, int HEAP>
    __device__ __forceinline__ void child_rt(const uint32_t (&x)[H], uint32_t (&b)[H]) const {
        if constexpr (H == 1 && G > 1)
            cross_rt<G, HEAP>(x[0], b[0]);
        else
            local_rt<H, HEAP>(x, b);
    }
};

// Out-of-line size-8 routines: one copy of each in the kernel image.  _x: one element per lane
// (G >= 8); _l2: two elements per lane (G = 4).
template <int G, int LOG2PAR, bool EXT, uint32_t PAT>
__device__ __noinline__ uint32_t node8_x(uint32_t v, int gl, uint32_t satp, uint32_t satn) {
    const LeafCtx<G, LOG2PAR, EXT> c{gl, satp, satn, 0u};
    uint32_t b;
    c.template cross_ct<8, PAT, false>(v, b);
    return b;
}
template <int G, int LOG2PAR, bool EXT>
__device__ __noinline__ uint32_t node8_rt_x(uint32_t v, uint32_t sub, int gl, uint32_t satp, uint32_t satn) {
    const LeafCtx<G, LOG2PAR, EXT> c{gl, satp, satn, sub};
    uint32_t b;
    c.template cross_rt<8, 0>(v, b);
    return b;
}
template <int G, int LOG2PAR, bool EXT, uint32_t PAT>
__device__ __noinline__ uint2 node8_l2(uint32_t a0, uint32_t a1, int gl, uint32_t satp, uint32_t satn) {
    const LeafCtx<G, LOG2PAR, EXT> c{gl, satp, satn, 0u};
    const uint32_t a[2] = {a0, a1};
    uint32_t b[2];
    c.template local_ct<2, PAT, false>(a, b);
    return make_uint2(b[0], b[1]);
}
template <int G, int LOG2PAR, bool EXT>
__device__ __noinline__ uint2 node8_rt_l2(uint32_t a0, uint32_t a1, uint32_t sub, int gl, uint32_t satp, uint32_t satn) {
    const LeafCtx<G, LOG2PAR, EXT> c{gl, satp, satn, sub};
    const uint32_t a[2] = {a0, a1};
    uint32_t b[2];
    c.template local_rt<2, 0>(a, b);
    return make_uint2(b[0], b[1]);
}

// W > 1 (with G = 32): W warps -- the whole CTA -- walk ONE frame pair together.  Nodes above the register
// subtree are spread over all G * W lanes with a CTA barrier behind every op; the subtree itself is walked by
// warp 0.  For batches too small to fill the GPU with one warp per pair (large trees).
template <int G, int LOG2PAR, bool EXT, int W = 1>
struct FastDecoder : LeafCtx<G, LOG2PAR, EXT> {
    using LeafCtx<G, LOG2PAR, EXT>::gl;
    using LeafCtx<G, LOG2PAR, EXT>::satp;
    using LeafCtx<G, LOG2PAR, EXT>::satn;
    static constexpr unsigned FULL = 0xFFFFFFFFu;
    static constexpr int GX = G * W;  // lanes that share the nodes above the register subtree
    static constexpr int S = 8 * G;
    static constexpr int LOG2S = (G == 1) ? 3 : (G == 2) ? 4 : (G == 4) ? 5 : (G == 8) ? 6 : (G == 16) ? 7 : 8;
    static constexpr int DESC_WORDS = (2 * (S - 1) + 31) / 32;
    static constexpr int FLAG_WORDS = (S + 31) / 32;

    const FastParams& p;
    int gx;         // lane index among the GX lanes of the frame pair (= gl when W == 1)
    bool sub_warp;  // this warp walks the register subtrees
    uint32_t sm_alpha_a, sm_beta_a;  // shared-window byte addresses
    uint8_t* sm_alpha_p;             // generic pointers to the same regions (register subtree I/O)
    uint8_t* sm_beta_p;
    uint8_t* gl_alpha;               // workspace: level l at byte offset 2 << l
    uint8_t* gl_beta;                // workspace: absolute element positions
    const int8_t* llrA;
    const int8_t* llrB;
    bool spec_ok;      // pattern-specialised size-8 routines allowed for the current subtree
    uint32_t sub8_pc;  // schedule index of the per-size-8-node descriptors of the current subtree
    uint32_t desc[DESC_WORDS];
    uint32_t flagw[FLAG_WORDS];  // raw information flags of the current subtree

    __device__ FastDecoder(const FastParams& p_) : p(p_) {}
    __device__ __forceinline__ void bar() const {
        if constexpr (W == 1)
            __syncwarp();
        else
            __syncthreads();
    }

    // ------------------------------------------------------------ storage
    template <bool SM>
    __device__ __forceinline__ MemRef<SM> aref(int l) const {  // alpha[l]: 2^l cells of 2 bytes
        MemRef<SM> r;
        r.a = sm_alpha_a + (2u << l);
        r.g = gl_alpha + (2u << l);
        return r;
    }
    template <bool SM>
    __device__ __forceinline__ MemRef<SM> bref(uint32_t o) const {  // partial sums of a node at offset o
        MemRef<SM> r;
        r.a = sm_beta_a + (o & ((2u << p.lsb) - 1u));
        r.g = gl_beta + o;
        return r;
    }
    __device__ __forceinline__ bool a_sm(int l) const { return (uint32_t)l <= p.lsa; }
    __device__ __forceinline__ bool b_sm(int l) const { return (uint32_t)l <= p.lsb; }

    // ------------------------------------------------------------ memory-level operations
    __device__ __forceinline__ static uint4 f8(const uint4 wa, const uint4 wb) {
        uint32_t a[8], b[8], r[8];
        unpack8(wa, a);
        unpack8(wb, b);
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = f16x2(a[j], b[j]);
        return pack8(r);
    }
    template <bool ZERO>
    __device__ __forceinline__ uint4 g8(const uint4 wa, const uint4 wb, const uint2 wm) const {
        uint32_t a[8], b[8], m[8], r[8];
        unpack8(wa, a);
        unpack8(wb, b);
        if (!ZERO) {
            expand4(wm.x, m);
            expand4(wm.y, m + 4);
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
            r[j] = ZERO ? g016x2_sat(a[j], b[j], satp, satn) : g16x2_sat(a[j], b[j], m[j], satp, satn);
        return pack8(r);
    }
    // 8 channel LLRs of both frames starting at element e -> 8 cells
    __device__ __forceinline__ uint4 root8(uint32_t e) const {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(llrA + e));
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(llrB + e));
        return rows_to_cells(a, b);
    }

    // alpha[l-1] = f(alpha[l] lower half, upper half); two chunks of 8 cells in flight per lane
    template <bool ROOT, bool ASM, bool DSM>
    __device__ __forceinline__ void f_t(int l) {
        const uint32_t h = 1u << (l - 1);
        const MemRef<ASM> src = aref<ASM>(l);
        const MemRef<DSM> dst = aref<DSM>(l - 1);
        uint32_t e = 8u * gx;
        for (; e + 8u * GX < h; e += 16u * GX) {
            const uint32_t e1 = e + 8u * GX;
            const uint4 a0 = ROOT ? root8(e) : src.ld128(2 * e), b0 = ROOT ? root8(e + h) : src.ld128(2 * (e + h));
            const uint4 a1 = ROOT ? root8(e1) : src.ld128(2 * e1), b1 = ROOT ? root8(e1 + h) : src.ld128(2 * (e1 + h));
            dst.st128(2 * e, f8(a0, b0));
            dst.st128(2 * e1, f8(a1, b1));
        }
        if (e < h) {
            const uint4 a0 = ROOT ? root8(e) : src.ld128(2 * e), b0 = ROOT ? root8(e + h) : src.ld128(2 * (e + h));
            dst.st128(2 * e, f8(a0, b0));
        }
        bar();
    }
    __device__ __forceinline__ void op_f(int l) {
        if ((uint32_t)l == p.log2n) return a_sm(l - 1) ? f_t<true, false, true>(l) : f_t<true, false, false>(l);
        if (a_sm(l)) return f_t<false, true, true>(l);
        return a_sm(l - 1) ? f_t<false, false, true>(l) : f_t<false, false, false>(l);
    }

    template <bool ZERO, bool ROOT, bool ASM, bool DSM, bool BSM>
    __device__ __forceinline__ void g_t(int l, uint32_t o) {
        const uint32_t h = 1u << (l - 1);
        const MemRef<ASM> src = aref<ASM>(l);
        const MemRef<DSM> dst = aref<DSM>(l - 1);
        const MemRef<BSM> bs = bref<BSM>(o);
        const uint2 z = make_uint2(0u, 0u);
        uint32_t e = 8u * gx;
        for (; e + 8u * GX < h; e += 16u * GX) {
            const uint32_t e1 = e + 8u * GX;
            const uint4 a0 = ROOT ? root8(e) : src.ld128(2 * e), b0 = ROOT ? root8(e + h) : src.ld128(2 * (e + h));
            const uint4 a1 = ROOT ? root8(e1) : src.ld128(2 * e1), b1 = ROOT ? root8(e1 + h) : src.ld128(2 * (e1 + h));
            const uint2 m0 = ZERO ? z : bs.ld64(e), m1 = ZERO ? z : bs.ld64(e1);
            dst.st128(2 * e, g8<ZERO>(a0, b0, m0));
            dst.st128(2 * e1, g8<ZERO>(a1, b1, m1));
        }
        if (e < h) {
            const uint4 a0 = ROOT ? root8(e) : src.ld128(2 * e), b0 = ROOT ? root8(e + h) : src.ld128(2 * (e + h));
            dst.st128(2 * e, g8<ZERO>(a0, b0, ZERO ? z : bs.ld64(e)));
        }
        bar();
    }
    template <bool ZERO>
    __device__ __forceinline__ void op_g(int l, uint32_t o) {
        const bool bsm = b_sm(l - 1), dsm = a_sm(l - 1);
        if ((uint32_t)l == p.log2n) {
            if (dsm) return bsm ? g_t<ZERO, true, false, true, true>(l, o) : g_t<ZERO, true, false, true, false>(l, o);
            return bsm ? g_t<ZERO, true, false, false, true>(l, o) : g_t<ZERO, true, false, false, false>(l, o);
        }
        if (a_sm(l)) return bsm ? g_t<ZERO, false, true, true, true>(l, o) : g_t<ZERO, false, true, true, false>(l, o);
        if (dsm) return bsm ? g_t<ZERO, false, false, true, true>(l, o) : g_t<ZERO, false, false, true, false>(l, o);
        return bsm ? g_t<ZERO, false, false, false, true>(l, o) : g_t<ZERO, false, false, false, false>(l, o);
    }

    // node (l,o) := (left ^ right, right) of its children; COPY: the left child is all-frozen.
    // CSM / DSM: children / node kept in shared memory.  When they differ the node moves.
    template <bool COPY, bool CSM, bool DSM>
    __device__ __forceinline__ void h_t(int l, uint32_t o) {
        const uint32_t h = 1u << (l - 1);
        const MemRef<CSM> cl = bref<CSM>(o), cr = bref<CSM>(o + h);
        const MemRef<DSM> d = bref<DSM>(o);
        for (uint32_t e = 8u * gx; e < h; e += 8u * GX) {
            uint2 x = cr.ld64(e);
            if (CSM != DSM) d.st64(h + e, x);
            if (!COPY) {
                const uint2 y = cl.ld64(e);
                x.x ^= y.x;
                x.y ^= y.y;
            }
            d.st64(e, x);
        }
        bar();
    }
    template <bool COPY>
    __device__ __forceinline__ void op_h(int l, uint32_t o) {
        if (b_sm(l)) return h_t<COPY, true, true>(l, o);
        return b_sm(l - 1) ? h_t<COPY, true, false>(l, o) : h_t<COPY, false, false>(l, o);
    }
    template <bool DSM>
    __device__ __forceinline__ void r0_t(int l, uint32_t o) {
        const MemRef<DSM> d = bref<DSM>(o);
        for (uint32_t e = 8u * gx; e < (1u << l); e += 8u * GX) d.st64(e, make_uint2(0u, 0u));
        bar();
    }
    __device__ __forceinline__ void op_r0(int l, uint32_t o) { return b_sm(l) ? r0_t<true>(l, o) : r0_t<false>(l, o); }

    // hard decision of node (l,o); returns true (warp-uniform) if some LLR of some pair was 0
    template <bool ROOT, bool ASM, bool DSM>
    __device__ __forceinline__ bool hd_t(int l, uint32_t o) {
        const MemRef<ASM> src = aref<ASM>(l);
        const MemRef<DSM> d = bref<DSM>(o);
        uint32_t z = 0;
        for (uint32_t e = 8u * gx; e < (1u << l); e += 8u * GX) {
            const uint4 w = ROOT ? root8(e) : src.ld128(2 * e);
            z |= ((w.x - 0x01010101u) & ~w.x) | ((w.y - 0x01010101u) & ~w.y) | ((w.z - 0x01010101u) & ~w.z) |
                 ((w.w - 0x01010101u) & ~w.w);
            d.st64(e, cells_to_hd_bytes(w));
        }
        bool any;
        if constexpr (W == 1) {
            any = __any_sync(FULL, (z & 0x80808080u) != 0u);
            __syncwarp();
        } else {
            any = __syncthreads_or((z & 0x80808080u) != 0u) != 0;
        }
        return any;
    }
    __device__ __forceinline__ bool op_hd(int l, uint32_t o) {
        const bool dsm = b_sm(l);
        if ((uint32_t)l == p.log2n) return dsm ? hd_t<true, false, true>(l, o) : hd_t<true, false, false>(l, o);
        if (a_sm(l)) return dsm ? hd_t<false, true, true>(l, o) : hd_t<false, true, false>(l, o);
        return dsm ? hd_t<false, false, true>(l, o) : hd_t<false, false, false>(l, o);
    }

    // ------------------------------------------------------------ register subtree (sizes S .. 16)
    // Nodes of size 8 and below are handled by the out-of-line routines of LeafCtx (node8_*), one
    // copy each, so that the unrolled code of the upper part of the subtree stays small enough for
    // the instruction cache.
    template <int HEAP>
    __device__ __forceinline__ uint32_t ntype() const {
        return (desc[(2 * HEAP) >> 5] >> ((2 * HEAP) & 31)) & 3u;
    }
    // the size-8 node with index K inside the subtree (heap index S/8 - 1 + K)
    template <int K, int R8>
    __device__ __forceinline__ void node8(const uint32_t (&x)[R8], uint32_t (&b)[R8]) {
        // 0x100: no specialised routine (schedule built without rate-1 pruning) -> descriptor-driven
        const uint32_t pat = spec_ok ? ((flagw[(8 * K) >> 5] >> ((8 * K) & 31)) & 0xFFu) : 0x100u;
        if constexpr (G >= 8) {
            uint32_t r;
            switch (pat) {
#define X(P)                                                              \
    case P:                                                               \
        r = node8_x<G, LOG2PAR, EXT, P>(x[0], this->gl, this->satp, this->satn); \
        break;
                SCPD_PAT8_CASES(X)
#undef X
                default:
                    r = node8_rt_x<G, LOG2PAR, EXT>(x[0], __ldg(p.sched + sub8_pc + K), this->gl, this->satp, this->satn);
                    break;
            }
            b[0] = r;
        } else if constexpr (G == 4) {
            uint2 r;
            switch (pat) {
#define X(P)                                                                     \
    case P:                                                                      \
        r = node8_l2<G, LOG2PAR, EXT, P>(x[0], x[1], this->gl, this->satp, this->satn); \
        break;
                SCPD_PAT8_CASES(X)
#undef X
                default:
                    r = node8_rt_l2<G, LOG2PAR, EXT>(x[0], x[1], __ldg(p.sched + sub8_pc + K), this->gl, this->satp,
                                                    this->satn);
                    break;
            }
            b[0] = r.x;
            b[1] = r.y;
        } else {  // G = 1, 2: not a performance target, inline descriptor-driven routine
            this->sub = __ldg(p.sched + sub8_pc + K);
            this->template local_rt<R8, 0>(x, b);
        }
    }
    // node of size R*G >= 16: lane holds elements gl + G*r, r < R
    template <int R, int HEAP>
    __device__ __forceinline__ void local(const uint32_t (&a)[R], uint32_t (&bout)[R]) {
        if constexpr (R * G == 8) {
            node8<HEAP - (S / 8 - 1), R>(a, bout);
        } else {
            const uint32_t t = ntype<HEAP>();
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
                    for (int r = 0; r < H; r++) x[r] = this->template g0fun<R * G>(a[r], a[r + H]);
                } else {
#pragma unroll
                    for (int r = 0; r < H; r++) x[r] = this->template gfun<R * G>(a[r], a[r + H], bl[r]);
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
    // node of size M > 8 spread one element per lane (only for G >= 16)
    template <int M, int HEAP>
    __device__ __forceinline__ void cross(uint32_t v, uint32_t& bout) {
        if constexpr (M == 8) {
            uint32_t x[1] = {v}, b[1];
            node8<HEAP - (S / 8 - 1), 1>(x, b);
            bout = b[0];
        } else {
            const uint32_t t = ntype<HEAP>();
            if (t == T_R0) {
                bout = 0u;
                return;
            }
            if (t == T_R1) {
                bout = signmask2(v);
                const uint32_t z = (this->gl < M) ? zacc2(0u, v) : 0u;
                if (!__any_sync(FULL, (z & 0x80008000u) != 0u)) return;
            }
            const uint32_t pv = __shfl_xor_sync(FULL, v, M / 2);
            const uint32_t tl = ntype<2 * HEAP + 1>(), tr = ntype<2 * HEAP + 2>();
            uint32_t bl = 0u, br = 0u;
            if (tl != T_R0) cross<M / 2, 2 * HEAP + 1>(f16x2(v, pv), bl);
            if (tr != T_R0) {
                const uint32_t ar = (tl == T_R0) ? this->template g0fun<M>(v, pv) : this->template gfun<M>(v, pv, bl);
                cross<M / 2, 2 * HEAP + 2>(ar, br);
            }
            const uint32_t up = __shfl_xor_sync(FULL, br, M / 2);
            bout = (this->gl & (M / 2)) ? up : (bl ^ br);
        }
    }
    template <int H, int HEAP>
    __device__ __forceinline__ void child(const uint32_t (&x)[H], uint32_t (&b)[H]) {
        if constexpr (H == 1 && G > 8)
            cross<G, HEAP>(x[0], b[0]);
        else
            local<H, HEAP>(x, b);
    }

    // subtree rooted at (LOG2S, o): alpha[LOG2S] (always in shared memory) -> partial sums of the node
    __device__ __forceinline__ void op_subtree(uint32_t o) {
        if (W == 1 || sub_warp) {
            const uint16_t* src = reinterpret_cast<const uint16_t*>(sm_alpha_p + (2u << LOG2S));
            uint32_t a[8], b[8];
#pragma unroll
            for (int r = 0; r < 8; r++) a[r] = prmt((uint32_t)src[G * r + gl], 0u, 0x9180);
            local<8, 0>(a, b);
            uint8_t* d = sm_beta_p + (o & ((2u << p.lsb) - 1u));
#pragma unroll
            for (int r = 0; r < 8; r++) d[G * r + gl] = (uint8_t)mask_to_byte(b[r]);
        }
        bar();
    }

    __device__ __forceinline__ void run() {
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
                case OP_R1: {  // hard decision; the explicit plain-SC ops that follow run only if an LLR was 0
                    const uint32_t skip = __ldg(p.sched + pc + 1);
                    pc += 1;
                    if (!op_hd(l, o)) pc += skip;
                    break;
                }
                case OP_SUB:
#pragma unroll
                    for (int k = 0; k < DESC_WORDS; k++) desc[k] = __ldg(p.sched + pc + 1 + k);
#pragma unroll
                    for (int k = 0; k < FLAG_WORDS; k++) flagw[k] = __ldg(p.sched + pc + 1 + DESC_WORDS + k);
                    spec_ok = op_nosat(w) != 0u;
                    sub8_pc = pc + 1 + DESC_WORDS + FLAG_WORDS;
                    pc += DESC_WORDS + FLAG_WORDS + S / 8;
                    op_subtree(o);
                    break;
                default: break;
            }
        }
    }

    // final partial sums of the root (one byte per element) -> packed bits of both frames
    template <bool SM>
    __device__ __forceinline__ void write_output_t(uint32_t* outA, uint32_t* outB) {
        const MemRef<SM> b = bref<SM>(0u);
        for (uint32_t w = gx; w < p.wpf; w += GX) {
            uint32_t ba = 0u, bb = 0u;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint4 c = b.ld128(32u * w + 16u * q);
                const uint32_t cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    ba |= (((cw[k] & 0x08040201u) * 0x01010101u) >> 24) << (16 * q + 4 * k);
                    bb |= (((cw[k] & 0x80402010u) * 0x01010101u) >> 28) << (16 * q + 4 * k);
                }
            }
            if (outA) outA[w] = ba;
            if (outB) outB[w] = bb;
        }
        bar();
    }
    __device__ __forceinline__ void write_output(uint32_t* outA, uint32_t* outB) {
        return b_sm((int)p.log2n) ? write_output_t<true>(outA, outB) : write_output_t<false>(outA, outB);
    }
};

template <int G, int LOG2PAR, bool EXT>
__global__ void __launch_bounds__(256) sc_decode_fast_kernel(const FastParams p) {
    extern __shared__ __align__(16) uint8_t smem_fast[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int GPW = 32 / G;
    const int nwarps = blockDim.x >> 5;
    const unsigned long long fp_per_cta = (unsigned long long)nwarps * GPW;
    const int grp = lane / G;
    const unsigned long long slot = (unsigned long long)warp * GPW + grp;

    FastDecoder<G, LOG2PAR, EXT> d(p);
    d.gl = lane % G;
    d.gx = d.gl;
    d.sub_warp = true;
    d.satp = p.satv * 0x00010001u;
    d.satn = ((0u - p.satv) & 0xFFFFu) * 0x00010001u;
    d.sm_alpha_p = smem_fast + slot * p.sm_stride;
    d.sm_beta_p = d.sm_alpha_p + 2u * p.sm_alpha_cells;
    d.sm_alpha_a = smem_u32(d.sm_alpha_p);
    d.sm_beta_a = smem_u32(d.sm_beta_p);
    d.gl_alpha = p.ws + ((unsigned long long)blockIdx.x * fp_per_cta + slot) * p.ws_stride;
    d.gl_beta = d.gl_alpha + 4ull * p.n;

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

// One CTA of W warps per frame pair (FastDecoder<32, ..., W>): shared memory and workspace hold one pair.
template <int LOG2PAR, bool EXT, int W>
__global__ void __launch_bounds__(32 * W) sc_decode_fast_coop_kernel(const FastParams p) {
    extern __shared__ __align__(16) uint8_t smem_fast[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FastDecoder<32, LOG2PAR, EXT, W> d(p);
    d.gl = lane;
    d.gx = (int)threadIdx.x;
    d.sub_warp = warp == 0;
    d.satp = p.satv * 0x00010001u;
    d.satn = ((0u - p.satv) & 0xFFFFu) * 0x00010001u;
    d.sm_alpha_p = smem_fast;
    d.sm_beta_p = d.sm_alpha_p + 2u * p.sm_alpha_cells;
    d.sm_alpha_a = smem_u32(d.sm_alpha_p);
    d.sm_beta_a = smem_u32(d.sm_beta_p);
    d.gl_alpha = p.ws + (unsigned long long)blockIdx.x * p.ws_stride;
    d.gl_beta = d.gl_alpha + 4ull * p.n;
    for (unsigned long long fp = blockIdx.x; fp < p.num_fp; fp += gridDim.x) {  // CTA-uniform
        const unsigned long long fa = 2 * fp;
        const bool has_b = fa + 1 < p.nframes;
        const unsigned long long fb = has_b ? fa + 1 : fa;
        d.llrA = p.llr + fa * p.n;
        d.llrB = p.llr + fb * p.n;
        d.run();
        d.write_output(p.xhat + fa * p.wpf, has_b ? p.xhat + fb * p.wpf : nullptr);
    }
}

}  // namespace scpd
