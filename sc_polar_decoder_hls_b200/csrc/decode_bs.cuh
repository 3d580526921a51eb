// decode_bs.cuh -- the bit-sliced SC decode kernel: one warp decodes a GROUP of 32 frames at once.
//
// Why: the decoder's arithmetic is Q-bit (Q = 6..8) and every frame of a batch runs the same
// control flow, so the natural SIMD axis on a GPU is the FRAME axis at bit granularity: a 32-bit
// register holds one bit plane of one LLR for 32 frames, f / g / h become a few dozen LOP3s per 32
// frames (f: 2P+1, g: 5P+2 for P magnitude planes) with no pack / unpack, no wasted SIMD half and
// partial sums that are already packed (see bs_arith.cuh).
//
// Layout of the work
//   * group = 32 consecutive frames; G lanes of a warp own one group from its channel planes (written by
//     bs_planes_kernel below) to the packed output words (a warp decodes 32/G groups in lock step).  Lane i of the G works on LLR
//     "slot" i, i + G, ... of the node being processed, so the f / g pair (i, i + n/2) is lane-local
//     for every node of 2G elements or more; smaller nodes use the first n/2 lanes and exchange
//     through shared memory.
//   * the host compiles the SC tree walk into one 32-bit word per node operation (bs_plan.h).  A subtree of
//     16 LLRs is ONE op that walks its 15 nodes from a word of 2-bit node types; ops on nodes of 32 / 64
//     LLRs have their level folded into the opcode (sizes, plane counts and shared-memory addresses are
//     immediates); larger nodes run streaming loops, optionally fused over up to three levels.
//   * alpha[l] (the 2^l LLRs a node of size 2^l receives): per LLR a vector of planes
//     {sign, m0, m1, ...}, stored as uint4 "plane quads": quad v of slot i at byte (v << l + i) * 16.
//     Levels <= lsa live in shared memory, larger ones in a per-group workspace (L2 / HBM); the channel
//     level has its own buffer, one set of planes per group of the batch.
//   * beta (partial sums): one word per code position (32 frames), natural order; nodes up to level
//     lsb in shared memory, larger ones in the workspace.
//   * bs_planes_kernel turns the int8 rows into planes by 32x32 bit transposes in registers (4 LLRs x 8
//     planes per transpose); the final partial sums are turned into packed rows the same way.
//
// Bit-exactness: see bs_arith.cuh for f / g / leaf rules in both number formats.  Width growth inside
// the PAR-wide leaf decoder when EXTENDED (Spec_P*_ext, functions.h:413-438...) is modelled by giving
// alpha[l], l < log2 PAR, one extra magnitude plane per level.  All-information nodes are replaced by
// the hard decision only when no LLR of any of the 32 frames is a CA2 zero; otherwise the explicit
// plain-SC ops the schedule carries behind the OP_R1 are executed (children try again).  SIGMAG never
// needs the fallback.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bs_arith.cuh"
#include "bs_plan.h"        // op-word encoding
#include "decode_fast.cuh"  // smem_u32 / lds128 / sts128 helpers
#include "schedule.h"

// register budget: CTAs of up to 128 threads, SCPD_BS_MINB of them per SM; SCPD_BS_U slots per lane and trip
// in the generic-level loops
#ifndef SCPD_BS_MINB
#define SCPD_BS_MINB 4
#endif
#ifndef SCPD_BS_U
#define SCPD_BS_U 4
#endif

namespace scpd {

struct BsParams {
    const uint32_t* sched;  // bit-sliced op words (bs_plan.h: bs_compile_schedule)
    uint32_t sched_words;   // > 0: the CTA keeps a copy of the schedule behind the group regions in shared memory
    const uint8_t* planes;  // alpha[log2n] of every group: (sign, |llr|) planes written by bs_planes_kernel
    unsigned long long planes_stride;  // bytes per group
    uint32_t* xhat;
    unsigned long long nframes, ngroups;
    uint32_t n, log2n, wpf;
    uint32_t lsa;          // alpha levels <= lsa in shared memory, above in the workspace
    uint32_t lsb;          // partial sums of nodes up to level lsb in shared memory (block of 2^(lsb+1) words)
    uint32_t sm_stride;    // bytes between the shared-memory regions of consecutive groups
    uint32_t sm_beta_off;  // byte offset of the partial-sum block inside a region
    uint8_t* ws;           // workspace, per resident group: alpha levels lsa < l < log2n, then n partial-sum words
    unsigned long long ws_stride;
    uint32_t ws_beta_off;
    uint32_t aoff[24];     // byte offset of alpha[l] inside the shared region (l <= lsa) or the workspace
    uint32_t prefetch;     // 1: g asks L2 for its whole source level up front (pays off for small trees)
};

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
#else
static inline uint32_t lds32(uint32_t a) { return *reinterpret_cast<const uint32_t*>(smem_fast + a); }
static inline void sts32(uint32_t a, uint32_t v) { *reinterpret_cast<uint32_t*>(smem_fast + a) = v; }
#endif

template <int FMT, int Q, int LOG2PAR, bool EXT, int G, int L>
__device__ __noinline__ void bs_sub(uint32_t sm_lane, uint32_t sm_blane, int ll, uint32_t desc, uint32_t heap, uint32_t ob);

// Node operations on the specialised levels (nodes of 2 .. 64 LLRs).  Everything lives in shared memory at
// offsets that are compile-time constants; level, plane counts and trip counts are immediates.
template <int FMT, int Q, int LOG2PAR, bool EXT, int G>
struct BsLow {
    static_assert(LOG2PAR >= 1 && LOG2PAR <= BS_LLOW, "leaf decoder must sit inside the specialised levels");
    static constexpr unsigned FULL = 0xFFFFFFFFu;
    static constexpr int PQ = Q - 1;  // magnitude planes of a saturated LLR
    // magnitude planes of alpha[l]: one more per level inside the un-saturated leaf decoder
    __host__ __device__ static constexpr int plevel(int l) { return PQ + ((EXT && l < LOG2PAR) ? LOG2PAR - l : 0); }
    __host__ __device__ static constexpr int nquads(int p) { return (p + 1 + 3) / 4; }
    // byte offset of alpha[L] in a group's shared region: levels 1, 2, ... laid end to end (bs_plan.h)
    __host__ __device__ static constexpr uint32_t aoff_low(int L) {
        uint32_t off = 0;
        for (int l = 1; l < L; l++) off += ((uint32_t)nquads(plevel(l)) * 16u) << l;
        return off;
    }

    int ll;             // lane in the group
    uint32_t sm_lane;   // shared-window byte address of the group region + 16 * ll
    uint32_t sm_blane;  // ... of the partial-sum block + 4 * ll

    // shared-memory forms with the level as an immediate: addr = byte address of the slot in quad 0
    template <int P, int L>
    __device__ __forceinline__ static void load_sm(uint32_t addr, bs::Val<P>& x) {
        constexpr int NV = nquads(P);
        uint32_t w[4 * NV];
#pragma unroll
        for (int v = 0; v < NV; v++) {
            const uint4 t = lds128(addr + ((uint32_t)v << L) * 16u);
            w[4 * v] = t.x;
            w[4 * v + 1] = t.y;
            w[4 * v + 2] = t.z;
            w[4 * v + 3] = t.w;
        }
        x.s = w[0];
#pragma unroll
        for (int k = 0; k < P; k++) x.m[k] = w[1 + k];
    }
    template <int P, int L>
    __device__ __forceinline__ static void store_sm(uint32_t addr, const bs::Val<P>& x) {
        constexpr int NV = nquads(P);
        uint32_t w[4 * NV];
        w[0] = x.s;
#pragma unroll
        for (int k = 1; k < 4 * NV; k++) w[k] = (k <= P) ? x.m[k - 1] : 0u;
#pragma unroll
        for (int v = 0; v < NV; v++)
            sts128(addr + ((uint32_t)v << L) * 16u, make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]));
    }

    // `ob` is the byte offset of the node inside the shared partial-sum block.  Lanes past the node read
    // harmless neighbouring data and only their stores are predicated off.
    // f / g run in phases (all loads, then the arithmetic, then the stores) over up to two slots per lane,
    // so that the shared-memory latency of one slot hides behind the other
    template <int L>
    __device__ __forceinline__ void f_low() {
        constexpr int P = plevel(L), PO = plevel(L - 1), H = 1 << (L - 1), IT = (H > G) ? H / G : 1, U = IT > 1 ? 2 : 1;
        const uint32_t src = sm_lane + aoff_low(L), dst = sm_lane + aoff_low(L - 1);
#pragma unroll
        for (int it = 0; it < IT; it += U) {
            bs::Val<P> a[U], b[U];
            bs::Val<PO> w[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                load_sm<P, L>(src + 16u * ((it + u) * G), a[u]);
                load_sm<P, L>(src + 16u * ((it + u) * G + H), b[u]);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                bs::Val<P> r;
                bs::f_op<P>(a[u], b[u], r);
                bs::widen<PO, P>(r, w[u]);
            }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (H >= G || ll < H) store_sm<PO, L - 1>(dst + 16u * ((it + u) * G), w[u]);
        }
        __syncwarp();
    }
    template <int L>
    __device__ __forceinline__ void g_low(uint32_t ob, bool zero) {  // zero: left child all-frozen, partial sums 0
        constexpr int P = plevel(L), PO = plevel(L - 1), H = 1 << (L - 1), IT = (H > G) ? H / G : 1, U = IT > 1 ? 2 : 1;
        const uint32_t src = sm_lane + aoff_low(L), dst = sm_lane + aoff_low(L - 1);
#pragma unroll
        for (int it = 0; it < IT; it += U) {
            bs::Val<P> a[U], b[U];
            bs::Val<PO> r[U];
            uint32_t u_[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                load_sm<P, L>(src + 16u * ((it + u) * G), a[u]);
                load_sm<P, L>(src + 16u * ((it + u) * G + H), b[u]);
                u_[u] = lds32(sm_blane + ob + 4u * ((it + u) * G));
                if (zero) u_[u] = 0u;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if constexpr (PO == P)
                    bs::g_sat<FMT, P>(a[u], b[u], u_[u], r[u]);
                else
                    bs::g_ext<P>(a[u], b[u], u_[u], r[u]);
            }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (H >= G || ll < H) store_sm<PO, L - 1>(dst + 16u * ((it + u) * G), r[u]);
        }
        __syncwarp();
    }
    template <int L>
    __device__ __forceinline__ void h_low(uint32_t ob, bool copy) {  // copy: left child all-frozen
        constexpr int H = 1 << (L - 1), IT = (H > G) ? H / G : 1;
#pragma unroll
        for (int it = 0; it < IT; it++) {
            uint32_t x = lds32(sm_blane + ob + 4u * (H + it * G));
            const uint32_t y = lds32(sm_blane + ob + 4u * (it * G));
            if (!copy) x ^= y;
            if (H >= G || ll < H) sts32(sm_blane + ob + 4u * (it * G), x);
        }
        __syncwarp();
    }
    template <int L>
    __device__ __forceinline__ void r0_low(uint32_t ob) {
        constexpr int N = 1 << L, IT = (N > G) ? N / G : 1;
#pragma unroll
        for (int it = 0; it < IT; it++)
            if (N >= G || ll < N) sts32(sm_blane + ob + 4u * (it * G), 0u);
        __syncwarp();
    }
    template <int L>
    __device__ __forceinline__ bool r1_low(uint32_t ob) {
        constexpr int P = plevel(L), N = 1 << L, IT = (N > G) ? N / G : 1;
        const uint32_t src = sm_lane + aoff_low(L);
        uint32_t z = 0u;
#pragma unroll
        for (int it = 0; it < IT; it++) {
            bs::Val<P> a;
            load_sm<P, L>(src + 16u * (it * G), a);
            uint32_t x = a.s;
            if (FMT == bs::FMT_CA2) {
                const uint32_t nz = bs::nonzero<P>(a);
                x &= nz;
                if (N >= G || ll < N) z |= ~nz;
            }
            if (N >= G || ll < N) sts32(sm_blane + ob + 4u * (it * G), x);
        }
        bool any = false;
        if (FMT == bs::FMT_CA2) any = __any_sync(FULL, z != 0u);
        __syncwarp();
        return any;
    }
    // 2-bit terminal (Spec_P2, functions.h:367-384): one lane per group, both LLRs of alpha[1].
    // t: node-type code of the schedule, 0 = flags (0,1), 1 = (0,0), 2 = (1,1), 3 = (1,0)
    __device__ __forceinline__ void p2_low(uint32_t t, uint32_t ob) {
        constexpr int P1 = plevel(1);
        if (ll == 0) {
            const uint32_t src = sm_lane + aoff_low(1);
            uint32_t x0 = 0u, x1 = 0u;
            if (t != T_R0) {
                bs::Val<P1> a, b;
                load_sm<P1, 1>(src, a);
                load_sm<P1, 1>(src + 16u, b);
                bs::p2_op<FMT, P1>(a, b, t == T_R1 ? 3u : (t == T_MIX ? 2u : 1u), x0, x1);
            }
            sts32(sm_blane + ob, x0);
            sts32(sm_blane + ob + 4u, x1);
        }
        __syncwarp();
    }
    // Fused walk of a subtree of 2^BS_LSUB LLRs: node types (2 bits per node, heap order) come from the
    // schedule, all branches are warp-uniform.  An all-information node whose LLRs contain a CA2 zero is
    // walked as a mixed node (its children are all-information again and take their own vote).
    // `types`: this node, left child, right child (2 bits each) for the out-of-line levels; `desc`: the whole subtree.
    template <int L>
    __device__ __forceinline__ void sub_body(uint32_t desc, uint32_t heap, uint32_t ob) {
        constexpr int H = 1 << (L - 1);
        const uint32_t t = (desc >> (2 * heap)) & 3u;
        if (t == T_R0) return r0_low<L>(ob);
        if (t == T_R1) {
            if (!r1_low<L>(ob)) return;
        }
        const uint32_t tl = (desc >> (2 * (2 * heap + 1))) & 3u, tr = (desc >> (2 * (2 * heap + 2))) & 3u;
        const bool left0 = tl == T_R0;
        if (!left0) {
            f_low<L>();
            sub_child<L - 1>(desc, 2 * heap + 1, ob);
        }
        if (tr == T_R0) {
            r0_low<L - 1>(ob + 4u * H);
        } else {
            g_low<L>(ob, left0);
            sub_child<L - 1>(desc, 2 * heap + 2, ob + 4u * H);
        }
        h_low<L>(ob, left0);
    }
    template <int L>
    __device__ __forceinline__ void sub_child(uint32_t desc, uint32_t heap, uint32_t ob) {
        if constexpr (L == 1)
            p2_low((desc >> (2 * heap)) & 3u, ob);
        else  // one out-of-line copy per level keeps the fused routine inside the instruction cache
            bs_sub<FMT, Q, LOG2PAR, EXT, G, L>(sm_lane, sm_blane, ll, desc, heap, ob);
    }
};

template <int FMT, int Q, int LOG2PAR, bool EXT, int G, int L>
__device__ __noinline__ void bs_sub(uint32_t sm_lane, uint32_t sm_blane, int ll, uint32_t desc, uint32_t heap, uint32_t ob) {
    BsLow<FMT, Q, LOG2PAR, EXT, G> low;
    low.ll = ll;
    low.sm_lane = sm_lane;
    low.sm_blane = sm_blane;
    low.template sub_body<L>(desc, heap, ob);
}

template <int FMT, int Q, int LOG2PAR, bool EXT, int G>
struct BsDecoder : BsLow<FMT, Q, LOG2PAR, EXT, G> {
    using Low = BsLow<FMT, Q, LOG2PAR, EXT, G>;
    using Low::FULL;
    using Low::PQ;
    using Low::ll;
    using Low::sm_lane;
    using Low::sm_blane;
    using Low::nquads;
    static constexpr int GPW = 32 / G;  // groups per warp
    static constexpr int NVQ = (Q + 3) / 4;

    const BsParams& p;
    int lane;           // lane in the warp
    uint8_t* smp_warp;  // generic pointer to the warp's first shared region
    uint8_t* smp_grp;   // ... to this lane's group region
    uint8_t* ws_warp;
    uint8_t* ws_grp;
    const uint8_t* pl_grp;  // channel planes of the group being decoded
    uint32_t bmask;     // byte mask of the shared partial-sum block
    uint32_t sm_sched;  // shared-window byte address of the schedule copy

    __device__ BsDecoder(const BsParams& p_) : p(p_) {}

    // ------------------------------------------------------------ storage of the generic levels
    // One code path for shared memory and workspace: arrays are reached through generic pointers (the
    // instruction-cache footprint of per-space variants cost more than the generic loads do).
    __device__ __forceinline__ uint8_t* aptr(int l) const {  // alpha[l]; the channel level is only read
        if ((uint32_t)l == p.log2n) return const_cast<uint8_t*>(pl_grp);
        return ((uint32_t)l <= p.lsa ? smp_grp : ws_grp) + p.aoff[l];
    }
    __device__ __forceinline__ uint8_t* bptr(int l, uint32_t o) const {  // partial sums of node (l, o)
        return (uint32_t)l <= p.lsb ? smp_grp + p.sm_beta_off + ((4u * o) & bmask) : ws_grp + p.ws_beta_off + 4ull * o;
    }
    // p0: address of the slot in plane quad 0; qs: bytes between the plane quads of the level (16 << l).
    // Callers keep p0 / qs in registers across a trip, so that the slots of a trip are immediate offsets.
    __device__ __forceinline__ static void load(const uint8_t* p0, size_t qs, bs::Val<PQ>& x) {
        uint32_t w[4 * NVQ];
#pragma unroll
        for (int v = 0; v < NVQ; v++) {
            const uint4 t = *reinterpret_cast<const uint4*>(p0 + v * qs);
            w[4 * v] = t.x;
            w[4 * v + 1] = t.y;
            w[4 * v + 2] = t.z;
            w[4 * v + 3] = t.w;
        }
        x.s = w[0];
#pragma unroll
        for (int k = 0; k < PQ; k++) x.m[k] = w[1 + k];
    }
    __device__ __forceinline__ static void store(uint8_t* p0, size_t qs, const bs::Val<PQ>& x) {
        uint32_t w[4 * NVQ];
        w[0] = x.s;
#pragma unroll
        for (int k = 1; k < 4 * NVQ; k++) w[k] = (k <= PQ) ? x.m[k - 1] : 0u;
#pragma unroll
        for (int v = 0; v < NVQ; v++)
            *reinterpret_cast<uint4*>(p0 + v * qs) = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
    }

    // ------------------------------------------------------------ generic levels (nodes of 128 LLRs and more)
    // U slots per lane and trip: all loads first, then the arithmetic, then the stores
    // alpha[l-1][i] = f(alpha[l][i], alpha[l][i + h])                          F_STATE my_module.h:373-445
    template <int U>
    __device__ __forceinline__ static void f_many(const uint8_t* src, uint8_t* dst, int l, uint32_t h, uint32_t i) {
        bs::Val<PQ> a[U], b[U], r[U];
        const size_t qs = (size_t)16 << l;
        const uint8_t* pa = src + (size_t)i * 16u;
        const uint8_t* pb = pa + (size_t)h * 16u;
        uint8_t* pd = dst + (size_t)i * 16u;
#pragma unroll
        for (int u = 0; u < U; u++) {
            load(pa + u * (G * 16), qs, a[u]);
            load(pb + u * (G * 16), qs, b[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) bs::f_op<PQ>(a[u], b[u], r[u]);
#pragma unroll
        for (int u = 0; u < U; u++) store(pd + u * (G * 16), qs >> 1, r[u]);
    }
    __device__ __forceinline__ void op_f(int l) {
        constexpr int U = SCPD_BS_U;
        const uint32_t h = 1u << (l - 1);
        const uint8_t* src = aptr(l);
        uint8_t* dst = aptr(l - 1);
        uint32_t i = ll;
        for (; i + (U - 1) * G < h; i += U * G) f_many<U>(src, dst, l, h, i);
        for (; i < h; i += G) f_many<1>(src, dst, l, h, i);
        __syncwarp();
    }
    // alpha[l-1][i] = g(alpha[l][i], alpha[l][i + h], beta[o + i]); zero: left child all-frozen, beta = 0
    //                                                                          G_STATE my_module.h:704-781
    template <int U>
    __device__ __forceinline__ static void g_many(const uint8_t* src, uint8_t* dst, const uint8_t* bs_, bool zero, int l,
                                                  uint32_t h, uint32_t i) {
        bs::Val<PQ> a[U], b[U], r[U];
        uint32_t u_[U];
        const size_t qs = (size_t)16 << l;
        const uint8_t* pa = src + (size_t)i * 16u;
        const uint8_t* pb = pa + (size_t)h * 16u;
        const uint8_t* pu = bs_ + (size_t)i * 4u;
        uint8_t* pd = dst + (size_t)i * 16u;
#pragma unroll
        for (int u = 0; u < U; u++) {
            load(pa + u * (G * 16), qs, a[u]);
            load(pb + u * (G * 16), qs, b[u]);
            u_[u] = zero ? 0u : *reinterpret_cast<const uint32_t*>(pu + u * (G * 4));
        }
#pragma unroll
        for (int u = 0; u < U; u++) bs::g_sat<FMT, PQ>(a[u], b[u], u_[u], r[u]);
#pragma unroll
        for (int u = 0; u < U; u++) store(pd + u * (G * 16), qs >> 1, r[u]);
    }
    // g reads a level that was written a whole subtree ago: when it lives in the workspace it has usually left
    // L2, so one lane asks for the whole array up front and only the first trip pays the DRAM latency.
    __device__ __forceinline__ void prefetch_alpha(const uint8_t* src, int l) const {
#if defined(__CUDACC__) && !defined(SCPD_BS_NO_PREFETCH)
        if (p.prefetch && (uint32_t)l > p.lsa && ll == 0) {
            const uint32_t bytes = ((uint32_t)NVQ * 16u) << l;
            for (uint32_t off = 0; off < bytes; off += 32768u) {
                const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + off), "r"(n));
            }
        }
#endif
    }
    __device__ __forceinline__ void op_g(int l, uint32_t o, bool zero) {
        constexpr int U = SCPD_BS_U;
        const uint32_t h = 1u << (l - 1);
        const uint8_t* src = aptr(l);
        prefetch_alpha(src, l);
        uint8_t* dst = aptr(l - 1);
        const uint8_t* bs_ = bptr(l - 1, o);
        uint32_t i = ll;
        for (; i + (U - 1) * G < h; i += U * G) g_many<U>(src, dst, bs_, zero, l, h, i);
        for (; i < h; i += G) g_many<1>(src, dst, bs_, zero, l, h, i);
        __syncwarp();
    }
    // Fused X(l) F(l-1) [F(l-2)]: a lane takes the DEPTH-1 levels of pairs that hang under its slot j, i.e. the
    // 2^(DEPTH-1) outputs j + k * (h >> (DEPTH-1)) of the first op, and carries them on in registers.  Every
    // level is still written (the g of that level needs it later) but nothing is read back.
    template <int DEPTH>
    __device__ __forceinline__ void op_fuse(int l, uint32_t o, uint32_t kind) {
        constexpr int W = 1 << (DEPTH - 1);  // outputs of the first op per lane and trip
        const uint32_t h = 1u << (l - 1), q = h >> (DEPTH - 1);
        const uint8_t* src = aptr(l);
        uint8_t* d1 = aptr(l - 1);
        uint8_t* d2 = aptr(l - 2);
        uint8_t* d3 = DEPTH == 3 ? aptr(l - 3) : nullptr;
        const uint8_t* bs_ = bptr(l - 1, o);
        const size_t qs = (size_t)16 << l, hs = (size_t)h * 16u, ks = (size_t)q * 16u;
        for (uint32_t j = ll; j < q; j += G) {
            bs::Val<PQ> a[W], b[W], r[W];
            uint32_t u_[W];
            const uint8_t* pa = src + (size_t)j * 16u;
#pragma unroll
            for (int k = 0; k < W; k++) {
                load(pa + k * ks, qs, a[k]);
                load(pa + k * ks + hs, qs, b[k]);
                u_[k] = kind == BSK_G ? *reinterpret_cast<const uint32_t*>(bs_ + 4u * (j + k * q)) : 0u;
            }
            if (kind == BSK_F) {
#pragma unroll
                for (int k = 0; k < W; k++) bs::f_op<PQ>(a[k], b[k], r[k]);
            } else {
#pragma unroll
                for (int k = 0; k < W; k++) bs::g_sat<FMT, PQ>(a[k], b[k], u_[k], r[k]);
            }
#pragma unroll
            for (int k = 0; k < W; k++) store(d1 + (size_t)j * 16u + k * ks, qs >> 1, r[k]);
            // F(l-1): pairs (x, x + h/2) = outputs k and k + W/2
            bs::Val<PQ> r2[W / 2];
#pragma unroll
            for (int k = 0; k < W / 2; k++) {
                bs::f_op<PQ>(r[k], r[k + W / 2], r2[k]);
                store(d2 + (size_t)j * 16u + k * ks, qs >> 2, r2[k]);
            }
            if constexpr (DEPTH == 3) {  // F(l-2): pair (x, x + h/4)
                bs::Val<PQ> r3;
                bs::f_op<PQ>(r2[0], r2[1], r3);
                store(d3 + (size_t)j * 16u, qs >> 3, r3);
            }
        }
        __syncwarp();
    }
    // node (l,o) := (left ^ right, right); copy: the left child is all-frozen          H_STATE my_module.h:903-932
    // When the children sit in the shared block and the node does not, the node moves to the workspace.
    __device__ __forceinline__ void op_h(int l, uint32_t o, bool copy) {
        const uint32_t h = 1u << (l - 1);
        const uint8_t* cl = bptr(l - 1, o);
        const uint8_t* cr = bptr(l - 1, o + h);
        uint8_t* d = bptr(l, o);
        const bool move = (uint32_t)(l - 1) <= p.lsb && (uint32_t)l > p.lsb;
        for (uint32_t i = 4u * ll; i < h; i += 4u * G) {  // h >= 64: a multiple of 4G, or 4 * ll < h selects the lanes
            uint4 x = *reinterpret_cast<const uint4*>(cr + 4u * i);
            if (move) *reinterpret_cast<uint4*>(d + 4u * (h + i)) = x;
            if (!copy) {
                const uint4 y = *reinterpret_cast<const uint4*>(cl + 4u * i);
                x.x ^= y.x;
                x.y ^= y.y;
                x.z ^= y.z;
                x.w ^= y.w;
            }
            *reinterpret_cast<uint4*>(d + 4u * i) = x;
        }
        __syncwarp();
    }
    __device__ __forceinline__ void op_r0(int l, uint32_t o) {
        uint8_t* d = bptr(l, o);
        for (uint32_t i = 4u * ll; i < (1u << l); i += 4u * G) {
            *reinterpret_cast<uint4*>(d + 4u * i) = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();
    }
    // hard decision of node (l,o); returns true (warp-uniform) if some LLR of some frame is a CA2 zero
    __device__ __forceinline__ bool op_hd(int l, uint32_t o) {
        const uint8_t* src = aptr(l);
        uint8_t* d = bptr(l, o);
        uint32_t z = 0u;
        for (uint32_t i = ll; i < (1u << l); i += G) {
            bs::Val<PQ> a;
            load(src + (size_t)i * 16u, (size_t)16 << l, a);
            uint32_t x = a.s;
            if (FMT == bs::FMT_CA2) {
                const uint32_t nz = bs::nonzero<PQ>(a);
                z |= ~nz;
                x &= nz;
            }
            *reinterpret_cast<uint32_t*>(d + 4u * i) = x;
        }
        bool any = false;
        if (FMT == bs::FMT_CA2) any = __any_sync(FULL, z != 0u);
        __syncwarp();
        return any;
    }

    __device__ __forceinline__ uint32_t fetch(uint32_t pc) const {
        return p.sched_words ? lds32(sm_sched + 4u * pc) : __ldg(p.sched + pc);
    }
    __device__ __forceinline__ void run() {
        uint32_t pc = 0;
        uint32_t w = fetch(pc);
        for (;;) {
            const uint32_t wn = fetch(pc + 1);  // the next word: an op, node types, or a skip count
            // Ops outside every rate-1 fallback region are executed by all warps of the CTA: meeting there
            // keeps the warps in the same stretch of code (one instruction-cache fill serves all of them).
            if (bs_op_sync(w)) __syncthreads();
            const int l = (int)bs_op_level(w);
            const uint32_t o = bs_op_offset(w);
            const uint32_t ob = (4u * o) & bmask;
            const uint32_t code = bs_op_code(w);
            uint32_t adv = 1;
            bool z;
            if (code == 8 * BS_LSUB + BSK_F) {  // by far the most frequent op
                this->template sub_body<BS_LSUB>(wn, 0u, ob);
                adv = 2;
            } else {
                switch (code) {
                    case 0: return;
                    case BS_NOP: break;
                    case BS_FUSE2 + BSK_F:
                    case BS_FUSE2 + BSK_G:
                    case BS_FUSE2 + BSK_G0: op_fuse<2>(l, o, code - BS_FUSE2); break;
                    case BS_FUSE3 + BSK_F:
                    case BS_FUSE3 + BSK_G:
                    case BS_FUSE3 + BSK_G0: op_fuse<3>(l, o, code - BS_FUSE3); break;
                    case 8 * BS_LSUB + BSK_R0: this->template r0_low<BS_LSUB>(ob); break;
#define BS_LOW(L)                                                           \
    case 8 * L + BSK_F: this->template f_low<L>(); break;                  \
    case 8 * L + BSK_G:                                                     \
    case 8 * L + BSK_G0: this->template g_low<L>(ob, code == 8 * L + BSK_G0); break;   \
    case 8 * L + BSK_H:                                                     \
    case 8 * L + BSK_HCOPY: this->template h_low<L>(ob, code == 8 * L + BSK_HCOPY); break; \
    case 8 * L + BSK_R0: this->template r0_low<L>(ob); break;              \
    case 8 * L + BSK_R1:                                                    \
        z = this->template r1_low<L>(ob);                                  \
        adv = 2;                                                            \
        if (!z) adv += wn;                                                  \
        break;
                    BS_LOW(5) BS_LOW(6)
#undef BS_LOW
                    case 8 * (BS_LLOW + 1) + BSK_F: op_f(l); break;
                    case 8 * (BS_LLOW + 1) + BSK_G:
                    case 8 * (BS_LLOW + 1) + BSK_G0: op_g(l, o, code == 8 * (BS_LLOW + 1) + BSK_G0); break;
                    case 8 * (BS_LLOW + 1) + BSK_H:
                    case 8 * (BS_LLOW + 1) + BSK_HCOPY: op_h(l, o, code == 8 * (BS_LLOW + 1) + BSK_HCOPY); break;
                    case 8 * (BS_LLOW + 1) + BSK_R0: op_r0(l, o); break;
                    case 8 * (BS_LLOW + 1) + BSK_R1:  // hard decision; the plain-SC ops behind it run only on a CA2 zero
                        z = op_hd(l, o);
                        adv = 2;
                        if (!z) adv += wn;
                        break;
                    default: break;
                }
            }
            pc += adv;
            w = (adv == 1) ? wn : fetch(pc);
        }
    }

    // final partial sums of the root (one word per code bit) -> packed rows of the 32 frames
    // (wrapper_out.h:31-33 laid end to end).  Lane L transposes the 32 code bits of word L, L + 32, ...
    __device__ __forceinline__ void write_output(unsigned long long g0) {
        const bool bsm = p.log2n <= p.lsb;
        for (int s = 0; s < GPW; s++) {
            const unsigned long long f0 = (g0 + s) * 32ull;
            const uint32_t nvalid = f0 >= p.nframes ? 0u : (p.nframes - f0 < 32ull ? (uint32_t)(p.nframes - f0) : 32u);
            const uint8_t* b = bsm ? smp_warp + (size_t)s * p.sm_stride + p.sm_beta_off
                                   : ws_warp + (unsigned long long)s * p.ws_stride + p.ws_beta_off;
            for (uint32_t wd = lane; wd < p.wpf; wd += 32u) {
                uint32_t a[32];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const uint4 t = *reinterpret_cast<const uint4*>(b + 128u * wd + 16u * q);
                    a[4 * q] = t.x;
                    a[4 * q + 1] = t.y;
                    a[4 * q + 2] = t.z;
                    a[4 * q + 3] = t.w;
                }
                bs::transpose32(a);
#pragma unroll
                for (int f = 0; f < 32; f++)
                    if ((uint32_t)f < nvalid) p.xhat[(f0 + f) * p.wpf + wd] = a[f];
            }
        }
        __syncwarp();
    }
    // g0: first of the GPW groups this warp decodes together
    __device__ __forceinline__ void decode_groups(unsigned long long g0) {
        // lane groups past the end of the batch decode the last group again and write nothing
        const unsigned long long g = g0 + (unsigned long long)(lane / G);
        pl_grp = p.planes + (g < p.ngroups ? g : p.ngroups - 1) * p.planes_stride;
        run();
        write_output(g0);
    }
};

// int8 rows -> bit planes: alpha[log2n] of every group, (sign, |llr|) in the plane-quad layout of the decode
// kernel (quad v of LLR i at byte ((v << log2n) + i) * 16).  One warp per (group, 128 LLRs): lane L reads the
// four LLRs 128c + 4L .. +3 of each of the 32 frames (coalesced 128-byte row segments), transposes the
// 32 x 32 bit matrix in registers and converts two's complement to sign-magnitude.      wrapper_in.h:30-42
template <int Q>
__global__ void __launch_bounds__(256) bs_planes_kernel(const int8_t* __restrict__ llr, unsigned long long nframes,
                                                        uint32_t n, uint32_t log2n, uint8_t* __restrict__ planes,
                                                        unsigned long long planes_stride) {
    constexpr int PQ = Q - 1, NVQ = (Q + 3) / 4;
    const int lane = threadIdx.x & 31;
    const unsigned long long chunks = n / 128u, ngroups = (nframes + 31) / 32;
    const unsigned long long ntasks = ngroups * chunks;
    const unsigned long long wstride = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    for (unsigned long long t = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < ntasks; t += wstride) {
        const unsigned long long g = t / chunks;
        const uint32_t c = (uint32_t)(t % chunks) * 128u;
        const unsigned long long f0 = g * 32ull;
        const uint32_t nvalid = nframes - f0 < 32ull ? (uint32_t)(nframes - f0) : 32u;
        const int8_t* row0 = llr + f0 * n + c + 4u * lane;
        uint32_t a[32];
#pragma unroll
        for (int f = 0; f < 32; f++)
            a[f] = ((uint32_t)f < nvalid) ? __ldg(reinterpret_cast<const uint32_t*>(row0 + (size_t)f * n)) : 0u;
        bs::transpose32(a);
        uint8_t* dst = planes + g * planes_stride;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t v[8] = {a[8 * k], a[8 * k + 1], a[8 * k + 2], a[8 * k + 3],
                                   a[8 * k + 4], a[8 * k + 5], a[8 * k + 6], a[8 * k + 7]};
            bs::Val<PQ> x;
            bs::from_int8_planes<PQ>(v, x);
            uint32_t w[4 * NVQ];
            w[0] = x.s;
#pragma unroll
            for (int j = 1; j < 4 * NVQ; j++) w[j] = (j <= PQ) ? x.m[j - 1] : 0u;
            const uint32_t i = c + 4u * lane + k;
#pragma unroll
            for (int q = 0; q < NVQ; q++)
                *reinterpret_cast<uint4*>(dst + (((size_t)q << log2n) + i) * 16u) =
                    make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        }
    }
}

// Every warp of a CTA walks the same schedule the same number of times (the barriers in run() rely on it);
// warps or lane groups without frames decode zeros and write nothing.
template <int FMT, int Q, int LOG2PAR, bool EXT, int G>
__global__ void __launch_bounds__(128, SCPD_BS_MINB) sc_decode_bs_kernel(const BsParams p) {
    extern __shared__ __align__(16) uint8_t smem_fast[];
    constexpr int GPW = 32 / G;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    BsDecoder<FMT, Q, LOG2PAR, EXT, G> d(p);
    d.lane = threadIdx.x & 31;
    d.ll = d.lane % G;
    const int sg = d.lane / G;
    d.smp_warp = smem_fast + (size_t)warp * GPW * p.sm_stride;
    d.smp_grp = d.smp_warp + (size_t)sg * p.sm_stride;
    const uint32_t sm_grp = smem_u32(d.smp_grp);
    d.sm_lane = sm_grp + 16u * d.ll;
    d.sm_blane = sm_grp + p.sm_beta_off + 4u * d.ll;
    d.bmask = (8u << p.lsb) - 4u;
    d.sm_sched = smem_u32(smem_fast + (size_t)nwarps * GPW * p.sm_stride);
    if (p.sched_words) {
        for (uint32_t i = threadIdx.x; i < p.sched_words; i += blockDim.x) sts32(d.sm_sched + 4u * i, __ldg(p.sched + i));
        __syncthreads();
    }
#if defined(__CUDACC__)
    // keep the per-lane addressing state in registers: without this ptxas rebuilds it from %tid and the
    // kernel parameters in front of every op
    asm volatile("" : "+r"(d.ll), "+r"(d.sm_lane), "+r"(d.sm_blane), "+r"(d.bmask));
#endif
    const unsigned long long slot = (unsigned long long)blockIdx.x * nwarps + warp;
    d.ws_warp = p.ws + slot * GPW * p.ws_stride;
    d.ws_grp = d.ws_warp + (unsigned long long)sg * p.ws_stride;
    const unsigned long long gpc = (unsigned long long)nwarps * GPW;  // groups per CTA and round
    for (unsigned long long base = (unsigned long long)blockIdx.x * gpc; base < p.ngroups; base += (unsigned long long)gridDim.x * gpc)
        d.decode_groups(base + (unsigned long long)warp * GPW);
}

}  // namespace scpd
