// ss_plan.h -- host side of the slot-sliced kernel (decode_ss.cuh): the static SC tree walk compiled into op
// words, and the memory plan (which LLR levels / partial sums stay in shared memory, where the rest sits in
// the per-warp workspace).  Shared by the C ABI (scpd_api.cu) and by the CPU emulation of the test tier.
//
// The reference walks the tree with an FSM and two packed stacks (my_module.h:228-263); every frame of a batch
// shares the frozen set, so here the walk is a straight-line program.  do_prunning's node classification
// (my_module.h:96-154) happens here as well: all-frozen (R0) and all-information (R1) nodes, nothing else.
//
// Op word: [5:0] opcode, [10:6] level l (node of 2^l LLRs), [31:11] node offset / 32.
//   SS_F / SS_G / SS_G0     alpha[l-1] = f / g / g(beta = 0) of alpha[l]                 (l >= 7)
//   SS_H / SS_HCOPY         beta(l,o) = (left ^ right, right) / (right, right)            (l >= 7)
//   SS_R0                   beta(l,o) = 0                                                   (l >= 6)
//   SS_R1 + skip count      hard decision of alpha[l]; the plain-SC ops that follow (skip count words) run only
//                           when some LLR of the node is a CA2 zero in some frame of the warp    (l >= 7)
//   SS_SUB + 4 words        a whole node of 64 LLRs from alpha[6]: word 0 = node types of (64, left 32, right 32), 2 bits
//                           each, [9:8] kind of a fused level-7 op, [10] pruning mode != NONE; words 1-2 = the 64
//                           information flags; word 3 = pattern ids of the eight 8-LLR nodes, 4 bits each
//   SS_XS + 4 words         SS_F / SS_G / SS_G0 at level 7 fused with the SS_SUB of the child it feeds: alpha[6]
//                           stays in registers for the child's f (it is still stored for the child's g)
//   SS_XF_F / _G / _G0      SS_F / SS_G / SS_G0 at level l fused with the SS_F(l-1) that opens the child: both levels are
//                           written (the g ops need them), alpha[l-1] is not read back.  Only where the three levels
//                           stream through global memory (large trees: the walk is bound by DRAM there)
// Node types: 0 mixed, 1 all-frozen, 2 all-information; for nodes of 2 LLRs the code is the flag pair itself:
// 0 = (0,1), 1 = (0,0), 2 = (1,1), 3 = (1,0).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace scpd {

enum : uint32_t {
    SS_END = 0,
    SS_F = 1,
    SS_G = 2,
    SS_G0 = 3,
    SS_H = 4,
    SS_HCOPY = 5,
    SS_R0 = 6,
    SS_R1 = 7,
    SS_SUB = 8,
    SS_XS = 9,      // level-7 op (kind in bits [31:30] of the first descriptor word's upper bits, see below) + SUB
    SS_XF_F = 10,   // SS_F / SS_G / SS_G0 at (l, o) + SS_F(l - 1) of the child it feeds
    SS_XF_G = 11,
    SS_XF_G0 = 12,
};
enum : uint32_t { SS_T_MIX = 0, SS_T_R0 = 1, SS_T_R1 = 2 };
// information-flag patterns of 8-LLR nodes with a specialised routine (bit i = flag of position i); index = pattern
// id.  These nine cover every frozen table shipped with the reference (Frozen_Bit_Tab/, Generated_Frozen_Bit/).
static const uint32_t SS_KNOWN8[9] = {0x00u, 0xFFu, 0xFEu, 0xE8u, 0x80u, 0xE0u, 0xFCu, 0xF8u, 0xC0u};

#if defined(__CUDACC__)
#define SS_HD __host__ __device__
#else
#define SS_HD
#endif
SS_HD static inline uint32_t ss_op_make(uint32_t code, uint32_t level, uint32_t offset) {
    return code | (level << 6) | ((offset >> 5) << 11);
}
SS_HD static inline uint32_t ss_op_code(uint32_t w) { return w & 63u; }
SS_HD static inline uint32_t ss_op_level(uint32_t w) { return (w >> 6) & 31u; }
SS_HD static inline uint32_t ss_op_word(uint32_t w) { return (w >> 11) & 0xFFFFFu; }  // node offset / 32 = first partial-sum word

struct SsStats {
    uint64_t n_ops = 0, n_f = 0, n_g = 0, n_r0 = 0, n_r1 = 0, n_sub = 0, n_sub32_mixed = 0, n_xf = 0;
};

struct SsBuilder {
    int log2n = 0, pruning = 0;
    int fuse = 1;  // (kept for the callers' signature: the level-7 op is always SS_XS)
    int xf_min = 0;  // > 0: fuse an f / g op at level l with the f that opens its child when l - 2 >= xf_min
    int lead_skip = 0;  // the f ops of the leftmost nodes of the top lead_skip levels are computed by ss_planes_kernel
    const uint8_t* flags = nullptr;
    std::vector<uint32_t> psum;
    std::vector<uint32_t> ops;
    SsStats st;

    uint32_t count(uint32_t o, uint32_t n) const { return psum[o + n] - psum[o]; }
    uint32_t type(uint32_t o, uint32_t n) const {
        if (n == 2) {
            const uint32_t f0 = flags[o] & 1u, f1 = flags[o + 1] & 1u;
            return (f0 == 0 && f1 == 1) ? 0u : (f0 == 0 && f1 == 0) ? 1u : (f0 == 1 && f1 == 1) ? 2u : 3u;
        }
        const uint32_t c = count(o, n);
        return (pruning >= 1 && c == 0) ? SS_T_R0 : (pruning >= 2 && c == n) ? SS_T_R1 : SS_T_MIX;
    }
    uint32_t flags32(uint32_t o) const {
        uint32_t v = 0;
        for (uint32_t i = 0; i < 32; i++) v |= (uint32_t)(flags[o + i] & 1u) << i;
        return v;
    }
    // pattern id of the 8-LLR node at offset o: 0 = all-frozen and pruned, 1.. = index into SS_KNOWN8,
    // 15 = decode from the flags at run time
    uint32_t id8(uint32_t o) const {
        const uint32_t fb = (flags32(o & ~31u) >> (o & 31u)) & 255u;
        if (pruning < 1) return 15u;  // mode NONE: every node is decoded
        if (fb == 0u) return 0u;
        for (uint32_t k = 1; k < sizeof(SS_KNOWN8) / sizeof(SS_KNOWN8[0]); k++)
            if (SS_KNOWN8[k] == fb) return k;
        return 15u;
    }
    // work estimate of the 32-LLR node (f / g element updates the half2 walker executes)
    void count32(uint32_t o, uint32_t n) {
        if (n == 2) return;
        const uint32_t t = type(o, n);
        if (t == SS_T_R0) return;
        const uint32_t h = n / 2;
        const bool l0 = pruning >= 1 && count(o, h) == 0, r0 = pruning >= 1 && count(o + h, h) == 0;
        if (!l0) {
            st.n_f += h;
            count32(o, h);
        }
        if (!r0) {
            st.n_g += h;
            count32(o + h, h);
        }
    }
    void push_sub_words(uint32_t o, uint32_t kind) {
        uint32_t w0 = type(o, 64) | (type(o, 32) << 2) | (type(o + 32, 32) << 4) | (kind << 8);
        if (pruning >= 1) w0 |= 1u << 10;
        ops.push_back(w0);
        ops.push_back(flags32(o));
        ops.push_back(flags32(o + 32));
        uint32_t ids = 0;
        for (uint32_t k = 0; k < 8; k++) ids |= id8(o + 8 * k) << (4 * k);
        ops.push_back(ids);
        st.n_sub++;
        for (uint32_t k = 0; k < 2; k++) {
            const uint32_t t = type(o + 32 * k, 32);
            if (t == SS_T_MIX) st.n_sub32_mixed++;
            if (t == SS_T_MIX && type(o, 64) == SS_T_MIX) count32(o + 32 * k, 32);
        }
        if (type(o, 64) == SS_T_MIX) {
            if (type(o, 32) != SS_T_R0) st.n_f += 32;
            if (type(o + 32, 32) != SS_T_R0) st.n_g += 32;
        }
    }
    // the op that produces alpha[l-1] of a child: kind = SS_F / SS_G / SS_G0 at (l, o); child at (l-1, oc)
    void emit_x(uint32_t kind, int l, uint32_t o, uint32_t oc) {
        const uint32_t h = 1u << (l - 1);
        if (kind == SS_F) st.n_f += h; else st.n_g += h;
        if (l == 7) {  // always fused: the plain f / g op pipelines four chunk pairs and does not exist at level 7
            // fused with the 64-LLR child (which is never all-frozen here: the caller prunes those)
            ops.push_back(ss_op_make(SS_XS, 7, oc));
            push_sub_words(oc, kind);
            xs_done = true;
            return;
        }
        // fused with the child's opening f: the child (l-1, oc) is a mixed node above the 64-LLR nodes whose left half is
        // not all-frozen, i.e. emit() will start it with SS_F(l-1, oc) -- which skip_f then drops
        if (xf_min > 0 && l - 2 >= xf_min && l - 2 >= 7 && !in_r1) {
            const uint32_t cc = count(oc, h);
            const bool mixed = !(pruning >= 1 && cc == 0) && !(pruning >= 2 && cc == h);
            if (mixed && !(pruning >= 1 && count(oc, h >> 1) == 0)) {
                ops.push_back(ss_op_make(kind == SS_F ? SS_XF_F : kind == SS_G ? SS_XF_G : SS_XF_G0, l, o));
                st.n_f += h >> 1;
                st.n_xf++;
                skip_f = true;
                return;
            }
        }
        ops.push_back(ss_op_make(kind, l, o));
    }
    bool xs_done = false;
    bool skip_f = false;  // the next emit() finds its opening f already done by an SS_XF_* op
    bool in_r1 = false;   // inside the plain-SC walk behind an SS_R1 (not fused: the exception path stays simple)

    void emit_child(int l, uint32_t o, bool r1) {
        if (xs_done) {  // the SS_XS in front already decoded this 64-LLR node
            xs_done = false;
            return;
        }
        if (r1)
            emit_r1(l, o);
        else
            emit(l, o);
    }
    // all-information node: hard decision, with its plain-SC walk behind it (children are all-information again)
    void emit_r1(int l, uint32_t o) {
        if (l == 6) {
            ops.push_back(ss_op_make(SS_SUB, 6, o));
            push_sub_words(o, 0);
            return;
        }
        ops.push_back(ss_op_make(SS_R1, l, o));
        st.n_r1++;
        const size_t at = ops.size();
        ops.push_back(0u);
        const uint32_t h = 1u << (l - 1);
        // work counters: the fallback is the exception, do not count it
        const SsStats keep = st;
        const bool was = in_r1;
        in_r1 = true;
        emit_x(SS_F, l, o, o);
        emit_child(l - 1, o, true);
        emit_x(SS_G, l, o, o + h);
        emit_child(l - 1, o + h, true);
        ops.push_back(ss_op_make(SS_H, l, o));
        in_r1 = was;
        st = keep;
        ops[at] = (uint32_t)(ops.size() - at - 1);
    }
    void emit(int l, uint32_t o) {
        const uint32_t n = 1u << l, c = count(o, n);
        if (pruning >= 1 && c == 0) {
            ops.push_back(ss_op_make(SS_R0, l, o));
            st.n_r0++;
            return;
        }
        if (l == 6) {
            ops.push_back(ss_op_make(SS_SUB, 6, o));
            push_sub_words(o, 0);
            return;
        }
        if (pruning >= 2 && c == n) {
            emit_r1(l, o);
            return;
        }
        const uint32_t h = n >> 1;
        const bool left_r0 = pruning >= 1 && count(o, h) == 0;
        const bool right_r0 = pruning >= 1 && count(o + h, h) == 0;
        if (left_r0) {
            emit_x(SS_G0, l, o, o + h);
            emit_child(l - 1, o + h, false);
            ops.push_back(ss_op_make(SS_HCOPY, l, o));
            return;
        }
        if (skip_f)
            skip_f = false;  // done by the SS_XF_* op of the parent
        else if (o == 0 && l > log2n - lead_skip && !in_r1)
            st.n_f += h;     // done by the plane conversion (ss_prefuse_depth)
        else
            emit_x(SS_F, l, o, o);
        emit_child(l - 1, o, false);
        if (right_r0) {
            ops.push_back(ss_op_make(SS_R0, l - 1, o + h));
            st.n_r0++;
            ops.push_back(ss_op_make(SS_H, l, o));
            return;
        }
        emit_x(SS_G, l, o, o + h);
        emit_child(l - 1, o + h, false);
        ops.push_back(ss_op_make(SS_H, l, o));
    }
};

// flags: n bytes, 1 = information bit; log2n >= 7.  pruning: 0 none, 1 all-frozen nodes, 2 + all-information nodes.
static inline std::vector<uint32_t> ss_build_schedule(int log2n, int pruning, const uint8_t* flags, SsStats* stats,
                                                      int fuse = 1, int xf_min = 0, int lead_skip = 0) {
    SsBuilder b;
    b.log2n = log2n;
    b.pruning = pruning;
    b.fuse = fuse;
    b.xf_min = xf_min;
    b.lead_skip = lead_skip;
    b.flags = flags;
    const uint32_t n = 1u << log2n;
    b.psum.assign(n + 1, 0);
    for (uint32_t i = 0; i < n; i++) b.psum[i + 1] = b.psum[i] + (flags[i] ? 1u : 0u);
    b.emit(log2n, 0);
    b.ops.push_back(0u);
    b.ops.push_back(0u);
    b.st.n_ops = b.ops.size();
    if (stats) *stats = b.st;
    return b.ops;
}

// ---------------------------------------------------------------------------------------------- memory plan
// Units: uint4 (16 bytes).  Per lane an LLR chunk (32 consecutive code positions of one frame, Q bit planes) is
// two uint4 "quads"; quad q of chunk c of a level sits at uint4 index (2c + q) * 32 + lane, so a warp's access
// to one quad is 512 contiguous bytes (conflict-free in shared memory, coalesced in the workspace).
// Partial sums: word w (code positions 32w .. 32w+31 of the lane's frame) is component w & 3 of the uint4 at
// index (w >> 2) * 32 + lane.
struct SsPlan {
    uint32_t lsa = 0;    // alpha levels 6 .. lsa in shared memory, lsa+1 .. log2n-1 in the workspace
    uint32_t lwin = 0;   // partial sums of nodes below level lwin in the shared window (2^lwin code positions)
    uint32_t ltm = 0;    // this alpha level (above lsa) lives in tensor memory instead; 0 = none
    uint32_t tm_cols = 0;  // tensor-memory columns per warp (8 per chunk: 2^(ltm - 2))
    uint32_t aoff[24] = {0};
    uint32_t sm_beta_off = 0, sm_stride = 0;  // per warp
    uint32_t ws_beta_off = 0;
    unsigned long long ws_stride = 0;  // per warp
    // the smallest workspace level (rewritten 2^(log2n - lhot) times per task) may live in its own dense array, one block
    // of hot_stride uint4 per warp, so that an L2 persisting window can cover exactly those lines; 0 = in the workspace
    uint32_t lhot = 0, hot_stride = 0;
    uint32_t win_words = 0;
};
// uint4 per 32-frame task: the channel planes, then alpha[log2n - s] of the leftmost node for s = 1 .. pre (the leading
// f chain computed by ss_planes_kernel); off[s] = where level log2n - s starts
static inline size_t ss_planes_quads(int log2n, int pre = 0, uint32_t* off = nullptr) {
    size_t q = 0;
    for (int s = 0; s <= pre; s++) {
        if (off) off[s] = (uint32_t)q;
        q += (size_t)64u << (log2n - s - 5);
    }
    return q;
}
// How many leading ops of the schedule (built without fused ops and with lead_skip = 0) are F(log2n, 0), F(log2n - 1, 0),
// ... whose result neither lives on chip (level above lsa) nor is smaller than a unit of the plane kernel: those are
// computed by ss_planes_kernel; the schedule is then built again with lead_skip = that depth.
static inline int ss_prefuse_depth(const std::vector<uint32_t>& ops, int log2n, uint32_t lsa, int max_depth) {
    int d = 0;
    while (d < max_depth && d < 3 && (size_t)d < ops.size() && ss_op_code(ops[d]) == SS_F && (int)ss_op_level(ops[d]) == log2n - d &&
           ss_op_word(ops[d]) == 0u && log2n - d - 1 > (int)lsa && log2n - d - 1 >= 8 && log2n >= 8)
        d++;
    return d;
}

// smem_per_warp in bytes.  Returns false when even the minimum (alpha[6] + a 256-position window) does not fit.
// tm_cols_avail: tensor-memory columns a warp may use (512 / ceil(warps per CTA / 4)), 0 = do not use tensor memory;
// want_ltm: -1 = the largest level above lsa that fits, 0 = none, else that level.
static inline bool ss_make_plan(int log2n, size_t smem_per_warp, SsPlan* out, int force_lsa = -1, int force_lwin = -1,
                                uint32_t tm_cols_avail = 0, int want_ltm = 0, bool hot = false) {
    SsPlan p;
    auto a_quads = [](int l) { return (size_t)64u << (l - 5); };
    auto win_quads = [&](int lwin) { return (size_t)32u * std::max<size_t>(1, (std::min<size_t>((size_t)1 << lwin, (size_t)1 << log2n) / 32 + 3) / 4); };
    int lwin = std::min(log2n + 1, 10);
    int lsa = 6;
    auto total = [&](int la, int lw) {
        size_t s = win_quads(lw);
        for (int l = 6; l <= la; l++) s += a_quads(l);
        return s * 16;
    };
    if (total(lsa, 8) > smem_per_warp) return false;
    while (lwin > 8 && total(lsa, lwin) > smem_per_warp) lwin--;
    while (lsa + 1 <= log2n - 1 && total(lsa + 1, lwin) <= smem_per_warp) lsa++;
    // a whole small frame's partial sums in the window when there is room
    if (lwin == 10 && log2n >= 10 && log2n <= 11 && total(lsa, log2n + 1) <= smem_per_warp) lwin = log2n + 1;
    if (force_lsa >= 6) lsa = std::min(std::max(6, force_lsa), std::max(6, log2n - 1));
    if (force_lwin >= 8) lwin = std::min(force_lwin, log2n + 1);
    if (lsa > log2n - 1) lsa = std::max(6, log2n - 1);
    if (total(lsa, lwin) > smem_per_warp) return false;
    p.lsa = (uint32_t)lsa;
    p.lwin = (uint32_t)lwin;
    size_t off = 0;
    for (int l = 6; l <= lsa && l <= log2n - 1; l++) {
        p.aoff[l] = (uint32_t)off;
        off += a_quads(l);
    }
    p.sm_beta_off = (uint32_t)off;
    off += win_quads(lwin);
    p.sm_stride = (uint32_t)off;
    p.win_words = (uint32_t)(std::min<size_t>((size_t)1 << lwin, (size_t)1 << log2n) / 32);
    int ltm = want_ltm;
    if (want_ltm < 0)  // the largest level that fits: every level costs the same traffic per LLR, the largest holds the most
        for (ltm = log2n - 1; ltm > lsa && ((uint32_t)1 << (ltm - 2)) > tm_cols_avail;) ltm--;
    if (ltm <= lsa || ltm < 8 || ltm > log2n - 1 || ((uint32_t)1 << (ltm - 2)) > tm_cols_avail) ltm = 0;
    p.ltm = (uint32_t)ltm;
    p.tm_cols = ltm ? (uint32_t)1 << (ltm - 2) : 0u;
    size_t woff = 0;
    for (int l = lsa + 1; l <= log2n - 1; l++) {
        if (l == ltm) continue;
        if (hot && p.lhot == 0) {  // the first (smallest) workspace level
            p.lhot = (uint32_t)l;
            p.hot_stride = (uint32_t)a_quads(l);
            p.aoff[l] = 0;
            continue;
        }
        p.aoff[l] = (uint32_t)woff;
        woff += a_quads(l);
    }
    p.ws_beta_off = (uint32_t)woff;
    woff += (size_t)32u * ((((size_t)1 << log2n) / 32 + 3) / 4);
    p.ws_stride = woff;
    *out = p;
    return true;
}

}  // namespace scpd
