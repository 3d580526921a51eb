// bs_plan.h -- host-side memory plan of the bit-sliced kernel (decode_bs.cuh): which alpha levels and
// how much of the partial sums stay in shared memory, and where everything else sits in the workspace.
// Shared by the C ABI (scpd_api.cu) and by the CPU warp emulator of the test tier.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "schedule.h"

namespace scpd {

// ---- op words of the bit-sliced kernel: [5:0] code = kind + 8 * min(level, BS_LLOW + 1), [10:6] level,
// [30:11] offset of the node in the frame.  Level 1 holds the 2-bit terminals: code 8 + flags (f0 | f1 << 1),
// and 8 + BSK_R0 for a pruned all-frozen pair.  Word count and order equal the generic schedule's, so the
// skip counts behind OP_R1 stay valid.
#define BS_LLOW 6
#define BS_LSUB 4  // nodes of 2^BS_LSUB LLRs are decoded by one fused routine (code 8 * BS_LSUB + BSK_F + node types)
enum : uint32_t { BSK_END = 0, BSK_F = 1, BSK_G = 2, BSK_G0 = 3, BSK_H = 4, BSK_HCOPY = 5, BSK_R0 = 6, BSK_R1 = 7 };
// Fused ops (generic levels only): X(l) immediately followed by F(l-1) [and F(l-2)] runs as one pass that keeps
// the freshly computed LLRs in registers, so alpha[l-1] (and alpha[l-2]) are written but not read back.  Codes
// 8 + kind (two levels) and 16 + kind (three levels), kind = F / G / G0 of the first op; the subsumed F words
// become BS_NOP so that word offsets (skip counts) stay valid.
#define BS_FUSE2 8
#define BS_FUSE3 16
#define BS_NOP (8 * (BS_LLOW + 1))
SCPD_HD static inline uint32_t bs_op_code(uint32_t w) { return w & 63u; }
SCPD_HD static inline uint32_t bs_op_level(uint32_t w) { return (w >> 6) & 31u; }
SCPD_HD static inline uint32_t bs_op_offset(uint32_t w) { return (w >> 11) & 0xFFFFFu; }
SCPD_HD static inline uint32_t bs_op_sync(uint32_t w) { return w >> 31; }  // all warps of the CTA meet before this op

// ops: output of build_schedule(..., log2sub = BS_LSUB, r1_mode = 1 or 2).  Returns false on an op the
// bit-sliced kernel has no code for (OP_P1: PAR = 1).
// sync_every: every sync_every-th op that lies outside all rate-1 fallback regions (those are the ops every
// warp executes, whatever its data) gets the CTA-barrier flag; 0 = none.
static inline void bs_fuse_chains(std::vector<uint32_t>* sched, int max_depth);
static inline bool bs_compile_schedule(const std::vector<uint32_t>& ops, std::vector<uint32_t>* out, int sync_every = 0,
                                       int fuse_depth = 3) {
    out->clear();
    out->reserve(ops.size());
    size_t fallback_end = 0;  // ops[i] with i < fallback_end are inside a fallback region
    int since_sync = 0;
    for (size_t i = 0; i < ops.size(); i++) {
        const uint32_t w = ops[i], l = op_level(w), o = op_offset(w);
        uint32_t sync = 0;
        if (sync_every > 0 && i >= fallback_end && op_code(w) != OP_END && ++since_sync >= sync_every) {
            sync = 1u << 31;
            since_sync = 0;
        }
        if (op_code(w) == OP_R1 && i >= fallback_end) fallback_end = i + 2 + ops[i + 1];
        uint32_t kind;
        switch (op_code(w)) {
            case OP_END: out->push_back(0u); continue;
            case OP_F: kind = BSK_F; break;
            case OP_G: kind = BSK_G; break;
            case OP_G0: kind = BSK_G0; break;
            case OP_H: kind = BSK_H; break;
            case OP_HCOPY: kind = BSK_HCOPY; break;
            case OP_R0: kind = BSK_R0; break;
            case OP_R1: kind = BSK_R1; break;
            case OP_SUB:
                if (l != BS_LSUB) return false;
                out->push_back((BSK_F + 8u * l) | (l << 6) | (o << 11) | sync);
                out->push_back(ops[++i]);  // node types, 2 bits per node in heap order (15 nodes)
                continue;
            default: return false;
        }
        if (l < BS_LSUB || (l == BS_LSUB && kind != BSK_R0)) return false;  // everything smaller sits inside an OP_SUB
        const uint32_t lc = l > BS_LLOW ? BS_LLOW + 1 : l;
        out->push_back((kind + 8u * lc) | (l << 6) | (o << 11) | sync);
        if (op_code(w) == OP_R1) out->push_back(ops[++i]);  // skip count, verbatim
    }
    if (fuse_depth >= 2) bs_fuse_chains(out, fuse_depth);
    return true;
}

// X(l) F(l-1) [F(l-2)] -> one fused op + BS_NOP words.  Only ops of the generic levels (> BS_LLOW) take part,
// and the last F of a chain must still be a generic-level op.  Data words (node types, skip counts) follow
// OP_SUB / R1 words and are stepped over.
static inline void bs_fuse_chains(std::vector<uint32_t>* sched, int max_depth) {
    std::vector<uint32_t>& s = *sched;
    const uint32_t gen = 8u * (BS_LLOW + 1);
    auto is_data_follow = [&](uint32_t w) {
        const uint32_t c = bs_op_code(w);
        return c == 8u * BS_LSUB + BSK_F || (c & 7u) == BSK_R1;
    };
    for (size_t i = 0; i + 1 < s.size(); i++) {
        const uint32_t w = s[i], c = bs_op_code(w), l = bs_op_level(w);
        if (c == 0) continue;
        if (is_data_follow(w) && c != 0) {
            i++;  // skip the data word
            continue;
        }
        if (c != gen + BSK_F && c != gen + BSK_G && c != gen + BSK_G0) continue;
        const uint32_t kind = c - gen;
        auto is_f = [&](size_t j, uint32_t lev) {
            return j < s.size() && bs_op_code(s[j]) == gen + BSK_F && bs_op_level(s[j]) == lev && lev > BS_LLOW;
        };
        if (!is_f(i + 1, l - 1)) continue;
        int depth = 2;
        if (max_depth >= 3 && is_f(i + 2, l - 2)) depth = 3;
        s[i] = (w & ~63u) | ((depth == 3 ? BS_FUSE3 : BS_FUSE2) + kind);
        for (int k = 1; k < depth; k++) s[i + k] = (s[i + k] & ~63u) | BS_NOP;
        i += depth - 1;
    }
}

struct BsPlan {
    uint32_t lsa = 0, lsb = 0;
    uint32_t sm_beta_off = 0, sm_stride = 0;  // bytes, per warp
    uint32_t ws_beta_off = 0;
    unsigned long long ws_stride = 0;  // bytes, per warp
    uint32_t aoff[24] = {0};
};

static inline int bs_planes(int q, int log2par, int ext, int l) {  // magnitude planes of alpha[l]
    return (q - 1) + ((ext && l < log2par) ? log2par - l : 0);
}
static inline size_t bs_planes_bytes(int q, int log2n) { return ((size_t)((q + 3) / 4) * 16u) << log2n; }  // per group
static inline size_t bs_alpha_bytes(int q, int log2par, int ext, int l) {
    const int quads = (bs_planes(q, log2par, ext, l) + 1 + 3) / 4;
    return ((size_t)quads * 16u) << l;
}

// smem_per_warp: shared-memory bytes one frame group may use.  force_lsa / force_lsb >= 0 pin the split (tests).
// Returns false if even the minimum (levels <= max(log2par, 5) resident) does not fit.
static inline bool bs_make_plan(int log2n, int q, int log2par, int ext, size_t smem_per_warp, BsPlan* out,
                                int force_lsa = -1, int force_lsb = -1, int lanes_per_group = 32) {
    const int lmin = std::min(log2n, BS_LLOW);
    const size_t n = (size_t)1 << log2n;
    auto a_total = [&](int lsa) {
        size_t s = 0;
        for (int l = 1; l <= lsa; l++) s += bs_alpha_bytes(q, log2par, ext, l);
        return s;
    };
    auto b_bytes = [&](int lsb) { return std::min<size_t>((size_t)8u << lsb, 4 * n); };  // 2^(lsb+1) words
    int lsa = log2n - 1, lsb = log2n;  // alpha[log2n] (the channel planes) has its own buffer
    while (a_total(lsa) + b_bytes(lsb) > smem_per_warp && (lsa > lmin || lsb > lmin)) {
        const size_t drop_a = lsa > lmin ? bs_alpha_bytes(q, log2par, ext, lsa) : 0;
        const size_t drop_b = lsb > lmin ? b_bytes(lsb) - b_bytes(lsb - 1) : 0;
        if (drop_a >= drop_b && lsa > lmin)
            lsa--;
        else
            lsb--;
    }
    if (force_lsa >= 0) lsa = std::min(log2n - 1, std::max(lmin, force_lsa));
    if (force_lsb >= 0) lsb = std::min(log2n, std::max(lmin, force_lsb));
    if (a_total(lsa) + b_bytes(lsb) > smem_per_warp) return false;
    BsPlan p;
    p.lsa = (uint32_t)lsa;
    p.lsb = (uint32_t)lsb;
    size_t off = 0;
    for (int l = 1; l <= lsa; l++) {
        p.aoff[l] = (uint32_t)off;
        off += bs_alpha_bytes(q, log2par, ext, l);
    }
    p.sm_beta_off = (uint32_t)off;
    off += b_bytes(lsb);
    off += 256;  // lanes past a small node read (never write) up to 4 * (32 + 31) bytes behind it
    p.sm_stride = (uint32_t)((off + 127) & ~(size_t)127);
    if (lanes_per_group < 32) p.sm_stride += 4u * (uint32_t)lanes_per_group;  // groups of a warp on distinct banks
    size_t woff = 0;
    for (int l = lsa + 1; l < log2n; l++) {
        p.aoff[l] = (uint32_t)woff;
        woff += bs_alpha_bytes(q, log2par, ext, l);
    }
    p.ws_beta_off = (uint32_t)woff;
    woff += 4 * n;
    p.ws_stride = (woff + 255) & ~(size_t)255;
    *out = p;
    return true;
}

}  // namespace scpd
