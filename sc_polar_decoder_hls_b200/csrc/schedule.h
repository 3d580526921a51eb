// schedule.h -- the static SC tree walk, compiled once per frozen set on the host.
//
// The reference walks the tree with an FSM and two packed stacks (my_module.h:228-263, states
// INIT/F/R/G/H; functions.h:11-37).  All frames of a batch share the frozen set, so here the walk
// is a straight-line program ("schedule") of node operations that every warp executes in lock
// step; do_prunning's node classification (my_module.h:96-154) happens here too, at build time.
//
// One op = one uint32:  [3:0] opcode  [8:4] level l (node size n = 2^l)  [9] nosat
//                       [11:10] leaf flags  [31:12] offset o of the node in the frame.
#pragma once
#include <cstdint>
#include <vector>

#ifdef __CUDACC__
#define SCPD_HD __host__ __device__
#else
#define SCPD_HD
#endif

namespace scpd {

enum Op : uint32_t {
    OP_END = 0,
    OP_F = 1,      // alpha[l-1][i] = f(alpha[l][i], alpha[l][i+n/2])            F_STATE  :373-445
    OP_G = 2,      // alpha[l-1][i] = g(alpha[l][i], alpha[l][i+n/2], beta[o+i]) G_STATE  :704-781
    OP_G0 = 3,     // same with beta = 0 (left child all-frozen)
    OP_H = 4,      // beta[o+i] ^= beta[o+n/2+i], i < n/2                        H_STATE  :903-932
    OP_HCOPY = 5,  // beta[o+i]  = beta[o+n/2+i]   (left child all-frozen: 0 ^ x)
    OP_R0 = 6,     // beta[o..o+n) = 0            (all-frozen node)
    OP_R1 = 7,     // all-information node: hard decision, plain-SC fallback if an LLR is 0.  In fast-kernel
                   // schedules the next word is the length of the fallback ops that follow (skipped
                   // when no LLR of the node is 0)
    OP_P2 = 8,     // two-bit terminal, Spec_P2 functions.h:367-384; leaf flags = (f0, f1)
    OP_P1 = 9,     // one-bit terminal, Spec_P1 functions.h:355-364; leaf flag bit0 = f
    OP_SUB = 10,   // register subtree of size 2^l (fast kernel); followed by its node-type words
    // reference-pruning mode (SCPD_PRUNE_REF_LEVEL2; the reference built with PRUNING_LEVEL 2), node of size 2^l at o:
    OP_REP = 11,   // left child is a repetition node: its bits = sign of the saturating sum of f   F_REP_STATE :1292-1390
    OP_GR1 = 12,   // right child is all-information: its bits = sign of g (bit 9: beta = 0)         G_R1_STATE  :1571-1642
    OP_GSPC = 13,  // right child is a single parity check: signs of g, least reliable one flipped
                   // when the parity is odd (bit 9: beta = 0)                                       G_SPC_STATE :1737-1842
};

SCPD_HD static inline uint32_t op_make(uint32_t opc, uint32_t level, uint32_t offset, uint32_t nosat = 0,
                               uint32_t lf = 0) {
    return opc | (level << 4) | (nosat << 9) | (lf << 10) | (offset << 12);
}
SCPD_HD static inline uint32_t op_code(uint32_t w) { return w & 15u; }
SCPD_HD static inline uint32_t op_level(uint32_t w) { return (w >> 4) & 31u; }
SCPD_HD static inline uint32_t op_nosat(uint32_t w) { return (w >> 9) & 1u; }
SCPD_HD static inline uint32_t op_lf(uint32_t w) { return (w >> 10) & 3u; }
SCPD_HD static inline uint32_t op_offset(uint32_t w) { return w >> 12; }

struct ScheduleStats {
    uint64_t n_ops = 0;
    uint64_t n_f = 0;   // f element updates per frame
    uint64_t n_g = 0;   // g element updates per frame
    uint64_t n_r0 = 0;  // pruned all-frozen nodes
    uint64_t n_r1 = 0;  // pruned all-information nodes
    uint64_t n_rep = 0; // repetition nodes decided by their sum (reference-pruning mode)
    uint64_t n_spc = 0; // single-parity-check nodes decided by signs + one flip (reference-pruning mode)
};

struct ScheduleBuilder {
    int log2n, log2par, extended, pruning;
    int log2sub = -1;  // >= 0: stop at nodes of this size and emit OP_SUB + descriptor (fast kernel)
    // bit-sliced kernel: 1 = every OP_R1 is followed by a skip count and by the plain-SC ops of the node
    // whose children are OP_R1 again (a zero LLR only unfolds the path it sits on); 2 = skip count 0 and
    // no fallback ops (SIGMAG, where the hard decision is always identical to plain SC)
    int r1_mode = 0;
    bool bs_sub = false;  // OP_SUB carries only the node-type words (bit-sliced kernel)
    const uint8_t* flags;
    std::vector<uint32_t> psum;  // prefix sums of flags
    std::vector<uint32_t> ops;
    ScheduleStats st;

    uint32_t count(uint32_t o, uint32_t n) const { return psum[o + n] - psum[o]; }
    // g of a node of size n is the un-saturated Function_G_ext iff the node lies inside the
    // PAR-wide leaf decoder and EXTENDED == 1 (Spec_P*_ext, functions.h:413-438 ...).
    uint32_t nosat(int l) const { return (extended && l <= log2par) ? 1u : 0u; }

    // Node types of a register subtree, 2 bits per node in heap order (root 0, children 2i+1, 2i+2),
    // sizes 2^l .. 2.  0 = mixed, 1 = all-frozen, 2 = all-information; for the size-2 nodes the code
    // is the flag pair itself: 0 = (0,1), 1 = (0,0), 2 = (1,1), 3 = (1,0).
    void emit_subtree(int l, uint32_t o) {
        const uint32_t nodes = (1u << l) - 1u;
        std::vector<uint32_t> words((2 * nodes + 31) / 32, 0u);
        uint32_t heap = 0;
        for (int d = 0; d < l; d++) {
            const uint32_t sz = 1u << (l - d);
            for (uint32_t k = 0; k < (1u << d); k++, heap++) {
                const uint32_t c = count(o + k * sz, sz);
                uint32_t t;
                if (sz == 2) {
                    const uint32_t f0 = flags[o + k * 2] & 1u, f1 = flags[o + k * 2 + 1] & 1u;
                    t = (f0 == 0 && f1 == 1) ? 0u : (f0 == 0 && f1 == 0) ? 1u : (f0 == 1 && f1 == 1) ? 2u : 3u;
                } else {
                    t = (pruning >= 1 && c == 0) ? 1u : (pruning >= 2 && c == sz) ? 2u : 0u;
                    if (c != 0 && c != sz) {
                        // work estimate only (the kernel skips f when the left child is all-frozen)
                        st.n_f += (count(o + k * sz, sz / 2) != 0 || pruning == 0) ? sz / 2 : 0;
                        st.n_g += sz / 2;
                    }
                }
                words[(2 * heap) >> 5] |= t << ((2 * heap) & 31);
            }
        }
        // bit 9: the pattern-specialised small-node routines (which prune internally) may be used
        ops.push_back(op_make(OP_SUB, l, o, pruning >= 2 ? 1u : 0u));
        for (uint32_t w : words) ops.push_back(w);
        if (bs_sub) return;
        push_flag_words(l, o, false);
        push_sub8_words(l, o, false);
    }
    // one word per size-8 node of the subtree: the types of its 7 nodes (sizes 8, 4, 4, 2, 2, 2, 2)
    // in local heap order, 2 bits each, same coding as above
    void push_sub8_words(int l, uint32_t o, bool all_ones) {
        for (uint32_t k = 0; k < (1u << l) / 8; k++) {
            uint32_t w = 0, heap = 0;
            for (int d = 0; d < 3; d++) {
                const uint32_t sz = 8u >> d;
                for (uint32_t j = 0; j < (1u << d); j++, heap++) {
                    const uint32_t off = o + 8 * k + j * sz;
                    uint32_t t;
                    if (all_ones) {
                        t = 2u;
                    } else if (sz == 2) {
                        const uint32_t f0 = flags[off] & 1u, f1 = flags[off + 1] & 1u;
                        t = (f0 == 0 && f1 == 1) ? 0u : (f0 == 0 && f1 == 0) ? 1u : (f0 == 1 && f1 == 1) ? 2u : 3u;
                    } else {
                        const uint32_t c = count(off, sz);
                        t = (pruning >= 1 && c == 0) ? 1u : (pruning >= 2 && c == sz) ? 2u : 0u;
                    }
                    w |= t << (2 * heap);
                }
            }
            ops.push_back(w);
        }
    }
    // raw information flags of the subtree (bit i = element o + i), ceil(2^l / 32) words: lets the
    // kernel pick a pattern-specialised routine for the small nodes
    void push_flag_words(int l, uint32_t o, bool all_ones) {
        const uint32_t sz = 1u << l;
        for (uint32_t w0 = 0; w0 < sz; w0 += 32) {
            uint32_t v = 0;
            for (uint32_t b = 0; b < 32 && w0 + b < sz; b++) v |= (all_ones ? 1u : (flags[o + w0 + b] & 1u)) << b;
            ops.push_back(v);
        }
    }

    // un-pruned walk of an all-information node down to the register subtrees (rate-1 fallback)
    void emit_plain(int l, uint32_t o) {
        if (l == log2sub) {
            const uint32_t nodes = (1u << l) - 1u;
            ops.push_back(op_make(OP_SUB, l, o, 1u));
            for (uint32_t w = 0; w < (2 * nodes + 31) / 32; w++) ops.push_back(0xAAAAAAAAu);  // every node type 2
            push_flag_words(l, o, true);
            push_sub8_words(l, o, true);
            return;
        }
        ops.push_back(op_make(OP_F, l, o));
        emit_plain(l - 1, o);
        ops.push_back(op_make(OP_G, l, o, nosat(l)));
        emit_plain(l - 1, o + (1u << (l - 1)));
        ops.push_back(op_make(OP_H, l, o));
    }

    // all-information node of the bit-sliced schedules (r1_mode 1 / 2)
    void emit_r1_tree(int l, uint32_t o) {
        if (l == log2sub) {  // the fused subtree routine does its own hard decision / zero vote
            emit_subtree(l, o);
            return;
        }
        if (l == 1) {
            ops.push_back(op_make(OP_P2, 1, o, 0, 3u));
            return;
        }
        ops.push_back(op_make(OP_R1, l, o));
        st.n_r1++;
        const size_t at = ops.size();
        ops.push_back(0u);
        if (r1_mode == 2) return;
        ops.push_back(op_make(OP_F, l, o));
        emit_r1_tree(l - 1, o);
        ops.push_back(op_make(OP_G, l, o, nosat(l)));
        emit_r1_tree(l - 1, o + (1u << (l - 1)));
        ops.push_back(op_make(OP_H, l, o));
        ops[at] = (uint32_t)(ops.size() - at - 1);
    }

    // ---- reference-pruning mode: the reference's own PRUNING_LEVEL 2 (config.h:16-30 as checked in: R1, REP, SPC, H0).
    // NOT plain SC (tests/test_oracle.py).  do_prunning (my_module.h:61-166) types every PAR-wide word of the frozen
    // table; F_STATE / G_STATE (:337-545) accumulate the words of a child into its type:
    enum NodeType { T_R0, T_R1, T_REP, T_SPC, T_RN };
    NodeType word_type(uint32_t o) const {
        const uint32_t p = 1u << log2par, c = count(o, p);
        if (c == 0) return T_R0;
        if (c == p) return T_R1;
        if (c == 1 && flags[o + p - 1]) return T_REP;
        if (c == p - 1 && !flags[o]) return T_SPC;
        return T_RN;
    }
    NodeType child_type(uint32_t o, uint32_t n) const {  // n >= PAR
        const uint32_t p = 1u << log2par, c = count(o, n);
        if (c == 0) return T_R0;
        if (c == n) return T_R1;
        if (count(o, n - p) == 0 && word_type(o + n - p) == T_REP) return T_REP;        // R0 words, then one REP word
        if (count(o + p, n - p) == n - p && word_type(o) == T_SPC) return T_SPC;        // one SPC word, then R1 words
        return T_RN;
    }
    void emit_l2(int l, uint32_t o) {
        if (l <= log2par) {  // the PAR-wide leaf stays Spec_Polar_Decoder (library.h:170-198); "& 0" pruning is value-neutral
            const int keep = pruning;
            pruning = 1;
            emit(l, o);
            pruning = keep;
            return;
        }
        const uint32_t n = 1u << l, h = n >> 1;
        if (count(o, n) == 0) {
            ops.push_back(op_make(OP_R0, l, o));
            st.n_r0++;
            return;
        }
        // INIT (:285-336) enters F_STATE / G_STATE for the root's children whatever their type
        const bool root = l == log2n;
        NodeType tl = child_type(o, h), tr = child_type(o + h, h);
        if (root && tl != T_R0) tl = T_RN;
        if (root && tr != T_R0) tr = T_RN;
        const uint32_t zb = tl == T_R0 ? 1u : 0u;  // H0: the left child is skipped, g runs on beta = 0
        if (tl == T_REP) {
            ops.push_back(op_make(OP_REP, l, o));
            st.n_rep++;
            st.n_f += h;
        } else if (tl != T_R0) {
            ops.push_back(op_make(OP_F, l, o));
            st.n_f += h;
            emit_l2(l - 1, o);
        }
        if (tr == T_R0) {
            ops.push_back(op_make(OP_R0, l - 1, o + h));
            st.n_r0++;
        } else if (tr == T_R1) {
            ops.push_back(op_make(OP_GR1, l, o, zb));
            st.n_r1++;
            st.n_g += h;
        } else if (tr == T_SPC) {
            ops.push_back(op_make(OP_GSPC, l, o, zb));
            st.n_spc++;
            st.n_g += h;
        } else {
            ops.push_back(op_make(zb ? OP_G0 : OP_G, l, o, nosat(l)));
            st.n_g += h;
            emit_l2(l - 1, o + h);
        }
        ops.push_back(op_make(zb ? OP_HCOPY : OP_H, l, o));
    }

    void emit(int l, uint32_t o) {
        const uint32_t n = 1u << l;
        const uint32_t c = count(o, n);
        if (pruning >= 1 && c == 0) {
            ops.push_back(op_make(OP_R0, l, o));
            st.n_r0++;
            return;
        }
        if (l == 0) {  // only reachable when PAR == 1
            ops.push_back(op_make(OP_P1, 0, o, 0, flags[o] & 1u));
            return;
        }
        if (l == 1 && log2par >= 1) {
            ops.push_back(op_make(OP_P2, 1, o, 0, (flags[o] & 1u) | ((flags[o + 1] & 1u) << 1)));
            return;
        }
        if (pruning >= 2 && c == n && r1_mode != 0) {
            emit_r1_tree(l, o);
            return;
        }
        if (pruning >= 2 && c == n) {
            ops.push_back(op_make(OP_R1, l, o));
            st.n_r1++;
            if (log2sub >= 0) {  // explicit plain-SC fallback of this node, skipped unless an LLR is 0
                const size_t at = ops.size();
                ops.push_back(0u);
                emit_plain(l, o);
                ops[at] = (uint32_t)(ops.size() - at - 1);
            }
            return;
        }
        if (l == log2sub) {
            emit_subtree(l, o);
            return;
        }
        const uint32_t h = n >> 1;
        const bool left_r0 = pruning >= 1 && count(o, h) == 0;
        const bool right_r0 = pruning >= 1 && count(o + h, h) == 0;
        if (left_r0) {
            ops.push_back(op_make(OP_G0, l, o, nosat(l)));
            st.n_g += h;
            emit(l - 1, o + h);
            ops.push_back(op_make(OP_HCOPY, l, o));
            return;
        }
        ops.push_back(op_make(OP_F, l, o));
        st.n_f += h;
        emit(l - 1, o);
        if (right_r0) {
            ops.push_back(op_make(OP_R0, l - 1, o + h));
            ops.push_back(op_make(OP_H, l, o));  // left ^ 0; also moves the node between storage spaces
            st.n_r0++;
            return;
        }
        ops.push_back(op_make(OP_G, l, o, nosat(l)));
        st.n_g += h;
        emit(l - 1, o + h);
        ops.push_back(op_make(OP_H, l, o));
    }
};

// flags: n bytes, 1 = information bit.  Returns the op list terminated by OP_END.
static inline std::vector<uint32_t> build_schedule(int log2n, int log2par, int extended, int pruning,
                                                   const uint8_t* flags, ScheduleStats* stats,
                                                   int log2sub = -1, int r1_mode = 0) {
    ScheduleBuilder b;
    b.log2sub = log2sub;
    b.r1_mode = r1_mode;
    b.bs_sub = r1_mode != 0;
    b.log2n = log2n;
    b.log2par = log2par;
    b.extended = extended;
    b.pruning = pruning;
    b.flags = flags;
    const uint32_t n = 1u << log2n;
    b.psum.assign(n + 1, 0);
    for (uint32_t i = 0; i < n; i++) b.psum[i + 1] = b.psum[i] + (flags[i] ? 1u : 0u);
    if (pruning == 3) b.emit_l2(log2n, 0);  // SCPD_PRUNE_REF_LEVEL2 (raw-pattern kernel only)
    else b.emit(log2n, 0);
    b.ops.push_back(op_make(OP_END, 0, 0));
    b.ops.push_back(op_make(OP_END, 0, 0));  // kernels prefetch one word ahead
    b.st.n_ops = b.ops.size();
    if (stats) *stats = b.st;
    return b.ops;
}

}  // namespace scpd
