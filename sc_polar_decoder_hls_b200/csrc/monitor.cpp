// monitor.cpp -- host-side counterpart of the reference's function x level monitor.
//
// The reference's RTL testbench histograms FSM states by function and tree level
// (src/rtl_simu_testbench/sc_monitor/sc_monitor.h:50-441, fed by the Fct_ID / N_value ports of
// my_module.h:21-30).  The walk is static for a given frozen set, so the same matrix follows from the
// table alone: node visits per frame and the trip counts of the pipelined loops f_loop / g_loop /
// h_loop (my_module.h:373,704,903: NB_ITER = N_REG >> 1 PAR-wide words per node), one per leaf decode
// (R_STATE :592-595).  With SCPD_PRUNE_NONE this is the reference at PRUNING_LEVEL 0; with pruning it
// is the walk the decode kernels of this library execute (schedule.h).
#include <cstring>
#include <vector>

#include "../../include/scpd.h"

namespace {
struct Walker {
    uint32_t log2par, pruning;
    const uint32_t* psum;
    scpd_stage_matrix* m;
    const uint8_t* flags;
    uint32_t count(uint32_t o, uint32_t n) const { return psum[o + n] - psum[o]; }
    void visit(int fn, int l, uint64_t iters) {
        m->visits[fn][l] += 1;
        m->iterations[fn][l] += iters;
    }
    // the reference FSM at PRUNING_LEVEL 2 (schedule.h emit_l2; my_module.h:337-545 child types)
    enum { T_R0, T_R1, T_REP, T_SPC, T_RN };
    int word_type(uint32_t o) const {
        const uint32_t p = 1u << log2par, c = count(o, p);
        return c == 0 ? T_R0 : c == p ? T_R1 : (c == 1 && flags[o + p - 1]) ? T_REP : (c == p - 1 && !flags[o]) ? T_SPC : T_RN;
    }
    int child_type(uint32_t o, uint32_t n) const {
        const uint32_t p = 1u << log2par, c = count(o, n);
        if (c == 0) return T_R0;
        if (c == n) return T_R1;
        if (count(o, n - p) == 0 && word_type(o + n - p) == T_REP) return T_REP;
        if (count(o + p, n - p) == n - p && word_type(o) == T_SPC) return T_SPC;
        return T_RN;
    }
    void walk_l2(int l, uint32_t o, int top) {
        const uint32_t n = 1u << l;
        if ((uint32_t)l <= log2par) return visit(SCPD_STAGE_R, l, 1);
        const uint64_t words = n >> log2par;
        const uint32_t h = n >> 1;
        int tl = child_type(o, h), tr = child_type(o + h, h);
        if (l == top) tl = tr = T_RN;  // INIT enters F_STATE / G_STATE whatever the children are
        if (tl == T_R0) visit(SCPD_STAGE_R0, l - 1, 0);
        else if (tl == T_REP) visit(SCPD_STAGE_REP, l, words >> 1);
        else {
            visit(SCPD_STAGE_F, l, words >> 1);
            walk_l2(l - 1, o, top);
        }
        if (tr == T_R1) visit(SCPD_STAGE_R1, l, words >> 1);
        else if (tr == T_SPC) visit(SCPD_STAGE_SPC, l, words >> 1);
        else {  // ELAG_RARE = 0: an all-frozen right child is walked like any other
            visit(SCPD_STAGE_G, l, words >> 1);
            walk_l2(l - 1, o + h, top);
        }
        visit(SCPD_STAGE_H, l, words >> 1);
    }
    void walk(int l, uint32_t o) {
        const uint32_t n = 1u << l, c = count(o, n);
        const uint64_t words = (n >> log2par) ? (n >> log2par) : 1;  // N_REG of this node
        if (pruning >= SCPD_PRUNE_R0 && c == 0) return visit(SCPD_STAGE_R0, l, 0);
        if ((uint32_t)l <= log2par) return visit(SCPD_STAGE_R, l, 1);
        if (pruning >= SCPD_PRUNE_R0_R1 && c == n) return visit(SCPD_STAGE_R1, l, words);
        const uint32_t h = n >> 1;
        const bool left_r0 = pruning >= SCPD_PRUNE_R0 && count(o, h) == 0;
        if (left_r0)
            visit(SCPD_STAGE_R0, l - 1, 0);
        else {
            visit(SCPD_STAGE_F, l, words >> 1);
            walk(l - 1, o);
        }
        if (!left_r0 && pruning >= SCPD_PRUNE_R0 && count(o + h, h) == 0)
            visit(SCPD_STAGE_R0, l - 1, 0);
        else {
            visit(SCPD_STAGE_G, l, words >> 1);
            walk(l - 1, o + h);
        }
        visit(SCPD_STAGE_H, l, words >> 1);
    }
};
}  // namespace

extern "C" int scpd_stage_profile(const scpd_config* cfg, const uint8_t* flags, scpd_stage_matrix* out) {
    if (!cfg || !flags || !out) return SCPD_E_ARG;
    const uint32_t n = cfg->n, par = cfg->par;
    if (n < 2 || (n & (n - 1)) || n > (1u << 20) || par < 1 || (par & (par - 1)) || 2 * par > n ||
        cfg->pruning > SCPD_PRUNE_REF_LEVEL2 || (cfg->pruning == SCPD_PRUNE_REF_LEVEL2 && (par < 2 || par > 256)))
        return SCPD_E_CONFIG;
    std::memset(out, 0, sizeof *out);
    std::vector<uint32_t> psum(n + 1, 0);
    for (uint32_t i = 0; i < n; i++) psum[i + 1] = psum[i] + (flags[i] ? 1u : 0u);
    int log2n = 0, log2par = 0;
    while ((1u << log2n) < n) log2n++;
    while ((1u << log2par) < par) log2par++;
    Walker w{(uint32_t)log2par, cfg->pruning, psum.data(), out, flags};
    if (cfg->pruning == SCPD_PRUNE_REF_LEVEL2) w.walk_l2(log2n, 0, log2n);
    else w.walk(log2n, 0);
    for (int f = 0; f < SCPD_STAGE_FUNCS; f++)
        for (int l = 0; l < 32; l++) out->total_iterations += out->iterations[f][l];
    return SCPD_OK;
}
