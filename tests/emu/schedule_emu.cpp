// schedule_emu.cpp -- TEST INFRASTRUCTURE.  A sequential host interpreter of the product's
// schedule (sc_polar_decoder_hls_b200/csrc/schedule.h) using the same int16 formulas and the
// same node/offset bookkeeping as the CUDA kernels (decode_generic.cuh).  It lets the CPU-only
// test tier check the schedule compiler, the pruning rules and the rate-1 fallback walk against
// the oracle without a GPU.  It is NOT linked into libscpd.so and is never a decode fallback.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../sc_polar_decoder_hls_b200/csrc/schedule.h"

using namespace scpd;

namespace {
struct Emu {
    int log2n, log2par, extended, satv;
    std::vector<std::vector<int>> alpha;  // alpha[l] has 2^l entries
    std::vector<uint8_t> beta;
    const int8_t* llr;
    uint64_t fallbacks = 0;
    bool force_fallback = false;

    int ld(int l, uint32_t i) const { return l == log2n ? (int)llr[i] : alpha[l][i]; }
    static int f(int a, int b) { return std::max(a + b, 0) - std::max(a, b); }
    int g(int a, int b, int s, bool nosat) const {
        int r = s ? b - a : b + a;
        if (!nosat) r = std::max(-satv, std::min(satv, r));
        return r;
    }
    bool nosat(int l) const { return extended && l <= log2par; }
    void op_f(int l) {
        uint32_t h = 1u << (l - 1);
        for (uint32_t i = 0; i < h; i++) alpha[l - 1][i] = f(ld(l, i), ld(l, i + h));
    }
    void op_g(int l, uint32_t o, bool ns, bool zero) {
        uint32_t h = 1u << (l - 1);
        for (uint32_t i = 0; i < h; i++) alpha[l - 1][i] = g(ld(l, i), ld(l, i + h), zero ? 0 : beta[o + i], ns);
    }
    void op_h(int l, uint32_t o, bool copy) {
        uint32_t h = 1u << (l - 1);
        for (uint32_t i = 0; i < h; i++) beta[o + i] = copy ? beta[o + h + i] : (beta[o + i] ^ beta[o + h + i]);
    }
    void op_p2(uint32_t o, uint32_t lf) {
        int a = ld(1, 0), b = ld(1, 1);
        int u0 = ((a < 0) ^ (b < 0)) & (int)(lf & 1);
        int s = u0 ? b - a : b + a;
        int u1 = (s < 0) & (int)((lf >> 1) & 1);
        beta[o] = (uint8_t)(u0 ^ u1);
        beta[o + 1] = (uint8_t)u1;
    }
    void op_p1(uint32_t o, uint32_t lf) { beta[o] = (uint8_t)((ld(0, 0) < 0) & (int)(lf & 1)); }
    bool op_hd(int l, uint32_t o) {
        uint32_t n = 1u << l;
        bool zero = false;
        for (uint32_t i = 0; i < n; i++) {
            int v = ld(l, i);
            zero |= (v == 0);
            beta[o + i] = v < 0;
        }
        return zero;
    }
    void generic_sc(int l, uint32_t o) {
        const int tl = log2par >= 1 ? 1 : 0;
        const uint32_t nt = 1u << (l - tl);
        for (uint32_t t = 0; t < nt; t++) {
            const uint32_t to = o + (t << tl);
            if (t == 0) {
                for (int lv = l; lv > tl; lv--) op_f(lv);
            } else {
                const int lv0 = __builtin_ctz(t) + tl + 1;
                op_g(lv0, to & ~((1u << lv0) - 1u), nosat(lv0), false);
                for (int lv = lv0 - 1; lv > tl; lv--) op_f(lv);
            }
            if (tl)
                op_p2(to, 3u);
            else
                op_p1(to, 1u);
            const int ones = __builtin_ctz(~t);
            for (int j = 1; j <= ones && tl + j <= l; j++) {
                const int lv = tl + j;
                op_h(lv, to + (1u << tl) - (1u << lv), false);
            }
        }
    }
    void run(const std::vector<uint32_t>& ops) {
        for (size_t pc = 0;; pc++) {
            uint32_t w = ops[pc];
            uint32_t opc = op_code(w), o = op_offset(w);
            int l = (int)op_level(w);
            if (opc == OP_END) break;
            switch (opc) {
                case OP_F: op_f(l); break;
                case OP_G: op_g(l, o, op_nosat(w), false); break;
                case OP_G0: op_g(l, o, op_nosat(w), true); break;
                case OP_H: op_h(l, o, false); break;
                case OP_HCOPY: op_h(l, o, true); break;
                case OP_R0: std::fill(beta.begin() + o, beta.begin() + o + (1u << l), 0); break;
                case OP_R1:
                    if (op_hd(l, o) || force_fallback) {
                        fallbacks++;
                        generic_sc(l, o);
                    }
                    break;
                case OP_P2: op_p2(o, op_lf(w)); break;
                case OP_P1: op_p1(o, op_lf(w)); break;
            }
        }
    }
};
}  // namespace

extern "C" {
// returns number of rate-1 fallbacks taken (>= 0) or -1 on bad arguments
long long emu_decode(int log2n, int log2par, int llr_bits, int extended, int pruning, const uint8_t* flags,
                     const int8_t* llr, size_t nframes, uint8_t* xhat, int force_fallback,
                     uint64_t* n_ops, uint64_t* n_fg) {
    if (log2n < 1 || log2par < 0 || log2par >= log2n) return -1;
    ScheduleStats st;
    std::vector<uint32_t> ops = build_schedule(log2n, log2par, extended, pruning, flags, &st);
    if (n_ops) *n_ops = st.n_ops;
    if (n_fg) *n_fg = st.n_f + st.n_g;
    Emu e;
    e.log2n = log2n;
    e.log2par = log2par;
    e.extended = extended;
    e.satv = (1 << (llr_bits - 1)) - 1;
    e.force_fallback = force_fallback != 0;
    e.alpha.resize(log2n + 1);
    for (int l = 0; l <= log2n; l++) e.alpha[l].assign(1u << l, 0);
    const size_t n = (size_t)1 << log2n;
    e.beta.assign(n, 0);
    for (size_t f = 0; f < nframes; f++) {
        e.llr = llr + f * n;
        std::fill(e.beta.begin(), e.beta.end(), 0xAA);  // poison: every bit must be written
        e.run(ops);
        std::memcpy(xhat + f * n, e.beta.data(), n);
    }
    return (long long)e.fallbacks;
}
}
