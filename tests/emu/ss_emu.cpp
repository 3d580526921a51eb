// ss_emu.cpp -- TEST INFRASTRUCTURE.  Runs the product's slot-sliced decode kernel source (csrc/decode_ss.cuh,
// compiled as plain C++ through tests/emu/fake_cuda/cuda_runtime.h) on the CPU.  Every lane of that kernel owns
// one frame and shares nothing with its neighbours but the schedule, so the emulation is a loop over lanes; the
// warp vote of the all-information shortcut degenerates to the lane's own flag (both branches give the same bits,
// which is what the tests check).
// fp16x2 arithmetic is emulated with exact float pairs (decode_ss.cuh, host branch).  Never a decode path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cuda_runtime.h"
// keep this include after the fake runtime
#include "../../sc_polar_decoder_hls_b200/csrc/decode_ss.cuh"

namespace cuda_emu {
thread_local LaneCtx* cur = nullptr;
uint32_t collective_exchange(uint32_t v, int) { return v; }
uint32_t collective_ballot(bool p) { return p ? 1u : 0u; }
void collective_sync() {}
int cta_barrier_or(int p) { return p; }
}  // namespace cuda_emu

using namespace scpd;

template <int Q, int LOG2PAR, bool EXT>
static int run(int log2n, int pruning, const uint8_t* flags, const int8_t* llr, size_t nframes, uint32_t* xhat,
               size_t smem_per_warp, int force_lsa, int force_lwin, int fuse, uint64_t* stats, int ltm, int max_pre, int xf) {
    constexpr int P = Q - 1;
    const uint32_t n = 1u << log2n;
    SsStats st;
    std::vector<uint32_t> sched = ss_build_schedule(log2n, pruning, flags, &st, fuse);
    if (stats) {
        stats[0] = st.n_ops;
        stats[1] = st.n_f;
        stats[2] = st.n_g;
        stats[3] = st.n_sub32_mixed;
    }
    SsPlan plan;
    if (!ss_make_plan(log2n, smem_per_warp, &plan, force_lsa, force_lwin, ltm ? 512u : 0u, ltm, (xf & 2) != 0)) return 1;
    if (ltm > 0 && plan.ltm != (uint32_t)ltm) return 3;
    const size_t ntasks = (nframes + 31) / 32;
    // the leading f ops are computed with the planes (scpd_api.cu: plan_ss / decode_ss)
    const int pre = ss_prefuse_depth(sched, log2n, plan.lsa, max_pre);
    // xf: fused SS_XF_* ops wherever the three levels involved live in global memory (scpd_api.cu: large trees only)
    sched = ss_build_schedule(log2n, pruning, flags, &st, fuse, (xf & 1) ? (int)std::max(plan.lsa, plan.ltm) + 1 : 0, pre);
    if (stats) {
        stats[4] = (uint64_t)pre;
        stats[5] = st.n_xf;
    }
    SsPre pre_off;
    const size_t pl_stride = ss_planes_quads(log2n, pre, pre_off.off);
    std::vector<uint4> planes(ntasks * pl_stride, uint4{0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu});
    // ss_planes_kernel, lane by lane
    for (size_t f = 0; f < nframes; f++) {
        uint4* dst = planes.data() + (f / 32) * pl_stride + (f % 32);
        if (n >= 256) {
            for (uint32_t u = 0; u < n / 256; u++) {
                const int8_t* row = llr + f * n;
                switch (pre) {
                    case 1: ss::planes_unit<P, 1>(row, true, n / 32, u, dst, pre_off); break;
                    case 2: ss::planes_unit<P, 2>(row, true, n / 32, u, dst, pre_off); break;
                    case 3: ss::planes_unit<P, 3>(row, true, n / 32, u, dst, pre_off); break;
                    default: ss::planes_unit<P, 0>(row, true, n / 32, u, dst, pre_off); break;
                }
            }
            continue;
        }
        for (uint32_t c = 0; c < n / 32; c++) {
            uint32_t v[8];
            std::memcpy(v, llr + f * n + 32u * c, 32);
            bs::Val<P> x;
            ss::chunk_planes<P>(v, x);
            uint32_t o[8];
            o[0] = x.s;
            for (int k = 1; k < 8; k++) o[k] = k <= P ? x.m[k - 1] : 0u;
            dst[(2u * c) * 32u] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[(2u * c + 1u) * 32u] = make_uint4(o[4], o[5], o[6], o[7]);
        }
    }
    SsParams p;
    std::memset(&p, 0, sizeof p);
    p.sched = sched.data();
    p.sched_words = 0;
    p.planes = planes.data();
    p.planes_stride = pl_stride;
    p.xhat = xhat;
    p.nframes = nframes;
    p.ntasks = ntasks;
    p.n = n;
    p.log2n = (uint32_t)log2n;
    p.wpf = n / 32;
    p.lsa = plan.lsa;
    p.lwin = plan.lwin;
    p.ltm = plan.ltm;
    p.tm_cols = plan.tm_cols;
    p.win_words = plan.win_words;
    p.sm_stride = plan.sm_stride;
    p.sm_beta_off = plan.sm_beta_off;
    p.ws_stride = plan.ws_stride;
    p.ws_beta_off = plan.ws_beta_off;
    for (int l = 0; l < 24; l++) p.aoff[l] = plan.aoff[l];
    p.lpre = (uint32_t)(log2n - pre);
    for (int s = 1; s <= pre; s++) p.poff[log2n - s] = pre_off.off[s];
    std::vector<uint4> smem(plan.sm_stride, uint4{0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu});
    std::vector<uint4> ws(plan.ws_stride ? plan.ws_stride : 1, uint4{0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu});
    p.ws = ws.data();
    std::vector<uint4> hotbuf(plan.hot_stride ? plan.hot_stride : 1, uint4{0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu});
    p.hot = hotbuf.data();
    p.lhot = plan.lhot;
    p.hot_stride = plan.hot_stride;
    std::vector<uint32_t> tmem((size_t)32 * 512, 0xDEADBEEFu);  // 512 columns per lane
    for (size_t task = 0; task < ntasks; task++) {
        for (int lane = 0; lane < 32; lane++) {
            SsThread<Q, LOG2PAR, EXT, false, true> t(p);
            t.bind(smem.data() + lane, ws.data() + lane, hotbuf.data() + lane);
            t.sched = sched.data();
            t.tm = tmem.data() + (size_t)lane * 512;
            t.pl = planes.data() + task * pl_stride + lane;
            t.run();
            t.write_output(task * 32ull + lane);
        }
    }
    return 0;
}

extern "C" int ss_emu_decode(int log2n, int q, int log2par, int ext, int pruning, const uint8_t* flags, const int8_t* llr,
                             size_t nframes, uint32_t* xhat, size_t smem_per_warp, int force_lsa, int force_lwin, int fuse,
                             uint64_t* stats, int ltm, int max_pre, int xf) {
#define SS_CASE(Q, LP, E)                          \
    if (q == Q && log2par == LP && ext == (E ? 1 : 0)) \
        return run<Q, LP, E>(log2n, pruning, flags, llr, nframes, xhat, smem_per_warp, force_lsa, force_lwin, fuse, stats, ltm, max_pre, xf);
    SS_CASE(8, 4, true)
    SS_CASE(8, 4, false)
    SS_CASE(6, 4, true)
    SS_CASE(7, 4, true)
    SS_CASE(8, 2, true)
    SS_CASE(8, 3, true)
    SS_CASE(6, 2, false)
#undef SS_CASE
    return 2;
}
