// warp_emu.cpp -- TEST INFRASTRUCTURE.  Runs the product's fast decode kernel source
// (csrc/decode_fast.cuh, compiled as C++ through tests/emu/fake_cuda/cuda_runtime.h) on the CPU:
// one ucontext fiber per lane, warp collectives resolved when all 32 lanes have arrived.
// Used to debug kernel logic and to cover it in the CPU-only test tier.  Never a decode path.
#include <ucontext.h>

#include <cstdio>
#include <type_traits>
#include <cstdlib>
#include <vector>

#include "cuda_runtime.h"
// keep this include after the fake runtime
#include "../../sc_polar_decoder_hls_b200/csrc/decode_fast.cuh"
#include "../../sc_polar_decoder_hls_b200/csrc/decode_bs.cuh"
#include "../../sc_polar_decoder_hls_b200/csrc/decode_raw.cuh"
#include "../../sc_polar_decoder_hls_b200/csrc/bs_plan.h"
namespace scpd {
uint32_t s_frame[32768 + 2];  // the dynamic shared memory of the counter kernels (count.cuh)
}
#include "../../sc_polar_decoder_hls_b200/csrc/count.cuh"

namespace scpd {
__attribute__((aligned(16))) uint8_t smem_fast[232448];
}

namespace cuda_emu {
thread_local LaneCtx* cur = nullptr;
namespace {
constexpr int W = 32;
constexpr int MAXW = 8;  // warps per emulated CTA
// One CTA: nw warps of 32 fibers.  Warp collectives are resolved per warp, __syncthreads over all fibers.
struct Cta {
    int nw = 1;
    ucontext_t main_ctx;
    ucontext_t ctx[MAXW * W];
    std::vector<char> stacks[MAXW * W];
    LaneCtx lanes[MAXW * W];
    bool done[MAXW * W];
    int running = 0;
    uint32_t exch[MAXW][2][W];
    unsigned long long arrivals[MAXW][2];
    unsigned long long phase[MAXW * W];
    uint32_t cta_exch[2][MAXW * W];
    unsigned long long cta_arrivals[2] = {0, 0};
    unsigned long long cta_phase[MAXW * W];
    void (*body)(void*) = nullptr;
    void* arg = nullptr;
};
Cta* g_cta = nullptr;
void yield_lane() {
    Cta* c = g_cta;
    const int me = c->running;
    swapcontext(&c->ctx[me], &c->main_ctx);
}
int arrive(uint32_t v) {  // warp collective: returns buffer index; blocks until the 32 lanes of the warp have arrived
    Cta* c = g_cta;
    const int me = c->running, wi = me / W;
    const unsigned long long ph = c->phase[me]++;
    const int buf = (int)(ph & 1);
    c->exch[wi][buf][me % W] = v;
    c->arrivals[wi][buf]++;
    while (c->arrivals[wi][buf] < (unsigned long long)W * (ph / 2 + 1)) yield_lane();
    return buf;
}
int cta_arrive(uint32_t v) {  // CTA barrier over all fibers
    Cta* c = g_cta;
    const int me = c->running;
    const unsigned long long ph = c->cta_phase[me]++;
    const int buf = (int)(ph & 1);
    c->cta_exch[buf][me] = v;
    c->cta_arrivals[buf]++;
    while (c->cta_arrivals[buf] < (unsigned long long)W * c->nw * (ph / 2 + 1)) yield_lane();
    return buf;
}
void trampoline() {
    Cta* c = g_cta;
    const int me = c->running;
    cur = &c->lanes[me];
    c->body(c->arg);
    c->done[me] = true;
    swapcontext(&c->ctx[me], &c->main_ctx);
}
}  // namespace
uint32_t collective_exchange(uint32_t v, int src_lane) {
    const int wi = g_cta->running / W;
    const int buf = arrive(v);
    return g_cta->exch[wi][buf][src_lane & 31];
}
uint32_t collective_ballot(bool pred) {
    const int wi = g_cta->running / W;
    const int buf = arrive(pred ? 1u : 0u);
    uint32_t r = 0;
    for (int i = 0; i < W; i++) r |= (g_cta->exch[wi][buf][i] & 1u) << i;
    return r;
}
void collective_sync() { (void)arrive(0); }
int cta_barrier_or(int pred) {
    const int buf = cta_arrive(pred ? 1u : 0u);
    uint32_t r = 0;
    for (int i = 0; i < W * g_cta->nw; i++) r |= g_cta->cta_exch[buf][i];
    return r != 0;
}

// run warps [first_warp, first_warp + nw) of a CTA of `body` concurrently to completion
void run_cta(void (*body)(void*), void* arg, dim3 block, dim3 grid, dim3 bdim, int first_warp, int nw) {
    Cta* c = new Cta();
    g_cta = c;
    c->nw = nw;
    c->body = body;
    c->arg = arg;
    std::memset(c->arrivals, 0, sizeof c->arrivals);
    for (int i = 0; i < W * nw; i++) {
        c->stacks[i].resize(1 << 18);
        c->done[i] = false;
        c->phase[i] = 0;
        c->cta_phase[i] = 0;
        c->lanes[i].tid.x = first_warp * 32 + i;
        c->lanes[i].bid = block;
        c->lanes[i].gdim = grid;
        c->lanes[i].bdim = bdim;
        c->lanes[i].lane = i % W;
        getcontext(&c->ctx[i]);
        c->ctx[i].uc_stack.ss_sp = c->stacks[i].data();
        c->ctx[i].uc_stack.ss_size = c->stacks[i].size();
        c->ctx[i].uc_link = &c->main_ctx;
        makecontext(&c->ctx[i], (void (*)())trampoline, 0);
    }
    for (;;) {
        bool all = true;
        for (int i = 0; i < W * nw; i++) {
            if (c->done[i]) continue;
            all = false;
            c->running = i;
            cur = &c->lanes[i];
            swapcontext(&c->main_ctx, &c->ctx[i]);
        }
        if (all) break;
    }
    delete c;
    g_cta = nullptr;
}
// one warp on its own (kernels whose warps do not synchronise with each other run warp after warp)
void run_warp(void (*body)(void*), void* arg, dim3 block, dim3 grid, dim3 bdim, int warp_index) {
    run_cta(body, arg, block, grid, bdim, warp_index, 1);
}
}  // namespace cuda_emu

using namespace scpd;

namespace {
struct Launch {
    FastParams p;
    int g, log2par, ext;
};
template <int G, int LP, bool EXT>
void body_t(void* a) {
    sc_decode_fast_kernel<G, LP, EXT>(static_cast<Launch*>(a)->p);
}
void (*pick(int g, int lp, int ext))(void*) {
#define CASE(G, LP, E) \
    if (g == G && lp == LP && ext == (E ? 1 : 0)) return body_t<G, LP, E>;
    CASE(1, 3, true) CASE(2, 4, true) CASE(4, 4, true) CASE(8, 4, true) CASE(16, 4, true) CASE(32, 4, true)
    CASE(8, 4, false) CASE(8, 2, true) CASE(8, 6, true) CASE(4, 2, true) CASE(2, 1, true) CASE(4, 4, false)
#undef CASE
    return nullptr;
}
}  // namespace

extern "C" {
// Emulates the fast kernel with `warps` warps per CTA and `grid` CTAs.  lsa/lsb as in FastParams
// (pass -1 for "everything in shared memory").  Returns 0, or -1 if the variant is not compiled in.
int emu_fast_decode(int g, int log2n, int log2par, int llr_bits, int extended, int pruning, const uint8_t* flags,
                    const int8_t* llr, size_t nframes, uint32_t* xhat, int lsa, int lsb, int warps, int grid) {
    void (*body)(void*) = pick(g, log2par, extended);
    if (!body) return -1;
    const int log2s = 3;
    int ls = log2s;
    for (int x = g; x > 1; x >>= 1) ls++;
    if (log2par > ls || log2n < ls + 1) return -2;
    ScheduleStats st;
    std::vector<uint32_t> sched = build_schedule(log2n, log2par, extended, pruning, flags, &st, ls);
    const uint32_t n = 1u << log2n;
    Launch L;
    L.g = g;
    FastParams& p = L.p;
    p.sched = sched.data();
    p.llr = llr;
    p.xhat = xhat;
    p.nframes = nframes;
    p.num_fp = (nframes + 1) / 2;
    p.n = n;
    p.log2n = (uint32_t)log2n;
    p.wpf = n / 32;
    p.satv = (1u << (llr_bits - 1)) - 1u;
    if (lsa < 0) lsa = log2n - 1;
    if (lsb < 0) lsb = log2n;
    if (lsa < ls) lsa = ls;
    if (lsb < ls) lsb = ls;
    p.lsa = (uint32_t)lsa;
    p.lsb = (uint32_t)lsb;
    p.sm_alpha_cells = 2u << lsa;
    const size_t beta_bytes = std::min<size_t>((size_t)2u << lsb, n);
    size_t stride = ((size_t)p.sm_alpha_cells * 2 + beta_bytes + 127) & ~(size_t)127;
    if (g < 8) stride += 16 * g;
    p.sm_stride = (uint32_t)stride;
    const int gpw = 32 / g;
    const size_t fp_per_cta = (size_t)warps * gpw;
    if (stride * fp_per_cta > sizeof(smem_fast)) return -3;
    p.ws_stride = ((unsigned long long)n * 5ull + 255ull) & ~255ull;
    std::vector<uint8_t> ws((size_t)grid * fp_per_cta * p.ws_stride + 256, 0xCD);
    p.ws = reinterpret_cast<uint8_t*>(((uintptr_t)ws.data() + 255) & ~(uintptr_t)255);
    dim3 gd, bd;
    gd.x = (unsigned)grid;
    bd.x = (unsigned)warps * 32;
    for (int b = 0; b < grid; b++) {
        std::memset(smem_fast, 0xEE, sizeof(smem_fast));
        dim3 bi;
        bi.x = (unsigned)b;
        for (int w = 0; w < warps; w++) cuda_emu::run_warp(body, &L, bi, gd, bd, w);
    }
    return 0;
}

// Emulates the CTA-cooperative variant of the fast kernel (W warps walk one frame pair together; the fibers of
// all W warps run concurrently with a real CTA barrier).  Returns 0, -1 if W is not compiled in.
int emu_fast_coop_decode(int w, int log2n, int log2par, int llr_bits, int extended, int pruning, const uint8_t* flags,
                         const int8_t* llr, size_t nframes, uint32_t* xhat, int lsa, int lsb, int grid) {
    struct CL {
        FastParams p;
    } L;
    void (*body)(void*) = nullptr;
    if (log2par == 4 && extended) {
        if (w == 2) body = [](void* a) { sc_decode_fast_coop_kernel<4, true, 2>(static_cast<CL*>(a)->p); };
        if (w == 4) body = [](void* a) { sc_decode_fast_coop_kernel<4, true, 4>(static_cast<CL*>(a)->p); };
        if (w == 8) body = [](void* a) { sc_decode_fast_coop_kernel<4, true, 8>(static_cast<CL*>(a)->p); };
    }
    if (!body) return -1;
    const int ls = 8;  // register subtree of 8 * 32 elements
    if (log2n < ls + 1) return -2;
    ScheduleStats st;
    std::vector<uint32_t> sched = build_schedule(log2n, log2par, extended, pruning, flags, &st, ls);
    const uint32_t n = 1u << log2n;
    FastParams& p = L.p;
    p.sched = sched.data();
    p.llr = llr;
    p.xhat = xhat;
    p.nframes = nframes;
    p.num_fp = (nframes + 1) / 2;
    p.n = n;
    p.log2n = (uint32_t)log2n;
    p.wpf = n / 32;
    p.satv = (1u << (llr_bits - 1)) - 1u;
    if (lsa < 0) lsa = log2n - 1;
    if (lsb < 0) lsb = log2n;
    if (lsa < ls) lsa = ls;
    if (lsb < ls) lsb = ls;
    p.lsa = (uint32_t)lsa;
    p.lsb = (uint32_t)lsb;
    p.sm_alpha_cells = 2u << lsa;
    const size_t beta_bytes = std::min<size_t>((size_t)2u << lsb, n);
    const size_t stride = ((size_t)p.sm_alpha_cells * 2 + beta_bytes + 127) & ~(size_t)127;
    p.sm_stride = (uint32_t)stride;
    if (stride > sizeof(smem_fast)) return -3;
    p.ws_stride = ((unsigned long long)n * 5ull + 255ull) & ~255ull;
    std::vector<uint8_t> ws((size_t)grid * p.ws_stride + 256, 0xCD);
    p.ws = reinterpret_cast<uint8_t*>(((uintptr_t)ws.data() + 255) & ~(uintptr_t)255);
    dim3 gd, bd;
    gd.x = (unsigned)grid;
    bd.x = (unsigned)w * 32;
    for (int b = 0; b < grid; b++) {
        std::memset(smem_fast, 0xEE, sizeof(smem_fast));
        dim3 bi;
        bi.x = (unsigned)b;
        cuda_emu::run_cta(body, &L, bi, gd, bd, 0, w);
    }
    return 0;
}

// CTA mode of emu_bs_decode: all warps of a CTA run concurrently with a real CTA barrier, the schedule is copied into
// shared memory by the CTA (BsParams::sched_words) and a barrier op is compiled in front of every sync_every-th op.
static int g_bs_cta = 0, g_bs_sync = 0;
void emu_bs_cta_mode(int on, int sync_every) {
    g_bs_cta = on;
    g_bs_sync = sync_every;
}

// Emulates the bit-sliced kernel (decode_bs.cuh).  fmt 0 = CA2, 1 = SIGMAG; g = lanes per frame group.
// lsa/lsb < 0: planned from smem_per_group.  Returns 0, -1 if the variant is not compiled into the
// emulator, -3 if it does not fit.
int emu_bs_decode(int fmt, int g, int log2n, int log2par, int llr_bits, int extended, int pruning, const uint8_t* flags,
                  const int8_t* llr, size_t nframes, uint32_t* xhat, int lsa, int lsb, int smem_per_group, int warps,
                  int grid, int fuse) {
    struct BL {
        BsParams p;
    } L;
    void (*body)(void*) = nullptr;
#define BCASE(F, Q, LP, E, GG)                                                                    \
    if (fmt == F && llr_bits == Q && log2par == LP && extended == (E ? 1 : 0) && g == GG)         \
        body = [](void* a) { sc_decode_bs_kernel<F, Q, LP, E, GG>(static_cast<BL*>(a)->p); };
    BCASE(0, 8, 4, true, 32) BCASE(0, 8, 4, true, 16) BCASE(0, 8, 4, true, 8) BCASE(0, 8, 4, false, 8)
    BCASE(0, 6, 4, true, 16) BCASE(0, 8, 2, true, 32) BCASE(0, 7, 6, true, 8) BCASE(1, 6, 4, true, 8)
    BCASE(1, 8, 4, true, 32) BCASE(1, 6, 4, false, 16) BCASE(1, 7, 6, true, 32) BCASE(0, 8, 1, true, 8)
#undef BCASE
    if (!body) return -1;
    if (log2n < 7) return -2;
    ScheduleStats st;
    std::vector<uint32_t> ops = build_schedule(log2n, log2par, extended, pruning, flags, &st, BS_LSUB, fmt == 0 ? 1 : 2);
    std::vector<uint32_t> sched;
    if (!bs_compile_schedule(ops, &sched, g_bs_cta ? g_bs_sync : 0, fuse)) return -4;
    BsPlan plan;
    if (!bs_make_plan(log2n, llr_bits, log2par, extended, (size_t)smem_per_group, &plan, lsa, lsb, g)) return -3;
    const int gpw = 32 / g;
    if ((size_t)plan.sm_stride * warps * gpw + (g_bs_cta ? sched.size() * 4 : 0) > sizeof(smem_fast)) return -3;
    BsParams& p = L.p;
    p.sched = sched.data();
    p.prefetch = 0;
    p.sched_words = g_bs_cta ? (uint32_t)sched.size() : 0u;  // warp-after-warp emulation cannot share a CTA's schedule copy
    p.xhat = xhat;
    p.nframes = nframes;
    p.ngroups = (nframes + 31) / 32;
    p.n = 1u << log2n;
    p.log2n = (uint32_t)log2n;
    p.wpf = p.n / 32;
    // channel planes: the transposing kernel, one emulated warp per CTA
    const size_t pl_stride = bs_planes_bytes(llr_bits, log2n);
    std::vector<uint8_t> planes((size_t)p.ngroups * pl_stride + 256, 0xAB);
    uint8_t* pl = reinterpret_cast<uint8_t*>(((uintptr_t)planes.data() + 255) & ~(uintptr_t)255);
    {
        struct PL {
            const int8_t* llr;
            unsigned long long nframes;
            uint32_t n, log2n;
            uint8_t* planes;
            unsigned long long stride;
            int q;
        } P{llr, nframes, p.n, (uint32_t)log2n, pl, pl_stride, llr_bits};
        void (*pbody)(void*) = [](void* a) {
            PL* x = static_cast<PL*>(a);
            if (x->q == 6) bs_planes_kernel<6>(x->llr, x->nframes, x->n, x->log2n, x->planes, x->stride);
            else if (x->q == 7) bs_planes_kernel<7>(x->llr, x->nframes, x->n, x->log2n, x->planes, x->stride);
            else bs_planes_kernel<8>(x->llr, x->nframes, x->n, x->log2n, x->planes, x->stride);
        };
        dim3 pg, pb, pi;
        pg.x = 3;
        pb.x = 32;
        for (unsigned b = 0; b < pg.x; b++) {
            pi.x = b;
            cuda_emu::run_warp(pbody, &P, pi, pg, pb, 0);
        }
    }
    p.planes = pl;
    p.planes_stride = pl_stride;
    p.lsa = plan.lsa;
    p.lsb = plan.lsb;
    p.sm_stride = plan.sm_stride;
    p.sm_beta_off = plan.sm_beta_off;
    p.ws_stride = plan.ws_stride;
    p.ws_beta_off = plan.ws_beta_off;
    for (int l = 0; l < 24; l++) p.aoff[l] = plan.aoff[l];
    std::vector<uint8_t> ws((size_t)grid * warps * gpw * plan.ws_stride + 256, 0xCD);
    p.ws = reinterpret_cast<uint8_t*>(((uintptr_t)ws.data() + 255) & ~(uintptr_t)255);
    dim3 gd, bd;
    gd.x = (unsigned)grid;
    bd.x = (unsigned)warps * 32;
    for (int b = 0; b < grid; b++) {
        std::memset(smem_fast, 0xEE, sizeof(smem_fast));
        dim3 bi;
        bi.x = (unsigned)b;
        if (g_bs_cta)
            cuda_emu::run_cta(body, &L, bi, gd, bd, 0, warps);
        else
            for (int w = 0; w < warps; w++) cuda_emu::run_warp(body, &L, bi, gd, bd, w);
    }
    return 0;
}

// Emulates the raw-pattern kernel (decode_raw.cuh): any format / LLR_BITS / PAR / EXTENDED.  g = lanes per
// frame (8 or 32); ls < 0: every alpha level in shared memory.  pruning is clamped to R0 as scpd_create does.
int emu_raw_decode(int fmt, int g, int log2n, int log2par, int llr_bits, int extended, int pruning, const uint8_t* flags,
                   const int8_t* llr, size_t nframes, uint32_t* xhat, int ls, int beta_in_smem, int warps, int grid) {
    struct RL {
        RawParams p;
    } L;
    void (*body)(void*) = nullptr;
    if (g == 8) body = [](void* a) { sc_decode_raw_kernel<8>(static_cast<RL*>(a)->p); };
    if (g == 32) body = [](void* a) { sc_decode_raw_kernel<32>(static_cast<RL*>(a)->p); };
    if (!body) return -1;
    ScheduleStats st;
    // pruning 3 = SCPD_PRUNE_REF_LEVEL2 (REP / R1 / SPC ops); otherwise only all-frozen nodes are pruned here
    std::vector<uint32_t> sched = build_schedule(log2n, log2par, extended, pruning == 3 ? 3 : pruning > 1 ? 1 : pruning, flags, &st);
    RawParams& p = L.p;
    p.sched = sched.data();
    p.llr = llr;
    p.xhat = xhat;
    p.nframes = nframes;
    p.n = 1u << log2n;
    p.log2n = (uint32_t)log2n;
    p.wpf = p.n >= 32 ? p.n / 32 : 1;
    p.q = (uint32_t)llr_bits;
    p.sigmag = fmt ? 1u : 0u;
    p.log2par = (uint32_t)log2par;
    if (ls < 0 || ls > log2n - 1) ls = log2n - 1;
    if (log2n == 1) ls = 0;
    p.ls = (uint32_t)ls;
    p.beta_in_smem = beta_in_smem ? 1u : 0u;
    p.sm_words_per_frame = (2u << ls) + (beta_in_smem ? p.wpf : 0u);
    p.ws_words_per_frame = (unsigned long long)p.n + p.wpf;
    const size_t f_per_cta = (size_t)warps * (32 / g);
    if ((size_t)p.sm_words_per_frame * 4 * f_per_cta > sizeof(smem_fast)) return -3;
    std::vector<uint32_t> ws((size_t)grid * f_per_cta * p.ws_words_per_frame + 16, 0xCDCDCDCDu);
    p.ws = ws.data();
    dim3 gd, bd;
    gd.x = (unsigned)grid;
    bd.x = (unsigned)warps * 32;
    for (int b = 0; b < grid; b++) {
        std::memset(smem_fast, 0xEE, sizeof(smem_fast));
        dim3 bi;
        bi.x = (unsigned)b;
        for (int w = 0; w < warps; w++) cuda_emu::run_warp(body, &L, bi, gd, bd, w);
    }
    return 0;
}

// The compiled op words of the bit-sliced kernel (host logic under test).  Returns the number of words,
// or a negative code; at most `cap` words are copied to `out`.
long long emu_bs_schedule(int fmt, int log2n, int log2par, int extended, int pruning, const uint8_t* flags, int fuse,
                          uint32_t* out, size_t cap, uint32_t* generic_words) {
    ScheduleStats st;
    std::vector<uint32_t> ops = build_schedule(log2n, log2par, extended, pruning, flags, &st, BS_LSUB, fmt == 0 ? 1 : 2);
    std::vector<uint32_t> sched;
    if (!bs_compile_schedule(ops, &sched, 0, fuse)) return -4;
    if (generic_words) *generic_words = (uint32_t)ops.size();
    for (size_t i = 0; i < sched.size() && i < cap; i++) out[i] = sched[i];
    return (long long)sched.size();
}

// Element functions of bs_arith.cuh on (sign, magnitude) inputs, 32 cases per call word (host logic under
// test against the oracle's sco_f / sco_g / sco_g_ext).  op: 0 = f, 1 = g saturated, 2 = g un-saturated.
// sa/ma/sb/mb/u: `count` entries; outputs sign and magnitude.  P = q - 1 magnitude planes.
int emu_bs_prim(int fmt, int q, int op, size_t count, const uint8_t* sa, const uint32_t* ma, const uint8_t* sb,
                const uint32_t* mb, const uint8_t* u, uint8_t* so, uint32_t* mo) {
    auto run = [&](auto tag) {
        constexpr int P = decltype(tag)::value;
        for (size_t base = 0; base < count; base += 32) {
            bs::Val<P> a{}, b{};
            uint32_t uw = 0;
            for (size_t k = 0; k < 32 && base + k < count; k++) {
                a.s |= (uint32_t)(sa[base + k] & 1u) << k;
                b.s |= (uint32_t)(sb[base + k] & 1u) << k;
                uw |= (uint32_t)(u[base + k] & 1u) << k;
                for (int p = 0; p < P; p++) {
                    a.m[p] |= ((ma[base + k] >> p) & 1u) << k;
                    b.m[p] |= ((mb[base + k] >> p) & 1u) << k;
                }
            }
            bs::Val<P + 1> r{};
            if (op == 0) {
                bs::Val<P> t;
                bs::f_op<P>(a, b, t);
                bs::widen<P + 1, P>(t, r);
            } else if (op == 1) {
                bs::Val<P> t;
                if (fmt == 0)
                    bs::g_sat<bs::FMT_CA2, P>(a, b, uw, t);
                else
                    bs::g_sat<bs::FMT_SM, P>(a, b, uw, t);
                bs::widen<P + 1, P>(t, r);
            } else {
                bs::g_ext<P>(a, b, uw, r);
            }
            for (size_t k = 0; k < 32 && base + k < count; k++) {
                so[base + k] = (uint8_t)((r.s >> k) & 1u);
                uint32_t m = 0;
                for (int p = 0; p <= P; p++) m |= ((r.m[p] >> k) & 1u) << p;
                mo[base + k] = m;
            }
        }
    };
    switch (q) {
        case 5: run(std::integral_constant<int, 4>{}); break;
        case 6: run(std::integral_constant<int, 5>{}); break;
        case 7: run(std::integral_constant<int, 6>{}); break;
        case 8: run(std::integral_constant<int, 7>{}); break;
        default: return -1;
    }
    return 0;
}
}  // extern "C"

// ---- count.cuh under the emulator: counters and polar transform with the frame in registers (wpl = words per lane,
// one of 1 2 4 8 16 32) or in shared memory (wpl = 0).  Returns 0, -1 for an unknown variant.
namespace {
struct CountArgs {
    uint32_t wpf, n, k;
    unsigned long long nframes;
    const uint32_t *xhat, *ref, *mask;
    int per_frame;
    unsigned long long* counters;
    uint32_t* uhat;
};
template <int WPL>
void count_body(void* a) {
    const CountArgs& c = *static_cast<CountArgs*>(a);
    count_all_kernel<WPL>(c.wpf, c.n, c.k, c.nframes, c.xhat, c.ref, c.per_frame, c.mask, c.counters);
}
void count_smem_body(void* a) {
    const CountArgs& c = *static_cast<CountArgs*>(a);
    count_all_smem_kernel(c.wpf, c.n, c.k, c.nframes, c.xhat, c.ref, c.per_frame, c.mask, c.counters);
}
template <int WPL>
void transform_body(void* a) {
    const CountArgs& c = *static_cast<CountArgs*>(a);
    polar_transform_reg_kernel<WPL>(c.wpf, c.nframes, c.xhat, c.uhat);
}
void transform_smem_body(void* a) {
    const CountArgs& c = *static_cast<CountArgs*>(a);
    polar_transform_smem_kernel(c.wpf, c.nframes, c.xhat, c.uhat);
}
int run_count_grid(void (*body)(void*), CountArgs* c, int grid, int warps) {
    if (!body || warps > 8) return -1;
    dim3 gd, bd;
    gd.x = (unsigned)grid;
    bd.x = (unsigned)warps * 32;
    for (int b = 0; b < grid; b++) {
        dim3 bi;
        bi.x = (unsigned)b;
        cuda_emu::run_cta(body, c, bi, gd, bd, 0, warps);
    }
    return 0;
}
}  // namespace
extern "C" {
int emu_count_all(int wpl, uint32_t wpf, uint32_t n, uint32_t k, unsigned long long nframes, const uint32_t* xhat,
                  const uint32_t* ref, int per_frame, const uint32_t* mask, unsigned long long* counters, int grid) {
    CountArgs c{wpf, n, k, nframes, xhat, ref, mask, per_frame, counters, nullptr};
    void (*body)(void*) = wpl == 0 ? count_smem_body : wpl == 1 ? count_body<1> : wpl == 2 ? count_body<2> : wpl == 4 ? count_body<4>
                          : wpl == 8 ? count_body<8> : wpl == 16 ? count_body<16> : wpl == 32 ? count_body<32> : nullptr;
    if (wpl == 0 && wpf + 2 > sizeof(scpd::s_frame) / 4) return -2;
    return run_count_grid(body, &c, grid, 8);
}
int emu_polar_transform(int wpl, uint32_t wpf, unsigned long long nframes, const uint32_t* xhat, uint32_t* uhat, int grid) {
    CountArgs c{wpf, 0, 0, nframes, xhat, nullptr, nullptr, 0, nullptr, uhat};
    void (*body)(void*) = wpl == 0 ? transform_smem_body : wpl == 1 ? transform_body<1> : wpl == 2 ? transform_body<2>
                          : wpl == 4 ? transform_body<4> : wpl == 8 ? transform_body<8> : wpl == 16 ? transform_body<16>
                          : wpl == 32 ? transform_body<32> : nullptr;
    if (wpl == 0 && wpf + 2 > sizeof(scpd::s_frame) / 4) return -2;
    return run_count_grid(body, &c, grid, 8);
}
}
