// scpd_api.cu -- C ABI (include/scpd.h): handle management, launch plumbing, testbench harness.
// Host logic only; the kernels are in decode_*.cuh / harness.cuh.  No CPU decode path exists.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/scpd.h"
#include "bs_plan.h"
#include "decode_bs.cuh"
#include "decode_fast.cuh"
#include "decode_generic.cuh"
#include "decode_raw.cuh"
#include "decode_ss.cuh"
#include "harness.cuh"
#include "internal.h"
#include "schedule.h"

namespace scpd {

static thread_local std::string g_last_error;
int set_error(int status, const std::string& msg) {
    g_last_error = msg;
    return status;
}
static int cuda_fail(cudaError_t e, const char* what) {
    return set_error(SCPD_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CUDA_TRY(expr)                                   \
    do {                                                 \
        cudaError_t _e = (expr);                         \
        if (_e != cudaSuccess) return cuda_fail(_e, #expr); \
    } while (0)

// ---------------------------------------------------------------- xorshift128 jump matrices
struct Bits128 {
    uint32_t w[4];
};
static Bits128 xs_step(Bits128 s) {  // sc_xorshift128.h:78-85 on (x,y,z,w) = w[0..3]
    uint32_t t = s.w[0];
    t ^= t << 11;
    t ^= t >> 8;
    Bits128 r;
    r.w[0] = s.w[1];
    r.w[1] = s.w[2];
    r.w[2] = s.w[3];
    r.w[3] = s.w[3] ^ (s.w[3] >> 19) ^ t;
    return r;
}
static Bits128 mat_apply(const std::vector<Bits128>& cols, const Bits128& v) {
    Bits128 r = {{0, 0, 0, 0}};
    for (int b = 0; b < 128; b++)
        if ((v.w[b >> 5] >> (b & 31)) & 1u)
            for (int k = 0; k < 4; k++) r.w[k] ^= cols[b].w[k];
    return r;
}
static const std::vector<Bits128>& jump_table_host() {
    static std::vector<Bits128> table;  // [64][128]
    static std::once_flag once;
    std::call_once(once, [] {
        table.resize(64 * 128);
        for (int b = 0; b < 128; b++) {
            Bits128 e = {{0, 0, 0, 0}};
            e.w[b >> 5] = 1u << (b & 31);
            table[b] = xs_step(e);
        }
        for (int j = 1; j < 64; j++) {
            std::vector<Bits128> prev(table.begin() + (j - 1) * 128, table.begin() + j * 128);
            for (int b = 0; b < 128; b++) table[j * 128 + b] = mat_apply(prev, prev[b]);  // T^(2^j) = T^(2^(j-1)) squared
        }
    });
    return table;
}
static std::mutex g_jt_mutex;
static std::map<int, uint4*> g_jt_dev;
static int jump_table_device(int device, XsJumpTable* out) {
    std::lock_guard<std::mutex> lk(g_jt_mutex);
    auto it = g_jt_dev.find(device);
    if (it == g_jt_dev.end()) {
        const auto& h = jump_table_host();
        uint4* d = nullptr;
        CUDA_TRY(cudaMalloc(&d, h.size() * sizeof(Bits128)));
        const cudaError_t ce = cudaMemcpy(d, h.data(), h.size() * sizeof(Bits128), cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) {
            cudaFree(d);
            return cuda_fail(ce, "jump table upload");
        }
        it = g_jt_dev.emplace(device, d).first;
    }
    out->cols = it->second;
    return SCPD_OK;
}

}  // namespace scpd

using namespace scpd;

struct FastPlan {
    int group = 0, warps = 4, log2s = 0, ctas_per_sm = 1;
    int coop = 0;  // > 0: CTA-cooperative variant, `coop` warps walk one frame pair (group = 32, one pair per CTA)
    std::vector<uint32_t> sched_host;
    ScheduleStats stats;
    uint32_t* d_sched = nullptr;
    uint32_t lsa = 0, lsb = 0, sm_alpha_cells = 0, sm_stride = 0;
    size_t smem_bytes = 0;
    unsigned long long ws_stride = 0;
};

struct scpd_decoder {
    scpd_config cfg;
    int device = 0;
    int log2n = 0, log2par = 0;
    uint32_t wpf = 1;
    std::vector<uint32_t> sched_host;
    std::vector<uint8_t> info_flags;  // the frozen table the handle was created with (1 = information bit)
    ScheduleStats stats;
    uint32_t* d_sched = nullptr;
    // generic-kernel layout
    int group = 32;  // lanes per frame pair
    int warps_per_cta = 4;
    uint32_t ls = 0, beta_in_smem = 1, sm_words_per_fp = 0;
    unsigned long long ws_words_per_fp = 0;
    size_t smem_bytes = 0;
    int ctas_per_sm = 1, num_sms = 1;
    uint32_t* d_ws = nullptr;
    size_t ws_bytes = 0;
    // fast-kernel plans (decode_fast.cuh); fast.group == 0: not available for this configuration.
    // fast_wide: the same kernel with wider lane groups (16, 32 lanes per frame pair), for batches too small to
    // fill the GPU with the default group width
    FastPlan fast;
    std::vector<FastPlan> fast_wide;
    uint8_t* d_fast_ws = nullptr;
    size_t fast_ws_bytes = 0;
    // bit-sliced kernel plan (decode_bs.cuh); bs_ok == false: not available for this configuration
    bool bs_ok = false;
    int bs_group = 8, bs_warps = 2, bs_ctas_per_sm = 1;
    bool bs_sched_smem = false;
    bool kernel_pinned = false;  // SCPD_KERNEL was set when the handle was created
    // test / profiling overrides, read once in scpd_create (INTEGRATION.md lists them): nothing reads the environment later
    unsigned fast_wide_lanes = 256;
    uint32_t bs_prefetch = 0;
    unsigned long long bs_min_groups = 0;
    size_t host_chunk_mb = 256;
    BsPlan bs_plan;
    std::vector<uint32_t> bs_sched_host;
    ScheduleStats bs_stats;
    uint32_t* d_bs_sched = nullptr;
    size_t bs_smem_bytes = 0;
    uint8_t* d_bs_ws = nullptr;
    size_t bs_ws_bytes = 0;
    uint8_t* d_bs_planes = nullptr;
    size_t bs_planes_bytes = 0;
    // slot-sliced kernel plan (decode_ss.cuh): lane = frame; ss_ok == false: not available for this configuration
    int ber_batch_mb = 2048;  // LLR staging per batch of the Monte-Carlo loop
    bool ss_l2persist = false;  // L2 persisting window over the hot-level array
    uint4* d_ss_hot = nullptr;
    size_t ss_hot_bytes = 0;
    int ss_pre = 0;  // leading f levels computed by the plane conversion
    bool ss_xf = false;  // the schedule holds fused SS_XF_* ops (kernel instantiation with XF)
    bool ss_ok = false;
    int ss_warps = 16;
    int ss_max_log2n = 15;
    unsigned long long ss_min_tasks = 0;
    bool ss_sched_smem = false;
    SsPlan ss_plan;
    std::vector<uint32_t> ss_sched_host;
    SsStats ss_stats;
    uint32_t* d_ss_sched = nullptr;
    uint4* d_ss_ws = nullptr;
    size_t ss_ws_bytes = 0;
    uint4* d_ss_planes = nullptr;
    size_t ss_planes_bytes = 0;
    unsigned long long* d_ss_prof = nullptr;  // scpd_stage_timing: [2][6][32] cycles / visits, then [2][32] R1 votes / fallbacks
    // raw-pattern kernel plan (decode_raw.cuh): the configurations outside the in-range int16x2 / bit-sliced
    // datapaths.  raw_only: it is the only kernel of this handle
    bool raw_ok = false, raw_only = false;
    int raw_group = 32, raw_ctas_per_sm = 1;
    std::vector<uint32_t> raw_sched_host;
    ScheduleStats raw_stats;
    uint32_t* d_raw_sched = nullptr;
    uint32_t raw_ls = 0, raw_beta_in_smem = 1, raw_sm_words = 0;
    unsigned long long raw_ws_words = 0;
    size_t raw_smem_bytes = 0;
    uint32_t* d_raw_ws = nullptr;
    size_t raw_ws_bytes = 0;
    // timing of the dominant kernel (scpd_kernel_timing)
    bool timing = false;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
    // pipeline of scpd_decode_host: copy-in, compute and copy-out streams over double-buffered staging
    cudaStream_t st_in = nullptr, st_comp = nullptr, st_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_dec[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    int8_t* d_llr2[2] = {nullptr, nullptr};
    uint32_t* d_xhat2[2] = {nullptr, nullptr};
    size_t pipe_frames = 0;
    int pipe_llr_bufs = 0;                 // LLR staging buffers behind d_llr2 (2, or 1 for the large trees of the Monte-Carlo loop)
    cudaEvent_t ev_llr_free = nullptr;     // recorded by decode_ss once the plane conversion has consumed the LLRs
    bool llr_free_recorded = false;
    // staging for scpd_decode_host / scpd_run_ber
    int8_t* d_llr = nullptr;
    uint32_t* d_xhat = nullptr;
    size_t stage_frames = 0;
    unsigned long long* d_counters = nullptr;
    // scratch of scpd_run_ber_ex, kept between calls: per-frame reference words (double-buffered), difference words,
    // information mask, ten counters
    uint32_t *d_ber_ref[2] = {nullptr, nullptr}, *d_ber_diff = nullptr, *d_ber_mask = nullptr;
    size_t ber_ref_bytes[2] = {0, 0}, ber_diff_bytes = 0;
    unsigned long long* d_ber_cnt = nullptr;
    uint64_t launches = 0;
    char last_kernel[96] = "";  // what the last scpd_decode launched (scpd_last_kernel_name)
};

static const void* generic_kernel_ptr(int group) {
    switch (group) {
        case 8: return (const void*)sc_decode_generic_kernel<8>;
        case 16: return (const void*)sc_decode_generic_kernel<16>;
        default: return (const void*)sc_decode_generic_kernel<32>;
    }
}

typedef void (*fast_kernel_t)(const FastParams);
static fast_kernel_t fast_kernel_ptr(int group, int log2par, int ext) {
    if (log2par == 4 && ext == 1) {
        switch (group) {
            case 2: return sc_decode_fast_kernel<2, 4, true>;
            case 4: return sc_decode_fast_kernel<4, 4, true>;
            case 8: return sc_decode_fast_kernel<8, 4, true>;
            case 16: return sc_decode_fast_kernel<16, 4, true>;
            case 32: return sc_decode_fast_kernel<32, 4, true>;
            default: return nullptr;
        }
    }
    if (group == 8) {
        if (log2par == 4 && ext == 0) return sc_decode_fast_kernel<8, 4, false>;
        if (log2par == 2 && ext == 1) return sc_decode_fast_kernel<8, 2, true>;
        if (log2par == 6 && ext == 1) return sc_decode_fast_kernel<8, 6, true>;
    }
    return nullptr;
}

static fast_kernel_t fast_coop_kernel_ptr(int log2par, int ext, int w) {
    if (log2par == 4 && ext == 1) {
        if (w == 4) return sc_decode_fast_coop_kernel<4, true, 4>;
        if (w == 8) return sc_decode_fast_coop_kernel<4, true, 8>;
    }
    return nullptr;
}

static int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

typedef void (*bs_kernel_t)(const BsParams);
// Instantiated (format, LLR_BITS, log2 PAR, EXTENDED, lanes per frame group) combinations of the
// bit-sliced kernel.  g == 0: "is there any instantiation for this configuration".
static bs_kernel_t bs_kernel_ptr(int fmt, int q, int log2par, int ext, int g) {
#define BS_K(F, Q, LP, E, GG)                                                                  \
    if (fmt == F && q == Q && log2par == LP && ext == (E ? 1 : 0) && (g == GG || g == 0)) \
        return sc_decode_bs_kernel<F, Q, LP, E, GG>;
    // the BASELINE setting (CA2, Q = 8, PAR = 16, EXTENDED) in every group width
    BS_K(0, 8, 4, true, 32)
    BS_K(0, 8, 4, true, 16)
    BS_K(0, 8, 4, true, 8)
#ifndef SCPD_FAST_BUILD
    // CA2: other quantisations / leaf widths of script/script_tests.sh
    BS_K(0, 8, 4, false, 32)
    BS_K(0, 7, 4, true, 32)
    BS_K(0, 6, 4, true, 32)
    BS_K(0, 6, 4, true, 16)
    BS_K(0, 6, 4, false, 32)
    BS_K(0, 8, 2, true, 32)
    BS_K(0, 8, 6, true, 32)
    BS_K(0, 7, 6, true, 32)
    BS_K(0, 6, 6, true, 32)
    BS_K(0, 7, 1, true, 32)
    // SIGMAG (the reference's checked-in default is SIGMAG, LLR_BITS 6: config.h:2,11)
    BS_K(1, 6, 4, true, 32)
    BS_K(1, 6, 4, true, 16)
    BS_K(1, 6, 4, false, 32)
    BS_K(1, 6, 6, true, 32)
    BS_K(1, 7, 6, true, 32)
    BS_K(1, 7, 4, true, 32)
    BS_K(1, 8, 6, true, 32)
    BS_K(1, 8, 4, true, 32)
    BS_K(1, 8, 4, false, 32)
#else
    BS_K(1, 6, 4, true, 32)
#endif
#undef BS_K
    return nullptr;
}

typedef void (*ss_kernel_t)(const SsParams);
// Instantiated (LLR_BITS, log2 PAR, EXTENDED) combinations of the slot-sliced kernel (CA2 only).
// prof: the instrumented build (scpd_stage_timing), instantiated for the BASELINE setting only
static ss_kernel_t ss_kernel_ptr(int q, int log2par, int ext, bool prof = false, bool xf = false) {
    // the instrumented build knows the fused ops as well (its code size does not matter)
    if (prof) return (q == 8 && log2par == 4 && ext == 1) ? sc_decode_ss_kernel<8, 4, true, true, true> : nullptr;
    if (xf) {  // the build with the fused SS_XF_* ops (large trees)
        if (q == 8 && log2par == 4 && ext == 1) return sc_decode_ss_kernel<8, 4, true, false, true>;
#ifndef SCPD_FAST_BUILD
        if (q == 8 && log2par == 4 && ext == 0) return sc_decode_ss_kernel<8, 4, false, false, true>;
        if (q == 7 && log2par == 4 && ext == 1) return sc_decode_ss_kernel<7, 4, true, false, true>;
        if (q == 6 && log2par == 4 && ext == 1) return sc_decode_ss_kernel<6, 4, true, false, true>;
#endif
        return nullptr;
    }
#define SS_K(Q, LP, E) \
    if (q == Q && log2par == LP && ext == (E ? 1 : 0)) return sc_decode_ss_kernel<Q, LP, E, false>;
    SS_K(8, 4, true)
#ifndef SCPD_FAST_BUILD
    SS_K(8, 4, false)
    SS_K(7, 4, true)
    SS_K(6, 4, true)
    SS_K(6, 4, false)
    SS_K(8, 2, true)
    SS_K(8, 3, true)
    SS_K(7, 5, true)
    SS_K(6, 5, true)
#endif
#undef SS_K
    return nullptr;
}

// Decide whether the slot-sliced kernel applies and lay out its shared memory / workspace.
static int plan_ss(scpd_decoder* d, const uint8_t* flags) {
    d->ss_ok = false;
    const char* ksel = std::getenv("SCPD_KERNEL");
    if (ksel && std::strcmp(ksel, "ss") != 0 && std::strcmp(ksel, "auto") != 0) return SCPD_OK;
    if (d->cfg.format != SCPD_FMT_CA2 || d->log2n < 7) return SCPD_OK;
    const int q = (int)d->cfg.llr_bits;
    // the un-saturated leaf sums must stay exact in fp16 (11 significant bits): (2^(Q-1) - 1) 2^log2PAR <= 2047
    if (d->cfg.extended && (((1u << (q - 1)) - 1u) << d->log2par) > 2047u) return SCPD_OK;
    ss_kernel_t k = ss_kernel_ptr(q, d->log2par, (int)d->cfg.extended);
    if (!k) return SCPD_OK;
    d->ss_sched_host = ss_build_schedule(d->log2n, (int)d->cfg.pruning, flags, &d->ss_stats, 1);
    d->ss_warps = std::max(1, std::min(SCPD_SS_THREADS / 32, env_int("SCPD_SS_WARPS", SCPD_SS_THREADS / 32)));
    d->ss_sched_smem = d->ss_sched_host.size() <= (size_t)env_int("SCPD_SS_SCHED_SMEM_WORDS", 2048);
    const size_t total = (size_t)227 * 1024 - (d->ss_sched_smem ? d->ss_sched_host.size() * 4 : 0);
    const size_t per_warp = std::min<size_t>(total / d->ss_warps, (size_t)env_int("SCPD_SS_SMEM_KB", 1024) * 1024);
    // one alpha level in tensor memory: 512 columns shared by the warps that sit on the same 32 TMEM lanes
    const uint32_t tm_cols = 512u / (uint32_t)((d->ss_warps + 3) / 4);
    // the smallest workspace level in its own dense array under an L2 persisting window (decode_ss): its lines are
    // rewritten 2^(log2n - level) times per task and otherwise keep falling out of L2 into DRAM (47 % of the walk's DRAM
    // writes at c2).  Measured: c2 443 -> 448, c3 336 -> 339, c1 382 -> 363 (not bound by DRAM): on from N = 4096.  The
    // device's default set-aside (24.9 MB on B200) holds the array (19.4 MB for 2368 resident warps); a larger one costs the
    // rest of the traffic dearly (40 MB: 404, 80 MB: 243 Gb/s at c2), so it is only raised when it is smaller than the array.
    d->ss_l2persist = env_int("SCPD_SS_L2PERSIST", d->log2n >= 12 ? 1 : 0) != 0;
    if (!ss_make_plan(d->log2n, per_warp - 16, &d->ss_plan, env_int("SCPD_SS_LSA", -1), env_int("SCPD_SS_LWIN", -1), tm_cols,
                      env_int("SCPD_SS_LTM", -1), d->ss_l2persist))
        return SCPD_OK;
    if (d->ss_l2persist && d->ss_plan.lhot) {
        int max_persist = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, d->device));
        const size_t want = (size_t)d->num_sms * d->ss_warps * d->ss_plan.hot_stride * 16;
        size_t cur = 0;
        CUDA_TRY(cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize));
        const size_t aside = std::min<size_t>((size_t)max_persist, std::max<size_t>(want, (size_t)env_int("SCPD_SS_L2PERSIST_MB", 0) << 20));
        if (env_int("SCPD_VERBOSE", 0)) fprintf(stderr, "[scpd] L2 persisting: max %d B, hot array %zu B, set-aside %zu B (was %zu)\n", max_persist, want, aside, cur);
        if (cur < aside && (size_t)max_persist >= want) CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, aside));
        if ((size_t)max_persist < want) d->ss_l2persist = false;  // the array still exists, without the window
    }
    const size_t smem = (size_t)d->ss_plan.sm_stride * 16 * d->ss_warps + (d->ss_sched_smem ? d->ss_sched_host.size() * 4 : 0) + 16;
    CUDA_TRY(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // the leading f ops of the walk depend on the channel alone: the plane conversion computes them (decode_ss.cuh).
    // Measured (profiles/tuning_r2.md): c2 385 -> 419 -> 424 Gb/s for 0 / 1 / 2 levels (the walk is DRAM-bound there and
    // reads the planes once less), c1 362 -> 357 -> 349 (its level 9 would leave tensor memory): on from N = 2048
    d->ss_pre = ss_prefuse_depth(d->ss_sched_host, d->log2n, d->ss_plan.lsa,
                                 env_int("SCPD_SS_PRE", d->log2n >= 12 ? 2 : d->log2n == 11 ? 1 : 0));
    // large trees (the walk is bound by DRAM): an f / g op whose result streams through global memory is fused with the f
    // that opens the child, so that level is not read back (SS_XF_*, a separate instantiation of the kernel)
    d->ss_xf = d->log2n >= env_int("SCPD_SS_XF_MIN_LOG2N", 15) && ss_kernel_ptr(q, d->log2par, (int)d->cfg.extended, false, true);
    const int xf_min = d->ss_xf ? (int)std::max(d->ss_plan.lsa, d->ss_plan.ltm) + 1 : 0;
    d->ss_sched_host = ss_build_schedule(d->log2n, (int)d->cfg.pruning, flags, &d->ss_stats, 1, xf_min,
                                         d->ss_pre);
    if (d->ss_xf)
        CUDA_TRY(cudaFuncSetAttribute((const void*)ss_kernel_ptr(q, d->log2par, (int)d->cfg.extended, false, true),
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    d->ss_ok = true;
    // with the fused ops the slot-sliced kernel also wins on the largest trees (c4 275 vs 250, c5 101 vs 99 Gb/s); without
    // that instantiation the frame-sliced kernel keeps N >= 2^16
    d->ss_max_log2n = env_int("SCPD_SS_MAX_LOG2N", d->ss_xf ? 20 : 15);
    // measured crossover with the int16x2 kernel (tools/r2_small.sh): between 8192 and 12288 frames at c1, c2 and c3
    d->ss_min_tasks = (unsigned long long)env_int("SCPD_SS_MIN_TASKS", 5 * d->num_sms / 2);
    if (env_int("SCPD_VERBOSE", 0))
        fprintf(stderr, "[scpd] slot-sliced kernel: %d warps/CTA, alpha levels 6..%u and partial sums below level %u in smem, "
                "%u B/warp, alpha level %u in tensor memory (%u columns/warp), workspace %llu B/warp, %zu schedule words (%s)\n",
                d->ss_warps, d->ss_plan.lsa, d->ss_plan.lwin, d->ss_plan.sm_stride * 16u, d->ss_plan.ltm, d->ss_plan.tm_cols,
                d->ss_plan.ws_stride * 16ull, d->ss_sched_host.size(), d->ss_sched_smem ? "shared" : "global");
    return SCPD_OK;
}

// Decide whether the bit-sliced kernel applies and lay out its shared memory / workspace.
static int plan_bs(scpd_decoder* d, const uint8_t* flags) {
    d->bs_ok = false;
    // SCPD_KERNEL = bs | fast | generic pins the kernel family (tests, profiling); default: the first that applies
    const char* ksel = std::getenv("SCPD_KERNEL");
    if (ksel && std::strcmp(ksel, "bs") != 0 && std::strcmp(ksel, "auto") != 0 && d->cfg.format == SCPD_FMT_CA2)
        return SCPD_OK;
    // lanes per frame group: narrow groups pay off while the tree is small (measured: profiles/README.md)
    int g = env_int("SCPD_BS_GROUP", d->log2n <= 11 ? 16 : 32);
    bs_kernel_t k = bs_kernel_ptr((int)d->cfg.format, (int)d->cfg.llr_bits, d->log2par, (int)d->cfg.extended, g);
    if (!k) {
        g = 32;
        k = bs_kernel_ptr((int)d->cfg.format, (int)d->cfg.llr_bits, d->log2par, (int)d->cfg.extended, g);
    }
    if (!k || d->log2n < 7) return SCPD_OK;
    d->bs_group = g;
    d->bs_warps = std::max(1, std::min(4, env_int("SCPD_BS_WARPS", 4)));
    const size_t per_group = (size_t)env_int("SCPD_BS_SMEM_KB", 7) * 1024;
    if (!bs_make_plan(d->log2n, (int)d->cfg.llr_bits, d->log2par, (int)d->cfg.extended, per_group, &d->bs_plan,
                      env_int("SCPD_BS_LSA", -1), env_int("SCPD_BS_LSB", -1), g))
        return SCPD_OK;
    {
        const std::vector<uint32_t> ops = build_schedule(d->log2n, d->log2par, (int)d->cfg.extended, (int)d->cfg.pruning,
                                                         flags, &d->bs_stats, BS_LSUB, d->cfg.format == SCPD_FMT_CA2 ? 1 : 2);
        // fused X(l) F(l-1) F(l-2) passes pay off once the spilled levels dominate (measured: +8 % at N = 2^15,
        // +10 % at 2^17, a loss at 2^12 and below where the read-back hits L2 anyway)
        if (!bs_compile_schedule(ops, &d->bs_sched_host, env_int("SCPD_BS_SYNC", 0),
                                 env_int("SCPD_BS_FUSE", d->log2n >= 14 ? 3 : 0)))
            return SCPD_OK;
    }
    // a schedule of up to 8 KB is copied into shared memory by every CTA (op fetches then never miss)
    d->bs_sched_smem = d->bs_sched_host.size() <= (size_t)env_int("SCPD_BS_SCHED_SMEM_WORDS", 2048);
    d->bs_smem_bytes = (size_t)d->bs_plan.sm_stride * d->bs_warps * (32 / g) +
                       (d->bs_sched_smem ? d->bs_sched_host.size() * 4 : 0);
    if (d->bs_smem_bytes > 227 * 1024) return SCPD_OK;
    CUDA_TRY(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->bs_smem_bytes));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k, d->bs_warps * 32, d->bs_smem_bytes));
    if (occ < 1) return SCPD_OK;
    d->bs_ctas_per_sm = occ;
    d->bs_ok = true;
    if (env_int("SCPD_VERBOSE", 0))
        fprintf(stderr, "[scpd] bit-sliced kernel: %d lanes/group, %d warps/CTA, alpha levels <=%u and partial sums <=%u "
                "in smem, %u B/group, %d CTAs/SM, workspace %llu B/group, %zu schedule words\n", g, d->bs_warps, d->bs_plan.lsa,
                d->bs_plan.lsb, d->bs_plan.sm_stride, occ, d->bs_plan.ws_stride, d->bs_sched_host.size());
    return SCPD_OK;
}

// Size the shared-memory / workspace layout of the fast kernel with g lanes per frame pair.  fp->group stays 0
// when the kernel does not apply to this configuration.
static int plan_fast_g(scpd_decoder* d, const uint8_t* flags, int g, FastPlan* fp, int coop = 0) {
    fp->group = 0;
    fp->coop = coop;
    if (coop) g = 32;
    fast_kernel_t k = coop ? fast_coop_kernel_ptr(d->log2par, (int)d->cfg.extended, coop)
                           : fast_kernel_ptr(g, d->log2par, (int)d->cfg.extended);
    if (!k) return SCPD_OK;
    const int log2s = ilog2(8 * g);
    if (d->log2par > log2s || d->log2n < log2s + 1 || d->cfg.n < 32) return SCPD_OK;  // leaf must sit inside the register subtree
    fp->log2s = log2s;
    fp->sched_host = build_schedule(d->log2n, d->log2par, (int)d->cfg.extended, (int)d->cfg.pruning, flags,
                                    &fp->stats, log2s);
    fp->warps = coop ? coop : std::max(1, std::min(8, env_int("SCPD_WARPS", 8)));
    const int gpw = 32 / g;
    const int fp_per_cta = coop ? 1 : fp->warps * gpw;
    const size_t budget = (size_t)env_int("SCPD_SMEM_KB", 100) * 1024;
    const size_t per_fp = coop ? budget / 8 : budget / fp_per_cta;  // cooperative: one pair per CTA, many CTAs per SM
    const uint32_t n = d->cfg.n;
    // Shared memory per frame pair: alpha levels log2s..lsa as cells (level l at cell (1 << l): 2 << lsa
    // cells), then the partial-sum block (one byte per element, 2^(lsb+1) bytes, at most n).
    // Start with everything resident and push the largest level out to the workspace until it fits.
    int lsa = d->log2n - 1, lsb = d->log2n;
    auto a_bytes = [&](int l) { return (size_t)(2u << l) * 2; };
    auto b_bytes = [&](int l) { return std::min<size_t>((size_t)2u << l, n); };
    while (a_bytes(lsa) + b_bytes(lsb) > per_fp && (lsa > log2s || lsb > log2s)) {
        const size_t drop_a = lsa > log2s ? a_bytes(lsa) - a_bytes(lsa - 1) : 0;
        const size_t drop_b = lsb > log2s ? b_bytes(lsb) - b_bytes(lsb - 1) : 0;
        if (drop_a >= drop_b && lsa > log2s)
            lsa--;
        else
            lsb--;
    }
    const size_t alpha_bytes = a_bytes(lsa);
    const size_t beta_bytes = b_bytes(lsb);
    size_t stride = alpha_bytes + beta_bytes;
    stride = (stride + 127) & ~(size_t)127;
    {   // spread the groups of a warp over distinct banks (SCPD_BANK_PAD overrides, bytes)
        const int pad = env_int("SCPD_BANK_PAD", g < 8 ? 16 * g : 32);
        stride += (size_t)pad;
    }
    fp->lsa = (uint32_t)lsa;
    fp->lsb = (uint32_t)lsb;
    fp->sm_alpha_cells = (uint32_t)(2u << lsa);
    fp->sm_stride = (uint32_t)stride;
    fp->smem_bytes = stride * fp_per_cta;
    fp->ws_stride = ((unsigned long long)n * 5ull + 255ull) & ~255ull;  // alpha 4n bytes + partial sums n bytes
    CUDA_TRY(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp->smem_bytes));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k, fp->warps * 32, fp->smem_bytes));
    if (occ < 1) return SCPD_OK;  // does not fit: stay on the generic kernel
    fp->ctas_per_sm = occ;
    fp->group = g;
    if (env_int("SCPD_VERBOSE", 0))
        fprintf(stderr, "[scpd] fast kernel G=%d warps/CTA=%d%s: alpha levels <=%d and partial sums <=%d in smem, "
                "%zu B/pair, %zu B/CTA, %d CTAs/SM, workspace %llu B/pair\n", g, fp->warps, coop ? " (one pair per CTA)" : "", lsa, lsb, stride,
                fp->smem_bytes, occ, fp->ws_stride);
    return SCPD_OK;
}

// Decide whether the fast kernel applies; default 8 lanes per frame pair, plus the wider groups for small batches.
static int plan_fast(scpd_decoder* d, const uint8_t* flags) {
    d->fast.group = 0;
    d->fast_wide.clear();
    const char* ksel = std::getenv("SCPD_KERNEL");
    if (ksel && std::strcmp(ksel, "generic") == 0) return SCPD_OK;
    if (d->cfg.format != SCPD_FMT_CA2 || d->cfg.llr_bits > 8 || d->log2par < 1) return SCPD_OK;
    const bool pinned_group = std::getenv("SCPD_GROUP") != nullptr || std::getenv("SCPD_COOP") != nullptr;
    int rc = plan_fast_g(d, flags, env_int("SCPD_GROUP", 8), &d->fast, env_int("SCPD_COOP", 0));
    if (rc == SCPD_OK && !d->fast.group) rc = plan_fast_g(d, flags, 8, &d->fast);
    if (rc != SCPD_OK || !d->fast.group || pinned_group) return rc;
    for (int g = 2 * d->fast.group; g <= 32; g *= 2) {
        FastPlan w;
        rc = plan_fast_g(d, flags, g, &w);
        if (rc != SCPD_OK) return rc;
        if (w.group) d->fast_wide.push_back(std::move(w));
    }
    // beyond a warp per pair: the CTA-cooperative variant.  Measured gain from N = 2^15 up (N = 2^19, 1024 frames:
    // 17 -> 26 Gb/s); at N = 4096 the register subtrees dominate and it loses 10 %
    for (int cw = 4; cw <= 8 && d->log2n >= 14; cw *= 2) {
        FastPlan w;
        rc = plan_fast_g(d, flags, 32, &w, cw);
        if (rc != SCPD_OK) return rc;
        if (w.group) d->fast_wide.push_back(std::move(w));
    }
    return SCPD_OK;
}

// The lane-group width for a batch of num_fp frame pairs: widen the group while the batch leaves most of the
// GPU's lanes idle (measured, profiles/tuning_r1.md: N = 2^19, 1024 frames: 5.8 / 11.0 Gb/s with 8 / 16 lanes).
static const FastPlan& pick_fast(const scpd_decoder* d, unsigned long long num_fp) {
    const FastPlan* best = &d->fast;
    const unsigned long long lanes_wanted = (unsigned long long)d->num_sms * d->fast_wide_lanes;
    for (const FastPlan& w : d->fast_wide)
        if (num_fp * (unsigned)(best->coop ? 32 * best->coop : best->group) < lanes_wanted) best = &w;
    return *best;
}

static const void* raw_kernel_ptr(int group) {
    return group == 8 ? (const void*)sc_decode_raw_kernel<8> : (const void*)sc_decode_raw_kernel<32>;
}

// Shared-memory / workspace layout of the raw-pattern kernel: one 32-bit word per LLR and frame.
static int plan_raw(scpd_decoder* d, const uint8_t* flags) {
    d->raw_ok = false;
    // all-frozen nodes are pruned (identical for any input); the rate-1 shortcut is not (decode_raw.cuh)
    const int raw_pruning = d->cfg.pruning == SCPD_PRUNE_REF_LEVEL2 ? 3 : std::min<int>((int)d->cfg.pruning, SCPD_PRUNE_R0);
    d->raw_sched_host = build_schedule(d->log2n, d->log2par, (int)d->cfg.extended, raw_pruning, flags, &d->raw_stats);
    d->raw_group = d->log2n <= 7 ? 8 : 32;
    const int f_per_cta = d->warps_per_cta * (32 / d->raw_group);
    const size_t per_frame_words = 100 * 1024 / 4 / f_per_cta;  // two CTAs per SM
    int ls = 0;
    for (int l = 0; l <= d->log2n - 1; l++)
        if ((size_t)(2u << l) <= per_frame_words) ls = l;
    d->raw_ls = (uint32_t)ls;
    size_t words = (size_t)(2u << ls);
    d->raw_beta_in_smem = (words + d->wpf <= per_frame_words) ? 1u : 0u;
    if (d->raw_beta_in_smem) words += d->wpf;
    d->raw_sm_words = (uint32_t)words;
    d->raw_smem_bytes = words * 4 * f_per_cta;
    const bool need_ws = ls < d->log2n - 1 || !d->raw_beta_in_smem;
    d->raw_ws_words = need_ws ? (unsigned long long)d->cfg.n + d->wpf : 0ull;
    const void* k = raw_kernel_ptr(d->raw_group);
    CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->raw_smem_bytes));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, d->warps_per_cta * 32, d->raw_smem_bytes));
    if (occ < 1) return set_error(SCPD_E_CUDA, "raw-pattern decode kernel does not fit on an SM");
    d->raw_ctas_per_sm = occ;
    d->raw_ok = true;
    if (env_int("SCPD_VERBOSE", 0))
        fprintf(stderr, "[scpd] raw-pattern kernel: %d lanes/frame, alpha levels <=%d in smem, %zu B/CTA, %d CTAs/SM\n",
                d->raw_group, ls, d->raw_smem_bytes, occ);
    return SCPD_OK;
}

static int plan_layout(scpd_decoder* d) {
    const char* env = std::getenv("SCPD_GROUP");
    d->group = 32;
    if (env) {
        int g = std::atoi(env);
        if (g == 8 || g == 16 || g == 32) d->group = g;
    }
    const int gpw = 32 / d->group;
    const int fp_per_cta = d->warps_per_cta * gpw;
    const size_t budget = 100 * 1024;  // two CTAs per SM
    const size_t per_fp_words = budget / 4 / fp_per_cta;
    const uint32_t n = d->cfg.n;
    // largest ls <= log2n-1 whose alpha levels 0..ls fit
    int ls = 0;
    for (int l = 0; l <= d->log2n - 1; l++)
        if ((size_t)(2u << l) <= per_fp_words) ls = l;
    if (d->log2n == 1) ls = 0;
    d->ls = (uint32_t)ls;
    size_t words = (size_t)(2u << ls);
    d->beta_in_smem = (words + 2 * d->wpf <= per_fp_words) ? 1u : 0u;
    if (d->beta_in_smem) words += 2 * d->wpf;
    d->sm_words_per_fp = (uint32_t)words;
    d->smem_bytes = words * 4 * fp_per_cta;
    const bool need_ws_alpha = (int)d->ls < d->log2n - 1;
    d->ws_words_per_fp = (need_ws_alpha || !d->beta_in_smem) ? (unsigned long long)n + 2ull * d->wpf : 0ull;
    const void* k = generic_kernel_ptr(d->group);
    CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_bytes));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, d->warps_per_cta * 32, d->smem_bytes));
    if (occ < 1) return set_error(SCPD_E_CUDA, "decode kernel does not fit on an SM");
    d->ctas_per_sm = occ;
    return SCPD_OK;
}

extern "C" int scpd_create(const scpd_config* cfg, const uint8_t* flags, int device, scpd_decoder** out) {
    if (!cfg || !flags || !out) return set_error(SCPD_E_ARG, "scpd_create: null argument");
    *out = nullptr;
    const uint32_t n = cfg->n, par = cfg->par;
    if (!is_pow2(n) || n < 2) return set_error(SCPD_E_CONFIG, "n must be a power of two >= 2");
    if (n > (1u << 20)) return set_error(SCPD_E_UNSUPPORTED, "n above 2^20 not supported");
    if (!is_pow2(par) || 2 * par > n)
        return set_error(SCPD_E_CONFIG, "par must be a power of two with 2*par <= n (my_module.h:294)");
    uint32_t k = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (flags[i] > 1) return set_error(SCPD_E_CONFIG, "info flags must be 0 or 1");
        k += flags[i];
    }
    if (k != cfg->k) return set_error(SCPD_E_CONFIG, "k differs from the number of information flags");
    if (cfg->format != SCPD_FMT_CA2 && cfg->format != SCPD_FMT_SIGMAG)
        return set_error(SCPD_E_CONFIG, "format must be SCPD_FMT_CA2 or SCPD_FMT_SIGMAG");
    if (cfg->pruning > SCPD_PRUNE_REF_LEVEL2) return set_error(SCPD_E_CONFIG, "unknown pruning mode");
    if (cfg->pruning == SCPD_PRUNE_REF_LEVEL2 && (par < 2 || par > 256))
        return set_error(SCPD_E_CONFIG, "SCPD_PRUNE_REF_LEVEL2 needs par in 2..256 (ADD_TREE_FUNCTION, library.h:72-93)");
    if (cfg->extended > 1) return set_error(SCPD_E_CONFIG, "extended must be 0 or 1");
    if (cfg->llr_bits < 5 || cfg->llr_bits > 9)
        return set_error(SCPD_E_CONFIG, "llr_bits outside 5..9 (range swept by script/script_tests.sh)");
    if (par > 512) return set_error(SCPD_E_UNSUPPORTED, "par above 512 not supported");
    const int log2par = ilog2(par);
    // Outside the in-range int16x2 / bit-sliced datapaths -> the raw-pattern kernel (decode_raw.cuh):
    //  * llr_bits 5: the +-31 quantiser alphabet wraps modulo 2^5 in the reference (wrapper_in.h:33-34)
    //  * un-saturated leaves wider than int16 (Q + log2 PAR bits incl. the final sum)
    //  * SIGMAG without a bit-sliced instantiation for (llr_bits, par, extended), or n < 128
    const char* ksel_raw = std::getenv("SCPD_KERNEL");
    const bool raw_only =
        cfg->llr_bits < 6 || (cfg->extended && cfg->llr_bits + (uint32_t)log2par > 16) ||
        (cfg->format == SCPD_FMT_SIGMAG &&
         (!bs_kernel_ptr(SCPD_FMT_SIGMAG, (int)cfg->llr_bits, log2par, (int)cfg->extended, 0) || n < 128)) ||
        cfg->pruning == SCPD_PRUNE_REF_LEVEL2 ||  // the reference's REP / SPC decoder: ops of decode_raw.cuh only
        (ksel_raw && std::strcmp(ksel_raw, "raw") == 0);

    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    scpd_decoder* d = new (std::nothrow) scpd_decoder();
    if (!d) return set_error(SCPD_E_NOMEM, "out of host memory");
    d->cfg = *cfg;
    d->info_flags.assign(flags, flags + n);
    d->device = device;
    d->log2n = ilog2(n);
    d->log2par = log2par;
    d->wpf = n >= 32 ? n / 32 : 1;
    d->sched_host = build_schedule(d->log2n, d->log2par, (int)cfg->extended,
                                   std::min<int>((int)cfg->pruning, SCPD_PRUNE_R0_R1), flags, &d->stats);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete d;
        return cuda_fail(e, "cudaGetDeviceProperties");
    }
    d->num_sms = prop.multiProcessorCount;
    d->kernel_pinned = std::getenv("SCPD_KERNEL") != nullptr;
    d->fast_wide_lanes = (unsigned)env_int("SCPD_FAST_WIDE_LANES", 256);
    // measured: +5 % at N = 1024 (latency of the read-back), -2 % at N = 4096 (already short of DRAM bandwidth)
    d->bs_prefetch = (uint32_t)env_int("SCPD_BS_PREFETCH", d->log2n <= 11 ? 1 : 0);
    d->bs_min_groups = (unsigned long long)env_int("SCPD_BS_MIN_GROUPS", d->log2n <= 11 ? 1536 : d->log2n <= 13 ? 768 : d->log2n <= 17 ? 512 : 256);
    d->host_chunk_mb = (size_t)env_int("SCPD_HOST_CHUNK_MB", 256);
    d->ber_batch_mb = std::max(1, env_int("SCPD_BER_BATCH_MB", 2048));
    d->raw_only = raw_only;
    int rc = SCPD_OK;
    if (!raw_only) {
        rc = plan_layout(d);
        if (rc == SCPD_OK) rc = plan_fast(d, flags);
        if (rc == SCPD_OK) rc = plan_bs(d, flags);
        if (rc == SCPD_OK) rc = plan_ss(d, flags);
        if (rc == SCPD_OK && cfg->format == SCPD_FMT_SIGMAG && !d->bs_ok) d->raw_only = true;
    }
    // SIGMAG handles keep the raw-pattern kernel beside the bit-sliced one for mis-aligned LLR buffers
    if (rc == SCPD_OK && (d->raw_only || cfg->format == SCPD_FMT_SIGMAG)) rc = plan_raw(d, flags);
    if (rc != SCPD_OK) {
        delete d;
        return rc;
    }
    e = cudaMalloc(&d->d_sched, d->sched_host.size() * sizeof(uint32_t));
    if (e == cudaSuccess)
        e = cudaMemcpy(d->d_sched, d->sched_host.data(), d->sched_host.size() * sizeof(uint32_t),
                       cudaMemcpyHostToDevice);
    auto upload_fast = [&](FastPlan& fp) {
        if (e != cudaSuccess || !fp.group) return;
        e = cudaMalloc(&fp.d_sched, fp.sched_host.size() * sizeof(uint32_t));
        if (e == cudaSuccess)
            e = cudaMemcpy(fp.d_sched, fp.sched_host.data(), fp.sched_host.size() * sizeof(uint32_t),
                           cudaMemcpyHostToDevice);
    };
    upload_fast(d->fast);
    for (FastPlan& w : d->fast_wide) upload_fast(w);
    if (e == cudaSuccess && d->raw_ok) {
        e = cudaMalloc(&d->d_raw_sched, d->raw_sched_host.size() * sizeof(uint32_t));
        if (e == cudaSuccess)
            e = cudaMemcpy(d->d_raw_sched, d->raw_sched_host.data(), d->raw_sched_host.size() * sizeof(uint32_t),
                           cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && d->bs_ok) {
        e = cudaMalloc(&d->d_bs_sched, d->bs_sched_host.size() * sizeof(uint32_t));
        if (e == cudaSuccess)
            e = cudaMemcpy(d->d_bs_sched, d->bs_sched_host.data(), d->bs_sched_host.size() * sizeof(uint32_t),
                           cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && d->ss_ok) {
        e = cudaMalloc(&d->d_ss_sched, d->ss_sched_host.size() * sizeof(uint32_t));
        if (e == cudaSuccess)
            e = cudaMemcpy(d->d_ss_sched, d->ss_sched_host.data(), d->ss_sched_host.size() * sizeof(uint32_t),
                           cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMalloc(&d->d_counters, 6 * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        scpd_destroy(d);
        return cuda_fail(e, "schedule upload");
    }
    *out = d;
    return SCPD_OK;
}

extern "C" void scpd_destroy(scpd_decoder* d) {
    if (!d) return;
    cudaSetDevice(d->device);
    cudaDeviceSynchronize();  // stream-ordered scratch (grow) is released with plain cudaFree below
    cudaFree(d->d_sched);
    cudaFree(d->d_ws);
    cudaFree(d->fast.d_sched);
    for (FastPlan& w : d->fast_wide) cudaFree(w.d_sched);
    cudaFree(d->d_fast_ws);
    cudaFree(d->d_bs_sched);
    cudaFree(d->d_bs_ws);
    cudaFree(d->d_bs_planes);
    cudaFree(d->d_raw_sched);
    cudaFree(d->d_raw_ws);
    cudaFree(d->d_ss_sched);
    cudaFree(d->d_ss_ws);
    cudaFree(d->d_ss_hot);
    cudaFree(d->d_ss_planes);
    cudaFree(d->d_ss_prof);
    for (int b = 0; b < 2; b++) {
        cudaFree(d->d_llr2[b]);
        cudaFree(d->d_xhat2[b]);
        if (d->ev_in[b]) cudaEventDestroy(d->ev_in[b]);
        if (d->ev_dec[b]) cudaEventDestroy(d->ev_dec[b]);
        if (d->ev_out[b]) cudaEventDestroy(d->ev_out[b]);
    }
    if (d->ev_llr_free) cudaEventDestroy(d->ev_llr_free);
    if (d->ev_k0) cudaEventDestroy(d->ev_k0);
    if (d->ev_k1) cudaEventDestroy(d->ev_k1);
    if (d->st_in) cudaStreamDestroy(d->st_in);
    if (d->st_comp) cudaStreamDestroy(d->st_comp);
    if (d->st_out) cudaStreamDestroy(d->st_out);
    cudaFree(d->d_llr);
    cudaFree(d->d_xhat);
    cudaFree(d->d_counters);
    cudaFree(d->d_ber_ref[0]);
    cudaFree(d->d_ber_ref[1]);
    cudaFree(d->d_ber_diff);
    cudaFree(d->d_ber_mask);
    cudaFree(d->d_ber_cnt);
    delete d;
}

// Handle-owned scratch grows in stream order (cudaFreeAsync / cudaMallocAsync on the caller's stream): scpd_decode
// stays asynchronous even when a larger batch than any before arrives.  Work queued earlier on the same stream keeps
// the old buffer until it has run; a handle is used from one stream at a time (scpd.h).
static int grow(void** ptr, size_t* have, size_t need, cudaStream_t st) {
    if (need <= *have) return SCPD_OK;
    if (*ptr) CUDA_TRY(cudaFreeAsync(*ptr, st));
    *ptr = nullptr;
    *have = 0;
    CUDA_TRY(cudaMallocAsync(ptr, need, st));
    *have = need;
    return SCPD_OK;
}

static int decode_fast(scpd_decoder* d, const int8_t* d_llr, size_t nframes, uint32_t* d_xhat, cudaStream_t st) {
    const unsigned long long num_fp = (nframes + 1) / 2;
    const FastPlan& fp = pick_fast(d, num_fp);
    const int g = fp.group, gpw = 32 / g;
    const unsigned long long fp_per_cta = fp.coop ? 1ull : (unsigned long long)fp.warps * gpw;
    unsigned long long grid = (num_fp + fp_per_cta - 1) / fp_per_cta;
    const unsigned long long max_grid = (unsigned long long)d->num_sms * fp.ctas_per_sm;
    if (grid > max_grid) grid = max_grid;
    const size_t ws_need = (size_t)(grid * fp_per_cta * fp.ws_stride);
    {
        const int grc = grow((void**)&d->d_fast_ws, &d->fast_ws_bytes, ws_need, st);
        if (grc) return grc;
    }
    FastParams p;
    p.sched = fp.d_sched;
    p.llr = d_llr;
    p.xhat = d_xhat;
    p.nframes = nframes;
    p.num_fp = num_fp;
    p.n = d->cfg.n;
    p.log2n = (uint32_t)d->log2n;
    p.wpf = d->wpf;
    p.satv = (1u << (d->cfg.llr_bits - 1)) - 1u;
    p.lsa = fp.lsa;
    p.lsb = fp.lsb;
    p.sm_alpha_cells = fp.sm_alpha_cells;
    p.sm_stride = fp.sm_stride;
    p.ws = d->d_fast_ws;
    p.ws_stride = fp.ws_stride;
    fast_kernel_t k = fp.coop ? fast_coop_kernel_ptr(d->log2par, (int)d->cfg.extended, fp.coop)
                              : fast_kernel_ptr(g, d->log2par, (int)d->cfg.extended);
    if (d->timing) CUDA_TRY(cudaEventRecord(d->ev_k0, st));
    k<<<dim3((unsigned)grid), dim3((unsigned)(fp.warps * 32)), fp.smem_bytes, st>>>(p);
    if (fp.coop)
        snprintf(d->last_kernel, sizeof d->last_kernel, "sc_decode_fast_coop_kernel (int16x2, %d warps per frame pair)", fp.coop);
    else
        snprintf(d->last_kernel, sizeof d->last_kernel, "sc_decode_fast_kernel (int16x2, %d lanes per frame pair)", g);
    if (d->timing) CUDA_TRY(cudaEventRecord(d->ev_k1, st));
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SCPD_OK;
}

static int decode_bs(scpd_decoder* d, const int8_t* d_llr, size_t nframes, uint32_t* d_xhat, cudaStream_t st) {
    const unsigned long long ngroups = (nframes + 31) / 32;
    const unsigned long long gpc = (unsigned long long)d->bs_warps * (32 / d->bs_group);  // groups per CTA
    unsigned long long grid = (ngroups + gpc - 1) / gpc;
    const unsigned long long max_grid = (unsigned long long)d->num_sms * d->bs_ctas_per_sm;
    if (grid > max_grid) grid = max_grid;
    const size_t ws_need = (size_t)(grid * gpc * d->bs_plan.ws_stride);
    {
        const int grc = grow((void**)&d->d_bs_ws, &d->bs_ws_bytes, ws_need, st);
        if (grc) return grc;
    }
    const size_t pl_stride = bs_planes_bytes((int)d->cfg.llr_bits, d->log2n);
    const size_t pl_need = (size_t)ngroups * pl_stride;
    {
        const int grc = grow((void**)&d->d_bs_planes, &d->bs_planes_bytes, pl_need, st);
        if (grc) return grc;
    }
    {   // int8 rows -> bit planes, one warp per (group, 128 LLRs)
        const unsigned long long tasks = ngroups * (d->cfg.n / 128u);
        const unsigned blocks = (unsigned)std::min<unsigned long long>((tasks + 7) / 8, (unsigned long long)d->num_sms * 32);
        switch (d->cfg.llr_bits) {
            case 6: bs_planes_kernel<6><<<blocks, 256, 0, st>>>(d_llr, nframes, d->cfg.n, (uint32_t)d->log2n, d->d_bs_planes, pl_stride); break;
            case 7: bs_planes_kernel<7><<<blocks, 256, 0, st>>>(d_llr, nframes, d->cfg.n, (uint32_t)d->log2n, d->d_bs_planes, pl_stride); break;
            default: bs_planes_kernel<8><<<blocks, 256, 0, st>>>(d_llr, nframes, d->cfg.n, (uint32_t)d->log2n, d->d_bs_planes, pl_stride); break;
        }
        d->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    BsParams p;
    p.sched = d->d_bs_sched;
    p.sched_words = d->bs_sched_smem ? (uint32_t)d->bs_sched_host.size() : 0u;
    p.planes = d->d_bs_planes;
    p.planes_stride = pl_stride;
    p.xhat = d_xhat;
    p.nframes = nframes;
    p.ngroups = ngroups;
    p.n = d->cfg.n;
    p.log2n = (uint32_t)d->log2n;
    p.wpf = d->wpf;
    p.lsa = d->bs_plan.lsa;
    p.lsb = d->bs_plan.lsb;
    p.sm_stride = d->bs_plan.sm_stride;
    p.sm_beta_off = d->bs_plan.sm_beta_off;
    p.ws = d->d_bs_ws;
    p.ws_stride = d->bs_plan.ws_stride;
    p.ws_beta_off = d->bs_plan.ws_beta_off;
    for (int l = 0; l < 24; l++) p.aoff[l] = d->bs_plan.aoff[l];
    p.prefetch = d->bs_prefetch;
    bs_kernel_t k = bs_kernel_ptr((int)d->cfg.format, (int)d->cfg.llr_bits, d->log2par, (int)d->cfg.extended, d->bs_group);
    if (d->timing) CUDA_TRY(cudaEventRecord(d->ev_k0, st));
    k<<<dim3((unsigned)grid), dim3((unsigned)(d->bs_warps * 32)), d->bs_smem_bytes, st>>>(p);
    snprintf(d->last_kernel, sizeof d->last_kernel, "sc_decode_bs_kernel (bit-sliced, %d lanes per 32-frame group)", d->bs_group);
    if (d->timing) CUDA_TRY(cudaEventRecord(d->ev_k1, st));
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SCPD_OK;
}

// slot-sliced kernel: lane = frame, 32 frames per warp ("task")
static int decode_ss(scpd_decoder* d, const int8_t* d_llr, size_t nframes, uint32_t* d_xhat, cudaStream_t st) {
    const unsigned long long ntasks = (nframes + 31) / 32;
    // warps per CTA: as many as the plan allows once every SM has that many tasks, fewer for small batches
    const int warps = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((unsigned long long)d->ss_warps,
                                                                                      (ntasks + d->num_sms - 1) / d->num_sms));
    const unsigned long long grid = std::min<unsigned long long>((unsigned long long)d->num_sms, (ntasks + warps - 1) / warps);
    int rc = grow((void**)&d->d_ss_ws, &d->ss_ws_bytes, (size_t)(grid * warps * d->ss_plan.ws_stride * 16ull) + 16, st);
    if (rc) return rc;
    const size_t hot_bytes = (size_t)(grid * warps) * d->ss_plan.hot_stride * 16;
    rc = grow((void**)&d->d_ss_hot, &d->ss_hot_bytes, hot_bytes + 16, st);
    if (rc) return rc;
    SsPre pre;
    const size_t pl_stride = ss_planes_quads(d->log2n, d->ss_pre, pre.off);
    rc = grow((void**)&d->d_ss_planes, &d->ss_planes_bytes, (size_t)ntasks * pl_stride * 16, st);
    if (rc) return rc;
    {
        const unsigned long long units = ntasks * std::max<unsigned long long>(1, d->cfg.n / 256u);
        const unsigned blocks = (unsigned)std::min<unsigned long long>((units + 7) / 8, (unsigned long long)d->num_sms * 32);
#define SS_PLANES(Q, D) ss_planes_kernel<Q, D><<<blocks, 256, 0, st>>>(d_llr, nframes, d->cfg.n, d->d_ss_planes, pl_stride, pre)
#define SS_PLANES_Q(Q)                   \
    switch (d->ss_pre) {                 \
        case 1: SS_PLANES(Q, 1); break;  \
        case 2: SS_PLANES(Q, 2); break;  \
        case 3: SS_PLANES(Q, 3); break;  \
        default: SS_PLANES(Q, 0); break; \
    }
        switch (d->cfg.llr_bits) {
            case 6: SS_PLANES_Q(6) break;
            case 7: SS_PLANES_Q(7) break;
            default: SS_PLANES_Q(8) break;
        }
#undef SS_PLANES_Q
#undef SS_PLANES
        d->launches++;
        CUDA_TRY(cudaGetLastError());
        if (d->ev_llr_free) {  // the walk reads planes only: the caller's LLR buffer is free from here (scpd_run_ber_ex)
            CUDA_TRY(cudaEventRecord(d->ev_llr_free, st));
            d->llr_free_recorded = true;
        }
    }
    SsParams p;
    p.sched = d->d_ss_sched;
    p.sched_words = d->ss_sched_smem ? (uint32_t)d->ss_sched_host.size() : 0u;
    p.planes = d->d_ss_planes;
    p.planes_stride = pl_stride;
    p.xhat = d_xhat;
    p.nframes = nframes;
    p.ntasks = ntasks;
    p.n = d->cfg.n;
    p.log2n = (uint32_t)d->log2n;
    p.wpf = d->wpf;
    p.lsa = d->ss_plan.lsa;
    p.lwin = d->ss_plan.lwin;
    p.win_words = d->ss_plan.win_words;
    p.ltm = d->ss_plan.ltm;
    p.tm_cols = d->ss_plan.tm_cols;
    p.sm_stride = d->ss_plan.sm_stride;
    p.sm_beta_off = d->ss_plan.sm_beta_off;
    p.ws = d->d_ss_ws;
    p.ws_stride = d->ss_plan.ws_stride;
    p.ws_beta_off = d->ss_plan.ws_beta_off;
    for (int l = 0; l < 24; l++) p.aoff[l] = d->ss_plan.aoff[l];
    p.hot = d->d_ss_hot;
    p.lhot = d->ss_plan.lhot;
    p.hot_stride = d->ss_plan.hot_stride;
    p.lpre = (uint32_t)(d->log2n - d->ss_pre);
    for (int l = 0; l < 24; l++) p.poff[l] = 0u;
    for (int s = 1; s <= d->ss_pre; s++) p.poff[d->log2n - s] = pre.off[s];
    p.prof = d->d_ss_prof;
    const size_t smem = (size_t)d->ss_plan.sm_stride * 16 * warps + (d->ss_sched_smem ? d->ss_sched_host.size() * 4 : 0) + 16;
    ss_kernel_t k = ss_kernel_ptr((int)d->cfg.llr_bits, d->log2par, (int)d->cfg.extended, d->d_ss_prof != nullptr, d->ss_xf);
    if (d->timing) CUDA_TRY(cudaEventRecord(d->ev_k0, st));
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((unsigned)(warps * 32));
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        int nattr = 0;
        if (d->ss_l2persist && d->ss_plan.lhot && hot_bytes) {
            attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
            attr[0].val.accessPolicyWindow.base_ptr = d->d_ss_hot;
            attr[0].val.accessPolicyWindow.num_bytes = hot_bytes;
            attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
            attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            nattr = 1;
        }
        cfg.attrs = attr;
        cfg.numAttrs = nattr;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, k, p));
    }
    snprintf(d->last_kernel, sizeof d->last_kernel, "sc_decode_ss_kernel (slot-sliced, lane per frame, %d warps/CTA)", warps);
    if (d->timing) CUDA_TRY(cudaEventRecord(d->ev_k1, st));
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SCPD_OK;
}

static int decode_raw(scpd_decoder* d, const int8_t* d_llr, size_t nframes, uint32_t* d_xhat, cudaStream_t st) {
    const unsigned long long f_per_cta = (unsigned long long)d->warps_per_cta * (32 / d->raw_group);
    unsigned long long grid = (nframes + f_per_cta - 1) / f_per_cta;
    const unsigned long long max_grid = (unsigned long long)d->num_sms * d->raw_ctas_per_sm;
    if (grid > max_grid) grid = max_grid;
    const size_t ws_need = (size_t)(grid * f_per_cta * d->raw_ws_words * 4ull);
    {
        const int grc = grow((void**)&d->d_raw_ws, &d->raw_ws_bytes, ws_need, st);
        if (grc) return grc;
    }
    RawParams p;
    p.sched = d->d_raw_sched;
    p.llr = d_llr;
    p.xhat = d_xhat;
    p.nframes = nframes;
    p.n = d->cfg.n;
    p.log2n = (uint32_t)d->log2n;
    p.wpf = d->wpf;
    p.q = d->cfg.llr_bits;
    p.sigmag = d->cfg.format == SCPD_FMT_SIGMAG ? 1u : 0u;
    p.log2par = (uint32_t)d->log2par;
    p.ls = d->raw_ls;
    p.beta_in_smem = d->raw_beta_in_smem;
    p.sm_words_per_frame = d->raw_sm_words;
    p.ws = d->d_raw_ws;
    p.ws_words_per_frame = d->raw_ws_words;
    const dim3 g((unsigned)grid), b((unsigned)(d->warps_per_cta * 32));
    if (d->timing) CUDA_TRY(cudaEventRecord(d->ev_k0, st));
    if (d->raw_group == 8)
        sc_decode_raw_kernel<8><<<g, b, d->raw_smem_bytes, st>>>(p);
    else
        sc_decode_raw_kernel<32><<<g, b, d->raw_smem_bytes, st>>>(p);
    snprintf(d->last_kernel, sizeof d->last_kernel, "sc_decode_raw_kernel (raw W-bit patterns, %d lanes per frame)", d->raw_group);
    if (d->timing) CUDA_TRY(cudaEventRecord(d->ev_k1, st));
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SCPD_OK;
}

extern "C" int scpd_decode(scpd_decoder* d, const int8_t* d_llr, size_t nframes, uint32_t* d_xhat, void* stream) {
    if (!d) return set_error(SCPD_E_ARG, "scpd_decode: null decoder");
    if (nframes == 0) return SCPD_OK;
    if (!d_llr || !d_xhat) return set_error(SCPD_E_ARG, "scpd_decode: null buffer");
    CUDA_TRY(cudaSetDevice(d->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (d->raw_only) return decode_raw(d, d_llr, nframes, d_xhat, st);
    // The slot-sliced kernel (a lane per frame) is the default for CA2 once the batch gives every SM a few warps
    // (measured, profiles/tuning_r2.md: 386 / 443 / 331 / 275 / 101 Gb/s against 231 / 288 / 287 / 250 / 99 at c1 ... c5);
    // configurations without the instantiation that fuses f / g with the child's f use it up to N = 2^15
    if (d->ss_ok && (reinterpret_cast<uintptr_t>(d_llr) & 15u) == 0 &&
        (d->kernel_pinned || (d->log2n <= d->ss_max_log2n && (nframes + 31) / 32 >= d->ss_min_tasks)))
        return decode_ss(d, d_llr, nframes, d_xhat, st);
    // One warp walks the tree of a 32-frame group alone, so the bit-sliced kernel needs many groups to fill the
    // GPU; below the measured crossover (profiles/tuning_r1.md: about 49 k frames at N = 1024, 24 k at N = 4096,
    // 16 k at N = 32768 and 131072, 8 k at N = 2^19) the int16x2 kernel is faster: 2 frames per lane group, and
    // lane groups that widen to 16 / 32 lanes as the batch shrinks (pick_fast).
    const unsigned long long bs_min_groups = d->bs_min_groups;
    const bool bs_small = d->fast.group && d->cfg.format == SCPD_FMT_CA2 && !d->kernel_pinned &&
                          (nframes + 31) / 32 < bs_min_groups;
    if (d->bs_ok && !bs_small && (reinterpret_cast<uintptr_t>(d_llr) & 3u) == 0)
        return decode_bs(d, d_llr, nframes, d_xhat, st);
    if (d->cfg.format != SCPD_FMT_CA2) return decode_raw(d, d_llr, nframes, d_xhat, st);  // mis-aligned LLR buffer
    if (d->fast.group && (reinterpret_cast<uintptr_t>(d_llr) & 7u) == 0) return decode_fast(d, d_llr, nframes, d_xhat, st);
    const int gpw = 32 / d->group;
    const unsigned long long fp_per_cta = (unsigned long long)d->warps_per_cta * gpw;
    const unsigned long long num_fp = (nframes + 1) / 2;
    unsigned long long grid = (num_fp + fp_per_cta - 1) / fp_per_cta;
    const unsigned long long max_grid = (unsigned long long)d->num_sms * d->ctas_per_sm;
    if (grid > max_grid) grid = max_grid;
    const size_t ws_need = (size_t)(grid * fp_per_cta * d->ws_words_per_fp * 4ull);
    {
        const int grc = grow((void**)&d->d_ws, &d->ws_bytes, ws_need, st);
        if (grc) return grc;
    }
    DecodeParams p;
    p.sched = d->d_sched;
    p.llr = d_llr;
    p.xhat = d_xhat;
    p.nframes = nframes;
    p.num_fp = num_fp;
    p.n = d->cfg.n;
    p.log2n = (uint32_t)d->log2n;
    p.wpf = d->wpf;
    p.satv = (1u << (d->cfg.llr_bits - 1)) - 1u;
    p.log2par = (uint32_t)d->log2par;
    p.extended = d->cfg.extended;
    p.ls = d->ls;
    p.beta_in_smem = d->beta_in_smem;
    p.sm_words_per_fp = d->sm_words_per_fp;
    p.ws = d->d_ws;
    p.ws_words_per_fp = d->ws_words_per_fp;
    const dim3 g((unsigned)grid), b((unsigned)(d->warps_per_cta * 32));
    if (d->timing) CUDA_TRY(cudaEventRecord(d->ev_k0, st));
    switch (d->group) {
        case 8: sc_decode_generic_kernel<8><<<g, b, d->smem_bytes, st>>>(p); break;
        case 16: sc_decode_generic_kernel<16><<<g, b, d->smem_bytes, st>>>(p); break;
        default: sc_decode_generic_kernel<32><<<g, b, d->smem_bytes, st>>>(p); break;
    }
    snprintf(d->last_kernel, sizeof d->last_kernel, "sc_decode_generic_kernel (int16x2, %d lanes per frame pair)", d->group);
    if (d->timing) CUDA_TRY(cudaEventRecord(d->ev_k1, st));
    d->launches++;
    CUDA_TRY(cudaGetLastError());
    return SCPD_OK;
}

static int ensure_stage(scpd_decoder* d, size_t nframes) {
    if (nframes <= d->stage_frames) return SCPD_OK;
    cudaFree(d->d_llr);
    cudaFree(d->d_xhat);
    d->d_llr = nullptr;
    d->d_xhat = nullptr;
    d->stage_frames = 0;
    CUDA_TRY(cudaMalloc(&d->d_llr, nframes * (size_t)d->cfg.n));
    CUDA_TRY(cudaMalloc(&d->d_xhat, nframes * (size_t)d->wpf * 4));
    d->stage_frames = nframes;
    return SCPD_OK;
}

static int ensure_pipeline(scpd_decoder* d, size_t chunk, int llr_bufs = 2) {
    if (!d->st_in) {
        CUDA_TRY(cudaStreamCreateWithFlags(&d->st_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&d->st_comp, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&d->st_out, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            CUDA_TRY(cudaEventCreateWithFlags(&d->ev_in[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&d->ev_dec[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&d->ev_out[b], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaEventCreateWithFlags(&d->ev_llr_free, cudaEventDisableTiming));
    }
    if (chunk <= d->pipe_frames && llr_bufs <= d->pipe_llr_bufs) return SCPD_OK;
    for (int b = 0; b < 2; b++) {
        cudaFree(d->d_llr2[b]);
        cudaFree(d->d_xhat2[b]);
        d->d_llr2[b] = nullptr;
        d->d_xhat2[b] = nullptr;
    }
    d->pipe_frames = 0;
    d->pipe_llr_bufs = 0;
    for (int b = 0; b < 2; b++) {
        if (b < llr_bufs) CUDA_TRY(cudaMalloc(&d->d_llr2[b], chunk * (size_t)d->cfg.n));
        CUDA_TRY(cudaMalloc(&d->d_xhat2[b], chunk * (size_t)d->wpf * 4));
    }
    d->pipe_frames = chunk;
    d->pipe_llr_bufs = llr_bufs;
    return SCPD_OK;
}

// Host-buffer entry point.  The batch is cut into chunks (about 256 MiB of LLRs each) that flow through a
// three-stage pipeline -- H2D copy, decode, D2H copy, one stream each, double-buffered staging -- so that
// the PCIe transfers of neighbouring chunks overlap the decode and each other (full duplex).
extern "C" int scpd_decode_host(scpd_decoder* d, const int8_t* h_llr, size_t nframes, uint32_t* h_xhat) {
    if (!d) return set_error(SCPD_E_ARG, "scpd_decode_host: null decoder");
    if (nframes == 0) return SCPD_OK;
    if (!h_llr || !h_xhat) return set_error(SCPD_E_ARG, "scpd_decode_host: null buffer");
    CUDA_TRY(cudaSetDevice(d->device));
    const size_t n = d->cfg.n, row_out = (size_t)d->wpf * 4;
    size_t chunk = (d->host_chunk_mb << 20) / n;
    chunk = std::max<size_t>(chunk, (size_t)64 * 32 * (size_t)d->num_sms / 16);  // >= 4 frame groups per SM per chunk
    chunk = std::max<size_t>(32, chunk & ~(size_t)31);
    if (chunk > nframes) chunk = nframes;
    int rc = ensure_pipeline(d, chunk);
    if (rc) return rc;
    size_t i = 0;
    for (size_t f0 = 0; f0 < nframes; f0 += chunk, i++) {
        const int b = (int)(i & 1);
        const size_t nb = std::min(chunk, nframes - f0);
        if (i >= 2) CUDA_TRY(cudaStreamWaitEvent(d->st_in, d->ev_dec[b], 0));  // decode i-2 has consumed this buffer
        CUDA_TRY(cudaMemcpyAsync(d->d_llr2[b], h_llr + f0 * n, nb * n, cudaMemcpyHostToDevice, d->st_in));
        CUDA_TRY(cudaEventRecord(d->ev_in[b], d->st_in));
        CUDA_TRY(cudaStreamWaitEvent(d->st_comp, d->ev_in[b], 0));
        if (i >= 2) CUDA_TRY(cudaStreamWaitEvent(d->st_comp, d->ev_out[b], 0));  // copy-out i-2 has drained this buffer
        rc = scpd_decode(d, d->d_llr2[b], nb, d->d_xhat2[b], d->st_comp);
        if (rc) {  // nothing may still be writing into the caller's buffers when the error is returned
            cudaStreamSynchronize(d->st_in);
            cudaStreamSynchronize(d->st_comp);
            cudaStreamSynchronize(d->st_out);
            return rc;
        }
        CUDA_TRY(cudaEventRecord(d->ev_dec[b], d->st_comp));
        CUDA_TRY(cudaStreamWaitEvent(d->st_out, d->ev_dec[b], 0));
        CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(h_xhat) + f0 * row_out, d->d_xhat2[b], nb * row_out,
                                 cudaMemcpyDeviceToHost, d->st_out));
        CUDA_TRY(cudaEventRecord(d->ev_out[b], d->st_out));
    }
    CUDA_TRY(cudaStreamSynchronize(d->st_out));
    CUDA_TRY(cudaStreamSynchronize(d->st_comp));
    CUDA_TRY(cudaStreamSynchronize(d->st_in));
    return SCPD_OK;
}

extern "C" int scpd_kernel_timing(scpd_decoder* d, int enable) {
    if (!d) return set_error(SCPD_E_ARG, "scpd_kernel_timing: null decoder");
    CUDA_TRY(cudaSetDevice(d->device));
    if (enable && !d->ev_k0) {
        CUDA_TRY(cudaEventCreate(&d->ev_k0));
        CUDA_TRY(cudaEventCreate(&d->ev_k1));
    }
    d->timing = enable != 0;
    return SCPD_OK;
}
extern "C" int scpd_last_kernel_ms(scpd_decoder* d, float* ms) {
    if (!d || !ms) return set_error(SCPD_E_ARG, "scpd_last_kernel_ms: null argument");
    if (!d->timing || !d->ev_k1) return set_error(SCPD_E_ARG, "scpd_last_kernel_ms: timing is off");
    CUDA_TRY(cudaSetDevice(d->device));
    CUDA_TRY(cudaEventSynchronize(d->ev_k1));
    CUDA_TRY(cudaEventElapsedTime(ms, d->ev_k0, d->ev_k1));
    return SCPD_OK;
}
extern "C" const char* scpd_kernel_name(const scpd_decoder* d) {
    static thread_local char buf[96];
    if (!d) return "";
    if (d->raw_only)
        snprintf(buf, sizeof buf, "sc_decode_raw_kernel (raw W-bit patterns, %d lanes per frame)", d->raw_group);
    else if (d->ss_ok && (d->log2n <= d->ss_max_log2n || !d->bs_ok))
        snprintf(buf, sizeof buf, "sc_decode_ss_kernel (slot-sliced, lane per frame, %d warps/CTA)", d->ss_warps);
    else if (d->bs_ok)
        snprintf(buf, sizeof buf, "sc_decode_bs_kernel (bit-sliced, %d lanes per 32-frame group, %d warps/CTA)", d->bs_group,
                 d->bs_warps);
    else if (d->fast.group)
        snprintf(buf, sizeof buf, "sc_decode_fast_kernel (int16x2, %d lanes per frame pair)", d->fast.group);
    else
        snprintf(buf, sizeof buf, "sc_decode_generic_kernel (int16x2, %d lanes per frame pair)", d->group);
    return buf;
}

extern "C" const char* scpd_last_kernel_name(const scpd_decoder* d) { return d ? d->last_kernel : ""; }

// One CTA per frame with the frame in shared memory (N > 32768): grid and dynamic shared memory, 0 blocks = does not fit
static unsigned smem_frame_grid(const void* kernel, uint32_t wpf, size_t nframes, int num_sms, size_t* smem) {
    *smem = ((size_t)wpf + 2) * 4;  // the frame, and two words for the sums of count_all_smem_kernel
    if (*smem > 200u * 1024u) return 0;
    if (*smem > 48u * 1024u && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    const unsigned long long per_sm = std::max<unsigned long long>(1, std::min<unsigned long long>(8, (220u * 1024u) / *smem));
    return (unsigned)std::min<unsigned long long>(nframes, per_sm * (unsigned long long)num_sms);
}
// u = x F^(x)n of whole frames: in registers (one pass over memory) for 32 <= n <= 32768, in shared memory above that,
// through the output buffer for n < 32 (and frames beyond the shared memory of an SM)
static cudaError_t launch_polar_transform(uint32_t wpf, uint32_t n, size_t nframes, const uint32_t* src, uint32_t* dst,
                                          int num_sms, cudaStream_t st) {
    const unsigned blocks = (unsigned)std::min<unsigned long long>((nframes + 7) / 8, (unsigned long long)num_sms * 8);
    size_t smem = 0;
    unsigned grid = 0;
#define TR(W) polar_transform_reg_kernel<W><<<blocks, 256, 0, st>>>(wpf, nframes, src, dst)
    if (n >= 32 && wpf > 1024 && (grid = smem_frame_grid((const void*)polar_transform_smem_kernel, wpf, nframes, num_sms, &smem)) != 0)
        polar_transform_smem_kernel<<<grid, 256, smem, st>>>(wpf, nframes, src, dst);
    else if (n < 32 || wpf > 1024) polar_transform_kernel<<<(unsigned)((nframes + 3) / 4), 128, 0, st>>>(wpf, n, nframes, src, dst);
    else if (wpf <= 32) TR(1);
    else if (wpf <= 64) TR(2);
    else if (wpf <= 128) TR(4);
    else if (wpf <= 256) TR(8);
    else if (wpf <= 512) TR(16);
    else TR(32);
#undef TR
    return cudaGetLastError();
}
// all ten counters of the Monte-Carlo loop from one pass over x^ (n >= 32); false = not available for this n
static bool launch_count_all(uint32_t wpf, uint32_t n, uint32_t k, size_t nframes, const uint32_t* xhat, const uint32_t* ref,
                             int per_frame, const uint32_t* mask, unsigned long long* cnt, int num_sms, cudaStream_t st) {
    if (n < 32) return false;
    if (wpf > 1024) {
        size_t smem = 0;
        const unsigned grid = smem_frame_grid((const void*)count_all_smem_kernel, wpf, nframes, num_sms, &smem);
        if (!grid) return false;
        count_all_smem_kernel<<<grid, 256, smem, st>>>(wpf, n, k, nframes, xhat, ref, per_frame, mask, cnt);
        return true;
    }
    const unsigned blocks = (unsigned)std::min<unsigned long long>((nframes + 7) / 8, (unsigned long long)num_sms * 8);
#define CA(W) count_all_kernel<W><<<blocks, 256, 0, st>>>(wpf, n, k, nframes, xhat, ref, per_frame, mask, cnt)
    if (wpf <= 32) CA(1);
    else if (wpf <= 64) CA(2);
    else if (wpf <= 128) CA(4);
    else if (wpf <= 256) CA(8);
    else if (wpf <= 512) CA(16);
    else CA(32);
#undef CA
    return true;
}

extern "C" int scpd_extract_info(scpd_decoder* d, const uint32_t* d_xhat, size_t nframes, uint32_t* d_uhat,
                                 void* stream) {
    if (!d) return set_error(SCPD_E_ARG, "scpd_extract_info: null decoder");
    if (nframes == 0) return SCPD_OK;
    if (!d_xhat || !d_uhat) return set_error(SCPD_E_ARG, "scpd_extract_info: null buffer");
    CUDA_TRY(cudaSetDevice(d->device));
    CUDA_TRY(launch_polar_transform(d->wpf, d->cfg.n, nframes, d_xhat, d_uhat, d->num_sms, (cudaStream_t)stream));
    d->launches++;
    return SCPD_OK;
}

// Measured function x level matrix: with timing on, warp 0 of every CTA of the slot-sliced kernel adds the
// SM-clock cycles it spends in each schedule op to a [function][level] histogram.
extern "C" int scpd_stage_timing(scpd_decoder* d, int enable) {
    if (!d) return set_error(SCPD_E_ARG, "scpd_stage_timing: null decoder");
    ss_kernel_t kp = d->ss_ok ? ss_kernel_ptr((int)d->cfg.llr_bits, d->log2par, (int)d->cfg.extended, true) : nullptr;
    if (!kp)
        return set_error(SCPD_E_UNSUPPORTED, "scpd_stage_timing: the instrumented slot-sliced kernel exists for CA2, LLR_BITS 8, PAR 16, EXTENDED only");
    CUDA_TRY(cudaSetDevice(d->device));
    CUDA_TRY(cudaFuncSetAttribute((const void*)kp, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_TRY(cudaDeviceSynchronize());
    if (enable && !d->d_ss_prof) {
        CUDA_TRY(cudaMalloc(&d->d_ss_prof, (2 * 6 + 2) * 32 * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemset(d->d_ss_prof, 0, (2 * 6 + 2) * 32 * sizeof(unsigned long long)));
    } else if (!enable && d->d_ss_prof) {
        cudaFree(d->d_ss_prof);
        d->d_ss_prof = nullptr;
    }
    return SCPD_OK;
}
extern "C" int scpd_stage_time(scpd_decoder* d, uint64_t cycles[6][32], uint64_t visits[6][32]) {
    if (!d || !cycles || !visits) return set_error(SCPD_E_ARG, "scpd_stage_time: null argument");
    if (!d->d_ss_prof) return set_error(SCPD_E_ARG, "scpd_stage_time: timing is off (scpd_stage_timing)");
    CUDA_TRY(cudaSetDevice(d->device));
    CUDA_TRY(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(2 * 6 * 32);
    CUDA_TRY(cudaMemcpy(h.data(), d->d_ss_prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemset(d->d_ss_prof, 0, h.size() * sizeof(unsigned long long)));
    for (int f = 0; f < 6; f++)
        for (int l = 0; l < 32; l++) {
            cycles[f][l] = h[f * 32 + l];
            visits[f][l] = h[192 + f * 32 + l];
        }
    return SCPD_OK;
}

// R1 shortcut statistics of the profiled warps since the last call (the PROF build of the slot-sliced kernel): votes[l] =
// R1 nodes of 2^l LLRs reached (l = 5: 32-LLR children, 6: 64-LLR nodes, above: SS_R1 ops), fallbacks[l] = those that a zero
// LLR somewhere in the warp's 32 frames sent to the full walk instead (CA2: hd(0) = 0 whatever the sign)
extern "C" int scpd_r1_votes(scpd_decoder* d, uint64_t votes[32], uint64_t fallbacks[32]) {
    if (!d || !votes || !fallbacks) return set_error(SCPD_E_ARG, "scpd_r1_votes: null argument");
    if (!d->d_ss_prof) return set_error(SCPD_E_ARG, "scpd_r1_votes: timing is off (scpd_stage_timing)");
    CUDA_TRY(cudaSetDevice(d->device));
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned long long h[64];
    CUDA_TRY(cudaMemcpy(h, d->d_ss_prof + 384, sizeof h, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemset(d->d_ss_prof + 384, 0, sizeof h));
    for (int l = 0; l < 32; l++) {
        votes[l] = h[l];
        fallbacks[l] = h[32 + l];
    }
    return SCPD_OK;
}

// Input contract of scpd_decode: |llr| <= 2^(LLR_BITS-1) - 1 (the reference's quantiser alphabet is +-31 whatever
// LLR_BITS is, main.cpp:16-18; wider inputs up to the internal saturation are accepted).  Values outside are undefined
// behaviour in the fast kernels -- they keep them as they are or drop high magnitude bits, depending on the kernel --
// and wrap modulo 2^LLR_BITS in the reference (wrapper_in.h:33-34); this pass lets an integrator check a batch.
extern "C" int scpd_validate_llr(scpd_decoder* d, const int8_t* d_llr, size_t nframes, uint64_t* h_out_of_range, void* stream) {
    if (!d || !h_out_of_range) return set_error(SCPD_E_ARG, "scpd_validate_llr: null argument");
    *h_out_of_range = 0;
    if (nframes == 0) return SCPD_OK;
    if (!d_llr) return set_error(SCPD_E_ARG, "scpd_validate_llr: null buffer");
    CUDA_TRY(cudaSetDevice(d->device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(d->d_counters, 0, sizeof(unsigned long long), st));
    const int limit = (1 << (d->cfg.llr_bits - 1)) - 1;
    count_out_of_range_kernel<<<(unsigned)d->num_sms * 8u, 256, 0, st>>>(d_llr, (unsigned long long)nframes * d->cfg.n, limit, d->d_counters);
    CUDA_TRY(cudaGetLastError());
    unsigned long long h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, d->d_counters, sizeof h, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *h_out_of_range = h;
    d->launches++;
    return SCPD_OK;
}

extern "C" int scpd_get_config(const scpd_decoder* d, scpd_config* out) {
    if (!d || !out) return set_error(SCPD_E_ARG, "scpd_get_config: null argument");
    *out = d->cfg;
    return SCPD_OK;
}
extern "C" int scpd_schedule_stats(const scpd_decoder* d, uint64_t* n_ops, uint64_t* n_fg) {
    if (!d) return set_error(SCPD_E_ARG, "scpd_schedule_stats: null decoder");
    if (!d->raw_only && d->ss_ok && d->log2n <= d->ss_max_log2n) {
        if (n_ops) *n_ops = d->ss_stats.n_ops;
        if (n_fg) *n_fg = d->ss_stats.n_f + d->ss_stats.n_g;
        return SCPD_OK;
    }
    const ScheduleStats& st = d->raw_only ? d->raw_stats : d->bs_ok ? d->bs_stats : d->fast.group ? d->fast.stats : d->stats;
    if (n_ops) *n_ops = st.n_ops;
    if (n_fg) *n_fg = st.n_f + st.n_g;
    return SCPD_OK;
}
extern "C" uint64_t scpd_launch_count(const scpd_decoder* d) { return d ? d->launches : 0; }

extern "C" float scpd_sigma(float ebn0_db, float rate) {  // main.cpp:91-98
    return 1.0f / sqrtf(2.f * rate * powf(10.f, ebn0_db / 10.f));
}

// 2 = SFU approximations, libm-grade recomputation next to a quantiser bin edge (same LLRs; harness.cuh); 0 = libm-grade
// only; 1 = approximations only.  SCPD_CHANNEL_FAST sets the initial value (read once).
static std::atomic<int>& channel_mode() {
    static std::atomic<int> mode(std::max(0, std::min(2, env_int("SCPD_CHANNEL_FAST", 2))));
    return mode;
}
extern "C" int scpd_channel_mode(int mode) {
    if (mode < 0) return channel_mode().load();
    return channel_mode().exchange(std::min(mode, 2));
}

extern "C" int scpd_channel_generate(uint32_t n, uint64_t first_frame, size_t nframes, uint8_t seed, float sigma,
                                     const uint8_t* d_codeword, int per_frame, int8_t* d_llr, void* stream) {
    if (nframes == 0) return SCPD_OK;
    if (!d_llr) return set_error(SCPD_E_ARG, "scpd_channel_generate: null buffer");
    if (!is_pow2(n) || n < 2) return set_error(SCPD_E_CONFIG, "n must be a power of two >= 2");
    int device = 0;
    CUDA_TRY(cudaGetDevice(&device));
    XsJumpTable jt;
    int rc = jump_table_device(device, &jt);
    if (rc) return rc;
    // a warp per block of frames: 32768 LLRs when the frame is shorter than that (harness.cuh)
    const uint32_t fpb = n < 32768u ? 32768u / n : 1u;
    const int log2c = ilog2(fpb * n / 64u);  // >= 9: a lane's range is a whole number of 16-byte stores
    const unsigned long long nblk = (nframes + fpb - 1) / fpb;
    const unsigned grid = (unsigned)((nblk + 3) / 4);
    const int mode = channel_mode().load();
    const float guard = 2.5e-4f * std::max(sigma, 1.0f);  // harness.cuh: distance from a quantiser bin edge below which a sample is recomputed
#define SCPD_CHAN(M, C) \
    channel_kernel<M, C><<<grid, 128, 0, (cudaStream_t)stream>>>(n, first_frame, nframes, seed, sigma, guard, d_codeword, per_frame, d_llr, jt, log2c, fpb)
    if (d_codeword) {
        if (mode == 0) SCPD_CHAN(0, true); else if (mode == 1) SCPD_CHAN(1, true); else SCPD_CHAN(2, true);
    } else {
        if (mode == 0) SCPD_CHAN(0, false); else if (mode == 1) SCPD_CHAN(1, false); else SCPD_CHAN(2, false);
    }
#undef SCPD_CHAN
    CUDA_TRY(cudaGetLastError());
    return SCPD_OK;
}

extern "C" int scpd_count_errors(uint32_t n, size_t nframes, const uint32_t* d_xhat, const uint32_t* d_ref,
                                 int per_frame, uint64_t* d_counters, void* stream) {
    if (nframes == 0) return SCPD_OK;
    if (!d_xhat || !d_counters) return set_error(SCPD_E_ARG, "scpd_count_errors: null buffer");
    if (!is_pow2(n) || n < 2) return set_error(SCPD_E_CONFIG, "n must be a power of two >= 2");
    const uint32_t wpf = n >= 32 ? n / 32 : 1;
    unsigned long long blocks = (nframes + 7) / 8;
    int device = 0, sms = 148;
    CUDA_TRY(cudaGetDevice(&device));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (blocks > (unsigned long long)sms * 8) blocks = (unsigned long long)sms * 8;
    count_errors_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        wpf, n, nframes, d_xhat, d_ref, per_frame, (unsigned long long*)d_counters);
    CUDA_TRY(cudaGetLastError());
    return SCPD_OK;
}

// Monte-Carlo loop on the device: source -> channel -> decode -> count, batch after batch.  Generation of batch i + 1
// runs on a second stream while batch i is decoded and counted (double-buffered staging).
//   src_mode SCPD_SRC_CODEWORDS: h_codewords = ncw codewords of n bytes (0/1); frame g of the stream sends codeword
//            g % ncw (ncw = 3 is the reference's sc_encoder for N in {8, 512, 1024}, sc_encoder.h:91-113); ncw = 0 /
//            NULL: the all-zero codeword (what the reference sends for any other N, :105-110)
//   src_mode SCPD_SRC_RANDOM: K random information bits per frame, encoded on the device (x = u F^(x)n, natural
//            order, frozen positions 0)
// counters[0..5] as scpd_count_errors (codeword bits, what the reference counts); [6..9] information bits:
// errors, frames in error, bits, frames (u^ = x^ F^(x)n on the information positions).
extern "C" int scpd_run_ber_ex(scpd_decoder* d, float ebn0_db, float rate, uint64_t first_frame, uint64_t nframes,
                               uint8_t seed, int src_mode, const uint8_t* h_codewords, uint32_t ncw, uint64_t payload_seed,
                               uint64_t h_counters[10]) {
    if (!d || !h_counters) return set_error(SCPD_E_ARG, "scpd_run_ber_ex: null argument");
    if (src_mode != SCPD_SRC_CODEWORDS && src_mode != SCPD_SRC_RANDOM) return set_error(SCPD_E_ARG, "scpd_run_ber_ex: unknown source mode");
    if (src_mode == SCPD_SRC_CODEWORDS && (!h_codewords || ncw == 0)) {
        h_codewords = nullptr;
        ncw = 0;
    }
    CUDA_TRY(cudaSetDevice(d->device));
    const uint32_t n = d->cfg.n, wpf = d->wpf;
    std::memset(h_counters, 0, 10 * sizeof(uint64_t));
    // batch so that LLR staging stays around 2 GiB per buffer, but never fewer than 8 tasks of 32 frames per SM; a whole
    // number of rounds of the slot-sliced kernel's resident warps, so that no decode of the loop ends on a ragged round
    size_t batch = (size_t)(((size_t)d->ber_batch_mb << 20) / n);
    batch = std::max<size_t>(batch, (size_t)8 * 32 * (size_t)d->num_sms);
    batch &= ~(size_t)31;
    if (d->ss_ok && d->log2n <= d->ss_max_log2n) {
        const size_t round = (size_t)d->num_sms * (size_t)d->ss_warps * 32;
        if (batch >= round) batch = batch / round * round;
    }
    // Large trees: a batch of the size above leaves the slot-sliced kernel with half its warps (c5: 37 888 frames for the
    // 75 776 the machine holds) and two 20 GB LLR buffers.  The walk reads planes only, so ONE LLR buffer does -- batch
    // i + 1 is generated into it as soon as the plane conversion of batch i is through (ev_llr_free), overlapping the
    // walk exactly as two buffers would -- and the batch grows towards a full round, as far as 85 % of the free memory
    // carries LLRs, planes, workspace (about n bytes per resident frame) and output; the frames of the call are then
    // split evenly over the batches.
    int llr_bufs = 2;
    if (d->ss_ok && d->log2n <= d->ss_max_log2n) {
        const size_t round = (size_t)d->num_sms * (size_t)d->ss_warps * 32;
        size_t free_b = 0, total_b = 0;
        if (batch < round && nframes > batch && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            SsPre pre;
            const size_t per_frame = (size_t)n + (ss_planes_quads(d->log2n, d->ss_pre, pre.off) * 16) / 32 + 3 * (size_t)wpf * 4;
            free_b += d->pipe_frames * ((size_t)d->pipe_llr_bufs * n + 2 * (size_t)wpf * 4) + d->ss_planes_bytes + d->ss_ws_bytes;
            const double budget = 0.85 * (double)free_b - 1.1 * (double)round * (double)n;
            size_t fit = budget > 0 ? (size_t)(budget / (double)per_frame) : 0;
            fit = std::min(fit, round) & ~(size_t)31;
            if (fit > batch) {
                const size_t nb = (size_t)((nframes + fit - 1) / fit);
                batch = (size_t)(((nframes + nb - 1) / nb + 31) & ~(uint64_t)31);
                llr_bufs = 1;
            }
        }
    }
    if (batch > nframes) batch = (size_t)nframes;
    if (batch == 0) return SCPD_OK;
    int rc = ensure_pipeline(d, batch, llr_bufs);
    if (rc) return rc;
    llr_bufs = d->pipe_llr_bufs;  // an earlier call may have left two
    const bool per_frame_ref = src_mode == SCPD_SRC_RANDOM || ncw > 1;
    struct Tmp {  // small per-call buffers, freed on every exit; the large scratch lives in the handle
        uint8_t* cw = nullptr;
        uint32_t* cws = nullptr;
        ~Tmp() {
            cudaFree(cw);
            cudaFree(cws);
        }
        uint32_t *mask = nullptr, *ref[2] = {nullptr, nullptr}, *diff = nullptr;
        unsigned long long* cnt = nullptr;
    } tmp;
    auto keep = [](void** ptr, size_t* have, size_t need) -> cudaError_t {  // grow a handle-owned buffer (synchronous)
        if (need <= *have) return cudaSuccess;
        cudaFree(*ptr);
        *ptr = nullptr;
        *have = 0;
        const cudaError_t e = cudaMalloc(ptr, need);
        if (e == cudaSuccess) *have = need;
        return e;
    };
    if (!d->d_ber_cnt) CUDA_TRY(cudaMalloc(&d->d_ber_cnt, 10 * sizeof(unsigned long long)));
    tmp.cnt = d->d_ber_cnt;
    if (!d->d_ber_mask) {  // information mask (packed flags) for the information-bit counters and the payload source
        std::vector<uint32_t> mask(wpf, 0);
        for (uint32_t i = 0; i < n; i++)
            if (d->info_flags[i]) mask[i >> 5] |= 1u << (i & 31);
        CUDA_TRY(cudaMalloc(&d->d_ber_mask, wpf * 4));
        CUDA_TRY(cudaMemcpy(d->d_ber_mask, mask.data(), wpf * 4, cudaMemcpyHostToDevice));
    }
    tmp.mask = d->d_ber_mask;
    if (ncw >= 1) {
        std::vector<uint32_t> packed((size_t)ncw * wpf, 0);
        for (uint32_t j = 0; j < ncw; j++)
            for (uint32_t i = 0; i < n; i++)
                if (h_codewords[(size_t)j * n + i] & 1) packed[(size_t)j * wpf + (i >> 5)] |= 1u << (i & 31);
        CUDA_TRY(cudaMalloc(&tmp.cws, packed.size() * 4));
        CUDA_TRY(cudaMemcpy(tmp.cws, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
        if (ncw == 1) {
            CUDA_TRY(cudaMalloc(&tmp.cw, n));
            CUDA_TRY(cudaMemcpy(tmp.cw, h_codewords, n, cudaMemcpyHostToDevice));
        }
    }
    if (per_frame_ref)
        for (int b = 0; b < 2; b++) {
            CUDA_TRY(keep((void**)&d->d_ber_ref[b], &d->ber_ref_bytes[b], batch * (size_t)wpf * 4));
            tmp.ref[b] = d->d_ber_ref[b];
        }
    cudaStream_t sg = d->st_in, sd = d->st_comp;  // generator stream, decode + count stream
    CUDA_TRY(cudaMemsetAsync(tmp.cnt, 0, 10 * sizeof(unsigned long long), sd));
    const float sigma = scpd_sigma(ebn0_db, rate);
    const unsigned gblocks = (unsigned)d->num_sms * 8u;
    size_t i = 0;
    for (uint64_t done = 0; done < nframes && rc == SCPD_OK; done += batch, i++) {
        const int b = (int)(i & 1);
        const size_t nb = (size_t)((nframes - done < batch) ? nframes - done : batch);
        if (i >= 2) CUDA_TRY(cudaStreamWaitEvent(sg, d->ev_out[b], 0));  // batch i - 2 has been decoded and counted
        int8_t* const llr_buf = d->d_llr2[llr_bufs == 2 ? b : 0];
        const uint8_t* cw_arg = tmp.cw;
        int cw_mode = 0;
        if (per_frame_ref) {
            if (src_mode == SCPD_SRC_RANDOM) {
                payload_kernel<<<gblocks, 256, 0, sg>>>(wpf, first_frame + done, nb, payload_seed, tmp.mask, tmp.ref[b]);
                CUDA_TRY(launch_polar_transform(wpf, n, nb, tmp.ref[b], tmp.ref[b], d->num_sms, sg));
            } else {
                ref_cycle_kernel<<<gblocks, 256, 0, sg>>>(wpf, first_frame + done, nb, tmp.cws, ncw, tmp.ref[b]);
            }
            CUDA_TRY(cudaGetLastError());
            cw_arg = reinterpret_cast<const uint8_t*>(tmp.ref[b]);
            cw_mode = 2;
            d->launches += src_mode == SCPD_SRC_RANDOM ? 2 : 1;
        }
        rc = scpd_channel_generate(n, first_frame + done, nb, seed, sigma, cw_arg, cw_mode, llr_buf, sg);
        if (rc) break;
        CUDA_TRY(cudaEventRecord(d->ev_in[b], sg));
        CUDA_TRY(cudaStreamWaitEvent(sd, d->ev_in[b], 0));
        d->llr_free_recorded = false;
        rc = scpd_decode(d, llr_buf, nb, d->d_xhat2[b], sd);
        if (rc) break;
        if (llr_bufs == 1) {  // the generator may overwrite the LLRs once they are planes; after the whole decode otherwise
            if (!d->llr_free_recorded) CUDA_TRY(cudaEventRecord(d->ev_llr_free, sd));
            CUDA_TRY(cudaStreamWaitEvent(sg, d->ev_llr_free, 0));
        }
        const uint32_t* ref = per_frame_ref ? tmp.ref[b] : tmp.cws;  // nullptr = all-zero
        if (launch_count_all(wpf, n, d->cfg.k, nb, d->d_xhat2[b], ref, per_frame_ref ? 1 : 0, tmp.mask, tmp.cnt, d->num_sms, sd)) {
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaEventRecord(d->ev_out[b], sd));
            d->launches += 2;  // channel, counters
            continue;
        }
        // frames the one-pass kernels do not take (n < 32: a frame is part of one word; frames beyond the shared memory of an SM)
        if (!tmp.diff) {  // x^ ^ x of a batch
            CUDA_TRY(keep((void**)&d->d_ber_diff, &d->ber_diff_bytes, batch * (size_t)wpf * 4));
            tmp.diff = d->d_ber_diff;
        }
        rc = scpd_count_errors(n, nb, d->d_xhat2[b], ref, per_frame_ref ? 1 : 0, (uint64_t*)tmp.cnt, sd);
        if (rc) break;
        // information bits: (x^ ^ x) F^(x)n on the information positions
        const unsigned long long total = (unsigned long long)nb * wpf;
        if (ref && per_frame_ref) {
            xor_words_kernel<<<gblocks, 256, 0, sd>>>(total, d->d_xhat2[b], ref, tmp.diff);
        } else if (ref) {  // one shared codeword: x^ ^ x row by row through the cycle kernel's indexing
            ref_cycle_kernel<<<gblocks, 256, 0, sd>>>(wpf, 0, nb, ref, 1, tmp.diff);
            xor_words_kernel<<<gblocks, 256, 0, sd>>>(total, d->d_xhat2[b], tmp.diff, tmp.diff);
        } else {
            CUDA_TRY(cudaMemcpyAsync(tmp.diff, d->d_xhat2[b], total * 4, cudaMemcpyDeviceToDevice, sd));
        }
        polar_transform_kernel<<<(unsigned)((nb + 3) / 4), 128, 0, sd>>>(wpf, n, nb, tmp.diff, tmp.diff);
        count_info_kernel<<<(unsigned)std::min<unsigned long long>((nb + 7) / 8, (unsigned long long)d->num_sms * 8), 256, 0, sd>>>(
            wpf, d->cfg.k, nb, tmp.diff, tmp.mask, tmp.cnt + 6);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(d->ev_out[b], sd));
        d->launches += 5;
    }
    cudaError_t e1 = cudaStreamSynchronize(sg), e2 = cudaStreamSynchronize(sd);  // also on the error paths: nothing in flight
    if (rc) return rc;
    if (e1 != cudaSuccess) return cuda_fail(e1, "generator stream");
    if (e2 != cudaSuccess) return cuda_fail(e2, "decode stream");
    CUDA_TRY(cudaMemcpy(h_counters, tmp.cnt, 10 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return SCPD_OK;
}

// sc_top_module of src/testbench with one stored codeword (or the all-zero one): the six counters of sc_error_counter
extern "C" int scpd_run_ber(scpd_decoder* d, float ebn0_db, float rate, uint64_t first_frame, uint64_t nframes,
                            uint8_t seed, const uint8_t* h_codeword, uint64_t h_counters[6]) {
    if (!d || !h_counters) return set_error(SCPD_E_ARG, "scpd_run_ber: null argument");
    uint64_t c[10];
    const int rc = scpd_run_ber_ex(d, ebn0_db, rate, first_frame, nframes, seed, SCPD_SRC_CODEWORDS, h_codeword,
                                   h_codeword ? 1u : 0u, 0, c);
    if (rc) return rc;
    std::memcpy(h_counters, c, 6 * sizeof(uint64_t));
    return SCPD_OK;
}

extern "C" const char* scpd_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* scpd_status_string(int s) {
    switch (s) {
        case SCPD_OK: return "ok";
        case SCPD_E_ARG: return "bad argument";
        case SCPD_E_CONFIG: return "invalid configuration";
        case SCPD_E_UNSUPPORTED: return "unsupported configuration";
        case SCPD_E_IO: return "i/o error";
        case SCPD_E_CUDA: return "cuda error";
        case SCPD_E_NOMEM: return "out of memory";
        default: return "unknown status";
    }
}
