// harness.cuh -- device versions of the reference testbench blocks (src/testbench/):
//   sc_xorshift128 (two streams) -> sc_awgn (Box-Muller) -> sc_bpsk + sc_adder -> sc_quantizer,
//   sc_error_counter, and the natural-order polar transform for information-bit extraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scpd {

// ---------------------------------------------------------------- xorshift128 jump-ahead
// The generator state (x,y,z,w) is 128 bits and the step (sc_xorshift128.h:78-85) is linear over
// GF(2): state' = T state.  jump[j] holds T^(2^j) as 128 columns of 4 words, so any stream
// position is reached with one matrix-vector product per set bit of the distance.
struct XsJumpTable {
    const uint4* cols;  // [64][128] columns
};

struct Xs128 {
    uint32_t x, y, z, w;
};
__device__ __forceinline__ uint32_t xs128_next(Xs128& s) {
    uint32_t t = s.x;
    t ^= t << 11;
    t ^= t >> 8;
    s.x = s.y;
    s.y = s.z;
    s.z = s.w;
    s.w ^= s.w >> 19;
    s.w ^= t;
    return s.w;
}
// Warp-cooperative state' = M state (all lanes hold the same state; all get the result).
__device__ __forceinline__ Xs128 xs128_matvec(const uint4* __restrict__ cols, Xs128 s, int lane) {
    // lane handles state bits lane, lane+32, lane+64, lane+96 (bit b of word b/32)
    uint4 acc = make_uint4(0, 0, 0, 0);
    const uint32_t sw[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if ((sw[k] >> lane) & 1u) {
            uint4 c = __ldg(cols + k * 32 + lane);
            acc.x ^= c.x;
            acc.y ^= c.y;
            acc.z ^= c.z;
            acc.w ^= c.w;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        acc.x ^= __shfl_xor_sync(0xFFFFFFFFu, acc.x, off);
        acc.y ^= __shfl_xor_sync(0xFFFFFFFFu, acc.y, off);
        acc.z ^= __shfl_xor_sync(0xFFFFFFFFu, acc.z, off);
        acc.w ^= __shfl_xor_sync(0xFFFFFFFFu, acc.w, off);
    }
    Xs128 r = {acc.x, acc.y, acc.z, acc.w};
    return r;
}
__device__ __forceinline__ Xs128 xs128_jump(XsJumpTable jt, Xs128 s, unsigned long long dist, int lane) {
    for (int j = 0; dist; j++, dist >>= 1)
        if (dist & 1ull) s = xs128_matvec(jt.cols + (size_t)j * 128, s, lane);
    return s;
}
__device__ __forceinline__ void xs128_seed(Xs128& a, Xs128& b, uint32_t seed) {
    const uint32_t m = (seed & 0xFFu) * 0x01010101u;  // sc_xorshift128.h:61-66, :99-104
    a.x = 0x12311178u & m;
    a.y = 0x65498732u | m;
    a.z = 0xFEDCAA01u ^ m;
    a.w = 0xF489A179u + m;
    b.x = 0x98765432u & m;
    b.y = 0x12345678u | m;
    b.z = 0xFCBADEFFu ^ m;
    b.w = 0x12121212u + m;
}
__device__ __forceinline__ float xs128_uniform(uint32_t w) {  // sc_xorshift128.h:86, no FMA contraction
    return __fsub_rn(1.0f, __fmul_rn(__uint2float_rn(w), 2.3283064365386963e-10f));
}
__device__ __forceinline__ int quantize_llr(float y) {  // sc_quantizer.h:77-80 with BETA=4, VSAT=+-31
    int iv = __float2int_rz(__fmul_rn(y, 4.0f));
    return max(-31, min(31, iv));
}

// One warp per BLOCK of F consecutive frames (F = 32768 / n for n < 32768, else 1): the block's F n / 2 draws of each
// stream are contiguous in the stream and its F n LLR bytes contiguous in memory.  Lane l produces the draws
// [l c, (l + 1) c) of the block, c = F n / 64 = 2^log2c, i.e. 2c consecutive LLR bytes, 16 at a time.  The jump to the
// block start and the 31-step chain that hands every lane its own stream position are paid once per block instead of
// once per frame (at n = 1024 they cost four times the draws themselves).
// fast = 0: logf / sqrtf / sincosf as libm-grade code (134 instructions per draw, the kernel is issue-bound at 85 %).
// fast = 2 (default, "guarded"): the SFU approximations (__logf, rsqrtf, __sincosf) first; they land within 1.6e-5 sigma
// quantiser steps of the libm-grade value when r1 <= 0.999 (tools/probe/chan_err.cu, 8e10 draws per sigma), so a sample
// further than 2.5e-4 max(sigma, 1) from a bin edge of the quantiser has its bin decided; the others (and r1 > 0.999, where
// the relative error of log explodes) are recomputed with the libm-grade code: same LLRs, about half the instructions.
// fast = 1: approximations only (a quantised LLR differs by one step on about 1e-5 of the samples).
// FAST (the mode above) and CW (a codeword is given; otherwise all-zero: every symbol +1) are compile-time: the kernel is
// bound by instruction issue, and the per-draw tests of run-time flags were a tenth of its instructions.
// MUFU.LG2 / MUFU.RSQ without the denormal pre-scaling __logf / rsqrtf wrap around them (four instructions each): the
// arguments here are r1 >= 2^-24 and t = -2 ln r1 >= 1.2e-7, or t = 0 for r1 = 1, which the guard sends to the libm-grade code
__device__ __forceinline__ float lg2_approx_ftz(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx_ftz(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <int FAST, bool CW>
__global__ void __launch_bounds__(128)
channel_kernel(uint32_t n, unsigned long long first_frame, unsigned long long nframes, uint32_t seed, float sigma, float guard,
               const uint8_t* __restrict__ codeword, int per_frame, int8_t* __restrict__ llr, XsJumpTable jt, int log2c,
               uint32_t fpb /* frames per block */) {
    constexpr int fast = FAST;
    const int lane = threadIdx.x & 31;
    const unsigned long long blk = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned long long f0 = blk * fpb;
    if (f0 >= nframes) return;
    Xs128 a, b;
    xs128_seed(a, b, seed);
    const unsigned long long start = (first_frame + f0) * (unsigned long long)(n / 2);
    a = xs128_jump(jt, a, start, lane);
    b = xs128_jump(jt, b, start, lane);
    {   // chain: lane l needs the state advanced by l * c draws; 31 cooperative products by T^c
        Xs128 ma = a, mb = b;
        const uint4* tc = jt.cols + (size_t)log2c * 128;
        for (int l = 1; l < 32; l++) {
            a = xs128_matvec(tc, a, lane);
            b = xs128_matvec(tc, b, lane);
            if (lane == l) {
                ma = a;
                mb = b;
            }
        }
        a = ma;
        b = mb;
    }
    const unsigned long long c = 1ull << log2c;
    const unsigned long long byte0 = 2ull * c * (unsigned long long)lane;  // first LLR byte of this lane inside the block
    const unsigned long long valid = (nframes - f0 < fpb ? nframes - f0 : (unsigned long long)fpb) * n;  // bytes of the block
    int8_t* out = llr + f0 * n;
    const float two_pi = __fmul_rn(2.0f, 3.14159265358979f);  // sc_awgn.h:61-62
    const uint32_t nmask = n - 1u;
    const bool aligned = (reinterpret_cast<uintptr_t>(llr) & 15u) == 0;
    // 16 bytes (8 draws) per store: a block is 32768 LLRs or one frame of more than that, so c >= 512 (the host checks)
    constexpr uint32_t per = 8u;
    for (unsigned long long d = 0; d < c; d += per) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        const unsigned long long i0 = byte0 + 2ull * d;  // byte offset inside the block
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) {
            float r1 = xs128_uniform(xs128_next(a));
            float r2 = xs128_uniform(xs128_next(b));
            r1 = fmaxf(r1, 5.9604644775390625e-08f);  // the reference has UB at r1 == 0
            const float y = __fmul_rn(two_pi, r2);  // sc_awgn.h:67
            float x, sn, cs;
            if (fast) {
                const float t = __fmul_rn(-1.3862943611198906f, lg2_approx_ftz(r1));  // -2 ln 2 * log2 r1
                x = __fmul_rn(t, rsqrt_approx_ftz(t));
                __sincosf(y, &sn, &cs);
            } else {
                x = sqrtf(__fmul_rn(-2.0f, logf(r1)));  // :68
                sincosf(y, &sn, &cs);
            }
            const float ph = __fmul_rn(x, sn), qu = __fmul_rn(x, cs);  // :74-77
            float s0 = 1.0f, s1 = 1.0f;  // sc_bpsk.h:53
            if (CW) {
                const unsigned long long i = i0 + 2u * k;
                const uint32_t pos = (uint32_t)i & nmask;
                if (per_frame == 2) {  // packed rows [frame][n / 32] (the reference words scpd_run_ber_ex builds on the device)
                    const uint32_t* row = reinterpret_cast<const uint32_t*>(codeword) + (f0 + i / n) * (n >= 32u ? n / 32u : 1u);
                    const uint32_t wv = row[pos >> 5] >> (pos & 31u);
                    s0 = (wv & 1u) ? -1.0f : 1.0f;
                    s1 = (wv & 2u) ? -1.0f : 1.0f;
                } else {
                    const uint8_t* cw = per_frame ? codeword + (f0 + i / n) * n : codeword;
                    s0 = cw[pos] ? -1.0f : 1.0f;
                    s1 = cw[pos + 1] ? -1.0f : 1.0f;
                }
            }
            float v0 = __fadd_rn(s0, __fmul_rn(ph, sigma)), v1 = __fadd_rn(s1, __fmul_rn(qu, sigma));  // sc_adder.h:139-140
            if (fast == 2) {  // guarded: is the quantiser bin of both samples decided whatever the approximation error?
                // distance to the nearest integer through the 1.5 * 2^23 trick (|t| < 2^22: two FADDs on the FMA pipe
                // instead of an FRND on the conversion pipe, which this kernel keeps almost as busy as the issue slots)
                const float t0 = __fmul_rn(v0, 4.0f), t1 = __fmul_rn(v1, 4.0f);
                const float n0 = __fsub_rn(__fadd_rn(t0, 12582912.0f), 12582912.0f);
                const float n1 = __fsub_rn(__fadd_rn(t1, 12582912.0f), 12582912.0f);
                const float dist = fminf(fabsf(__fsub_rn(t0, n0)), fabsf(__fsub_rn(t1, n1)));
                if (dist < guard || !(r1 <= 0.999f)) {
                    x = sqrtf(__fmul_rn(-2.0f, logf(r1)));
                    sincosf(y, &sn, &cs);
                    v0 = __fadd_rn(s0, __fmul_rn(__fmul_rn(x, sn), sigma));
                    v1 = __fadd_rn(s1, __fmul_rn(__fmul_rn(x, cs), sigma));
                }
            }
            const int q0 = quantize_llr(v0);
            const int q1 = quantize_llr(v1);
            w[k >> 1] |= (((uint32_t)q0 & 0xFFu) | (((uint32_t)q1 & 0xFFu) << 8)) << (16 * (k & 1));
        }
        if (i0 < valid) {
            if (aligned && i0 + 16u <= valid) {
                *reinterpret_cast<uint4*>(out + i0) = make_uint4(w[0], w[1], w[2], w[3]);
            } else {  // short lane ranges (n < 16 at the end of the batch) or a buffer that is not 16-byte aligned
                for (uint32_t k = 0; k < per; k++)
                    if (i0 + 2u * k < valid) {
                        const uint32_t h = w[k >> 1] >> (16 * (k & 1));
                        out[i0 + 2u * k] = (int8_t)(h & 0xFFu);
                        out[i0 + 2u * k + 1] = (int8_t)((h >> 8) & 0xFFu);
                    }
            }
        }
    }
}

// ---------------------------------------------------------------- error counter
// sc_error_counter.h:68-125.  One warp per frame; counters accumulated with one atomic per warp.
__global__ void __launch_bounds__(256)
count_errors_kernel(uint32_t wpf, uint32_t n, unsigned long long nframes, const uint32_t* __restrict__ xhat,
                    const uint32_t* __restrict__ ref, int per_frame, unsigned long long* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const unsigned long long wstride = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    unsigned long long be = 0, fe = 0, bew = 0, few = 0, nf = 0;
    const uint32_t tail = (n < 32) ? ((1u << n) - 1u) : 0xFFFFFFFFu;
    for (unsigned long long f = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < nframes;
         f += wstride) {
        uint32_t e = 0;
        const uint32_t* r = ref ? (per_frame ? ref + f * wpf : ref) : nullptr;
        for (uint32_t w = lane; w < wpf; w += 32) e += __popc((xhat[f * wpf + w] ^ (r ? r[w] : 0u)) & tail);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xFFFFFFFFu, e, off);
        be += e;
        fe += (e != 0);
        bew += (e & 1023u);  // sc_uint<10> err (:70-71)
        few += ((e & 1023u) != 0);
        nf += 1;
    }
    if (lane == 0 && nf) {
        atomicAdd(counters + 0, be);
        atomicAdd(counters + 1, fe);
        atomicAdd(counters + 2, nf * n);
        atomicAdd(counters + 3, nf);
        atomicAdd(counters + 4, bew);
        atomicAdd(counters + 5, few);
    }
}

// ---------------------------------------------------------------- input contract check
// counts the LLRs outside +-limit (16 bytes per thread and trip); the decode kernels assume |llr| <= 2^(Q-1) - 1
__global__ void __launch_bounds__(256)
count_out_of_range_kernel(const int8_t* __restrict__ llr, unsigned long long nbytes, int limit, unsigned long long* __restrict__ count) {
    unsigned long long bad = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const bool al = (reinterpret_cast<uintptr_t>(llr) & 15u) == 0;
    const unsigned long long nvec = al ? nbytes / 16 : 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(llr) + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int x = (int)(int8_t)(w[k] >> (8 * b));
                bad += (x > limit || x < -limit);
            }
    }
    for (unsigned long long i = nvec * 16 + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += stride) {
        const int x = llr[i];
        bad += (x > limit || x < -limit);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) bad += __shfl_xor_sync(0xFFFFFFFFu, bad, off);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(count, bad);
}

// ---------------------------------------------------------------- codeword sources of scpd_run_ber_ex
// ref[f] = packed codeword (first_frame + f) % ncw of `cws` ([ncw][wpf] words): the reference's sc_encoder replays its
// stored codewords in turn (sc_encoder.h:91-113: j = 0, 1, 2, 0, ...)
__global__ void __launch_bounds__(256)
ref_cycle_kernel(uint32_t wpf, unsigned long long first_frame, unsigned long long nframes, const uint32_t* __restrict__ cws,
                 uint32_t ncw, uint32_t* __restrict__ ref) {
    const unsigned long long total = nframes * wpf;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long f = i / wpf;
        ref[i] = cws[((first_frame + f) % ncw) * wpf + (uint32_t)(i % wpf)];
    }
}
// u[f] = random information word of frame first_frame + f: 32 bits per word from a counter-based generator
// (splitmix64 finaliser over (payload_seed, frame, word)), masked with the information flags; frozen positions are 0.
// The reference has no payload source; this one is ours and is restated in tests/ for the checks.
__device__ __forceinline__ uint32_t payload_word(unsigned long long seed, unsigned long long frame, uint32_t w) {
    unsigned long long z = seed * 0x9E3779B97F4A7C15ull + frame * 0xBF58476D1CE4E5B9ull + (unsigned long long)w * 0x94D049BB133111EBull;
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)z;
}
__global__ void __launch_bounds__(256)
payload_kernel(uint32_t wpf, unsigned long long first_frame, unsigned long long nframes, unsigned long long seed,
               const uint32_t* __restrict__ info_mask, uint32_t* __restrict__ u) {
    const unsigned long long total = nframes * wpf;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long f = i / wpf;
        const uint32_t w = (uint32_t)(i % wpf);
        u[i] = payload_word(seed, first_frame + f, w) & info_mask[w];
    }
}
__global__ void __launch_bounds__(256)
xor_words_kernel(unsigned long long total, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x)
        out[i] = a[i] ^ b[i];
}
// information-bit errors: d = (x^ ^ x) F^(x)n = u^ ^ u (the transform is linear and its own inverse); counters[0] +=
// popcount(d & info mask), [1] += frames with any, [2] += k per frame, [3] += frames
__global__ void __launch_bounds__(256)
count_info_kernel(uint32_t wpf, uint32_t k, unsigned long long nframes, const uint32_t* __restrict__ d,
                  const uint32_t* __restrict__ info_mask, unsigned long long* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const unsigned long long wstride = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    unsigned long long be = 0, fe = 0, nf = 0;
    for (unsigned long long f = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < nframes;
         f += wstride) {
        uint32_t e = 0;
        for (uint32_t w = lane; w < wpf; w += 32) e += __popc(d[f * wpf + w] & info_mask[w]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xFFFFFFFFu, e, off);
        be += e;
        fe += (e != 0);
        nf += 1;
    }
    if (lane == 0 && nf) {
        atomicAdd(counters + 0, be);
        atomicAdd(counters + 1, fe);
        atomicAdd(counters + 2, nf * k);
        atomicAdd(counters + 3, nf);
    }
}

// ---------------------------------------------------------------- u^ = x^ F^(x)n  (extra)
// One warp per frame, in place on the output buffer (uhat may alias nothing else).
__global__ void __launch_bounds__(128)
polar_transform_kernel(uint32_t wpf, uint32_t n, unsigned long long nframes, const uint32_t* __restrict__ xhat,
                       uint32_t* __restrict__ uhat) {
    const int lane = threadIdx.x & 31;
    const unsigned long long f = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= nframes) return;
    const uint32_t* src = xhat + f * wpf;
    uint32_t* dst = uhat + f * wpf;
    // strides inside a word
    for (uint32_t w = lane; w < wpf; w += 32) {
        uint32_t v = src[w];
        v ^= (v >> 1) & 0x55555555u;
        v ^= (v >> 2) & 0x33333333u;
        v ^= (v >> 4) & 0x0F0F0F0Fu;
        v ^= (v >> 8) & 0x00FF00FFu;
        v ^= (v >> 16) & 0x0000FFFFu;
        if (n < 32) {
            // only strides below n apply; redo with the right subset
            v = src[w];
            for (uint32_t h = 1; h < n; h <<= 1) {
                uint32_t m = 0;
                for (uint32_t b = 0; b < n; b++)
                    if (!(b & h)) m |= 1u << b;
                v ^= (v >> h) & m;
            }
        }
        dst[w] = v;
    }
    __syncwarp();
    for (uint32_t h = 1; h < wpf; h <<= 1) {  // word strides
        for (uint32_t w = lane; w < wpf; w += 32)
            if (!(w & h)) dst[w] ^= dst[w + h];
        __syncwarp();
    }
}

}  // namespace scpd

#include "count.cuh"
