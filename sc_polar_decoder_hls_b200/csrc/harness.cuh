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

// One warp per frame.  Lane l produces draws [l*c, (l+1)*c) of the frame, c = n/64 (n >= 64), so
// it writes 2c consecutive LLR bytes.  For n < 64 lane 0 produces the whole frame.
__global__ void __launch_bounds__(128)
channel_kernel(uint32_t n, unsigned long long first_frame, unsigned long long nframes, uint32_t seed, float sigma,
               const uint8_t* __restrict__ codeword, int per_frame, int8_t* __restrict__ llr, XsJumpTable jt,
               int log2c /* log2(n/64), or -1 */) {
    const int lane = threadIdx.x & 31;
    const unsigned long long f = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= nframes) return;
    Xs128 a, b;
    xs128_seed(a, b, seed);
    const unsigned long long start = (first_frame + f) * (unsigned long long)(n / 2);
    a = xs128_jump(jt, a, start, lane);
    b = xs128_jump(jt, b, start, lane);
    uint32_t ndraw = n / 2;
    if (log2c >= 0) {
        // chain: lane l needs the state advanced by l*c draws; 31 cooperative products by T^c
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
        ndraw = 1u << log2c;
    } else if (lane != 0) {
        return;
    }
    const uint32_t first_draw = (log2c >= 0) ? (uint32_t)lane * ndraw : 0u;
    const uint8_t* cw = codeword ? (per_frame ? codeword + f * n : codeword) : nullptr;
    int8_t* out = llr + f * n;
    const float two_pi = __fmul_rn(2.0f, 3.14159265358979f);  // sc_awgn.h:61-62
    for (uint32_t d = 0; d < ndraw; d++) {
        float r1 = xs128_uniform(xs128_next(a));
        float r2 = xs128_uniform(xs128_next(b));
        r1 = fmaxf(r1, 5.9604644775390625e-08f);  // the reference has UB at r1 == 0
        const float y = __fmul_rn(two_pi, r2);             // sc_awgn.h:67
        const float x = sqrtf(__fmul_rn(-2.0f, logf(r1)));  // :68
        float sn, cs;
        sincosf(y, &sn, &cs);
        const float ph = __fmul_rn(x, sn), qu = __fmul_rn(x, cs);  // :74-77
        const uint32_t i = 2u * (first_draw + d);
        const float s0 = (cw && cw[i]) ? -1.0f : 1.0f;  // sc_bpsk.h:53
        const float s1 = (cw && cw[i + 1]) ? -1.0f : 1.0f;
        const int q0 = quantize_llr(__fadd_rn(s0, __fmul_rn(ph, sigma)));  // sc_adder.h:139-140
        const int q1 = quantize_llr(__fadd_rn(s1, __fmul_rn(qu, sigma)));
        *reinterpret_cast<char2*>(out + i) = make_char2((signed char)q0, (signed char)q1);
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
