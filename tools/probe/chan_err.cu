// chan_err.cu -- how far the SFU approximations of the channel's Box-Muller step (harness.cuh, fast path) land from the
// libm-grade path, in units of the quantiser step: max |4 (s + sigma n_fast) - 4 (s + sigma n_precise)| over 2^36 draws
// of (r1, r2) on the 2^-32 grid the reference's uniform generator produces (sc_xorshift128.h:86).  Sizes the guard band
// of the guarded fast path (a sample closer than EPS to a bin edge is recomputed with logf / sqrtf / sincosf).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float uni(uint32_t w) { return __fsub_rn(1.0f, __fmul_rn(__uint2float_rn(w), 2.3283064365386963e-10f)); }

__global__ void probe(unsigned long long per_thread, float sigma, float r1_cut, float* max_err, unsigned long long* hist) {
    unsigned long long s = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345ull;
    float worst = 0.0f;
    const float two_pi = __fmul_rn(2.0f, 3.14159265358979f);
    for (unsigned long long i = 0; i < per_thread; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;   // xorshift64: covers the 32-bit grids of both uniforms
        float r1 = fmaxf(uni((uint32_t)s), 5.9604644775390625e-08f), r2 = uni((uint32_t)(s >> 32));
        if (r1 > r1_cut) continue;  // the guarded path recomputes these
        const float y = __fmul_rn(two_pi, r2);
        float sp, cp, sf, cf;
        const float xp = sqrtf(__fmul_rn(-2.0f, logf(r1)));
        sincosf(y, &sp, &cp);
        float l2, rs;  // the kernel's fast path (harness.cuh: lg2_approx_ftz, rsqrt_approx_ftz)
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(r1));
        const float t = __fmul_rn(-1.3862943611198906f, l2);
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(t));
        const float xf = __fmul_rn(t, rs);
        __sincosf(y, &sf, &cf);
        const float e0 = fabsf(4.0f * sigma * (xp * sp - xf * sf)), e1 = fabsf(4.0f * sigma * (xp * cp - xf * cf));
        const float e = fmaxf(e0, e1);
        worst = fmaxf(worst, e);
        int b = e > 0.0f ? min(31, max(0, 40 + (int)floorf(log2f(e)))) : 0;   // bucket 40 + log2(e)
        if ((i & 1023) == 0) atomicAdd(&hist[b], 1ull);
    }
    atomicMax((int*)max_err, __float_as_int(worst));
}

int main() {
    float* d_max; unsigned long long* d_hist;
    cudaMalloc(&d_max, 4); cudaMalloc(&d_hist, 32 * 8);
    for (float sigma : {0.5f, 1.0f, 2.0f}) {
        for (float cut : {1.0f, 0.999f}) {
            cudaMemset(d_max, 0, 4); cudaMemset(d_hist, 0, 32 * 8);
            probe<<<148 * 8, 256>>>(1ull << 18, sigma, cut, d_max, d_hist);   // 148 * 8 * 256 * 2^18 = 7.9e10 draws
            float m; unsigned long long h[32];
            cudaMemcpy(&m, d_max, 4, cudaMemcpyDeviceToHost); cudaMemcpy(h, d_hist, sizeof h, cudaMemcpyDeviceToHost);
            printf("sigma %.2f r1 <= %.3f: max error %.3e quantiser steps (2^%.1f);", sigma, cut, m, log2f(m));
            printf(" sampled histogram of log2(error):");
            for (int b = 10; b < 32; b++) if (h[b]) printf(" %d:%llu", b - 40, h[b]);
            printf("\n");
        }
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
