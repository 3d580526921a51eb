// int_peak_probe.cu -- measures sustained per-SM issue rates (lane-instructions / clk / SM) of the
// packed-integer and fp16x2 instructions the SC decode kernels are built from.  Denominator of the
// integer-ALU roofline (SURVEY.md 8d / BASELINE.md 2).  Build: nvcc -arch=sm_100a -O3 -o probe ...
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <string>

#define ILP 8
#define ITERS 4096

template <int OP>
__device__ __forceinline__ uint32_t step(uint32_t x, uint32_t y, uint32_t z) {
    if (OP == 0) return __vmaxs2(x, y);                       // VIMNMX.S16x2
    if (OP == 1) return __vadd2(x, y);                        // VIADD.16x2
    if (OP == 2) return __viaddmax_s16x2(x, y, z);            // VIADDMNMX.S16x2
    if (OP == 3) return (x ^ y) & z;                          // LOP3
    if (OP == 4) return __byte_perm(x, y, 0x5140);            // PRMT
    if (OP == 5) return x * y + z;                            // IMAD
    if (OP == 6) return x + y + z;                            // IADD3
    if (OP == 7) { __half2 a = *(__half2*)&x, b = *(__half2*)&y, c = *(__half2*)&z; __half2 r = __hfma2(a, b, c); return *(uint32_t*)&r; }  // HFMA2
    if (OP == 8) { __half2 a = *(__half2*)&x, b = *(__half2*)&y; __half2 r = __hmax2(a, b); return *(uint32_t*)&r; }  // HMNMX2
    if (OP == 9) { __half2 a = *(__half2*)&x, b = *(__half2*)&y; __half2 r = __hadd2(a, b); return *(uint32_t*)&r; }  // HADD2
    if (OP == 10) return __vimax3_s16x2(x, y, z);             // VIMNMX3
    if (OP == 11) return __shfl_xor_sync(0xFFFFFFFFu, x, 1) + y;  // SHFL (+IADD)
    if (OP == 12) return __vsub2(__viaddmax_s16x2(x, y, 0u), __vmaxs2(x, y));  // f (3 instr)
    if (OP == 13) { uint32_t an = __vsub2(x ^ z, z); return __vmaxs2(__viaddmin_s16x2(y, an, 0x007F007Fu), 0xFF81FF81u); }  // g sat (4 instr)
    if (OP == 14) return __vabsdiffs2(x, y);
    if (OP == 15) return (x >> 3) ^ y;                        // SHF + LOP
    if (OP == 16) return __vmins4(x, y);                      // emulated byte SIMD
    if (OP == 17) return __popc(x) + y;
    return x;
}

template <int OP>
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed) {
    uint32_t v[ILP];
    uint32_t y = seed * 0x9E3779B9u + threadIdx.x, z = seed ^ 0x00FF00FFu;
#pragma unroll
    for (int k = 0; k < ILP; k++) v[k] = seed + k * 0x01010101u + threadIdx.x;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) v[k] = step<OP>(v[k], y, z);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) acc ^= v[k];
    if (acc == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// shared-memory bandwidth: each lane LDS.32 / LDS.64 / LDS.128 in a loop
template <int VEC>
__global__ void __launch_bounds__(256) probe_lds(uint32_t* out, int iters) {
    __shared__ uint4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    uint32_t acc = 0;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; it++) {
        if (VEC == 1) acc += ((uint32_t*)buf)[(idx + it * 32) & 4095];
        if (VEC == 2) { uint2 t = ((uint2*)buf)[(idx + it * 32) & 2047]; acc += t.x ^ t.y; }
        if (VEC == 4) { uint4 t = buf[(idx + it * 32) & 1023]; acc += t.x ^ t.y ^ t.z ^ t.w; }
    }
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

struct Res { const char* name; double lanes_per_clk_sm; };

template <int OP>
static double run(const char* name, int instr_per_step, int sms, double clk_hz) {
    uint32_t* d; cudaMalloc(&d, 1 << 20);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = sms * 8, threads = 256;
    probe<OP><<<blocks, threads>>>(d, 1); cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe<OP><<<blocks, threads>>>(d, 2);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)blocks * threads * ILP * ITERS * instr_per_step;
    double per_clk_sm = ops / (ms * 1e-3) / clk_hz / sms;
    printf("%-28s %8.3f ms  %7.1f lane-instr/clk/SM  (%d instr/step)  %.3e lane-instr/s chip\n", name, ms, per_clk_sm,
           instr_per_step, ops / (ms * 1e-3));
    cudaFree(d);
    return per_clk_sm;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double clk = khz * 1e3;
    printf("device %s  SMs %d  clock %.0f MHz (nominal max; rates below assume it)\n", p.name, sms, clk / 1e6);
    run<0>("VIMNMX.S16x2 (vmaxs2)", 1, sms, clk);
    run<1>("VIADD.16x2 (vadd2)", 1, sms, clk);
    run<2>("VIADDMNMX.S16x2", 1, sms, clk);
    run<10>("VIMNMX3.S16x2", 1, sms, clk);
    run<3>("LOP3", 1, sms, clk);
    run<4>("PRMT", 1, sms, clk);
    run<5>("IMAD", 1, sms, clk);
    run<6>("IADD3", 1, sms, clk);
    run<7>("HFMA2", 1, sms, clk);
    run<8>("HMNMX2 (hmax2)", 1, sms, clk);
    run<9>("HADD2", 1, sms, clk);
    run<11>("SHFL+IADD", 2, sms, clk);
    run<12>("f = viaddmax,vmax,vsub", 3, sms, clk);
    run<13>("g = lop,vsub,viaddmin,vmax", 4, sms, clk);
    run<14>("vabsdiffs2", 1, sms, clk);
    run<15>("SHF+LOP3", 2, sms, clk);
    run<16>("vmins4 (emulated)", 1, sms, clk);
    run<17>("POPC+IADD", 2, sms, clk);
    // shared memory
    {
        uint32_t* d; cudaMalloc(&d, 4096);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        const int iters = 1 << 14;
        for (int vec : {1, 2, 4}) {
            float ms;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(a);
                if (vec == 1) probe_lds<1><<<sms * 8, 256>>>(d, iters);
                if (vec == 2) probe_lds<2><<<sms * 8, 256>>>(d, iters);
                if (vec == 4) probe_lds<4><<<sms * 8, 256>>>(d, iters);
                cudaEventRecord(b); cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms, a, b);
            }
            double bytes = (double)sms * 8 * 256 * iters * 4 * vec;
            printf("LDS.%-3d                      %8.3f ms  %7.1f B/clk/SM\n", 32 * vec, ms, bytes / (ms * 1e-3) / clk / sms);
        }
    }
    return 0;
}
