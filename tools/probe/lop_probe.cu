// lop_probe.cu -- issue rates of the bitwise/integer instructions a bit-sliced decoder is built from
// (inline PTX so that nothing is folded).  lane-instr / clk / SM at the nominal clock.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define ILP 8
#define ITERS 4096

template <int OP>
__device__ __forceinline__ void step(uint32_t (&v)[ILP]) {
#pragma unroll
    for (int k = 0; k < ILP; k++) {
        uint32_t& x = v[k];
        const uint32_t y = v[(k + 3) % ILP], z = v[(k + 5) % ILP];
        if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(y), "r"(z));
        if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z));
        if (OP == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
        if (OP == 3) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x) : "r"(y));
        if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(x) : "r"(y));
        if (OP == 5) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
        if (OP == 6) {  // LOP3 + IMAD alternating (two pipes?)
            if (k & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(y), "r"(z));
        }
        if (OP == 7) {  // LOP3 + VIMNMX alternating
            if (k & 1) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(y), "r"(z));
        }
        if (OP == 8) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
        if (OP == 9) asm volatile("shr.u32 %0, %0, 1;" : "+r"(x));
        if (OP == 10) {  // LOP3 + SHF alternating
            if (k & 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x) : "r"(y));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(y), "r"(z));
        }
        if (OP == 11) asm volatile("add.s16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
        if (OP == 12) {  // LOP3 + PRMT
            if (k & 1) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(x) : "r"(y));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(y), "r"(z));
        }
        if (OP == 13) asm volatile("mov.b32 %0, %1;" : "+r"(x) : "r"(y));
        if (OP == 14) asm volatile("sub.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
        // round 2: the fp16x2 ops of the slot-sliced kernel's 32-LLR walker, and whether they co-issue with LOP3
        if (OP == 15) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
        if (OP == 16) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
        if (OP == 17) asm volatile("mul.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
        if (OP == 18) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
        if (OP == 19) {  // LOP3 + HFMA2 alternating
            if (k & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(y), "r"(z));
        }
        if (OP == 20) {  // LOP3 + HADD2 alternating
            if (k & 1) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(y), "r"(z));
        }
        if (OP == 21) asm volatile("and.b32 %0, %0, %1;" : "+r"(x) : "r"(y));  // two-input logic op
        if (OP == 22) {  // 3 LOP3 : 1 HFMA2 (the kernel's mix is about 3 : 1)
            if ((k & 3) == 3) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x) : "r"(y), "r"(z));
        }
        if (OP == 23) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*reinterpret_cast<float*>(&x)) : "f"(__uint_as_float(y)), "f"(__uint_as_float(z)));
    }
}

template <int OP>
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed) {
    uint32_t v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) v[k] = seed * 0x9E3779B9u + k * 0x01010101u + threadIdx.x;
    for (int it = 0; it < ITERS; it++) step<OP>(v);
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) acc ^= v[k];
    if (acc == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
static void run(const char* name, int sms, double clk, int warps_per_sm) {
    uint32_t* d;
    cudaMalloc(&d, 1 << 22);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int threads = 32 * (warps_per_sm >= 8 ? 8 : warps_per_sm), blocks = sms * (warps_per_sm >= 8 ? warps_per_sm / 8 : 1);
    probe<OP><<<blocks, threads>>>(d, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe<OP><<<blocks, threads>>>(d, 2);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)blocks * threads * ILP * ITERS;
    printf("%-22s warps/SM %2d  %8.3f ms  %7.1f lane-instr/clk/SM\n", name, warps_per_sm, ms, ops / (ms * 1e-3) / clk / sms);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clk = khz * 1e3;
    printf("device %s SMs %d clock %.0f MHz\n", p.name, sms, clk / 1e6);
    for (int w : {4, 8, 16, 32, 64}) {
        run<0>("LOP3 (maj)", sms, clk, w);
    }
    run<1>("LOP3 (xor3)", sms, clk, 32);
    run<2>("IADD", sms, clk, 32);
    run<14>("ISUB", sms, clk, 32);
    run<3>("SHF.L.W", sms, clk, 32);
    run<9>("SHR imm", sms, clk, 32);
    run<4>("PRMT", sms, clk, 32);
    run<5>("IMAD", sms, clk, 32);
    run<8>("VIMNMX.S16x2", sms, clk, 32);
    run<11>("VIADD.16x2", sms, clk, 32);
    run<13>("MOV", sms, clk, 32);
    run<6>("LOP3+IMAD", sms, clk, 32);
    run<7>("LOP3+VIMNMX", sms, clk, 32);
    run<10>("LOP3+SHF", sms, clk, 32);
    run<12>("LOP3+PRMT", sms, clk, 32);
    run<21>("LOP (2 inputs)", sms, clk, 32);
    run<15>("HFMA2", sms, clk, 32);
    run<16>("HADD2", sms, clk, 32);
    run<17>("HMUL2", sms, clk, 32);
    run<18>("HMNMX2", sms, clk, 32);
    run<23>("FFMA", sms, clk, 32);
    run<19>("LOP3+HFMA2", sms, clk, 32);
    run<20>("LOP3+HADD2", sms, clk, 32);
    run<22>("3 LOP3 : 1 HFMA2", sms, clk, 32);
    return 0;
}
