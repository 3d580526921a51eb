// count.cuh -- error counters of the Monte-Carlo loop and the polar transform with the frame held on chip
// (sc_error_counter.h:68-125 for the codeword-bit counters; information bits through u^ = x^ F^(x)n).
// Split from harness.cuh so that the test emulator (tests/emu/warp_emu.cpp) can execute these kernels on the CPU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace scpd {

// ---------------------------------------------------------------- the same transform with the frame in registers
// One warp per frame, word lane + 32 j of the frame in v[j] (32 <= n <= 32 * 32 * WPL; lanes >= wpf hold zeros when
// wpf < 32).  Strides inside a word are shifts, word strides below 32 are lane exchanges, those from 32 up pair
// registers of the same lane: no pass over memory between the strides.
template <int WPL>
__device__ __forceinline__ void transform_regs(uint32_t (&v)[WPL], uint32_t wpf, int lane) {
#pragma unroll
    for (int j = 0; j < WPL; j++) {
        uint32_t x = v[j];
        x ^= (x >> 1) & 0x55555555u;
        x ^= (x >> 2) & 0x33333333u;
        x ^= (x >> 4) & 0x0F0F0F0Fu;
        x ^= (x >> 8) & 0x00FF00FFu;
        x ^= (x >> 16) & 0x0000FFFFu;
        v[j] = x;
    }
#pragma unroll
    for (int h = 1; h < 32; h <<= 1) {
        if ((uint32_t)h < wpf) {  // warp-uniform
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                const uint32_t t = __shfl_xor_sync(0xFFFFFFFFu, v[j], h);
                if (!(lane & h)) v[j] ^= t;
            }
        }
    }
#pragma unroll
    for (int hj = 1; hj < WPL; hj <<= 1)
#pragma unroll
        for (int j = 0; j < WPL; j++)
            if (!(j & hj)) v[j] ^= v[j + hj];
}

template <int WPL>
__global__ void __launch_bounds__(256)
polar_transform_reg_kernel(uint32_t wpf, unsigned long long nframes, const uint32_t* __restrict__ xhat, uint32_t* __restrict__ uhat) {
    const int lane = threadIdx.x & 31;
    const unsigned long long wstride = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    for (unsigned long long f = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < nframes;
         f += wstride) {
        uint32_t v[WPL];
#pragma unroll
        for (int j = 0; j < WPL; j++) v[j] = (lane + 32u * j < wpf) ? xhat[f * wpf + lane + 32u * j] : 0u;
        transform_regs<WPL>(v, wpf, lane);
#pragma unroll
        for (int j = 0; j < WPL; j++)
            if (lane + 32u * j < wpf) uhat[f * wpf + lane + 32u * j] = v[j];
    }
}

// ---------------------------------------------------------------- all ten counters of the Monte-Carlo loop in one pass
// count_errors_kernel (counters 0..5) and polar transform + count_info_kernel (counters 6..9) on d = x^ ^ x held in
// registers: x^ is read once and nothing is written.  ref: nullptr (all-zero codeword), one shared codeword, or one per
// frame.  The two per-frame sums travel through the warp reduction in one register (each is at most n <= 32768).
template <int WPL>
__global__ void __launch_bounds__(256)
count_all_kernel(uint32_t wpf, uint32_t n, uint32_t k, unsigned long long nframes, const uint32_t* __restrict__ xhat,
                 const uint32_t* __restrict__ ref, int per_frame, const uint32_t* __restrict__ info_mask,
                 unsigned long long* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const unsigned long long wstride = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    uint32_t m[WPL], r[WPL];
#pragma unroll
    for (int j = 0; j < WPL; j++) {
        const bool in = lane + 32u * j < wpf;
        m[j] = in ? info_mask[lane + 32u * j] : 0u;
        r[j] = (in && ref && !per_frame) ? ref[lane + 32u * j] : 0u;
    }
    unsigned long long be = 0, fe = 0, bew = 0, few = 0, bi = 0, fi = 0, nf = 0;
    for (unsigned long long f = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < nframes;
         f += wstride) {
        uint32_t v[WPL];
        uint32_t e = 0;
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            const bool in = lane + 32u * j < wpf;
            uint32_t x = in ? xhat[f * wpf + lane + 32u * j] : 0u;
            x ^= (in && per_frame) ? ref[f * wpf + lane + 32u * j] : r[j];
            v[j] = x;
            e += __popc(x);
        }
        transform_regs<WPL>(v, wpf, lane);
        uint32_t ei = 0;
#pragma unroll
        for (int j = 0; j < WPL; j++) ei += __popc(v[j] & m[j]);
        uint32_t both = e | (ei << 16);  // either sum is at most 32768 < 2^16
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) both += __shfl_xor_sync(0xFFFFFFFFu, both, off);
        e = both & 0xFFFFu;
        ei = both >> 16;
        be += e;
        fe += (e != 0);
        bew += (e & 1023u);  // sc_uint<10> err (sc_error_counter.h:70-71)
        few += ((e & 1023u) != 0);
        bi += ei;
        fi += (ei != 0);
        nf += 1;
    }
    if (lane == 0 && nf) {
        atomicAdd(counters + 0, be);
        atomicAdd(counters + 1, fe);
        atomicAdd(counters + 2, nf * n);
        atomicAdd(counters + 3, nf);
        atomicAdd(counters + 4, bew);
        atomicAdd(counters + 5, few);
        atomicAdd(counters + 6, bi);
        atomicAdd(counters + 7, fi);
        atomicAdd(counters + 8, nf * k);
        atomicAdd(counters + 9, nf);
    }
}

// ---------------------------------------------------------------- the same two kernels for frames beyond 32768 bits
// One CTA per frame, the frame in shared memory (wpf words: 64 KB at N = 2^19).  A warp reads 32 consecutive words, so the
// strides inside a word and the word strides below 32 happen in registers on the way in; the strides from 32 up are
// passes over shared memory (pairs (w, w + h), one __syncthreads each).  Requires wpf >= 32 (a multiple of 32).
__device__ __forceinline__ uint32_t transform_word_and_lanes(uint32_t x, int lane) {
    x ^= (x >> 1) & 0x55555555u;
    x ^= (x >> 2) & 0x33333333u;
    x ^= (x >> 4) & 0x0F0F0F0Fu;
    x ^= (x >> 8) & 0x00FF00FFu;
    x ^= (x >> 16) & 0x0000FFFFu;
#pragma unroll
    for (int h = 1; h < 32; h <<= 1) {
        const uint32_t t = __shfl_xor_sync(0xFFFFFFFFu, x, h);
        if (!(lane & h)) x ^= t;
    }
    return x;
}
__device__ __forceinline__ void transform_smem_strides(uint32_t* s, uint32_t wpf) {
    for (uint32_t h = 32; h < wpf; h <<= 1) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < wpf / 2; i += blockDim.x) {
            const uint32_t w = ((i & ~(h - 1u)) << 1) | (i & (h - 1u));
            s[w] ^= s[w + h];
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
polar_transform_smem_kernel(uint32_t wpf, unsigned long long nframes, const uint32_t* __restrict__ xhat, uint32_t* __restrict__ uhat) {
    extern __shared__ uint32_t s_frame[];
    const int lane = threadIdx.x & 31;
    for (unsigned long long f = blockIdx.x; f < nframes; f += gridDim.x) {
        for (uint32_t w = threadIdx.x; w < wpf; w += blockDim.x) s_frame[w] = transform_word_and_lanes(xhat[f * wpf + w], lane);
        transform_smem_strides(s_frame, wpf);
        for (uint32_t w = threadIdx.x; w < wpf; w += blockDim.x) uhat[f * wpf + w] = s_frame[w];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
count_all_smem_kernel(uint32_t wpf, uint32_t n, uint32_t k, unsigned long long nframes, const uint32_t* __restrict__ xhat,
                      const uint32_t* __restrict__ ref, int per_frame, const uint32_t* __restrict__ info_mask,
                      unsigned long long* __restrict__ counters) {
    extern __shared__ uint32_t s_frame[];  // wpf words of the frame, then the two per-frame sums
    unsigned int& s_e = s_frame[wpf];
    unsigned int& s_ei = s_frame[wpf + 1];
    const int lane = threadIdx.x & 31;
    unsigned long long be = 0, fe = 0, bew = 0, few = 0, bi = 0, fi = 0, nf = 0;  // thread 0's
    for (unsigned long long f = blockIdx.x; f < nframes; f += gridDim.x) {
        if (threadIdx.x == 0) s_e = s_ei = 0;
        uint32_t e = 0;
        for (uint32_t w = threadIdx.x; w < wpf; w += blockDim.x) {
            uint32_t x = xhat[f * wpf + w];
            if (ref) x ^= per_frame ? ref[f * wpf + w] : ref[w];
            e += __popc(x);
            s_frame[w] = transform_word_and_lanes(x, lane);
        }
        transform_smem_strides(s_frame, wpf);  // its first barrier also orders the reset of s_e / s_ei
        uint32_t ei = 0;
        for (uint32_t w = threadIdx.x; w < wpf; w += blockDim.x) ei += __popc(s_frame[w] & info_mask[w]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            e += __shfl_xor_sync(0xFFFFFFFFu, e, off);
            ei += __shfl_xor_sync(0xFFFFFFFFu, ei, off);
        }
        if (lane == 0) {
            atomicAdd(&s_e, e);
            atomicAdd(&s_ei, ei);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t te = s_e, ti = s_ei;
            be += te;
            fe += (te != 0);
            bew += (te & 1023u);  // sc_uint<10> err (sc_error_counter.h:70-71)
            few += ((te & 1023u) != 0);
            bi += ti;
            fi += (ti != 0);
            nf += 1;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && nf) {
        atomicAdd(counters + 0, be);
        atomicAdd(counters + 1, fe);
        atomicAdd(counters + 2, nf * n);
        atomicAdd(counters + 3, nf);
        atomicAdd(counters + 4, bew);
        atomicAdd(counters + 5, few);
        atomicAdd(counters + 6, bi);
        atomicAdd(counters + 7, fi);
        atomicAdd(counters + 8, nf * k);
        atomicAdd(counters + 9, nf);
    }
}

}  // namespace scpd
