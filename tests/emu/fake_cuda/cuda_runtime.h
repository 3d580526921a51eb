// fake cuda_runtime.h -- TEST INFRASTRUCTURE.  Lets the .cuh kernel sources of the product compile
// as plain C++ so that tests/emu/warp_emu.cpp can execute them on the CPU, one fiber per lane,
// with exact emulation of the warp collectives and of the packed-SIMD / PRMT intrinsics they use.
// It exists to debug and regression-test kernel logic without a GPU; it is never a decode path.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))
#define __shared__
#define __restrict__

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct char2 { signed char x, y; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline char2 make_char2(signed char x, signed char y) { return char2{x, y}; }

namespace cuda_emu {
struct LaneCtx { dim3 tid, bid, bdim, gdim; int lane; };
extern thread_local LaneCtx* cur;
uint32_t collective_exchange(uint32_t v, int src_lane);  // returns value contributed by src_lane
uint32_t collective_ballot(bool pred);
void collective_sync();
int cta_barrier_or(int pred);  // __syncthreads_or over the emulated CTA
}  // namespace cuda_emu
#define threadIdx (cuda_emu::cur->tid)
#define blockIdx (cuda_emu::cur->bid)
#define blockDim (cuda_emu::cur->bdim)
#define gridDim (cuda_emu::cur->gdim)

template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int __ffs(uint32_t v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline uint32_t __brev(uint32_t v) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
// the fibers of an emulated CTA run one at a time and switch only inside collectives: an atomic is a plain update
template <class T, class U>
static inline T atomicAdd(T* p, U v) { const T old = *p; *p = (T)(old + (T)v); return old; }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { cuda_emu::collective_sync(); }
// CTA barrier over the fibers of run_cta (kernels whose warps are emulated one after the other through run_warp
// see a CTA of one warp and must not rely on cross-warp state under emulation)
static inline void __syncthreads() { (void)cuda_emu::cta_barrier_or(0); }
static inline int __syncthreads_or(int pred) { return cuda_emu::cta_barrier_or(pred); }
static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src, int width = 32) {
    const int lane = cuda_emu::cur->lane;
    return cuda_emu::collective_exchange(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int m, int width = 32) {
    (void)width;
    return cuda_emu::collective_exchange(v, cuda_emu::cur->lane ^ m);
}
static inline uint32_t __ballot_sync(unsigned, int pred) { return cuda_emu::collective_ballot(pred != 0); }
static inline int __any_sync(unsigned, int pred) { return cuda_emu::collective_ballot(pred != 0) != 0; }

// prmt.b32 with sign-replicate selector bit, as the hardware does it
static inline uint32_t cuda_emu_prmt(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t src = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t b = (uint32_t)(src >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) b = (b & 0x80) ? 0xFF : 0x00;
        r |= b << (8 * i);
    }
    return r;
}
// the __byte_perm intrinsic only honours 3 selector bits per byte (nvcc masks with 0x7777)
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) { return cuda_emu_prmt(x, y, s & 0x7777u); }
static inline int16_t emu_lo(uint32_t v) { return (int16_t)(v & 0xFFFF); }
static inline int16_t emu_hi(uint32_t v) { return (int16_t)(v >> 16); }
static inline uint32_t emu_pk(int lo, int hi) { return ((uint32_t)lo & 0xFFFF) | ((uint32_t)hi << 16); }
static inline uint32_t __vmins2(uint32_t a, uint32_t b) { return emu_pk(std::min(emu_lo(a), emu_lo(b)), std::min(emu_hi(a), emu_hi(b))); }
static inline uint32_t __vmaxs2(uint32_t a, uint32_t b) { return emu_pk(std::max(emu_lo(a), emu_lo(b)), std::max(emu_hi(a), emu_hi(b))); }
static inline uint32_t __vadd2(uint32_t a, uint32_t b) { return emu_pk(emu_lo(a) + emu_lo(b), emu_hi(a) + emu_hi(b)); }
static inline uint32_t __vsub2(uint32_t a, uint32_t b) { return emu_pk(emu_lo(a) - emu_lo(b), emu_hi(a) - emu_hi(b)); }
static inline uint32_t __viaddmax_s16x2(uint32_t a, uint32_t b, uint32_t c) {
    return emu_pk(std::max<int>((int16_t)(emu_lo(a) + emu_lo(b)), emu_lo(c)), std::max<int>((int16_t)(emu_hi(a) + emu_hi(b)), emu_hi(c)));
}
static inline uint32_t __viaddmin_s16x2(uint32_t a, uint32_t b, uint32_t c) {
    return emu_pk(std::min<int>((int16_t)(emu_lo(a) + emu_lo(b)), emu_lo(c)), std::min<int>((int16_t)(emu_hi(a) + emu_hi(b)), emu_hi(c)));
}
