// bs_arith.cuh -- bit-sliced LLR arithmetic: one 32-bit register holds ONE BIT of the same LLR for
// 32 different frames ("a frame group"), so every LOP3 advances 32 frames at once and the Q-bit
// quantisation of the reference maps to Q bit planes instead of wasting a 16-bit SIMD half.
//
// A value is kept in sign-magnitude planes: s = sign plane, m[0..P) = magnitude planes, LSB first.
//   * SIGMAG builds (shared/src/functions.h:124-281, scalar.h:88-239) use exactly this form, -0 included.
//   * CA2 builds (functions.h:48-118, scalar.h:9-76) are two's complement in the reference; the same
//     numbers are represented here as (sign, |v|).  A zero may carry either sign internally; the
//     places where CA2 looks at the sign of a zero (hard decisions: hd(0) = 0) mask it with "m != 0".
//     The value -2^(Q-1) never occurs (quantiser +-31, symmetric saturation).
//
// Plain C++ (host + device) so that tests can run it on the CPU against the oracle.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BS_FN __host__ __device__ __forceinline__
#else
#define BS_FN inline
#endif

namespace scpd {
namespace bs {

enum : int { FMT_CA2 = 0, FMT_SM = 1 };

// One LOP3: bit k of LUT is the output for inputs (a, b, c) = (k >> 2 & 1, k >> 1 & 1, k & 1), i.e. LUT is the
// expression evaluated on a = 0xF0, b = 0xCC, c = 0xAA.  Spelling the carry / select steps of the plane
// arithmetic as single LOP3s matters: left to the compiler the saturating g came out at 63 LOP3 instead of 37.
template <uint32_t LUT>
BS_FN uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
#else
    uint32_t r = 0;
    for (int k = 0; k < 8; k++)
        if ((LUT >> k) & 1u) r |= ((k & 4) ? a : ~a) & ((k & 2) ? b : ~b) & ((k & 1) ? c : ~c);
    return r;
#endif
}
constexpr uint32_t LA = 0xF0u, LB = 0xCCu, LC = 0xAAu;  // to write LUTs as expressions

template <int P>
struct Val {
    uint32_t s;
    uint32_t m[P];
};

template <int P>
BS_FN uint32_t nonzero(const Val<P>& a) {
    uint32_t r = a.m[0];
#pragma unroll
    for (int i = 1; i < P; i++) r |= a.m[i];
    return r;
}

// widen with zero planes
template <int PO, int P>
BS_FN void widen(const Val<P>& a, Val<PO>& r) {
    static_assert(PO >= P, "widen");
    r.s = a.s;
#pragma unroll
    for (int i = 0; i < PO; i++) r.m[i] = (i < P) ? a.m[i] : 0u;
}

// f = (sa ^ sb, min(ma, mb))            F_function_C2 functions.h:48-61, F_function_SM :124-145
template <int P>
BS_FN void f_op(const Val<P>& a, const Val<P>& b, Val<P>& r) {
    uint32_t lt = ~a.m[0] & b.m[0];  // ma < mb, rippled from the LSB
#pragma unroll
    for (int i = 1; i < P; i++) lt = lop3<((~LA & LB) | (~(LA ^ LB) & LC)) & 0xFFu>(a.m[i], b.m[i], lt);
#pragma unroll
    for (int i = 0; i < P; i++) r.m[i] = lop3<((LA & LC) | (LB & ~LC)) & 0xFFu>(a.m[i], b.m[i], lt);
    r.s = a.s ^ b.s;
}

// First pass of b + (-1)^u a on magnitudes: D = signs differ (after applying u) -> mb + ~ma
// (ones' complement subtraction), else mb + ma.
template <int P>
BS_FN void addsub_pass1(const Val<P>& a, const Val<P>& b, uint32_t D, uint32_t (&T)[P], uint32_t& cout) {
    T[0] = lop3<(LA ^ LB ^ LC) & 0xFFu>(a.m[0], b.m[0], D);
    uint32_t c = ~T[0] & b.m[0];
#pragma unroll
    for (int i = 1; i < P; i++) {
        const uint32_t t = lop3<(LA ^ LB ^ LC) & 0xFFu>(a.m[i], b.m[i], D);
        T[i] = t ^ c;
        c = lop3<((LA & LB) | (~LA & LC)) & 0xFFu>(t, c, b.m[i]);  // t ? c : b
    }
    cout = c;
}

// g, un-saturated: magnitude grows by one plane.   G_extended_C2 functions.h:77-88, G_extended_SM :197-239
// Tie (equal magnitudes, opposite signs) gives (sign of the a-term, 0) as qfull_add_sub_sm does (scalar.h:196-225).
template <int P>
BS_FN void g_ext(const Val<P>& a, const Val<P>& b, uint32_t u, Val<P + 1>& r) {
    const uint32_t D = lop3<(LA ^ LB ^ LC) & 0xFFu>(a.s, u, b.s);
    uint32_t T[P], cout;
    addsub_pass1<P>(a, b, D, T, cout);
    uint32_t c2 = cout & D;         // end-around carry when mb > ma
    const uint32_t K = D & ~cout;   // mb <= ma: result is ~T
#pragma unroll
    for (int i = 0; i < P; i++) {
        r.m[i] = lop3<(LA ^ LB ^ LC) & 0xFFu>(T[i], c2, K);
        c2 &= T[i];
    }
    r.m[P] = cout & ~D;
    r.s = b.s ^ K;
}

// g saturated to +-(2^P - 1), CA2.                 G_function_C2 functions.h:63-75, qsat scalar.h:15-21
template <int P>
BS_FN void g_sat_ca2(const Val<P>& a, const Val<P>& b, uint32_t u, Val<P>& r) {
    const uint32_t D = lop3<(LA ^ LB ^ LC) & 0xFFu>(a.s, u, b.s);
    uint32_t T[P], cout;
    addsub_pass1<P>(a, b, D, T, cout);
    const uint32_t K = D ^ cout;     // D & ~cout: invert;  ~D & cout: overflow -> all ones
    const uint32_t ovf = ~D & cout;
    uint32_t c2 = cout;
#pragma unroll
    for (int i = 0; i < P; i++) {
        r.m[i] = lop3<((LC & (~LA | LB)) | (~LC & (LA ^ LB))) & 0xFFu>(T[i], c2, K);
        if (i + 1 < P) c2 = lop3<((LA & LB) | LC) & 0xFFu>(T[i], c2, ovf);
    }
    r.s = lop3<(LA ^ (LB & ~LC)) & 0xFFu>(b.s, D, cout);
}

// g with the magnitude clamped to 2^(P-1) - 1, SIGMAG (half range: qsat_sm<Q-1>, scalar.h:94-99;
// functions.h:186-191).  The top magnitude plane of the result is always 0.
template <int P>
BS_FN void g_sat_sm(const Val<P>& a, const Val<P>& b, uint32_t u, Val<P>& r) {
    Val<P + 1> e;
    g_ext<P>(a, b, u, e);
    const uint32_t ovf = e.m[P] | e.m[P - 1];
#pragma unroll
    for (int i = 0; i < P - 1; i++) r.m[i] = e.m[i] | ovf;
    r.m[P - 1] = 0u;
    r.s = e.s;
}

template <int FMT, int P>
BS_FN void g_sat(const Val<P>& a, const Val<P>& b, uint32_t u, Val<P>& r) {
    if (FMT == FMT_CA2)
        g_sat_ca2<P>(a, b, u, r);
    else
        g_sat_sm<P>(a, b, u, r);
}

// hard decision: sign bit; a CA2 zero decides 0
template <int FMT, int P>
BS_FN uint32_t hd(const Val<P>& a) {
    return FMT == FMT_CA2 ? (a.s & nonzero<P>(a)) : a.s;
}

// Spec_P2 (functions.h:367-384) with F_simplified / G_simplified (:90-118 CA2, :241-281 SIGMAG):
// the two partial sums (x0, x1) = (u0 ^ u1, u1) of a 2-bit node with information flags lf = f0 | f1 << 1.
template <int FMT, int P>
BS_FN void p2_op(const Val<P>& a, const Val<P>& b, uint32_t lf, uint32_t& x0, uint32_t& x1) {
    if (lf == 0u) {
        x0 = 0u;
        x1 = 0u;
    } else if (lf == 3u) {  // (u0 ^ u1, u1) = (hd(a), hd(b)) for every input
        x0 = hd<FMT, P>(a);
        x1 = hd<FMT, P>(b);
    } else if (lf == 1u) {  // only u0 is information
        x0 = hd<FMT, P>(a) ^ hd<FMT, P>(b);
        x1 = 0u;
    } else {  // only u1: sign of the exact sum b + a (CA2: sum 0 -> 0); SIGMAG: larger magnitude, tie -> a
        uint32_t lt = ~a.m[0] & b.m[0], ne = a.m[0] ^ b.m[0];
#pragma unroll
        for (int i = 1; i < P; i++) {
            lt = lop3<((~LA & LB) | (~(LA ^ LB) & LC)) & 0xFFu>(a.m[i], b.m[i], lt);
            ne = lop3<((LA ^ LB) | LC) & 0xFFu>(a.m[i], b.m[i], ne);
        }
        uint32_t u1;
        if (FMT == FMT_CA2) {
            const uint32_t sa = a.s & nonzero<P>(a), sb = b.s & nonzero<P>(b);
            // same sign: that sign; opposite: sign of the larger magnitude, 0 on a tie
            u1 = (~(sa ^ sb) & sa) | ((sa ^ sb) & ne & ((lt & sb) | (~lt & sa)));
        } else {
            u1 = (lt & b.s) | (~lt & a.s);
        }
        x0 = u1;
        x1 = u1;
    }
}

// two's complement planes v[0..8) of an int8 LLR -> (sign, |v|) with P magnitude planes
// (wrapper_in.h:34 / qconv_format scalar.h:229-239 for SIGMAG; |v| <= 2^P - 1 by the input contract)
template <int P>
BS_FN void from_int8_planes(const uint32_t (&v)[8], Val<P>& r) {
    const uint32_t s = v[7];
    uint32_t c = s;
#pragma unroll
    for (int i = 0; i < P; i++) {
        r.m[i] = lop3<(LA ^ LB ^ LC) & 0xFFu>(v[i], s, c);
        c = lop3<((LA ^ LB) & LC) & 0xFFu>(v[i], s, c);
    }
    r.s = s;
}

BS_FN uint32_t bperm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((src >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
#endif
}

// 32 x 32 bit-matrix transpose in registers: out[b] bit f = in[f] bit b.
BS_FN void transpose32(uint32_t (&a)[32]) {
#pragma unroll
    for (int k = 0; k < 16; k++) {  // 16-bit blocks
        const uint32_t lo = bperm(a[k], a[k + 16], 0x5410), hi = bperm(a[k], a[k + 16], 0x7632);
        a[k] = lo;
        a[k + 16] = hi;
    }
#pragma unroll
    for (int k0 = 0; k0 < 32; k0 += 16)
#pragma unroll
        for (int k = k0; k < k0 + 8; k++) {  // bytes
            const uint32_t lo = bperm(a[k], a[k + 8], 0x6240), hi = bperm(a[k], a[k + 8], 0x7351);
            a[k] = lo;
            a[k + 8] = hi;
        }
#pragma unroll
    for (int j = 4; j >= 1; j >>= 1) {
        const uint32_t m = (j == 4) ? 0x0F0F0F0Fu : (j == 2) ? 0x33333333u : 0x55555555u;
#pragma unroll
        for (int k = 0; k < 32; k++) {
            if (k & j) continue;
            const uint32_t x = a[k], y = a[k + j];
            a[k] = (x & m) | ((y << j) & ~m);
            a[k + j] = ((x >> j) & m) | (y & ~m);
        }
    }
}

}  // namespace bs
}  // namespace scpd
