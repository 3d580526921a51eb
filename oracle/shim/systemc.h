/*
 * systemc.h -- minimal stand-in for the Accellera SystemC datatypes and module macros, written
 * for this repository so that the reference's own headers
 *     shared/src/{scalar,vector,functions,library,sc_list_fct}.h, src/module/{my_module,wrapper_*}.h
 * compile UNMODIFIED, where they lie under /root/reference, with plain g++ (no SystemC / Vivado).
 *
 * TEST INFRASTRUCTURE ONLY: used by oracle/build_ref.py to produce oracle/_ref/*.so, the
 * "reference compiled natively" that pins oracle/sc_oracle.c.  Not part of the product.
 *
 * What is modelled (and nothing more):
 *   sc_bigint<W> / sc_biguint<W> / sc_bv<W> / sc_uint<W> / sc_int<W>
 *       exact W-bit values; construction/assignment truncates to W bits (modular), signed types
 *       sign-extend when widened; .range(h,l) (read/write), operator[] (read/write),
 *       operator, (concatenation, left operand = most significant), unary - ~ +,
 *       + - * / % & | ^ << >> and comparisons on values that fit 64 bits (evaluated exactly in
 *       int64, which is what SystemC's arbitrary-precision temporaries give for these widths),
 *       .to_int(), .or_reduce().
 *   sc_fifo_in/out, sc_in/out/sc_signal, SC_MODULE / SC_CTOR / SC_CTHREAD, wait():
 *       no simulation kernel.  FIFOs are plain queues; reading an empty FIFO throws
 *       sc_shim::fifo_empty, which is how a driver regains control from the module's
 *       `while(true)` thread bodies.  wait() calls an optional hook.
 */
#ifndef SC_SHIM_SYSTEMC_H
#define SC_SHIM_SYSTEMC_H

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iomanip>
#include <iostream>
#include <type_traits>

using namespace std;

#ifndef SC_SHIM_MAXBITS
#define SC_SHIM_MAXBITS 4160
#endif

namespace sc_shim {

typedef long long i64;
typedef unsigned long long u64;

struct fifo_empty {};
struct stop_request {};
extern void (*wait_hook)();

struct sc_tag {};
template <class T>
struct is_scx : std::is_base_of<sc_tag, typename std::decay<T>::type> {};
template <class T>
struct is_int : std::integral_constant<bool, std::is_integral<typename std::decay<T>::type>::value ||
                                                 std::is_enum<typename std::decay<T>::type>::value> {};
template <class T>
struct is_operand : std::integral_constant<bool, is_scx<T>::value || is_int<T>::value> {};

static inline u64 lowmask(int nbits) { return nbits >= 64 ? ~0ULL : ((1ULL << nbits) - 1ULL); }

/* ---- dynamic-width unsigned bit vector: result of .range() and of concatenation ---- */
struct dyn : sc_tag {
    enum { NW = (SC_SHIM_MAXBITS + 63) / 64 };
    int w;
    u64 d[NW];
    dyn() : w(0) {}
    explicit dyn(int width) : w(width) {
        if (width > SC_SHIM_MAXBITS) {
            fprintf(stderr, "sc_shim: width %d exceeds SC_SHIM_MAXBITS\n", width);
            abort();
        }
        for (int i = 0; i < nw(); i++) d[i] = 0;
    }
    dyn(const dyn& o) : w(o.w) {
        for (int i = 0; i < nw(); i++) d[i] = o.d[i];
    }
    dyn& operator=(const dyn& o) {
        w = o.w;
        for (int i = 0; i < nw(); i++) d[i] = o.d[i];
        return *this;
    }
    int nw() const { return (w + 63) >> 6; }
    bool bit(int i) const { return (i < w) && ((d[i >> 6] >> (i & 63)) & 1ULL); }
    void setbit(int i, bool v) {
        if (v)
            d[i >> 6] |= 1ULL << (i & 63);
        else
            d[i >> 6] &= ~(1ULL << (i & 63));
    }
    i64 to_i64() const { return w == 0 ? 0 : (i64)(d[0] & lowmask(w)); }
    const dyn& to_dyn() const { return *this; }
    int to_int() const { return (int)to_i64(); }
    u64 to_uint64() const { return (u64)to_i64(); }
    explicit operator bool() const {
        for (int i = 0; i < nw(); i++)
            if (d[i]) return true;
        return false;
    }
    i64 operator~() const { return ~to_i64(); }
    i64 operator-() const { return -to_i64(); }
};

/* bits [lo, lo+len) of a word array of nsrc bits -> dst (len bits, zero padded) */
static inline void extract_bits(const u64* src, int nsrc_words, int lo, int len, u64* dst) {
    int nd = (len + 63) >> 6;
    int ws = lo >> 6, bs = lo & 63;
    for (int i = 0; i < nd; i++) {
        u64 v = 0;
        int s = ws + i;
        if (s < nsrc_words) v = src[s] >> bs;
        if (bs && s + 1 < nsrc_words) v |= src[s + 1] << (64 - bs);
        dst[i] = v;
    }
    if (len & 63) dst[nd - 1] &= lowmask(len & 63);
}
/* write the low len bits of src into dst at bit offset lo */
static inline void deposit_bits(u64* dst, int lo, int len, const u64* src, int nsrc_words) {
    for (int i = 0; i < len; i++) {
        int s = i >> 6;
        bool v = (s < nsrc_words) && ((src[s] >> (i & 63)) & 1ULL);
        int p = lo + i;
        if (v)
            dst[p >> 6] |= 1ULL << (p & 63);
        else
            dst[p >> 6] &= ~(1ULL << (p & 63));
    }
}

template <int W, bool S>
struct sc_val;

/* ---- proxies ---- */
template <int W, bool S>
struct bitref : sc_tag {
    sc_val<W, S>* p;
    int i;
    bitref(sc_val<W, S>* p_, int i_) : p(p_), i(i_) {}
    bool get() const { return p->bit(i); }
    i64 to_i64() const { return get() ? 1 : 0; }
    dyn to_dyn() const {
        dyn r(1);
        r.d[0] = get() ? 1 : 0;
        return r;
    }
    operator bool() const { return get(); }
    bitref& operator=(const bitref& o) {
        p->setbit(i, o.get());
        return *this;
    }
    template <class T, typename std::enable_if<is_operand<T>::value, int>::type = 0>
    bitref& operator=(const T& v);
    i64 operator~() const { return ~to_i64(); }
};

template <int W, bool S>
struct rangeref : sc_tag {
    sc_val<W, S>* p;
    int hi, lo;
    rangeref(sc_val<W, S>* p_, int h, int l) : p(p_), hi(h), lo(l) {}
    int len() const { return hi - lo + 1; }
    dyn to_dyn() const {
        dyn r(len());
        extract_bits(p->d, sc_val<W, S>::NW, lo, len(), r.d);
        return r;
    }
    i64 to_i64() const { return to_dyn().to_i64(); }
    int to_int() const { return (int)to_i64(); }
    rangeref& operator=(const rangeref& o) {
        dyn v = o.to_dyn();
        deposit_bits(p->d, lo, len(), v.d, v.nw());
        return *this;
    }
    template <class T, typename std::enable_if<is_operand<T>::value, int>::type = 0>
    rangeref& operator=(const T& v);
    i64 operator~() const { return ~to_i64(); }
    i64 operator-() const { return -to_i64(); }
    rangeref<W, S> range(int h, int l) const { return rangeref<W, S>(p, lo + h, lo + l); }
};

/* raw little-endian two's complement words of any operand, sign/zero filled to nwords */
template <class T>
static inline void raw_words(const T& v, u64* out, int nwords,
                             typename std::enable_if<is_int<T>::value, int>::type = 0) {
    typedef typename std::decay<T>::type D;
    bool neg = std::is_signed<D>::value && ((i64)v < 0);
    out[0] = (u64)(i64)v;
    if (!std::is_signed<D>::value) out[0] = (u64)v;
    for (int i = 1; i < nwords; i++) out[i] = neg ? ~0ULL : 0ULL;
}

template <int W, bool S>
struct sc_val : sc_tag {
    enum { NW = (W + 63) / 64, width = W };
    u64 d[NW];

    void trunc() {
        if (W & 63) d[NW - 1] &= lowmask(W & 63);
    }
    void set_from_dyn(const dyn& v) {
        for (int i = 0; i < NW; i++) d[i] = (i < v.nw()) ? v.d[i] : 0ULL;
        trunc();
    }
    template <int W2, bool S2>
    void set_from_val(const sc_val<W2, S2>& o) {
        bool neg = S2 && o.bit(W2 - 1);
        for (int i = 0; i < NW; i++) {
            u64 v;
            if (i < sc_val<W2, S2>::NW) {
                v = o.d[i];
                if (i == sc_val<W2, S2>::NW - 1 && neg && (W2 & 63)) v |= ~lowmask(W2 & 63);
            } else
                v = neg ? ~0ULL : 0ULL;
            d[i] = v;
        }
        trunc();
    }

    sc_val() {
        for (int i = 0; i < NW; i++) d[i] = 0;
    }
    template <class T, typename std::enable_if<is_int<T>::value, int>::type = 0>
    sc_val(T v) {
        raw_words(v, d, NW);
        trunc();
    }
    template <int W2, bool S2>
    sc_val(const sc_val<W2, S2>& o) {
        set_from_val(o);
    }
    sc_val(const dyn& v) { set_from_dyn(v); }
    template <int W2, bool S2>
    sc_val(const rangeref<W2, S2>& r) {
        set_from_dyn(r.to_dyn());
    }
    template <int W2, bool S2>
    sc_val(const bitref<W2, S2>& b) {
        for (int i = 0; i < NW; i++) d[i] = 0;
        d[0] = b.get() ? 1 : 0;
        trunc();
    }
    template <class T, typename std::enable_if<is_operand<T>::value, int>::type = 0>
    sc_val& operator=(const T& v) {
        *this = sc_val(v);
        return *this;
    }

    bool bit(int i) const { return (d[i >> 6] >> (i & 63)) & 1ULL; }
    void setbit(int i, bool v) {
        if (v)
            d[i >> 6] |= 1ULL << (i & 63);
        else
            d[i >> 6] &= ~(1ULL << (i & 63));
    }
    i64 to_i64() const {
        static_assert(W <= 64, "sc_shim: arithmetic on a value wider than 64 bits is not modelled");
        u64 v = d[0];
        if (S && W < 64 && bit(W - 1)) v |= ~lowmask(W);
        return (i64)v;
    }
    dyn to_dyn() const {
        dyn r(W);
        for (int i = 0; i < NW; i++) r.d[i] = d[i];
        return r;
    }
    int to_int() const { return (int)to_i64(); }
    unsigned to_uint() const { return (unsigned)to_i64(); }
    i64 to_int64() const { return to_i64(); }
    u64 to_uint64() const { return (u64)to_i64(); }
    int length() const { return W; }
    bool or_reduce() const {
        for (int i = 0; i < NW; i++)
            if (d[i]) return true;
        return false;
    }

    rangeref<W, S> range(int h, int l) { return rangeref<W, S>(this, h, l); }
    dyn range(int h, int l) const {
        dyn r(h - l + 1);
        extract_bits(d, NW, l, h - l + 1, r.d);
        return r;
    }
    rangeref<W, S> operator()(int h, int l) { return range(h, l); }
    bitref<W, S> operator[](int i) { return bitref<W, S>(this, i); }
    sc_val<1, false> operator[](int i) const { return sc_val<1, false>(bit(i) ? 1 : 0); }
    template <class T, typename std::enable_if<is_scx<T>::value, int>::type = 0>
    bitref<W, S> operator[](const T& i) {
        return bitref<W, S>(this, (int)i.to_i64());
    }

    i64 operator~() const { return ~to_i64(); }
    i64 operator-() const { return -to_i64(); }
    sc_val operator+() const { return *this; }
    sc_val& operator++() {
        *this = sc_val(to_i64() + 1);
        return *this;
    }
    sc_val operator++(int) {
        sc_val t = *this;
        *this = sc_val(to_i64() + 1);
        return t;
    }
    sc_val& operator--() {
        *this = sc_val(to_i64() - 1);
        return *this;
    }
    sc_val operator--(int) {
        sc_val t = *this;
        *this = sc_val(to_i64() - 1);
        return t;
    }
};

template <class T>
static inline i64 num(const T& v, typename std::enable_if<is_int<T>::value, int>::type = 0) {
    return (i64)v;
}
template <class T>
static inline i64 num(const T& v, typename std::enable_if<is_scx<T>::value, int>::type = 0) {
    return v.to_i64();
}
template <class T>
static inline dyn as_dyn(const T& v, typename std::enable_if<is_scx<T>::value, int>::type = 0) {
    return v.to_dyn();
}

template <class T>
static inline dyn rhs_dyn(const T& v, int len, typename std::enable_if<is_int<T>::value, int>::type = 0) {
    dyn r(len < 64 ? 64 : len);
    raw_words(v, r.d, r.nw());
    return r;
}
template <class T>
static inline dyn rhs_dyn(const T& v, int, typename std::enable_if<is_scx<T>::value, int>::type = 0) {
    return v.to_dyn();
}

template <int W, bool S>
template <class T, typename std::enable_if<is_operand<T>::value, int>::type>
bitref<W, S>& bitref<W, S>::operator=(const T& v) {
    p->setbit(i, (num(v) & 1) != 0);
    return *this;
}
template <int W, bool S>
template <class T, typename std::enable_if<is_operand<T>::value, int>::type>
rangeref<W, S>& rangeref<W, S>::operator=(const T& v) {
    /* integral and narrow operands go through int64; wide ones through their bit pattern */
    dyn src = rhs_dyn(v, len());
    deposit_bits(p->d, lo, len(), src.d, src.nw());
    return *this;
}
/* ---- operators: exact in int64 for operands that fit (every use in the hot path does) ---- */
#define SC_SHIM_ARITH(op)                                                                               \
    template <class A, class B,                                                                        \
              typename std::enable_if<(is_scx<A>::value || is_scx<B>::value) && is_operand<A>::value && \
                                          is_operand<B>::value,                                        \
                                      int>::type = 0>                                                  \
    inline i64 operator op(const A& a, const B& b) {                                                   \
        return num(a) op num(b);                                                                       \
    }
SC_SHIM_ARITH(+)
SC_SHIM_ARITH(-)
SC_SHIM_ARITH(*)
SC_SHIM_ARITH(/)
SC_SHIM_ARITH(%)
SC_SHIM_ARITH(&)
SC_SHIM_ARITH(|)
SC_SHIM_ARITH(^)
SC_SHIM_ARITH(<<)
SC_SHIM_ARITH(>>)
#undef SC_SHIM_ARITH
#define SC_SHIM_CMP(op)                                                                                 \
    template <class A, class B,                                                                        \
              typename std::enable_if<(is_scx<A>::value || is_scx<B>::value) && is_operand<A>::value && \
                                          is_operand<B>::value,                                        \
                                      int>::type = 0>                                                  \
    inline bool operator op(const A& a, const B& b) {                                                  \
        return num(a) op num(b);                                                                       \
    }
SC_SHIM_CMP(<)
SC_SHIM_CMP(>)
SC_SHIM_CMP(<=)
SC_SHIM_CMP(>=)
SC_SHIM_CMP(==)
SC_SHIM_CMP(!=)
#undef SC_SHIM_CMP
#define SC_SHIM_COMPOUND(op)                                                                       \
    template <int W, bool S, class B, typename std::enable_if<is_operand<B>::value, int>::type = 0> \
    inline sc_val<W, S>& operator op##=(sc_val<W, S>& a, const B& b) {                             \
        a = sc_val<W, S>(a.to_i64() op num(b));                                                    \
        return a;                                                                                  \
    }
SC_SHIM_COMPOUND(+)
SC_SHIM_COMPOUND(-)
SC_SHIM_COMPOUND(*)
SC_SHIM_COMPOUND(&)
SC_SHIM_COMPOUND(|)
SC_SHIM_COMPOUND(^)
SC_SHIM_COMPOUND(<<)
SC_SHIM_COMPOUND(>>)
#undef SC_SHIM_COMPOUND

/* concatenation: (a, b) -> a is the most significant part */
template <class A, class B,
          typename std::enable_if<is_scx<A>::value && is_scx<B>::value, int>::type = 0>
inline dyn operator,(const A& a, const B& b) {
    dyn hi = as_dyn(a), lo = as_dyn(b);
    dyn r(hi.w + lo.w);
    for (int i = 0; i < lo.nw(); i++) r.d[i] = lo.d[i];
    deposit_bits(r.d, lo.w, hi.w, hi.d, hi.nw());
    return r;
}

template <int W, bool S>
inline std::ostream& operator<<(std::ostream& os, const sc_val<W, S>& v) {
    for (int i = W - 1; i >= 0; i--) os << (v.bit(i) ? '1' : '0');
    return os;
}

}  // namespace sc_shim

/* ---- the SystemC names ---- */
template <int W>
using sc_bigint = sc_shim::sc_val<W, true>;
template <int W>
using sc_biguint = sc_shim::sc_val<W, false>;
template <int W>
using sc_bv = sc_shim::sc_val<W, false>;
template <int W>
using sc_lv = sc_shim::sc_val<W, false>;

template <int W>
struct sc_uint : sc_shim::sc_val<W, false> {
    typedef sc_shim::sc_val<W, false> base;
    sc_uint() : base() {}
    template <class T, typename std::enable_if<sc_shim::is_operand<T>::value, int>::type = 0>
    sc_uint(const T& v) : base(v) {}
    template <class T, typename std::enable_if<sc_shim::is_operand<T>::value, int>::type = 0>
    sc_uint& operator=(const T& v) {
        base::operator=(base(v));
        return *this;
    }
    operator unsigned long long() const { return (unsigned long long)this->to_i64(); }
    sc_uint& operator++() {
        base::operator++();
        return *this;
    }
    sc_uint operator++(int) {
        sc_uint t = *this;
        base::operator++();
        return t;
    }
    sc_uint& operator--() {
        base::operator--();
        return *this;
    }
    sc_uint operator--(int) {
        sc_uint t = *this;
        base::operator--();
        return t;
    }
};
template <int W>
struct sc_int : sc_shim::sc_val<W, true> {
    typedef sc_shim::sc_val<W, true> base;
    sc_int() : base() {}
    template <class T, typename std::enable_if<sc_shim::is_operand<T>::value, int>::type = 0>
    sc_int(const T& v) : base(v) {}
    template <class T, typename std::enable_if<sc_shim::is_operand<T>::value, int>::type = 0>
    sc_int& operator=(const T& v) {
        base::operator=(base(v));
        return *this;
    }
    operator long long() const { return this->to_i64(); }
};

/* ---- module scaffolding (no simulation kernel) ---- */
struct sc_module_name {
    sc_module_name(const char* = "") {}
};
struct sc_module {
    void reset_signal_is(const void*, bool) {}
    template <class T>
    void reset_signal_is(const T&, bool) {}
    const char* name() const { return "sc_shim"; }
};
#define SC_MODULE(x) struct x : public sc_module
#define SC_CTOR(x) x(sc_module_name = sc_module_name())
#define SC_CTHREAD(fn, clk) \
    do {                    \
    } while (0)
#define SC_THREAD(fn) \
    do {              \
    } while (0)
#define SC_METHOD(fn) \
    do {              \
    } while (0)

inline void wait() {
    if (sc_shim::wait_hook) sc_shim::wait_hook();
}
inline void wait(int) { wait(); }

template <class T>
struct sc_signal {
    T v;
    sc_signal(const char* = "") : v() {}
    const T& read() const { return v; }
    void write(const T& x) { v = x; }
    operator const T&() const { return v; }
    sc_signal& operator=(const T& x) {
        v = x;
        return *this;
    }
    int pos() const { return 0; }
};
template <class T>
struct sc_in : sc_signal<T> {
    sc_in(const char* n = "") : sc_signal<T>(n) {}
};
template <class T>
struct sc_out : sc_signal<T> {
    sc_out(const char* n = "") : sc_signal<T>(n) {}
    using sc_signal<T>::operator=;
};
template <class T>
struct sc_fifo {
    std::deque<T> q;
    sc_fifo(const char* = "", int = 16) {}
    T read() {
        if (q.empty()) throw sc_shim::fifo_empty();
        T v = q.front();
        q.pop_front();
        return v;
    }
    void write(const T& v) { q.push_back(v); }
    int num_available() const { return (int)q.size(); }
};
template <class T>
struct sc_fifo_in : sc_fifo<T> {};
template <class T>
struct sc_fifo_out : sc_fifo<T> {};

/* shared/src/vector.h:32 calls its debug printer SHOW_SM before declaring it (vector.h:257);
 * g++ needs the template name visible at that point.  Forward declaration only. */
template <int P, int Q>
inline void SHOW_SM(sc_biguint<P * Q> a);

#endif /* SC_SHIM_SYSTEMC_H */
