/*
 * ref_driver.cpp -- drives the reference's OWN decoder (src/module/my_module.h with
 * wrapper_in.h / wrapper_out.h and shared/src/*.h, compiled unmodified where they lie under
 * /root/reference) through the systemc.h shim.  TEST INFRASTRUCTURE ONLY.
 *
 * One shared object per compile-time configuration (config.h + polar_parameters.h are
 * generated next to symlinks of the module headers by oracle/build_ref.py, because the
 * reference configures N / PAR / Q / format / EXTENDED with macros).
 *
 * There is no SystemC kernel: each SC_CTHREAD body is called as a plain function after its
 * input FIFO has been pre-loaded; the body returns control by throwing when it next reads an
 * empty FIFO (or, for do_prunning, from the wait() hook once `enable` is raised).
 */
#include "systemc.h"

namespace sc_shim {
void (*wait_hook)() = nullptr;
}

#define private public
#include "src/module/my_module.h"
#include "src/module/wrapper_in.h"
#include "src/module/wrapper_out.h"
#undef private

#include <memory>
#include <vector>

static my_module* g_dut = nullptr;
static void prunning_hook() {
    if (g_dut && g_dut->enable.read()) throw sc_shim::stop_request();
}

extern "C" {

/* compile-time configuration of this build: n, par, q, format(0 CA2,1 SIGMAG), extended, pruning */
void ref_config(int32_t out[6]) {
    out[0] = _NBITS;
    out[1] = PAR;
    out[2] = LLR_BITS;
#if defined CA2
    out[3] = 0;
#else
    out[3] = 1;
#endif
    out[4] = EXTENDED;
    out[5] = PRUNING_LEVEL;
}

/* Decode nframes frames exactly as the testbench pipeline quantiser -> wrapper_in -> my_module
 * -> wrapper_out does (sc_top_module.h:146-160).  llr: [nframes][N] int8; flags: N bytes
 * (1 = information); xhat: [nframes][N] bytes. */
int ref_decode(const uint8_t* flags, const int8_t* llr, size_t nframes, uint8_t* xhat) {
    std::unique_ptr<my_module> dut(new my_module("dut"));
    std::unique_ptr<wrapper_in> win(new wrapper_in("w_in"));
    std::unique_ptr<wrapper_out> wout(new wrapper_out("w_o"));

    /* frozen table: encoder.do_fb streams Frozen_Bits[0..N) once (sc_encoder.h:59-62) */
    for (long i = 0; i < _NBITS; i++) dut->FB.q.push_back((BIT)(int)(flags[i] & 1));
    g_dut = dut.get();
    sc_shim::wait_hook = prunning_hook;
    try {
        dut->do_prunning();
    } catch (sc_shim::stop_request&) {
    } catch (sc_shim::fifo_empty&) {
        sc_shim::wait_hook = nullptr;
        return -1;
    }
    sc_shim::wait_hook = nullptr;
    g_dut = nullptr;

    const size_t chunk = 64; /* frames per pass, bounds FIFO memory */
    for (size_t f0 = 0; f0 < nframes; f0 += chunk) {
        size_t nf = (nframes - f0 < chunk) ? nframes - f0 : chunk;
        /* sc_quantizer writes a short into sc_fifo<LLR> (sc_quantizer.h:80) */
        for (size_t i = 0; i < nf * (size_t)_NBITS; i++) win->e.q.push_back((LLR)(int)llr[f0 * _NBITS + i]);
        try {
            win->do_action();
        } catch (sc_shim::fifo_empty&) {
        }
        dut->e.q.swap(win->s.q);
        try {
            dut->do_action();
        } catch (sc_shim::fifo_empty&) {
        }
        if (dut->s.q.size() != nf * (size_t)N_DIVIDED) return -2;
        wout->e.q.swap(dut->s.q);
        try {
            wout->do_action();
        } catch (sc_shim::fifo_empty&) {
        }
        if (wout->s.q.size() != nf * (size_t)_NBITS) return -3;
        size_t k = f0 * (size_t)_NBITS;
        for (auto& b : wout->s.q) xhat[k++] = (uint8_t)b.to_int();
        wout->s.q.clear();
    }
    return 0;
}

/* Primitive entry points on one PAR-wide word; lanes passed as raw Q-bit patterns. */
static TYPE_LLRS pack_word(const uint32_t* lanes) {
    TYPE_LLRS w = 0;
    for (int i = 0; i < PAR; i++) w.range(LLR_BITS * (i + 1) - 1, LLR_BITS * i) = (sc_biguint<LLR_BITS>)(unsigned long long)lanes[i];
    return w;
}
static void unpack_word(const TYPE_LLRS& w, uint32_t* lanes) {
    for (int i = 0; i < PAR; i++) lanes[i] = (uint32_t)((sc_biguint<LLR_BITS>)w.range(LLR_BITS * (i + 1) - 1, LLR_BITS * i)).to_uint64();
}
void ref_pu_f(const uint32_t* a, const uint32_t* b, uint32_t* out) {
    TYPE_LLRS r = PU_FUNCTION_F<PAR, LLR_BITS>(pack_word(a), pack_word(b));
    unpack_word(r, out);
}
void ref_pu_g(const uint32_t* a, const uint32_t* b, const uint8_t* s, uint32_t* out) {
    TYPE_BITS sa = 0;
    for (int i = 0; i < PAR; i++) sa[i] = (int)(s[i] & 1);
    TYPE_LLRS r = PU_FUNCTION_G<PAR, LLR_BITS>(pack_word(a), pack_word(b), sa);
    unpack_word(r, out);
}
void ref_leaf(const uint32_t* a, const uint8_t* flags, uint8_t* bits) {
    TYPE_BITS fb = 0;
    for (int i = 0; i < PAR; i++) fb[i] = (int)(flags[i] & 1);
    TYPE_BITS r = Spec_Polar_Decoder<PAR, LLR_BITS>(pack_word(a), fb);
    for (int i = 0; i < PAR; i++) bits[i] = r.bit(i) ? 1 : 0;
}
/* wrapper_in's Adapt_format on one LLR (wrapper_in.h:33-34) */
uint32_t ref_adapt(int llr) {
    LLR value = (LLR)llr;
    LLR va = Adapt_format<LLR_BITS>(value);
    return (uint32_t)((sc_biguint<LLR_BITS>)va).to_uint64();
}
}
