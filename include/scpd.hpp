// scpd.hpp -- C++ host-side mirror of the reference's decoder interface over the C ABI (scpd.h).
//
// The reference exposes the decoder as SC_MODULE(my_module) with three ports
// (src/module/my_module.h:32-35):   FB  (frozen table, once)   e  (LLR words in)   s  (bit words out)
// wired by sc_top_module (src/testbench/sc_top_module.h:146-160).  PolarDecoder keeps that shape:
// the constructor is the FB port + do_prunning(), operator() / decode() is one e -> s transfer
// for a whole batch of frames, and run_ber() is the testbench loop of src/testbench/main.cpp.
// Header-only; link against libscpd.so.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "scpd.h"

namespace scpd {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};
inline void check(int status) {
    if (status != SCPD_OK) throw Error(status, std::string(scpd_status_string(status)) + ": " + scpd_last_error());
}

// Frozen-table loaders: Frozen_Bit_Generator/src/Writer.h:21-105.
inline std::vector<uint8_t> load_order_table(const std::string& path, uint32_t n, uint32_t k) {
    std::vector<uint8_t> f(n);
    check(scpd_frozen_load_order(path.c_str(), n, k, f.data()));
    return f;
}
inline std::vector<uint8_t> load_flag_table(const std::string& path, uint32_t n, uint32_t* k = nullptr) {
    std::vector<uint8_t> f(n);
    check(scpd_frozen_load_flags(path.c_str(), n, f.data(), k));
    return f;
}

struct BerCounters {  // sc_error_counter.h:36-40 (+ the reference's 10-bit wrapped variants)
    uint64_t bit_errors, frame_errors, bits, frames, bit_errors_wrapped, frame_errors_wrapped;
    double ber() const { return bits ? double(bit_errors) / double(bits) : 0.0; }
    double fer() const { return frames ? double(frame_errors) / double(frames) : 0.0; }
};

class PolarDecoder {
public:
    // config.h / polar_parameters.h macros -> scpd_config; info_flags: 1 = information bit
    PolarDecoder(const scpd_config& cfg, const std::vector<uint8_t>& info_flags, int device = 0) : cfg_(cfg) {
        if (info_flags.size() != cfg.n) throw Error(SCPD_E_ARG, "info_flags must have n entries");
        check(scpd_create(&cfg_, info_flags.data(), device, &h_));
    }
    ~PolarDecoder() { scpd_destroy(h_); }
    PolarDecoder(const PolarDecoder&) = delete;
    PolarDecoder& operator=(const PolarDecoder&) = delete;

    uint32_t n() const { return cfg_.n; }
    uint32_t k() const { return cfg_.k; }
    uint32_t words_per_frame() const { return cfg_.n >= 32 ? cfg_.n / 32 : 1; }

    // device buffers, asynchronous on `stream` (a cudaStream_t)
    void decode(const int8_t* d_llr, size_t nframes, uint32_t* d_xhat, void* stream = nullptr) {
        check(scpd_decode(h_, d_llr, nframes, d_xhat, stream));
    }
    // host buffers: what a host-only caller such as the reference testbench would use
    std::vector<uint32_t> operator()(const std::vector<int8_t>& llr) {
        const size_t nframes = llr.size() / cfg_.n;
        std::vector<uint32_t> out(nframes * words_per_frame());
        check(scpd_decode_host(h_, llr.data(), nframes, out.data()));
        return out;
    }
    BerCounters run_ber(float ebn0_db, float rate, uint64_t nframes, uint64_t first_frame = 0, uint8_t seed = 0xF0,
                        const uint8_t* codeword = nullptr) {
        uint64_t c[6];
        check(scpd_run_ber(h_, ebn0_db, rate, first_frame, nframes, seed, codeword, c));
        return BerCounters{c[0], c[1], c[2], c[3], c[4], c[5]};
    }
    // the other codeword sources (stored codewords in turn, random payloads) and the information-bit counters:
    // c[0..5] codeword bits as run_ber, c[6..9] information bits (errors, frames in error, bits, frames)
    std::vector<uint64_t> run_ber_ex(float ebn0_db, float rate, uint64_t nframes, uint64_t first_frame, uint8_t seed, int src_mode,
                                     const uint8_t* codewords, uint32_t ncw, uint64_t payload_seed) {
        std::vector<uint64_t> c(10);
        check(scpd_run_ber_ex(h_, ebn0_db, rate, first_frame, nframes, seed, src_mode, codewords, ncw, payload_seed, c.data()));
        return c;
    }
    // measurement aids (sc_monitor's role): duration of the tree-walk kernel of the last decode, kernel in use
    void kernel_timing(bool on) { check(scpd_kernel_timing(h_, on ? 1 : 0)); }
    float last_kernel_ms() {
        float ms = 0.f;
        check(scpd_last_kernel_ms(h_, &ms));
        return ms;
    }
    std::string kernel_name() const { return scpd_kernel_name(h_); }
    std::string last_kernel_name() const { return scpd_last_kernel_name(h_); }
    // sc_monitor's function x level matrix for this handle's table and pruning mode (host only)
    scpd_stage_matrix stage_profile(const std::vector<uint8_t>& info_flags) const {
        scpd_stage_matrix m;
        check(scpd_stage_profile(&cfg_, info_flags.data(), &m));
        return m;
    }
    scpd_decoder* handle() { return h_; }

private:
    scpd_config cfg_;
    scpd_decoder* h_ = nullptr;
};

}  // namespace scpd
