// frozen_io.cpp -- frozen-bit table loaders / writers (host, no CUDA).
//
// Counterpart of Frozen_Bit_Generator (main.cpp:11-56, src/Writer.h:21-171): the two on-disk
// formats of Frozen_Bit_Tab/ and Generated_Frozen_Bit/, the "affect" order file it writes back
// (Writer.h:75-80) and the polar_parameters.h it generates (Writer.h:110-162).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/scpd.h"
#include "internal.h"

namespace scpd {

static bool read_lines(const char* path, std::vector<std::string>& lines) {
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) return false;
    std::string s;
    while (std::getline(f, s)) lines.push_back(s);  // a trailing '\r' is left in place: it is whitespace below
    return true;
}

static bool parse_ints(const std::string& line, std::vector<long long>& out) {
    const char* p = line.c_str();
    while (*p) {
        while (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n') p++;
        if (!*p) break;
        char* e = nullptr;
        long long v = std::strtoll(p, &e, 10);
        if (e == p) return false;  // not a number
        out.push_back(v);
        p = e;
    }
    return true;
}

}  // namespace scpd

using namespace scpd;

extern "C" int scpd_frozen_load_order(const char* path, uint32_t n, uint32_t k, uint8_t* flags_out) {
    if (!path || !flags_out) return set_error(SCPD_E_ARG, "scpd_frozen_load_order: null argument");
    if (n == 0 || k > n) return set_error(SCPD_E_CONFIG, "scpd_frozen_load_order: need 0 < n and k <= n");
    std::vector<std::string> lines;
    if (!read_lines(path, lines)) return set_error(SCPD_E_IO, std::string("cannot open ") + path);
    // line 1: N of the table; lines 2-3 skipped; line 4: the order (Writer.h:35-58)
    if (lines.size() < 4) return set_error(SCPD_E_IO, std::string(path) + ": expected 4 lines (N, 0, 0, order)");
    std::vector<long long> hdr, order;
    if (!parse_ints(lines[0], hdr) || hdr.empty() || hdr[0] <= 0)
        return set_error(SCPD_E_IO, std::string(path) + ": first line is not a frame size");
    if (!parse_ints(lines[3], order)) return set_error(SCPD_E_IO, std::string(path) + ": order line is not numeric");
    // keep only channels below n, in order (Writer.h:64-71)
    std::vector<uint32_t> sub;
    sub.reserve(n);
    std::vector<uint8_t> seen(n, 0);
    for (long long v : order) {
        if (v < 0) return set_error(SCPD_E_IO, std::string(path) + ": negative channel index");
        if ((unsigned long long)v < n) {
            if (seen[(size_t)v]) return set_error(SCPD_E_IO, std::string(path) + ": channel index repeated");
            seen[(size_t)v] = 1;
            sub.push_back((uint32_t)v);
        }
    }
    if (sub.size() != n)
        return set_error(SCPD_E_IO, std::string(path) + ": order does not cover all channels below n");
    std::memset(flags_out, 0, n);
    for (uint32_t i = 0; i < k; i++) flags_out[sub[i]] = 1;  // the k most reliable carry information (Writer.h:86-89)
    return SCPD_OK;
}

extern "C" int scpd_frozen_load_flags(const char* path, uint32_t n, uint8_t* flags_out, uint32_t* k_out) {
    if (!path || !flags_out) return set_error(SCPD_E_ARG, "scpd_frozen_load_flags: null argument");
    std::vector<std::string> lines;
    if (!read_lines(path, lines)) return set_error(SCPD_E_IO, std::string("cannot open ") + path);
    if (lines.empty()) return set_error(SCPD_E_IO, std::string(path) + ": empty file");
    std::vector<long long> v;
    if (!parse_ints(lines[0], v)) return set_error(SCPD_E_IO, std::string(path) + ": flag line is not numeric");  // Writer.h:97-104
    if (v.size() != n) return set_error(SCPD_E_IO, std::string(path) + ": number of flags differs from n");
    uint32_t k = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (v[i] != 0 && v[i] != 1) return set_error(SCPD_E_IO, std::string(path) + ": flag is not 0 or 1");
        flags_out[i] = (uint8_t)v[i];
        k += (uint32_t)v[i];
    }
    if (k_out) *k_out = k;
    return SCPD_OK;
}

extern "C" int scpd_frozen_write_order(const char* path, uint32_t n, const uint32_t* order) {
    if (!path || !order) return set_error(SCPD_E_ARG, "scpd_frozen_write_order: null argument");
    std::ofstream f(path, std::ios::binary);
    if (!f.is_open()) return set_error(SCPD_E_IO, std::string("cannot create ") + path);
    f << n << "\n0\n0\n";  // Writer.h:76
    for (uint32_t i = 0; i < n; i++) f << order[i] << "    ";  // :77-79
    return f.good() ? SCPD_OK : set_error(SCPD_E_IO, std::string("write failed: ") + path);
}

extern "C" int scpd_frozen_write_flags(const char* path, uint32_t n, const uint8_t* flags) {
    if (!path || !flags) return set_error(SCPD_E_ARG, "scpd_frozen_write_flags: null argument");
    std::ofstream f(path, std::ios::binary);
    if (!f.is_open()) return set_error(SCPD_E_IO, std::string("cannot create ") + path);
    std::string s;
    s.reserve(2 * (size_t)n);
    for (uint32_t i = 0; i < n; i++) {
        s.push_back(flags[i] ? '1' : '0');
        s.push_back(' ');
    }
    f << s;
    return f.good() ? SCPD_OK : set_error(SCPD_E_IO, std::string("write failed: ") + path);
}

extern "C" int scpd_write_polar_parameters(const char* path, uint32_t n, uint32_t par, int en,
                                           const uint8_t* flags) {
    if (!path || !flags) return set_error(SCPD_E_ARG, "scpd_write_polar_parameters: null argument");
    if (n == 0 || (n & (n - 1)) || par == 0 || (par & (par - 1)) || par > n)
        return set_error(SCPD_E_CONFIG, "scpd_write_polar_parameters: n and par must be powers of two, par <= n");
    std::ofstream o(path, std::ios::binary);
    if (!o.is_open()) return set_error(SCPD_E_IO, std::string("cannot create ") + path);
    const int log2n = ilog2(n), log2p = ilog2(par);
    // Writer.h:110-122
    o << "#ifndef POLAR_HEADER_H\n#define POLAR_HEADER_H\n\n";
    o << "#define _NBITS       " << n << "\n";
    o << "#define _LOG2N       " << log2n << "\n";
    o << "#define _DEPTH       " << (log2n + 1) << "\n\n";
    o << "#define PAR          " << par << "\n";
    o << "#define LOG2_PAR     " << log2p << "\n";
    o << "#define N_DIVIDED    (_NBITS / PAR) \n";
    o << "#define DEPTH_DIV    " << (log2n - log2p + 1) << "\n\n";
    o << "#define COUNTER      sc_uint<_DEPTH>\n\n";
    std::string words;  // PAR-bit words, most significant flag first (Writer.h:133-137)
    for (uint32_t i = 0; i < n / par; i++) {
        for (uint32_t j = 0; j < par; j++) words.push_back(flags[((i + 1) * par - 1) - j] ? '1' : '0');
        words += "\", \"";
    }
    std::string bits;
    for (uint32_t i = 0; i < n; i++) {
        bits.push_back(flags[i] ? '1' : '0');
        bits += ", ";
    }
    if (en) {  // Writer.h:123-142
        o << "const sc_bv<PAR> Frozen_Bits[N_DIVIDED] = {\n   //" << bits << "\n     \"";
        o << words.substr(0, words.size() - 3);  // seekp(-3): drops the trailing `, "`
    } else {  // Writer.h:143-160
        o << "const sc_bv<1> Frozen_Bits[_NBITS] = {\n   // \"" << words << "\n     ";
        o << bits.substr(0, bits.size() - 2);  // seekp(-2): drops the trailing `, `
    }
    o << "\n};\n\n\n#endif // POLAR_HEADER_H\n";
    return o.good() ? SCPD_OK : set_error(SCPD_E_IO, std::string("write failed: ") + path);
}
