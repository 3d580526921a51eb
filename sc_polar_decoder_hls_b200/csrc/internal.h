// internal.h -- small host-side helpers shared by the translation units of libscpd.
#pragma once
#include <cstdint>
#include <string>

namespace scpd {

int set_error(int status, const std::string& msg);  // records msg for scpd_last_error(), returns status

static inline int ilog2(uint64_t v) {
    int l = 0;
    while ((1ull << (l + 1)) <= v) l++;
    return l;
}
static inline bool is_pow2(uint64_t v) { return v && !(v & (v - 1)); }

}  // namespace scpd
