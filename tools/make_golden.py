#!/usr/bin/env python3
"""Generate the committed fixtures from the read-only reference tree.

Run in the build container only (needs /root/reference):
    python tools/make_golden.py
Writes
  tests/golden/codewords.json      the 9 golden codewords of
                                   src/testbench/sc_encoder/sc_encoder.h:74-89
  tests/golden/frozen_index.json   N, K, source file and sha256 of every frozen table
  sc_polar_decoder_hls_b200/data/<name>.bits
                                   information-flag sets (1 = information bit), packed
                                   LSB-first, for the BASELINE.json configs and the golden
                                   codewords' codes.  The GPU box has no /root/reference, so
                                   bench.py / tests regenerate reference-format text files from
                                   these through the product's own writers.
"""
import hashlib
import json
import os
import re
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_order(path, k):
    tok = open(path).read().split()
    n = int(tok[0])
    idx = [int(x) for x in tok[3:3 + n] if int(x) < n]
    flags = [0] * n
    for i in idx[:k]:
        flags[i] = 1
    return flags


def load_flags(path):
    return [int(x) for x in open(path).read().split()]


def pack(flags):
    out = bytearray(len(flags) // 8 if len(flags) >= 8 else 1)
    for i, f in enumerate(flags):
        if f:
            out[i >> 3] |= 1 << (i & 7)
    return bytes(out)


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are already committed")
    src = open(f"{REF}/src/testbench/sc_encoder/sc_encoder.h").read()
    cws = {}
    for name, n in (("cw8x4", 8), ("cw512x256", 512), ("cw1024x512", 1024)):
        m = re.search(r"const bool %s\[3\]\[%d\] = \{(.*?)\};" % (name, n), src, re.S)
        rows = re.findall(r"\{([^{}]*)\}", m.group(1))
        assert len(rows) == 3
        cws[name] = ["".join(re.findall(r"[01]", r)) for r in rows]
        assert all(len(r) == n for r in cws[name])
    os.makedirs(f"{ROOT}/tests/golden", exist_ok=True)
    json.dump(cws, open(f"{ROOT}/tests/golden/codewords.json", "w"), indent=0)

    keep = {
        "FB_N8_K4": ("Frozen_Bit_Tab/FB_N8_K4.txt", 4),
        "FB_N512_K256": ("Frozen_Bit_Tab/FB_N512_K256.txt", 256),
        "FB_N1024_K512": ("Frozen_Bit_Tab/FB_N1024_K512.txt", 512),
        "FB_N2048_K1024": ("Frozen_Bit_Tab/FB_N2048_K1024.txt", 1024),
        "frozen_n_1024_k_512": ("Generated_Frozen_Bit/frozen_n_1024_k_512.txt", None),
        "frozen_n_4096_k_3072": ("Generated_Frozen_Bit/frozen_n_4096_k_3072.txt", None),
        "frozen_n_32768_k_29492_snr_4_5": ("Generated_Frozen_Bit/frozen_n_32768_k_29492_snr_4_5.txt", None),
        "frozen_n_131072_k_117964": ("Generated_Frozen_Bit/frozen_n_131072_k_117964.txt", None),
        "frozen_n_524288_k_262144": ("Generated_Frozen_Bit/frozen_n_524288_k_262144.txt", None),
    }
    os.makedirs(f"{ROOT}/sc_polar_decoder_hls_b200/data", exist_ok=True)
    index = {}
    for d in ("Frozen_Bit_Tab", "Generated_Frozen_Bit"):
        for fn in sorted(os.listdir(f"{REF}/{d}")):
            p = f"{REF}/{d}/{fn}"
            if d == "Frozen_Bit_Tab":
                m = re.match(r"FB_N(\d+)_K(\d+)\.txt", fn)
                n, k = int(m.group(1)), int(m.group(2))
                flags = load_order(p, k)
            else:
                m = re.match(r"frozen_n_(\d+)_k_(\d+)", fn)
                n, k = int(m.group(1)), int(m.group(2))
                flags = load_flags(p)
            assert len(flags) == n and sum(flags) == k, fn
            index[f"{d}/{fn}"] = {"n": n, "k": k, "sha256_packed": hashlib.sha256(pack(flags)).hexdigest()}
    for name, (rel, k) in keep.items():
        flags = load_order(f"{REF}/{rel}", k) if k is not None else load_flags(f"{REF}/{rel}")
        open(f"{ROOT}/sc_polar_decoder_hls_b200/data/{name}.bits", "wb").write(pack(flags))
    json.dump(index, open(f"{ROOT}/tests/golden/frozen_index.json", "w"), indent=1, sort_keys=True)
    print("wrote", len(index), "index entries,", len(keep), "packed sets, 9 codewords")


if __name__ == "__main__":
    main()
