#!/usr/bin/env python3
"""Build oracle/_ref/: the reference's own decoder compiled natively (TEST INFRASTRUCTURE).

The reference configures everything at compile time (src/module/config.h, polar_parameters.h),
so one shared object is built per configuration:

    oracle/_ref/refdec_n{N}_p{PAR}_q{Q}_{ca2|sm}_e{EXT}.so

Nothing is copied from /root/reference: the module headers are *symlinked* into a scratch tree
(so that their `#include "config.h"` / `#include "polar_parameters.h"` pick up the generated
files and `../../../shared/src/library.h` resolves), and compiled in place with g++ against
oracle/shim/systemc.h.  polar_parameters.h is produced by the reference's own
Frozen_Bit_Generator (compiled from its source into oracle/_ref/FB_Generator).

Usage: python oracle/build_ref.py [--ref /root/reference] [--only TAG ...] [--list]
"""
import argparse
import math
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

# (N, K, frozen source relative to the reference root, is_flag_file, PAR, Q, format, EXTENDED)
CONFIGS = [
    # golden-codeword codes
    (8, 4, "Frozen_Bit_Tab/FB_N8_K4.txt", 0, 2, 8, "ca2", 1),
    (8, 4, "Frozen_Bit_Tab/FB_N8_K4.txt", 0, 4, 6, "sm", 1),
    (512, 256, "Frozen_Bit_Tab/FB_N512_K256.txt", 0, 16, 8, "ca2", 1),
    (512, 256, "Frozen_Bit_Tab/FB_N512_K256.txt", 0, 64, 6, "sm", 1),
    # BASELINE config 1 and its neighbours (PAR / Q / format / EXTENDED sweep)
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 16, 8, "ca2", 1),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 16, 8, "ca2", 0),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 4, 8, "ca2", 1),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 64, 8, "ca2", 1),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 256, 8, "ca2", 1),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 16, 6, "ca2", 1),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 16, 5, "ca2", 1),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 16, 9, "ca2", 1),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 16, 6, "sm", 1),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 16, 8, "sm", 1),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 16, 6, "sm", 0),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 64, 7, "sm", 1),
    # configurations only the raw-pattern kernel covers: wrapping 5-bit alphabet in SIGMAG, 17-bit leaf, SIGMAG Q 9
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 16, 5, "sm", 1),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 256, 9, "ca2", 1),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 4, 9, "sm", 1),
    # the reference's checked-in default (config.h): SIGMAG, LLR_BITS 6, PRUNING_LEVEL 2 with R1 / REP / SPC / H0 --
    # NOT the plain-SC contract of this library; built to measure how far the reference's own pruning departs from it
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 16, 6, "sm", 1, 2),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 16, 8, "ca2", 1, 2),
    # more PRUNING_LEVEL 2 builds: they pin sco_decode_l2 (the restatement of that decoder) over PAR / Q / format / N
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 4, 8, "ca2", 1, 2),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 64, 8, "ca2", 1, 2),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 16, 8, "ca2", 0, 2),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 64, 7, "sm", 1, 2),
    (1024, 512, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0, 4, 9, "sm", 1, 2),
    (512, 256, "Frozen_Bit_Tab/FB_N512_K256.txt", 0, 16, 8, "ca2", 1, 2),
    (8, 4, "Frozen_Bit_Tab/FB_N8_K4.txt", 0, 2, 8, "ca2", 1, 2),
    (4096, 3072, "Generated_Frozen_Bit/frozen_n_4096_k_3072.txt", 1, 16, 8, "ca2", 1, 2),
    (32768, 29492, "Generated_Frozen_Bit/frozen_n_32768_k_29492_snr_4_5.txt", 1, 16, 8, "ca2", 1, 2),
    # BASELINE configs 2..5 at the headline setting
    (4096, 3072, "Generated_Frozen_Bit/frozen_n_4096_k_3072.txt", 1, 16, 8, "ca2", 1),
    (32768, 29492, "Generated_Frozen_Bit/frozen_n_32768_k_29492_snr_4_5.txt", 1, 16, 8, "ca2", 1),
    (131072, 117964, "Generated_Frozen_Bit/frozen_n_131072_k_117964.txt", 1, 16, 8, "ca2", 1),
    (524288, 262144, "Generated_Frozen_Bit/frozen_n_524288_k_262144.txt", 1, 16, 8, "ca2", 1),
]


def tag(cfg):
    n, k, src, isflag, par, q, fmt, ext = cfg[:8]
    pl = cfg[8] if len(cfg) > 8 else 0
    return f"n{n}_p{par}_q{q}_{fmt}_e{ext}" + (f"_pl{pl}" if pl else "")


def sh(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise SystemExit(f"command failed: {' '.join(cmd)}")
    return r.stdout


def symlink(src, dst):
    if os.path.islink(dst) or os.path.exists(dst):
        os.remove(dst)
    os.symlink(src, dst)


CONFIG_H = """
#define LLR_BITS		{q}
#define LLR_MAXV    	(+127)
#define LLR_MINV    	(-127)

#define LLR 			sc_bigint<LLR_BITS>
#define BIT 			sc_biguint<1>
#define TYPE_LLRS   	sc_bigint<PAR * LLR_BITS>
#define TYPE_BITS   	sc_biguint<PAR>

#define {fmt}

#define EXTENDED		{ext}

#define PRUNING_LEVEL 	{pl}

#define ELAG_R1			{on}
#define ELAG_REP		{on}
#define ELAG_SPC		{on}
#define ELAG_REP2		0
#define ELAG_SPC2		0
#define ELAG_RARE		0
#define ELAG_H0			{on}

#define _MONITORING_
"""


def build_one(ref, cfg, cxx):
    n, k, src, isflag, par, q, fmt, ext = cfg[:8]
    pl = cfg[8] if len(cfg) > 8 else 0  # PRUNING_LEVEL; > 0 switches the checked-in ELAG_* defaults on (config.h:16-30)
    t = tag(cfg)
    so = os.path.join(OUT, f"refdec_{t}.so")
    work = os.path.join(OUT, "work", t)
    mod = os.path.join(work, "tree", "a", "src", "module")
    os.makedirs(mod, exist_ok=True)
    # FB_Generator writes ../../Frozen_Bit_Tab/FB_N*_K*.txt relative to its cwd for order input
    cwd = os.path.join(work, "x", "y")
    os.makedirs(cwd, exist_ok=True)
    os.makedirs(os.path.join(work, "Frozen_Bit_Tab"), exist_ok=True)
    sh([os.path.join(OUT, "FB_Generator"), str(n), str(k), str(par), "0", os.path.join(ref, src), str(isflag),
        mod + "/"], cwd=cwd)
    # config.h in the reference's own macro vocabulary (the sweep scripts rewrite it the same way,
    # script/script_tests.sh:24-53)
    open(os.path.join(mod, "config.h"), "w").write(
        CONFIG_H.format(q=q, fmt="CA2" if fmt == "ca2" else "SIGMAG", ext=ext, pl=pl, on=1 if pl else 0))
    for h in ("my_module.h", "wrapper_in.h", "wrapper_out.h"):
        symlink(os.path.join(ref, "src", "module", h), os.path.join(mod, h))
    symlink(os.path.join(ref, "shared"), os.path.join(work, "tree", "shared"))
    depth_div = int(math.log2(n // par)) + 1
    maxbits = max(par * (q + 3), 8 * depth_div + 8, 128) + 64
    cmd = [cxx, "-std=c++14", "-O2", "-fPIC", "-shared", "-w", f"-DSC_SHIM_MAXBITS={maxbits}",
           # macros only the unused SC-List templates of sc_list_fct.h need to parse (no config defines them)
           "-DMAX_VAL=127", "-DL_SIZE=2", "-DLOG2_L=1",
           "-I", os.path.join(HERE, "shim"), "-I", os.path.join(work, "tree", "a"),
           "-o", so, os.path.join(HERE, "ref_driver.cpp")]
    sh(cmd)
    return so


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", nargs="*")
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--cxx", default=os.environ.get("CXX", "g++"))
    a = ap.parse_args()
    if a.list:
        for c in CONFIGS:
            print(tag(c))
        return
    if not os.path.isdir(a.ref):
        print("reference tree absent: keeping prebuilt oracle/_ref (if any)")
        return
    os.makedirs(OUT, exist_ok=True)
    fbg = os.path.join(OUT, "FB_Generator")
    if not os.path.exists(fbg):
        sh([a.cxx, "-std=c++11", "-O1", "-w", "-o", fbg, os.path.join(a.ref, "Frozen_Bit_Generator", "main.cpp")])
    for c in CONFIGS:
        if a.only and tag(c) not in a.only:
            continue
        so = os.path.join(OUT, f"refdec_{tag(c)}.so")
        if os.path.exists(so) and os.path.getmtime(so) > max(
                os.path.getmtime(os.path.join(HERE, "ref_driver.cpp")),
                os.path.getmtime(os.path.join(HERE, "shim", "systemc.h"))):
            continue
        print("building", tag(c), flush=True)
        build_one(a.ref, c, a.cxx)
    print("ok")


if __name__ == "__main__":
    main()
