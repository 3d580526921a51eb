#!/usr/bin/env python3
"""Throughput of the five BASELINE.json configurations on one GPU (one JSON line each), and the c5 BER/FER
sweep of configs[4] through scpd_run_ber.  Companion of bench.py (which is the contract line for configs[1]).

Every decode is spot-checked bit-for-bit against the oracle (tests/oracle_lib.py) on a few frames."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import oracle_lib as ol
import sc_polar_decoder_hls_b200 as scpd

SETS = {"c1": ("FB_N1024_K512", 1024, 512, 2.5, 1 << 20), "c2": ("frozen_n_4096_k_3072", 4096, 3072, 3.5, 1 << 20),
        "c3": ("frozen_n_32768_k_29492_snr_4_5", 32768, 29492, 4.5, 1 << 16),
        "c4": ("frozen_n_131072_k_117964", 131072, 117964, 4.5, 1 << 16),
        "c5": ("frozen_n_524288_k_262144", 524288, 262144, 2.0, 1 << 15)}

ap = argparse.ArgumentParser()
ap.add_argument("--cfgs", default="c1,c2,c3,c4,c5")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--sweep", action="store_true", help="also run the c5 BER/FER sweep 1..4 dB (2048 frames per point)")
a = ap.parse_args()
peak = 6552.3
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, KeyError, ValueError):
    pass
for cfg in a.cfgs.split(","):
    name, n, k, snr, frames = SETS[cfg]
    flags = scpd.packed_flags(name, n)
    dec = scpd.Decoder(n, k, flags)
    llr = scpd.channel_generate(n, frames, scpd.sigma(snr, k / n))
    out = torch.empty((frames, n // 32), dtype=torch.int32, device="cuda")
    for _ in range(2):
        dec.decode(llr, out)
    torch.cuda.synchronize()
    chk = 8 if n > 8192 else 64
    want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr[:chk].cpu().numpy(), threads=8)
    assert (out[:chk].cpu().numpy().view(np.uint32) == want).all(), cfg
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        dec.decode(llr, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    dec.kernel_timing(True)
    dec.decode(llr, out)
    kms = dec.last_kernel_ms()
    fps = frames / (ms * 1e-3)
    ops, fg = dec.schedule_stats()
    lane_instr = (n / 2) * np.log2(n) * (7 / 4 + 1 / 32)
    print(json.dumps({"config": cfg, "n": n, "k": k, "ebn0_db": snr, "frames": frames, "ms_per_decode": ms,
                      "tree_walk_kernel_ms": kms, "frames_per_s": fps, "info_gbps": fps * k / 1e9,
                      "coded_gbps": fps * n / 1e9, "kernel": dec.kernel_name, "fg_updates_per_frame": fg,
                      "hbm_roofline_frac": (n + n // 8) * frames / (kms * 1e-3) / 1e9 / peak,
                      "alu_roofline_frac": lane_instr * frames / (kms * 1e-3) / (148 * 62.0 * 1.965e9),
                      "parity": f"first {chk} frames bit-exact vs oracle"}), flush=True)
    del llr, out
    dec.close()
    torch.cuda.empty_cache()
if a.sweep:
    name, n, k, _, _ = SETS["c5"]
    dec = scpd.Decoder(n, k, scpd.packed_flags(name, n))
    for snr in (1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0):
        c = dec.run_ber(snr, k / n, 2048)
        print(json.dumps({"config": "c5 sweep", "ebn0_db": snr, "frames": c[3], "bit_errors": c[0], "frame_errors": c[1],
                          "ber": c[0] / max(1, c[2]), "fer": c[1] / max(1, c[3])}), flush=True)
