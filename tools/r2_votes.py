#!/usr/bin/env python3
"""Share of R1 (all-information) nodes that a zero LLR in the warp sends to the full walk, per node size (scpd_r1_votes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sc_polar_decoder_hls_b200 as scpd
SETS = {"c1": ("FB_N1024_K512", 1024, 512, 2.5), "c2": ("frozen_n_4096_k_3072", 4096, 3072, 3.5),
        "c3": ("frozen_n_32768_k_29492_snr_4_5", 32768, 29492, 4.5)}
for key in sys.argv[1:] or ["c1", "c2", "c3"]:
    name, n, k, snr = SETS[key]
    dec = scpd.Decoder(n, k, scpd.packed_flags(name, n))
    llr = scpd.channel_generate(n, 148 * 16 * 32, scpd.sigma(snr, k / n))
    dec.stage_timing(True)
    dec.decode(llr)
    torch.cuda.synchronize()
    v, f = dec.r1_votes()
    per = ", ".join(f"2^{l}: {int(f[l])}/{int(v[l])}" for l in range(32) if v[l])
    print(f"{key} Eb/N0 {snr} dB: {int(f.sum())} of {int(v.sum())} R1 votes fall back ({100.0 * f.sum() / max(1, v.sum()):.1f} %)  [{per}]")
    dec.close()
