#!/usr/bin/env python3
"""Quick device-timed throughput of scpd_decode for a config (development aid, not the bench)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sc_polar_decoder_hls_b200 as scpd

SETS = {"n2048": ("FB_N2048_K1024", 2048, 1024, 2.5), "c1": ("FB_N1024_K512", 1024, 512, 2.5), "c2": ("frozen_n_4096_k_3072", 4096, 3072, 3.5),
        "c3": ("frozen_n_32768_k_29492_snr_4_5", 32768, 29492, 4.5),
        "c4": ("frozen_n_131072_k_117964", 131072, 117964, 4.5),
        "c5": ("frozen_n_524288_k_262144", 524288, 262144, 2.0)}

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="c1")
ap.add_argument("--frames", type=int, default=1 << 16)
ap.add_argument("--prune", type=int, default=2)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--check", type=int, default=0, help="verify this many frames against the oracle")
a = ap.parse_args()
name, n, k, snr = SETS[a.cfg]
flags = scpd.packed_flags(name, n)
dec = scpd.Decoder(n, k, flags, pruning=a.prune)
llr = scpd.channel_generate(n, a.frames, scpd.sigma(snr, k / n))
out = torch.empty((a.frames, n // 32), dtype=torch.int32, device="cuda")
for _ in range(2):
    dec.decode(llr, out)
torch.cuda.synchronize()
if a.check:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import oracle_lib as ol
    idx = np.unique(np.concatenate([np.arange(min(a.check, a.frames)), np.arange(max(0, a.frames - a.check), a.frames)]))
    got = out.cpu().numpy().view(np.uint32)[idx]
    want = (ol.pack_bits(ol.decode_l2(n, 16, 8, 0, 1, flags, llr.cpu().numpy()[idx])) if a.prune == 3 else
            ol.decode_packed(n, 16, 8, 0, 1, flags, llr.cpu().numpy()[idx], threads=8))
    print("check", len(idx), "frames:", "OK" if (got == want).all() else "MISMATCH")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    dec.decode(llr, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
fps = a.frames / (ms * 1e-3)
print(f"{a.cfg} N={n} K={k} prune={a.prune} group={os.environ.get('SCPD_GROUP', '32')} frames={a.frames} "
      f"{ms:.3f} ms  {fps:.3e} frames/s  info {fps * k / 1e9:.2f} Gb/s  coded {fps * n / 1e9:.2f} Gb/s  "
      f"ops,fg={dec.schedule_stats()}")
