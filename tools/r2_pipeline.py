#!/usr/bin/env python3
"""Development aid: time the device channel generator alone and the whole Monte-Carlo loop (scpd_run_ber)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sc_polar_decoder_hls_b200 as scpd
SETS = {"c1": ("FB_N1024_K512", 1024, 512, 2.5), "c2": ("frozen_n_4096_k_3072", 4096, 3072, 3.5),
        "c3": ("frozen_n_32768_k_29492_snr_4_5", 32768, 29492, 4.5), "c4": ("frozen_n_131072_k_117964", 131072, 117964, 4.5)}
for key in sys.argv[1:] or ["c1", "c2"]:
    name, n, k, snr = SETS[key]
    nfr = min(1 << int(os.environ.get('NFR_LOG2', '20')), (1 << 34) // n)
    llr = scpd.channel_generate(n, nfr, scpd.sigma(snr, k / n))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        llr = scpd.channel_generate(n, nfr, scpd.sigma(snr, k / n))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    del llr
    torch.cuda.empty_cache()
    dec = scpd.Decoder(n, k, scpd.packed_flags(name, n))
    dec.run_ber(snr, k / n, nfr)
    t0 = time.perf_counter()
    cnt = dec.run_ber(snr, k / n, nfr)
    sec = time.perf_counter() - t0
    print(f"{key} frames={nfr} fast={os.environ.get('SCPD_CHANNEL_FAST','0')} channel {ms:.3f} ms ({nfr*n/2/ms/1e6:.1f} Gdraws/s)  run_ber {sec*1e3:.2f} ms = {nfr*k/sec/1e9:.1f} Gb/s info  counters {cnt[:2]}")
