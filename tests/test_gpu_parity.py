"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on identical
quantised LLRs.  Bit-exact is the bar (integer/bit path)."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

CONFIG_SETS = {
    "c1": ("FB_N1024_K512", 1024, 512, 2.5),
    "c2": ("frozen_n_4096_k_3072", 4096, 3072, 3.5),
    "c3": ("frozen_n_32768_k_29492_snr_4_5", 32768, 29492, 4.5),
    "c4": ("frozen_n_131072_k_117964", 131072, 117964, 4.5),
    "c5": ("frozen_n_524288_k_262144", 524288, 262144, 2.0),
}


@pytest.fixture(scope="module")
def scpd():
    import torch
    assert torch.cuda.is_available()
    import sc_polar_decoder_hls_b200 as m
    return m


_ORACLE_CACHE = {}


def _oracle(n, par, q, ext, flags, llr, fmt=0):
    """Oracle output is independent of pruning mode and kernel variant: compute once per input."""
    key = (n, par, q, ext, fmt, flags.tobytes(), llr.shape, hash(llr.tobytes()))
    if key not in _ORACLE_CACHE:
        if len(_ORACLE_CACHE) > 64:
            _ORACLE_CACHE.clear()
        _ORACLE_CACHE[key] = ol.decode_packed(n, par, q, fmt, ext, flags, llr, threads=8)
    return _ORACLE_CACHE[key]


_LLR_CACHE = {}


def _llrs(seed, n, nfr, k, snr):
    key = (seed, n, nfr, k, snr)
    if key not in _LLR_CACHE:
        _LLR_CACHE[key] = ol.test_llrs(np.random.default_rng(seed), n, nfr, k, snr)
    return _LLR_CACHE[key]


def _check(scpd, name, n, k, par, q, ext, prune, llr, via="device", fmt=0):
    import torch
    flags = scpd.packed_flags(name, n)
    dec = scpd.Decoder(n, k, flags, par=par, llr_bits=q, fmt=fmt, extended=ext, pruning=prune)
    if via == "host":
        got = dec.decode_host(llr)
    else:
        x = dec.decode(torch.from_numpy(llr).cuda())
        torch.cuda.synchronize()
        got = x.cpu().numpy().view(np.uint32)
    want = _oracle(n, par, q, ext, flags, llr, fmt)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (f"{bad.size} of {len(llr)} frames differ (first {bad[:5]}) {name} par={par} q={q} ext={ext} "
                           f"prune={prune} fmt={fmt}")
    dec.close()


@pytest.fixture(params=["auto", "generic", "raw", "fast2", "fast8", "fast16", "fast32", "coop4", "coop8", "bs8", "bs16", "bs32", "bs32ws",
                        "ss", "ss12", "ss_notm", "ss_tm8", "ss_xf", "ss_pre0", "ss_pre3"])
def kernel_mode(request, monkeypatch):
    """Selects the decode kernel through the library's environment switches (read in scpd_create):
    the bit-sliced kernel with 8 / 16 / 32 lanes per frame group (bs32ws: partial sums pushed out to the
    workspace early, one warp per CTA), the int16x2 kernel with 2 / 8 / 16 lanes per frame pair, the
    generic kernel, the raw-pattern kernel, and the library's own choice."""
    mode = request.param
    for v in ("SCPD_KERNEL", "SCPD_GROUP", "SCPD_COOP", "SCPD_BS_GROUP", "SCPD_BS_LSB", "SCPD_BS_WARPS", "SCPD_SS_WARPS",
              "SCPD_SS_LTM", "SCPD_SS_LSA", "SCPD_SS_LWIN", "SCPD_SS_XF_MIN_LOG2N", "SCPD_SS_PRE"):
        monkeypatch.delenv(v, raising=False)
    if mode.startswith("ss"):
        # the slot-sliced kernel (a lane per frame): default plan (16 warps per CTA, one LLR level in tensor memory),
        # 12 warps (other tensor-memory column split), no tensor memory, level 8 in tensor memory with a small
        # partial-sum window
        monkeypatch.setenv("SCPD_KERNEL", "ss")
        if mode == "ss12":
            monkeypatch.setenv("SCPD_SS_WARPS", "12")
        elif mode == "ss_notm":
            monkeypatch.setenv("SCPD_SS_LTM", "0")
        elif mode == "ss_tm8":
            monkeypatch.setenv("SCPD_SS_LTM", "8")
            monkeypatch.setenv("SCPD_SS_LWIN", "8")
        elif mode == "ss_xf":      # f / g fused with the child's opening f wherever three levels stream through global memory
            monkeypatch.setenv("SCPD_SS_XF_MIN_LOG2N", "10")
        elif mode == "ss_pre0":    # every f level computed by the walk
            monkeypatch.setenv("SCPD_SS_PRE", "0")
        elif mode == "ss_pre3":    # as many leading f levels as the plane conversion can take
            monkeypatch.setenv("SCPD_SS_PRE", "3")
        return mode
    if mode in ("generic", "raw"):
        monkeypatch.setenv("SCPD_KERNEL", mode)
    elif mode.startswith("coop"):
        monkeypatch.setenv("SCPD_KERNEL", "fast")
        monkeypatch.setenv("SCPD_COOP", mode[4:])
    elif mode.startswith("fast"):
        monkeypatch.setenv("SCPD_KERNEL", "fast")
        monkeypatch.setenv("SCPD_GROUP", mode[4:])
    elif mode.startswith("bs"):
        monkeypatch.setenv("SCPD_KERNEL", "bs")
        monkeypatch.setenv("SCPD_BS_GROUP", mode[2:4].rstrip("w"))
        if mode.endswith("ws"):
            monkeypatch.setenv("SCPD_BS_LSB", "6")
            monkeypatch.setenv("SCPD_BS_WARPS", "1")
    return mode


@pytest.mark.parametrize("key,nfr", [("c1", 600), ("c2", 150), ("c3", 8)])
@pytest.mark.parametrize("prune", [0, 1, 2])
def test_every_kernel_variant(scpd, kernel_mode, key, nfr, prune):
    if key == "c3" and kernel_mode not in ("auto", "bs32", "bs32ws", "fast8", "coop4", "ss", "ss_tm8", "ss_xf", "ss_pre3"):
        pytest.skip("the large tree is covered by one variant per kernel family")
    name, n, k, snr = CONFIG_SETS[key]
    llr = _llrs(21, n, nfr, k, snr).copy()
    llr[-1] = 0
    llr[-2][::7] = 0
    _check(scpd, name, n, k, 16, 8, 1, prune, llr)


def test_golden_codewords_noiseless(scpd):
    """The reference's 9 stored codewords (sc_encoder.h:74-89) at sigma = 0 (LLR = +-4)."""
    cws = ol.golden_codewords()
    for key, name, n, k in (("cw8x4", "FB_N8_K4", 8, 4), ("cw512x256", "FB_N512_K256", 512, 256),
                            ("cw1024x512", "FB_N1024_K512", 1024, 512)):
        flags = scpd.packed_flags(name, n)
        llr = np.where(cws[key] == 1, -4, 4).astype(np.int8)
        for par in (2, 16):
            if 2 * par > n:
                continue
            for prune in (0, 1, 2):
                dec = scpd.Decoder(n, k, flags, par=par, pruning=prune)
                got = ol.unpack_bits(dec.decode_host(llr), n)
                assert (got == cws[key]).all(), (key, par, prune)


@pytest.mark.parametrize("prune", [0, 1, 2])
@pytest.mark.parametrize("par,q,ext", [(16, 8, 1), (16, 8, 0), (4, 8, 1), (64, 8, 1), (256, 8, 1), (16, 6, 1),
                                       (16, 6, 0), (16, 7, 1), (64, 7, 1), (64, 6, 1), (16, 9, 1), (2, 7, 1), (1, 8, 1)])
def test_c1_sweep(scpd, par, q, ext, prune):
    name, n, k, snr = CONFIG_SETS["c1"]
    llr = _llrs(par * 100 + q * 10 + ext, n, 400, k, snr)
    _check(scpd, name, n, k, par, q, ext, prune, llr)


@pytest.mark.parametrize("prune", [0, 2])
@pytest.mark.parametrize("par,q,ext", [(16, 6, 1), (16, 6, 0), (64, 6, 1), (64, 7, 1), (16, 7, 1), (16, 8, 1), (16, 8, 0),
                                       (64, 8, 1)])
def test_sigmag_sweep(scpd, par, q, ext, prune):
    """SIGMAG number format (config.h:11; the reference's checked-in default is SIGMAG, LLR_BITS 6): signed
    zero, tie rule of qfull_add_sub_sm, half-range saturation of g."""
    for key, nfr in (("c1", 300), ("c2", 70)):
        name, n, k, snr = CONFIG_SETS[key]
        llr = _llrs(par + q + ext, n, nfr, k, snr).copy()
        llr[-1] = 0
        llr[-2][::3] = 0
        _check(scpd, name, n, k, par, q, ext, prune, llr, fmt=scpd.FMT_SIGMAG)


def test_sigmag_differs_from_ca2_and_unsupported_combinations(scpd):
    """The two formats give different codewords on noisy frames (so the format is part of the contract);
    a SIGMAG configuration without a bit-sliced instantiation runs on the raw-pattern kernel, and what no
    kernel covers is refused, not approximated."""
    name, n, k, snr = CONFIG_SETS["c1"]
    flags = scpd.packed_flags(name, n)
    llr = _llrs(77, n, 400, k, 1.0)
    out = []
    for fmt in (scpd.FMT_CA2, scpd.FMT_SIGMAG):
        dec = scpd.Decoder(n, k, flags, par=16, llr_bits=6, fmt=fmt, extended=1)
        out.append(dec.decode_host(llr))
    assert (out[0] != out[1]).any()
    dec = scpd.Decoder(n, k, flags, par=4, llr_bits=9, fmt=scpd.FMT_SIGMAG, extended=1)
    assert "raw" in dec.kernel_name
    assert (dec.decode_host(llr) == ol.decode_packed(n, 4, 9, 1, 1, flags, llr, threads=8)).all()
    n2 = 1 << 21
    with pytest.raises(scpd.ScpdError) as e:
        scpd.Decoder(n2, n2 // 2, np.tile(np.array([0, 1], np.uint8), n2 // 2))
    assert e.value.status == scpd.E_UNSUPPORTED


@pytest.mark.parametrize("fmt", [0, 1])
def test_raw_kernel_configurations(scpd, fmt):
    """Everything the reference can be configured to outside the in-range datapaths (decode_raw.cuh): LLR_BITS 5
    (the +-31 alphabet wraps modulo 32 on the sc_fifo<LLR> write), SIGMAG for arbitrary (LLR_BITS, PAR, EXTENDED),
    N below 128, un-saturated leaves wider than 16 bits, the whole int8 input range."""
    import torch
    rng = np.random.default_rng(50 + fmt)
    for key, nfr in (("c1", 300), ("c2", 60)):
        name, n, k, snr = CONFIG_SETS[key]
        flags = scpd.packed_flags(name, n)
        llr = _llrs(31 + fmt, n, nfr, k, snr).copy()
        llr[-1] = 0
        llr[-2] = rng.integers(-128, 128, n).astype(np.int8)
        cases = [(16, 5, 1), (16, 5, 0), (64, 5, 1), (256, 9, 1), (1, 5, 1)]
        if fmt == 1:
            cases += [(4, 9, 1), (256, 8, 1), (1, 6, 1), (2, 8, 0), (16, 9, 0), (16, 7, 0)]
        for par, q, ext in cases:
            for prune in (0, 2):
                dec = scpd.Decoder(n, k, flags, par=par, llr_bits=q, fmt=fmt, extended=ext, pruning=prune)
                assert "raw" in dec.kernel_name, (par, q, ext, dec.kernel_name)
                want = _oracle(n, par, q, ext, flags, llr, fmt)
                got = dec.decode(torch.from_numpy(llr).cuda()).cpu().numpy().view(np.uint32)
                assert (got == want).all(), (key, par, q, ext, prune)
                dec.close()
    for n in (4, 8, 32, 64):  # below the bit-sliced kernel's smallest tree
        for trial in range(3):
            flags = (rng.random(n) < rng.random()).astype(np.uint8)
            llr = rng.integers(-128, 128, size=(70, n)).astype(np.int8)
            for par, q in ((1, 6), (2, 5), (2, 8), (16, 6), (32, 9)):
                if 2 * par > n or (fmt == 0 and q > 5):
                    continue
                dec = scpd.Decoder(n, int(flags.sum()), flags, par=par, llr_bits=q, fmt=fmt, pruning=trial)
                assert "raw" in dec.kernel_name
                assert (dec.decode_host(llr) == ol.decode_packed(n, par, q, fmt, 1, flags, llr)).all(), (n, par, q, trial)
                dec.close()
    if fmt == 1:  # mis-aligned LLR buffer: the bit-sliced SIGMAG handle falls back to the raw-pattern kernel
        name, n, k, snr = CONFIG_SETS["c1"]
        flags = scpd.packed_flags(name, n)
        llr = _llrs(9, n, 100, k, snr)
        dec = scpd.Decoder(n, k, flags, llr_bits=6, fmt=1)
        assert "bit-sliced" in dec.kernel_name
        buf = torch.zeros(llr.size + 1, dtype=torch.int8, device="cuda")
        buf[1:] = torch.from_numpy(llr).cuda().flatten()
        got = dec.decode(buf[1:].view(100, n)).cpu().numpy().view(np.uint32)
        assert (got == ol.decode_packed(n, 16, 6, 1, 1, flags, llr)).all()


@pytest.mark.parametrize("fmt", [0, 1])
def test_reference_pruning_level2_mode(scpd, fmt):
    """SCPD_PRUNE_REF_LEVEL2: the decoder the reference is as checked in (PRUNING_LEVEL 2: R0 / R1 / REP / SPC / H0),
    bit-exact against sco_decode_l2, which tests/test_oracle.py pins on eleven builds of the reference's own sources.
    It is not plain SC: the same frames decoded with SCPD_PRUNE_R0_R1 differ on some of the noisy ones."""
    import torch
    rng = np.random.default_rng(90 + fmt)
    differs = 0
    for key, nfr in (("c1", 400), ("c2", 120), ("c3", 6)):
        name, n, k, snr = CONFIG_SETS[key]
        flags = scpd.packed_flags(name, n)
        cases = [(16, 8, 1), (16, 6, 1), (4, 8, 1), (64, 7, 0), (256, 9, 1), (2, 8, 1)] if key != "c3" else [(16, 8, 1)]
        for par, q, ext in cases:
            m = min(2 ** (q - 1) - 1, 127)
            llr = np.concatenate([ol.channel(n, nfr, ol.sigma(snr - 1.5, k / n)),
                                  rng.integers(-m, m + 1, size=(nfr // 4 + 1, n)).astype(np.int8),
                                  rng.integers(-2, 3, size=(nfr // 4 + 1, n)).astype(np.int8), np.zeros((1, n), np.int8)])
            dec = scpd.Decoder(n, k, flags, par=par, llr_bits=q, fmt=fmt, extended=ext, pruning=scpd.PRUNE_REF_LEVEL2)
            assert "raw" in dec.kernel_name
            want = ol.pack_bits(ol.decode_l2(n, par, q, fmt, ext, flags, llr))
            got = dec.decode(torch.from_numpy(llr).cuda()).cpu().numpy().view(np.uint32)
            assert (got == want).all(), (key, par, q, ext)
            assert (dec.decode_host(llr[:33]) == want[:33]).all()
            dec.close()
            if (par, q, ext) == (16, 8, 1) and key != "c3":
                plain = scpd.Decoder(n, k, flags, fmt=fmt).decode_host(llr)
                differs += int((plain != want).any(axis=1).sum())
    assert differs > 0
    # the Monte-Carlo loop in that mode: the reference testbench as it ships (c1, golden codeword cycle)
    name, n, k, snr = CONFIG_SETS["c1"]
    flags = scpd.packed_flags(name, n)
    cws = ol.golden_codewords()["cw1024x512"]
    q = 6 if fmt else 8
    dec = scpd.Decoder(n, k, flags, llr_bits=q, fmt=fmt, pruning=scpd.PRUNE_REF_LEVEL2)
    got = dec.run_ber(1.5, k / n, 3000, codeword=cws[0])
    llr = ol.channel(n, 3000, ol.sigma(1.5, k / n), codeword=cws[0])
    want = ol.count_errors(n, ol.decode_l2(n, 16, q, fmt, 1, flags, llr), ref=cws[0])
    assert list(got[:4]) == want[:4] and want[1] > 0


def test_c1_headline_many_frames(scpd):
    """>= 10^5 noisy frames of BASELINE config 1 at 2.5 dB (BASELINE.md parity gate), odd batch."""
    name, n, k, snr = CONFIG_SETS["c1"]
    import torch
    flags = scpd.packed_flags(name, n)
    nfr = 100001
    cw = ol.golden_codewords()["cw1024x512"]
    llr = scpd.channel_generate(n, nfr, scpd.sigma(snr, k / n), codeword=cw[1])
    dec = scpd.Decoder(n, k, flags)
    got = dec.decode(llr).cpu().numpy().view(np.uint32)
    want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr.cpu().numpy(), threads=8)
    assert (got == want).all()
    # and therefore identical error counts
    ref = ol.pack_bits(cw[1][None, :])[0]
    c_dev = scpd.count_errors(n, torch.from_numpy(got.view(np.int32)).cuda(), ref)
    c_or = ol.count_errors(n, ol.unpack_bits(want, n), cw[1])
    assert c_dev == c_or


@pytest.mark.parametrize("key,nfr", [("c2", 300), ("c3", 24), ("c4", 6), ("c5", 3)])
@pytest.mark.parametrize("prune", [0, 2])
def test_baseline_configs(scpd, key, nfr, prune):
    name, n, k, snr = CONFIG_SETS[key]
    llr = _llrs(5, n, nfr, k, snr)
    _check(scpd, name, n, k, 16, 8, 1, prune, llr)


@pytest.mark.parametrize("key,nfr", [("c2", 100000), ("c3", 4096), ("c4", 512), ("c5", 128)])
def test_baseline_configs_parity_gate(scpd, key, nfr):
    """the parity gate before any timing counts -- >= 10^5 channel frames at c2 (c1 has its own
    test), 4096 / 512 / 128 at c3 / c4 / c5 (the oracle needs seconds per frame of N = 2^19), device-generated LLRs at the
    configuration's Eb/N0, default kernel choice at that batch size, every frame compared."""
    import torch
    name, n, k, snr = CONFIG_SETS[key]
    flags = scpd.packed_flags(name, n)
    llr = scpd.channel_generate(n, nfr, scpd.sigma(snr, k / n))
    dec = scpd.Decoder(n, k, flags)
    got = dec.decode(llr).cpu().numpy().view(np.uint32)
    torch.cuda.synchronize()
    want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr.cpu().numpy(), threads=16)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (key, dec.last_kernel_name, bad[:5])


def test_zero_llr_fallback(scpd):
    """All-zero and sparse LLR frames force the rate-1 plain-SC fallback on every node (G3/G10)."""
    name, n, k, _ = CONFIG_SETS["c1"]
    rng = np.random.default_rng(11)
    llr = rng.integers(-31, 32, size=(64, n)).astype(np.int8)
    llr[rng.random(llr.shape) < 0.6] = 0
    llr[0] = 0
    llr[1] = -1
    llr[2] = 127
    llr[3] = -127
    for prune in (1, 2):
        _check(scpd, name, n, k, 16, 8, 1, prune, llr)


@pytest.mark.parametrize("group", ["16", "32"])
def test_bit_sliced_kernel_edge_cases(scpd, monkeypatch, group):
    """The batch-size dispatch sends small batches to the int16x2 kernel, so the edge cases are repeated here
    with the bit-sliced kernel pinned: arbitrary (non-polar) flag tables, zero-heavy and full-range LLRs
    (the CA2 rate-1 fallback on almost every node), ragged last groups, one-frame batches."""
    monkeypatch.setenv("SCPD_KERNEL", "bs")
    monkeypatch.setenv("SCPD_BS_GROUP", group)
    rng = np.random.default_rng(int(group))
    for n in (128, 256, 1024):
        for trial in range(3):
            flags = (rng.random(n) < rng.random()).astype(np.uint8)
            llr = rng.integers(-127, 128, size=(77, n)).astype(np.int8)
            llr[rng.random(llr.shape) < 0.3] = 0
            llr[0] = 0
            llr[1] = 127
            llr[2] = -127
            for prune in (0, 1, 2):
                dec = scpd.Decoder(n, int(flags.sum()), flags, par=16, pruning=prune)
                assert "bit-sliced" in dec.kernel_name
                assert (dec.decode_host(llr) == ol.decode_packed(n, 16, 8, 0, 1, flags, llr)).all(), (n, trial, prune)
                dec.close()
    name, n, k, snr = CONFIG_SETS["c1"]
    flags = scpd.packed_flags(name, n)
    for fmt, q in ((scpd.FMT_CA2, 8), (scpd.FMT_SIGMAG, 6)):
        dec = scpd.Decoder(n, k, flags, llr_bits=q, fmt=fmt)
        for nfr in (1, 2, 31, 32, 33, 65, 257):
            llr = ol.test_llrs(rng, n, nfr, k, snr)
            llr[-1][::3] = 0
            assert (dec.decode_host(llr) == ol.decode_packed(n, 16, q, fmt, 1, flags, llr)).all(), (fmt, nfr)
        dec.close()


def test_slot_sliced_kernel_edge_cases(scpd, monkeypatch):
    """The slot-sliced kernel pinned (small batches would otherwise go to the int16x2 kernel): arbitrary (non-polar)
    flag tables -- the run-time fallback of the fp16x2 walker --, zero-heavy and full-range LLRs (the CA2 rate-1
    fallback on almost every node), ragged last tasks, one-frame batches, the (LLR_BITS, PAR, EXTENDED) variants."""
    monkeypatch.setenv("SCPD_KERNEL", "ss")
    rng = np.random.default_rng(2)
    for n in (128, 256, 1024):
        for trial in range(3):
            flags = (rng.random(n) < rng.random()).astype(np.uint8)
            llr = rng.integers(-127, 128, size=(77, n)).astype(np.int8)
            llr[rng.random(llr.shape) < 0.3] = 0
            llr[0] = 0
            llr[1] = 127
            llr[2] = -127
            for prune in (0, 1, 2):
                dec = scpd.Decoder(n, int(flags.sum()), flags, par=16, pruning=prune)
                assert "slot-sliced" in dec.kernel_name
                assert (dec.decode_host(llr) == ol.decode_packed(n, 16, 8, 0, 1, flags, llr)).all(), (n, trial, prune)
                dec.close()
    name, n, k, snr = CONFIG_SETS["c1"]
    flags = scpd.packed_flags(name, n)
    dec = scpd.Decoder(n, k, flags)
    for nfr in (1, 2, 31, 32, 33, 65, 257, 1000):
        llr = ol.test_llrs(rng, n, nfr, k, snr)
        llr[-1][::3] = 0
        assert (dec.decode_host(llr) == ol.decode_packed(n, 16, 8, 0, 1, flags, llr)).all(), nfr
    assert "slot-sliced" in dec.last_kernel_name
    dec.close()
    for par, q, ext in ((16, 8, 0), (16, 7, 1), (16, 6, 1), (16, 6, 0), (4, 8, 1), (8, 8, 1), (32, 7, 1), (32, 6, 1)):
        ma = (1 << (q - 1)) - 1
        llr = ol.test_llrs(rng, n, 300, k, snr, maxabs=min(31, ma))
        llr[150:] = rng.integers(-ma, ma + 1, size=(150, n))
        llr[-1] = 0
        for prune in (0, 2):
            dec = scpd.Decoder(n, k, flags, par=par, llr_bits=q, extended=ext, pruning=prune)
            assert "slot-sliced" in dec.kernel_name, (par, q, ext)
            assert (dec.decode_host(llr) == ol.decode_packed(n, par, q, 0, ext, flags, llr, threads=8)).all(), (par, q, ext, prune)
            dec.close()


@pytest.mark.parametrize("n,xf_min,pre", [(4096, 12, 2), (4096, 12, 0), (8192, 13, 3), (32768, 15, 2), (32768, 99, 1), (65536, 15, 2)])
def test_slot_sliced_large_tree_ops_on_arbitrary_tables(scpd, monkeypatch, n, xf_min, pre):
    """The two things the slot-sliced kernel does beyond the plain schedule -- the leading f levels computed by the plane
    conversion (SCPD_SS_PRE) and f / g fused with the child's opening f (SS_XF_*, forced onto small trees here) -- on
    flag tables no construction produces: random densities, an all-frozen first half or quarter (no leading f), an
    all-information second half (R1 with its fallback walk), every pruning mode, LLRs with many CA2 zeros, ragged batch."""
    monkeypatch.setenv("SCPD_KERNEL", "ss")
    monkeypatch.setenv("SCPD_SS_XF_MIN_LOG2N", str(xf_min))
    monkeypatch.setenv("SCPD_SS_PRE", str(pre))
    rng = np.random.default_rng(n + xf_min + pre)
    nfr = 37 if n <= 8192 else 33
    for trial in range(4):
        dens = rng.random(n // 256).repeat(256)              # blocks of 256 positions with their own density
        flags = (rng.random(n) < dens).astype(np.uint8)
        if trial == 1:
            flags[:n // 2] = 0
        if trial == 2:
            flags[:n // 4] = 0
            flags[n // 2:] = 1
        if trial == 3:
            flags[n // 8:n // 4] = 1
        llr = rng.integers(-127, 128, size=(nfr, n)).astype(np.int8)
        llr[rng.random(llr.shape) < 0.2] = 0
        llr[0] = 0
        want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr, threads=8)
        for prune in (0, 1, 2):
            dec = scpd.Decoder(n, int(flags.sum()), flags, par=16, pruning=prune)
            assert "slot-sliced" in dec.kernel_name
            assert (dec.decode_host(llr) == want).all(), (trial, prune)
            dec.close()


@pytest.mark.parametrize("q,ext", [(8, 0), (7, 1), (6, 1)])
def test_slot_sliced_fused_ops_other_configurations(scpd, monkeypatch, q, ext):
    """The instantiations of the kernel with the fused SS_XF_* ops other than the headline one (LLR_BITS 6 / 7, EXTENDED 0),
    forced onto c2 and used by default at c3; full internal range of the saturation, zeros, every pruning mode."""
    monkeypatch.setenv("SCPD_KERNEL", "ss")
    rng = np.random.default_rng(q * 2 + ext)
    ma = (1 << (q - 1)) - 1
    for key, nfr, xf_min in (("c2", 70, "12"), ("c3", 6, "15")):
        monkeypatch.setenv("SCPD_SS_XF_MIN_LOG2N", xf_min)
        name, n, k, snr = CONFIG_SETS[key]
        flags = scpd.packed_flags(name, n)
        llr = ol.test_llrs(rng, n, nfr, k, snr, maxabs=min(31, ma))
        llr[nfr // 2:] = rng.integers(-ma, ma + 1, size=(nfr - nfr // 2, n))
        llr[-1][::3] = 0
        want = ol.decode_packed(n, 16, q, 0, ext, flags, llr, threads=8)
        for prune in (0, 2):
            dec = scpd.Decoder(n, k, flags, par=16, llr_bits=q, extended=ext, pruning=prune)
            assert "slot-sliced" in dec.kernel_name
            assert (dec.decode_host(llr) == want).all(), (key, prune)
            dec.close()


@pytest.mark.parametrize("key,nfr", [("c4", 40), ("c5", 9)])
def test_slot_sliced_kernel_on_the_largest_trees(scpd, monkeypatch, key, nfr):
    """c4 / c5 on the slot-sliced kernel with its fused ops (what scpd_decode launches for large batches of them; pinned
    here because the oracle needs a second per frame of N = 2^19): channel frames, zero-heavy frames, ragged task."""
    monkeypatch.setenv("SCPD_KERNEL", "ss")
    name, n, k, snr = CONFIG_SETS[key]
    flags = scpd.packed_flags(name, n)
    llr = _llrs(5, n, nfr, k, snr).copy()
    llr[-1] = 0
    llr[-2][::5] = 0
    dec = scpd.Decoder(n, k, flags)
    got = dec.decode_host(llr)
    assert "slot-sliced" in dec.last_kernel_name
    assert (got == _oracle(n, 16, 8, 1, flags, llr)).all()
    dec.close()


def test_slot_sliced_is_the_default_for_large_batches(scpd):
    """scpd_decode's kernel choice for CA2 up to N = 2^14: the slot-sliced kernel once the batch gives every SM a few
    warps, the int16x2 kernel below -- and the same bits whichever kernel runs."""
    import torch
    name, n, k, snr = CONFIG_SETS["c2"]
    flags = scpd.packed_flags(name, n)
    dec = scpd.Decoder(n, k, flags)
    llr = scpd.channel_generate(n, 32768, scpd.sigma(snr, k / n))
    want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr[:64].cpu().numpy(), threads=8)
    seen = {}
    for nfr in (32768, 2048):
        out = dec.decode(llr[:nfr])
        torch.cuda.synchronize()
        seen[nfr] = dec.last_kernel_name
        assert (out[:64].cpu().numpy().view(np.uint32) == want).all(), nfr
    assert "slot-sliced" in seen[32768] and "int16x2" in seen[2048], seen


def test_channel_guarded_fast_path_gives_the_same_llrs(scpd):
    """scpd_channel_mode: the default (2) evaluates Box-Muller with the SFU approximations and re-runs the libm-grade code
    for samples next to a quantiser bin edge; it must produce the LLRs of mode 0 (libm-grade only) sample for sample --
    here 2^29 of them over three noise levels, stored codewords and the all-zero one -- while mode 1 (approximations
    only) differs on a few, which is what the guard band is for.  Mode 0 itself is checked against the host chain."""
    import torch
    cw = ol.golden_codewords()["cw1024x512"]
    prev = scpd.channel_mode()
    assert prev == 2
    try:
        unguarded = 0
        for n, nfr, snr, cwsel in ((1024, 1 << 17, 0.5, cw[1]), (4096, 1 << 15, 3.5, None), (1024, 1 << 17, 6.0, cw[:2].repeat(1 << 16, axis=0))):
            out = {}
            for mode in (0, 2, 1):
                scpd.channel_mode(mode)
                out[mode] = scpd.channel_generate(n, nfr, scpd.sigma(snr, 0.5), first_frame=12345, codeword=cwsel)
            torch.cuda.synchronize()
            assert torch.equal(out[0], out[2]), (n, snr)
            unguarded += int((out[0] != out[1]).sum())
            host = ol.channel(n, 64, ol.sigma(snr, 0.5), first_frame=12345,
                              codeword=None if cwsel is None else (cwsel if cwsel.ndim == 1 else cwsel[:64]))
            assert (out[0][:64].cpu().numpy() == host).all()
        assert 0 < unguarded < 20000  # ~1e-5 of 4e8 samples
    finally:
        scpd.channel_mode(prev)


def test_validate_llr_counts_contract_violations(scpd):
    import torch
    name, n, k, snr = CONFIG_SETS["c1"]
    flags = scpd.packed_flags(name, n)
    rng = np.random.default_rng(9)
    llr = rng.integers(-31, 32, size=(77, n)).astype(np.int8)
    llr[3, 5] = 100
    llr[70, 1000] = -128
    llr[76, 1023] = -32
    for q, want in ((8, 1), (6, 3), (7, 2)):
        dec = scpd.Decoder(n, k, flags, llr_bits=q)
        t = torch.from_numpy(llr).cuda()
        assert dec.validate_llr(t) == want, q
        assert dec.validate_llr(t.flatten()[1:n * 76 + 1].view(76, n)) in (want, want - 1)  # mis-aligned view
        dec.close()


def test_ragged_and_empty_batches(scpd):
    name, n, k, snr = CONFIG_SETS["c1"]
    flags = scpd.packed_flags(name, n)
    dec = scpd.Decoder(n, k, flags)
    assert dec.decode_host(np.zeros((0, n), np.int8)).shape == (0, n // 32)
    rng = np.random.default_rng(3)
    for nfr in (1, 2, 3, 7, 8, 9, 31, 33, 257):
        llr = ol.test_llrs(rng, n, nfr, k, snr)
        got = dec.decode_host(llr)
        want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr)
        assert (got == want).all(), nfr


def test_host_path_equals_device_path(scpd):
    name, n, k, snr = CONFIG_SETS["c2"]
    rng = np.random.default_rng(4)
    llr = ol.test_llrs(rng, n, 130, k, snr)
    _check(scpd, name, n, k, 16, 8, 1, 2, llr, via="host")


def test_arbitrary_flag_patterns(scpd):
    """Frozen sets that are not polar-structured (any 0/1 table is a legal FB stream)."""
    rng = np.random.default_rng(9)
    import torch
    for n in (4, 8, 32, 128, 512):
        for trial in range(4):
            flags = (rng.random(n) < rng.random()).astype(np.uint8)
            llr = rng.integers(-31, 32, size=(50, n)).astype(np.int8)
            for par in (1, 2, 16):
                if 2 * par > n:
                    continue
                for prune in (0, 1, 2):
                    dec = scpd.Decoder(n, int(flags.sum()), flags, par=par, pruning=prune)
                    got = dec.decode_host(llr)
                    want = ol.decode_packed(n, par, 8, 0, 1, flags, llr)
                    assert (got == want).all(), (n, trial, par, prune)


def test_device_channel_matches_oracle_chain(scpd):
    """Integer RNG stream is exact by construction; float libm differs in the last ulp, so LLRs
    may differ by one quantisation step on a tiny fraction of samples (SURVEY 'Channel parity')."""
    for n, nfr, first in ((1024, 64, 0), (1024, 5, 100003), (64, 9, 17), (8, 40, 0), (4096, 6, 54321)):  # the oracle steps, the device jumps
        sig = scpd.sigma(2.5, 0.5)
        cw = (np.arange(n) % 3 == 0).astype(np.uint8)
        dev = scpd.channel_generate(n, nfr, sig, first_frame=first, codeword=cw).cpu().numpy()
        ora = ol.channel(n, nfr, ol.sigma(2.5, 0.5), first_frame=first, codeword=cw)
        diff = dev.astype(int) - ora.astype(int)
        assert np.abs(diff).max() <= 1, (n, first)
        assert (diff != 0).mean() < 2e-3, (n, first, (diff != 0).mean())


def test_run_ber_counts(scpd):
    """Whole device Monte-Carlo loop vs generate(device) -> oracle decode -> oracle count."""
    name, n, k, snr = CONFIG_SETS["c1"]
    flags = scpd.packed_flags(name, n)
    cw = ol.golden_codewords()["cw1024x512"][0]
    dec = scpd.Decoder(n, k, flags)
    nfr = 4000
    cnt = dec.run_ber(snr, k / n, nfr, first_frame=77, codeword=cw)
    llr = scpd.channel_generate(n, nfr, scpd.sigma(snr, k / n), first_frame=77, codeword=cw).cpu().numpy()
    want = ol.count_errors(n, ol.decode(n, 16, 8, 0, 1, flags, llr), cw)
    assert cnt == want
    assert cnt[3] == nfr and cnt[2] == nfr * n and 0 < cnt[1] < nfr // 10


def test_run_ber_ex_codeword_cycle_and_random_payload(scpd):
    """scpd_run_ber_ex against the same chain taken apart: (i) the reference's encoder for N = 1024 -- its three stored
    codewords in turn (sc_encoder.h:91-113) --, (ii) random information words encoded on the device.  Codeword-bit
    counters as the reference counts them and information-bit counters, against generate(device) -> oracle decode ->
    host counts; batches start at an odd stream position."""
    import torch
    name, n, k, snr = CONFIG_SETS["c1"]
    flags = scpd.packed_flags(name, n)
    cws = ol.golden_codewords()["cw1024x512"]
    dec = scpd.Decoder(n, k, flags)
    nfr, first = 3000, 77

    def host_counts(x):  # x: uint8 [nfr, n] codewords actually sent
        llr = scpd.channel_generate(n, nfr, scpd.sigma(snr, k / n), first_frame=first, codeword=x).cpu().numpy()
        xhat = ol.decode(n, 16, 8, 0, 1, flags, llr)
        want = ol.count_errors(n, xhat, x)
        du = ol.polar_transform(xhat ^ x)[:, flags == 1]
        e = du.sum(axis=1)
        return want + [int(e.sum()), int((e != 0).sum()), nfr * k, nfr]

    sent = cws[(first + np.arange(nfr)) % 3]
    got = dec.run_ber_ex(snr, k / n, nfr, first_frame=first, codewords=cws)
    assert got == host_counts(sent)
    assert 0 < got[1] < nfr // 10 and 0 < got[7] <= got[1]
    # one stored codeword and the all-zero codeword through the same entry point
    assert dec.run_ber_ex(snr, k / n, nfr, first_frame=first, codewords=cws[1])[:6] == dec.run_ber(snr, k / n, nfr, first_frame=first, codeword=cws[1])
    assert dec.run_ber_ex(snr, k / n, nfr, first_frame=first)[:6] == dec.run_ber(snr, k / n, nfr, first_frame=first)
    # random payloads: restate the source on the host, encode with the oracle's transform
    u = ol.unpack_bits(scpd.payload_words(5, first, nfr, flags), n)
    assert (u[:, flags == 0] == 0).all() and 0.45 < u[:, flags == 1].mean() < 0.55
    x = ol.polar_transform(u)
    got = dec.run_ber_ex(snr, k / n, nfr, first_frame=first, random_payload=True, payload_seed=5)
    assert got == host_counts(x)
    dec.close()


def test_run_ber_one_llr_buffer_for_large_trees(scpd):
    """scpd_run_ber_ex on a large tree with more frames than one batch: ONE LLR staging buffer, refilled by the generator as
    soon as the plane conversion of the previous batch is through, batches grown towards a full round and split evenly
    (c3: 80 000 frames = 2 x 40 000).  The counters must equal the sum over the two halves run as single batches (the
    path test_run_ber_counts pins on the oracle), also with per-frame references (random payloads), and the handle must
    still serve scpd_decode_host (two staging buffers) and the loop again afterwards."""
    name, n, k, snr = CONFIG_SETS["c3"]
    flags = scpd.packed_flags(name, n)
    nfr, first = 80000, 4321
    for kw in ({}, {"random_payload": True, "payload_seed": 3}):
        dec = scpd.Decoder(n, k, flags)  # a fresh handle: the first call of the handle is the one with one LLR buffer
        whole = dec.run_ber_ex(snr, k / n, nfr, first_frame=first, **kw)
        h1 = dec.run_ber_ex(snr, k / n, nfr // 2, first_frame=first, **kw)
        h2 = dec.run_ber_ex(snr, k / n, nfr // 2, first_frame=first + nfr // 2, **kw)
        assert whole == [a + b for a, b in zip(h1, h2)], kw
        assert whole[3] == nfr and 0 < whole[1] < nfr and 0 < whole[7] <= whole[1]
        dec.close()
    dec = scpd.Decoder(n, k, flags)
    assert dec.run_ber_ex(snr, k / n, nfr, first_frame=first, **kw) == whole
    llr = _llrs(5, n, 40, k, 4.5)
    assert (dec.decode_host(llr) == _oracle(n, 16, 8, 1, flags, llr)).all()
    assert dec.run_ber_ex(snr, k / n, nfr, first_frame=first, **kw) == whole
    dec.close()


@pytest.mark.parametrize("n", [8, 32, 256, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 524288])
def test_run_ber_counters_and_transform_every_frame_size(scpd, n):
    """The ten counters of scpd_run_ber_ex and scpd_extract_info for every instantiation of the register-resident
    kernels (count_all_kernel / polar_transform_reg_kernel: 1 ... 32 words per lane, fewer than 32 words per frame), for
    the shared-memory kernels of the frames above 32768 bits (with and without the opt-in above 48 KB) and for n < 32,
    which keeps the multi-pass kernels.  The decode is not what is
    tested here (the product's own output is the input of the host counts): all-zero codeword, one stored codeword,
    random payloads, on a table that is not polar-structured, at a noise level that leaves errors in most frames."""
    import torch
    rng = np.random.default_rng(n)
    flags = (rng.random(n) < 0.5).astype(np.uint8)
    flags[-1] = 1
    k = int(flags.sum())
    nfr, first, snr = (300 if n <= 4096 else 96 if n <= 65536 else 40), 1234567, 1.0
    dec = scpd.Decoder(n, k, flags, par=min(16, n // 2))

    def host_counts(x):  # x: None, [n] or [nfr, n]
        llr = scpd.channel_generate(n, nfr, scpd.sigma(snr, k / n), first_frame=first, codeword=x).cpu().numpy()
        xhat = ol.unpack_bits(dec.decode_host(llr), n)
        sent = np.zeros((nfr, n), np.uint8) if x is None else np.broadcast_to(x, (nfr, n))
        want = ol.count_errors(n, xhat, sent)
        e = ol.polar_transform(xhat ^ sent)[:, flags == 1].sum(axis=1)
        return want + [int(e.sum()), int((e != 0).sum()), nfr * k, nfr]

    assert dec.run_ber_ex(snr, k / n, nfr, first_frame=first) == host_counts(None)
    cw = ol.polar_transform((rng.integers(0, 2, n).astype(np.uint8) & flags)[None, :])[0]
    got = dec.run_ber_ex(snr, k / n, nfr, first_frame=first, codewords=cw)
    assert got == host_counts(cw)
    assert got[1] > 0 and got[7] > 0
    x = ol.polar_transform(ol.unpack_bits(scpd.payload_words(9, first, nfr, flags), n))
    assert dec.run_ber_ex(snr, k / n, nfr, first_frame=first, random_payload=True, payload_seed=9) == host_counts(x)
    # the transform on its own
    w = rng.integers(0, 1 << 32, size=(nfr, max(1, n // 32)), dtype=np.uint64).astype(np.uint32)
    if n < 32:
        w &= np.uint32((1 << n) - 1)
    got_u = dec.extract_info(torch.from_numpy(w.view(np.int32)).cuda()).cpu().numpy().view(np.uint32)
    assert (got_u == ol.pack_bits(ol.polar_transform(ol.unpack_bits(w, n)))).all()
    dec.close()


def test_stage_time_matrix(scpd):
    """scpd_stage_time: measured cycles per function x level.  Visits must equal the schedule's op counts
    times the number of profiled warps; cycles must be positive exactly where there are visits; the 64-LLR nodes (row R)
    and the f / g rows carry most of the time."""
    import torch
    name, n, k, snr = CONFIG_SETS["c1"]
    flags = scpd.packed_flags(name, n)
    dec = scpd.Decoder(n, k, flags)
    llr = scpd.channel_generate(n, 148 * 16 * 32, scpd.sigma(snr, k / n))
    dec.stage_timing(True)
    dec.decode(llr)
    torch.cuda.synchronize()
    cyc, vis = dec.stage_time()
    assert "slot-sliced" in dec.last_kernel_name
    assert ((cyc > 0) == (vis > 0)).all()
    warps = int(vis[scpd.STAGE_FUNCS.index("H"), 10])  # one H at the root per profiled warp and task
    assert warps >= 1 and vis[scpd.STAGE_FUNCS.index("R"), 6] == 15 * warps  # 15 nodes of 64 LLRs are decoded (c1, R0+R1)
    assert vis[scpd.STAGE_FUNCS.index("F"), 10] == warps and vis[scpd.STAGE_FUNCS.index("G"), 10] == warps
    share = cyc.sum(axis=1) / cyc.sum()
    assert share[scpd.STAGE_FUNCS.index("R")] > 0.3 and share[scpd.STAGE_FUNCS.index("G")] > share[scpd.STAGE_FUNCS.index("H")]
    cyc2, vis2 = dec.stage_time()
    assert cyc2.sum() == 0 and vis2.sum() == 0  # read clears
    dec.stage_timing(False)
    dec.close()


def test_r1_votes_count_the_zero_llr_fallbacks(scpd):
    """scpd_r1_votes: the measured side of the pruning statistics.  Every R1 node the profiled warps reach is a vote; a
    zero LLR in any of the warp's 32 frames turns it into a fallback (full walk).  LLRs without a zero: no fallback; all-zero
    LLRs: every vote falls back; channel LLRs (trunc(4 y) is 0 for |y| < 0.25): somewhere in between, and the same bits."""
    import torch
    name, n, k, snr = CONFIG_SETS["c1"]
    flags = scpd.packed_flags(name, n)
    dec = scpd.Decoder(n, k, flags)
    nfr = 148 * 16 * 32
    llr = scpd.channel_generate(n, nfr, scpd.sigma(snr, k / n))
    want = dec.decode(llr).clone()
    dec.stage_timing(True)
    shares = []
    for x in (torch.where(llr == 0, torch.ones_like(llr), llr), torch.zeros_like(llr), llr):
        got = dec.decode(x)
        torch.cuda.synchronize()
        v, f = dec.r1_votes()
        assert v.sum() > 0 and (f <= v).all() and v[:5].sum() == 0
        shares.append(float(f.sum()) / float(v.sum()))
    assert torch.equal(got, want)  # the profiled build decodes the same bits
    assert shares[0] == 0.0 and shares[1] == 1.0 and 0.0 < shares[2] < 1.0
    v, f = dec.r1_votes()
    assert v.sum() == 0 and f.sum() == 0  # read clears
    dec.stage_timing(False)
    dec.close()


def test_extract_info_and_roundtrip_full_size(scpd):
    """Size-independent property at BASELINE scale (c2, 2^17 frames = 512 MiB of LLRs): random
    information words -> polar transform -> noiseless LLRs -> decode returns the codeword, and
    u^ = x^ F^(x)n returns the information word (encode -> decode round trip)."""
    import torch
    name, n, k, _ = CONFIG_SETS["c2"]
    flags = scpd.packed_flags(name, n)
    dec = scpd.Decoder(n, k, flags)
    nfr = 1 << 17
    g = torch.Generator(device="cuda").manual_seed(1)
    u = torch.randint(0, 2, (nfr, n), device="cuda", dtype=torch.int32, generator=g)
    u *= torch.from_numpy(flags.astype(np.int32)).cuda()
    # pack u, transform on the device with the product's own butterfly
    w = torch.tensor([1 << i for i in range(31)] + [-(1 << 31)], device="cuda", dtype=torch.int64)
    upk = (u.view(nfr, n // 32, 32).to(torch.int64) * w).sum(-1).to(torch.int32)
    xpk = dec.extract_info(upk)  # F^(x)n is an involution: encode = same transform
    shifts = torch.arange(32, device="cuda", dtype=torch.int32)
    xbits = ((xpk.unsqueeze(-1) >> shifts) & 1).view(nfr, n)
    llr = (4 - 8 * xbits).to(torch.int8)
    del xbits, u
    xhat = dec.decode(llr)
    assert torch.equal(xhat, xpk)
    assert torch.equal(dec.extract_info(xhat), upk)
    # anchor the transform itself on the oracle for a few frames
    some = ol.unpack_bits(upk[:4].cpu().numpy().view(np.uint32), n)
    assert (ol.pack_bits(ol.polar_transform(some)) == xpk[:4].cpu().numpy().view(np.uint32)).all()


def test_dispatch_by_batch_size(scpd):
    """scpd_decode's kernel choice (profiles/tuning_r1.md, tuning_r2.md): the slot-sliced kernel from 2.5 tasks of 32 frames
    per SM, the int16x2 kernel below the crossover with lane groups that widen as the batch shrinks, a whole CTA per
    frame pair for the smallest batches of large trees -- and the same bits whichever kernel runs."""
    import torch
    name, n, k, snr = CONFIG_SETS["c3"]
    flags = scpd.packed_flags(name, n)
    dec = scpd.Decoder(n, k, flags)
    assert dec.last_kernel_name == ""
    llr = scpd.channel_generate(n, 16384, scpd.sigma(snr, k / n))
    seen = {}
    for nfr in (16384, 11000, 6000, 4096, 1024, 64):
        out = dec.decode(llr[:nfr])
        torch.cuda.synchronize()
        seen[nfr] = (dec.last_kernel_name, out[:8].cpu().numpy().view(np.uint32))
    assert "slot-sliced" in seen[16384][0]
    assert " 8 lanes" in seen[11000][0] and "16 lanes" in seen[6000][0] and "32 lanes" in seen[4096][0]
    assert "coop" in seen[1024][0] and "4 warps" in seen[1024][0] and "8 warps" in seen[64][0]
    want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr[:8].cpu().numpy(), threads=8)
    for nfr, (kname, got) in seen.items():
        assert (got == want).all(), (nfr, kname)
