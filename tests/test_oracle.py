"""CPU tier: the oracle against the reference's golden vectors and against the reference itself
(its own sources compiled natively into oracle/_ref by oracle/build_ref.py)."""
import ctypes
import os

import numpy as np
import pytest

import oracle_lib as ol
import sc_polar_decoder_hls_b200 as scpd

GOLD = (("cw8x4", "FB_N8_K4", 8, 4), ("cw512x256", "FB_N512_K256", 512, 256),
        ("cw1024x512", "FB_N1024_K512", 1024, 512))


def test_golden_codewords_are_codewords_of_their_frozen_sets():
    """u = x F^(x)n is zero on every frozen position only in natural order."""
    cws = ol.golden_codewords()
    for key, name, n, k in GOLD:
        flags = scpd.packed_flags(name, n)
        assert flags.sum() == k
        u = ol.polar_transform(cws[key])
        assert (u[:, flags == 0] == 0).all()


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("q", [5, 6, 8])
@pytest.mark.parametrize("ext", [0, 1])
def test_golden_codewords_decode_noiseless(fmt, q, ext):
    """sc_encoder.h:74-89 codewords through a noiseless channel (DEBUG_MAIN: LLR = +-4) come back."""
    cws = ol.golden_codewords()
    for key, name, n, k in GOLD:
        flags = scpd.packed_flags(name, n)
        llr = np.where(cws[key] == 1, -4, 4).astype(np.int8)
        for par in (1, 2, 4, 16, 64):
            if 2 * par > n:
                continue
            assert (ol.decode(n, par, q, fmt, ext, flags, llr) == cws[key]).all(), (key, par)


REF_TAGS = [("n8_p2_q8_ca2_e1", "FB_N8_K4", 300), ("n8_p4_q6_sm_e1", "FB_N8_K4", 300),
            ("n512_p16_q8_ca2_e1", "FB_N512_K256", 40), ("n512_p64_q6_sm_e1", "FB_N512_K256", 30),
            ("n1024_p16_q8_ca2_e1", "FB_N1024_K512", 40), ("n1024_p16_q8_ca2_e0", "FB_N1024_K512", 30),
            ("n1024_p4_q8_ca2_e1", "FB_N1024_K512", 30), ("n1024_p64_q8_ca2_e1", "FB_N1024_K512", 20),
            ("n1024_p256_q8_ca2_e1", "FB_N1024_K512", 20), ("n1024_p16_q6_ca2_e1", "FB_N1024_K512", 30),
            ("n1024_p16_q5_ca2_e1", "FB_N1024_K512", 30), ("n1024_p16_q9_ca2_e1", "FB_N1024_K512", 30),
            ("n1024_p16_q6_sm_e1", "FB_N1024_K512", 30), ("n1024_p16_q8_sm_e1", "FB_N1024_K512", 30),
            ("n1024_p16_q6_sm_e0", "FB_N1024_K512", 30), ("n1024_p64_q7_sm_e1", "FB_N1024_K512", 20),
            ("n1024_p16_q5_sm_e1", "FB_N1024_K512", 30), ("n1024_p256_q9_ca2_e1", "FB_N1024_K512", 20),
            ("n1024_p4_q9_sm_e1", "FB_N1024_K512", 30),
            ("n4096_p16_q8_ca2_e1", "frozen_n_4096_k_3072", 8),
            ("n32768_p16_q8_ca2_e1", "frozen_n_32768_k_29492_snr_4_5", 2)]


@pytest.mark.parametrize("tag,name,nfr", REF_TAGS)
def test_oracle_equals_reference_compiled_natively(tag, name, nfr):
    """The reference's my_module.h FSM + wrappers (compiled unmodified on the systemc.h shim) and
    the oracle produce the same codeword estimate on random and channel LLRs."""
    R = ol.ref_lib(tag)
    if R is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    cfg = (ctypes.c_int32 * 6)()
    R.ref_config(cfg)
    n, par, q, fmt, ext, _ = list(cfg)
    flags = scpd.packed_flags(name, n)
    rng = np.random.default_rng(abs(hash(tag)) % 1000)
    llr = ol.test_llrs(rng, n, nfr, int(flags.sum()))
    llr[-1] = rng.integers(-31, 32, size=n)
    llr[-1][rng.random(n) < 0.4] = 0  # exercise the zero / tie rules
    llr[-2] = rng.integers(-128, 128, size=n)  # beyond every LLR_BITS here: truncation on the sc_fifo<LLR> write
    assert (ol.ref_decode(R, flags, llr) == ol.decode(n, par, q, fmt, ext, flags, llr)).all()


@pytest.mark.parametrize("tag", ["n1024_p16_q8_ca2_e1", "n1024_p16_q6_sm_e1", "n1024_p16_q5_ca2_e1",
                                 "n1024_p64_q7_sm_e1", "n8_p4_q6_sm_e1"])
def test_primitives_equal_reference(tag):
    """PU_FUNCTION_F / PU_FUNCTION_G / Spec_Polar_Decoder / Adapt_format of the compiled reference
    against sco_f / sco_g / sco_leaf / sco_input, exhaustively over (a, b, s) for the element ops."""
    R = ol.ref_lib(tag)
    if R is None:
        pytest.skip("oracle/_ref not built")
    L = ol.lib()
    cfg = (ctypes.c_int32 * 6)()
    R.ref_config(cfg)
    n, par, q, fmt, ext, _ = list(cfg)
    R.ref_adapt.restype = ctypes.c_uint32
    for v in range(-128, 128):
        assert R.ref_adapt(v) == L.sco_input(fmt, q, v), v
    vals = np.arange(1 << q, dtype=np.uint32)
    pairs = np.array(np.meshgrid(vals, vals)).reshape(2, -1).T  # every (a, b) pattern pair
    pad = (-len(pairs)) % par
    pairs = np.concatenate([pairs, np.zeros((pad, 2), np.uint32)])
    a = np.ascontiguousarray(pairs[:, 0].reshape(-1, par))
    b = np.ascontiguousarray(pairs[:, 1].reshape(-1, par))
    out = np.zeros(par, np.uint32)
    step = max(1, len(a) // 400)  # all pairs for small Q, a dense sample for Q = 8
    for row in range(0, len(a), step if q >= 8 else 1):
        R.ref_pu_f(ol.P(a[row]), ol.P(b[row]), ol.P(out))
        want = [L.sco_f(fmt, q, int(x), int(y)) for x, y in zip(a[row], b[row])]
        assert out.tolist() == want
        for s in (0, 1):
            sb = np.full(par, s, np.uint8)
            R.ref_pu_g(ol.P(a[row]), ol.P(b[row]), ol.P(sb), ol.P(out))
            want = [L.sco_g(fmt, q, int(x), int(y), s) for x, y in zip(a[row], b[row])]
            assert out.tolist() == want
    rng = np.random.default_rng(1)
    bits_ref = np.zeros(par, np.uint8)
    bits_or = np.zeros(par, np.uint8)
    for _ in range(300):
        llr = rng.integers(0, 1 << q, size=par).astype(np.uint32)
        if rng.random() < 0.3:
            llr[rng.random(par) < 0.3] = 0
        fl = (rng.random(par) < rng.random()).astype(np.uint8)
        R.ref_leaf(ol.P(llr), ol.P(fl), ol.P(bits_ref))
        L.sco_leaf(fmt, ext, par, q, ol.P(llr), ol.P(fl), ol.P(bits_or))
        assert (bits_ref == bits_or).all()


def test_par_and_format_are_part_of_the_contract():
    """with EXTENDED=1 the decoded bits depend on PAR; CA2 and SIGMAG differ."""
    n, k = 1024, 512
    flags = scpd.packed_flags("FB_N1024_K512", n)
    rng = np.random.default_rng(3)
    llr = rng.integers(-31, 32, size=(200, n)).astype(np.int8)
    a = ol.decode(n, 4, 6, 0, 1, flags, llr)  # Q=6: +-31 inputs saturate at once, as in the survey's check
    b = ol.decode(n, 64, 6, 0, 1, flags, llr)
    assert (a != b).any(axis=1).sum() > 100
    a0 = ol.decode(n, 4, 6, 0, 0, flags, llr)
    b0 = ol.decode(n, 64, 6, 0, 0, flags, llr)
    assert (a0 == b0).all()
    big = rng.integers(-127, 128, size=(200, n)).astype(np.int8)  # Q=8 needs full-range inputs to show it
    assert (ol.decode(n, 4, 8, 0, 1, flags, big) != ol.decode(n, 64, 8, 0, 1, flags, big)).any()
    ch = ol.channel(n, 300, ol.sigma(2.5, 0.5))
    assert (ol.decode(n, 16, 8, 0, 1, flags, ch) != ol.decode(n, 16, 8, 1, 1, flags, ch)).any()


def test_rate1_hard_decision_rule():
    """SURVEY G10 / App. A.5 re-derived: on an isolated all-information node plain SC equals the hard
    decision when no LLR is zero (CA2) and always in SIGMAG; with CA2 zeros it does not."""
    n = 64
    flags = np.ones(n, np.uint8)
    rng = np.random.default_rng(5)
    nz = rng.integers(1, 32, size=(500, n)) * rng.choice([-1, 1], size=(500, n))
    nz = nz.astype(np.int8)
    assert (ol.decode(n, 16, 8, 0, 1, flags, nz) == (nz < 0)).all()
    z = nz.copy()
    z[rng.random(z.shape) < 0.11] = 0
    assert (ol.decode(n, 16, 8, 0, 1, flags, z) != (z < 0)).any(axis=1).sum() > 400
    sm = ol.decode(n, 16, 8, 1, 1, flags, z)
    assert (sm == (z < 0)).all()


@pytest.mark.parametrize("tag,q,fmt", [("n1024_p16_q6_sm_e1_pl2", 6, 1), ("n1024_p16_q8_ca2_e1_pl2", 8, 0)])
def test_reference_full_pruning_is_not_plain_sc(tag, q, fmt):
    """Why REP / SPC pruning is not one of this library's modes: the reference compiled with
    its checked-in PRUNING_LEVEL 2 (R1 / REP / SPC / H0, config.h:16-30) returns the golden codewords and agrees with
    plain SC on clean frames, but departs from it on noisy ones -- so it is a different decoder, not a shortcut."""
    R = ol.ref_lib(tag)
    if R is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    n, k = 1024, 512
    flags = scpd.packed_flags("FB_N1024_K512", n)
    cws = ol.golden_codewords()["cw1024x512"]
    assert (ol.ref_decode(R, flags, np.where(cws == 1, -4, 4).astype(np.int8)) == cws).all()
    clean = ol.channel(n, 100, ol.sigma(4.0, k / n))
    assert (ol.ref_decode(R, flags, clean) == ol.decode(n, 16, q, fmt, 1, flags, clean)).all()
    noisy = ol.channel(n, 200, ol.sigma(1.0, k / n))
    differ = (ol.ref_decode(R, flags, noisy) != ol.decode(n, 16, q, fmt, 1, flags, noisy)).any(axis=1)
    assert 0 < differ.sum() < len(noisy)


L2_BUILDS = [  # (oracle/_ref tag, n, k, table, PAR, Q, format, EXTENDED)
    ("n1024_p16_q8_ca2_e1_pl2", 1024, 512, "FB_N1024_K512", 16, 8, 0, 1),
    ("n1024_p16_q6_sm_e1_pl2", 1024, 512, "FB_N1024_K512", 16, 6, 1, 1),
    ("n1024_p4_q8_ca2_e1_pl2", 1024, 512, "FB_N1024_K512", 4, 8, 0, 1),
    ("n1024_p64_q8_ca2_e1_pl2", 1024, 512, "FB_N1024_K512", 64, 8, 0, 1),
    ("n1024_p16_q8_ca2_e0_pl2", 1024, 512, "FB_N1024_K512", 16, 8, 0, 0),
    ("n1024_p64_q7_sm_e1_pl2", 1024, 512, "FB_N1024_K512", 64, 7, 1, 1),
    ("n1024_p4_q9_sm_e1_pl2", 1024, 512, "FB_N1024_K512", 4, 9, 1, 1),
    ("n512_p16_q8_ca2_e1_pl2", 512, 256, "FB_N512_K256", 16, 8, 0, 1),
    ("n8_p2_q8_ca2_e1_pl2", 8, 4, "FB_N8_K4", 2, 8, 0, 1),
    ("n4096_p16_q8_ca2_e1_pl2", 4096, 3072, "frozen_n_4096_k_3072", 16, 8, 0, 1),
    ("n32768_p16_q8_ca2_e1_pl2", 32768, 29492, "frozen_n_32768_k_29492_snr_4_5", 16, 8, 0, 1),
]


@pytest.mark.parametrize("tag,n,k,tab,par,q,fmt,ext", L2_BUILDS, ids=[b[0] for b in L2_BUILDS])
def test_level2_restatement_matches_reference_pruning_level_2(tag, n, k, tab, par, q, fmt, ext):
    """sco_decode_l2 (REP / SPC / R1 / R0 shortcuts, my_module.h:1292-1842) against the reference's own sources built
    with PRUNING_LEVEL 2: noisy channel frames, the whole quantiser alphabet, the whole Q-bit range, near-zero LLRs
    (ties in the SPC minimum search, zero sums in REP) -- frames on which plain SC answers differently."""
    R = ol.ref_lib(tag)
    if R is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    flags = scpd.packed_flags(tab, n)
    rng = np.random.default_rng(n + par + q)
    nf = 64 if n <= 4096 else 6
    m = min(2 ** (q - 1) - 1, 127)
    inputs = {
        "noisy": ol.channel(n, nf, ol.sigma(1.0 if k / n < 0.6 else 3.0, k / n)),
        "alphabet": rng.integers(-31, 32, size=(nf, n)).astype(np.int8),
        "full": rng.integers(-m, m + 1, size=(nf, n)).astype(np.int8),
        "near_zero": rng.integers(-2, 3, size=(nf, n)).astype(np.int8),
        "zeros": np.zeros((2, n), np.int8),
    }
    departs = 0
    for name, llr in inputs.items():
        want = ol.ref_decode(R, flags, llr)
        assert (ol.decode_l2(n, par, q, fmt, ext, flags, llr) == want).all(), name
        departs += int((ol.decode(n, par, q, fmt, ext, flags, llr) != want).any(axis=1).sum())
    if n >= 512:
        assert departs > 0  # the inputs do exercise the shortcuts: plain SC disagrees on some of them
