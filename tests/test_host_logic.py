"""CPU tier: host logic of the product (schedule compiler, pruning rules, kernel source under the
warp emulator) against the oracle, and the C ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_lib as ol
import sc_polar_decoder_hls_b200 as scpd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIBS = {}


def _emu_lib(name):
    """Loaded lazily: the session fixture (conftest.py) rebuilds stale emulator libraries first."""
    if name not in _LIBS:
        _LIBS[name] = ctypes.CDLL(os.path.join(ROOT, "tests", "emu", name))
        if name == "libschedule_emu.so":
            _LIBS[name].emu_decode.restype = ctypes.c_longlong
    return _LIBS[name]


def _emu(flags, n, par, q, ext, prune, llr, force=0):
    out = np.zeros(llr.shape, np.uint8)
    nops, nfg = ctypes.c_uint64(), ctypes.c_uint64()
    fb = _emu_lib('libschedule_emu.so').emu_decode(int(np.log2(n)), int(np.log2(par)), q, ext, prune, ol.P(flags), ol.P(llr),
                        ctypes.c_size_t(len(llr)), ol.P(out), force, ctypes.byref(nops), ctypes.byref(nfg))
    assert fb >= 0
    return out, fb, nops.value, nfg.value


@pytest.mark.parametrize("name,n,k", [("FB_N8_K4", 8, 4), ("FB_N512_K256", 512, 256),
                                      ("FB_N1024_K512", 1024, 512), ("frozen_n_4096_k_3072", 4096, 3072)])
def test_schedule_interpreter_equals_oracle(name, n, k):
    """Every pruning mode of the schedule (and the forced rate-1 fallback walk) is bit-identical to
    plain SC, for PAR 1..256, Q 6/8/9, EXTENDED 0/1."""
    flags = scpd.packed_flags(name, n)
    rng = np.random.default_rng(7)
    for par in (1, 2, 4, 16, 64, 256):
        if 2 * par > n:
            continue
        for q, ext in ((8, 1), (8, 0), (6, 1), (9, 1)):
            if ext and q + int(np.log2(par)) > 16:
                continue
            nfr = 12 if n >= 1024 else 40
            llr = ol.test_llrs(rng, n, nfr, k, maxabs=min(31, (1 << (q - 1)) - 1))
            llr[-1][rng.random(n) < 0.5] = 0
            want = ol.decode(n, par, q, 0, ext, flags, llr)
            for prune, force in ((0, 0), (1, 0), (2, 0), (2, 1)):
                got, fb, nops, nfg = _emu(flags, n, par, q, ext, prune, llr, force)
                assert (got == want).all(), (par, q, ext, prune, force)


def test_schedule_pruning_reduces_work():
    n, k = 1024, 512
    flags = scpd.packed_flags("FB_N1024_K512", n)
    llr = np.zeros((1, n), np.int8)
    work = [_emu(flags, n, 16, 8, 1, p, llr)[3] for p in (0, 1, 2)]
    assert work[0] == n * 10 - n  # every f and g above the 2-bit terminals: N log2 N - N
    assert work[0] > work[1] > work[2]


def test_arbitrary_flag_tables():
    rng = np.random.default_rng(2)
    for n in (4, 8, 16, 64, 256):
        for _ in range(20):
            flags = (rng.random(n) < rng.random()).astype(np.uint8)
            llr = rng.integers(-31, 32, size=(30, n)).astype(np.int8)
            for par in (1, 2, 4, 16):
                if 2 * par > n:
                    continue
                want = ol.decode(n, par, 8, 0, 1, flags, llr)
                for prune in (0, 1, 2):
                    assert (_emu(flags, n, par, 8, 1, prune, llr)[0] == want).all()


def _wemu(g, flags, n, par, ext, prune, llr, lsa=-1, lsb=-1, warps=1, grid=1):
    out = np.zeros((len(llr), n // 32), np.uint32)
    rc = _emu_lib('libwarp_emu.so').emu_fast_decode(g, int(np.log2(n)), int(np.log2(par)), 8, ext, prune, ol.P(flags), ol.P(llr),
                              ctypes.c_size_t(len(llr)), ol.P(out), lsa, lsb, warps, grid)
    assert rc == 0, rc
    return out


@pytest.mark.parametrize("g,par,ext", [(1, 8, 1), (2, 16, 1), (4, 16, 1), (8, 16, 1), (16, 16, 1), (32, 16, 1),
                                       (8, 16, 0), (8, 4, 1), (8, 64, 1), (4, 4, 1), (2, 2, 1)])
def test_fast_kernel_source_under_warp_emulator(g, par, ext):
    """decode_fast.cuh executed lane by lane on the CPU (tests/emu/warp_emu.cpp): cell packing,
    register subtree, node-type descriptors, shared/workspace storage split, ragged batches."""
    n, k = 1024, 512
    flags = scpd.packed_flags("FB_N1024_K512", n)
    rng = np.random.default_rng(g * 10 + par)
    nfr = 2 * (32 // g) + 3  # more than one warp pass, odd tail
    llr = ol.test_llrs(rng, n, nfr, k)
    llr[0] = 0
    llr[1][::5] = 0
    want = ol.decode_packed(n, par, 8, 0, ext, flags, llr)
    for prune in (0, 1, 2):
        assert (_wemu(g, flags, n, par, ext, prune, llr) == want).all(), prune
    ls = 3 + int(np.log2(g))
    if ls + 2 <= 9:
        assert (_wemu(g, flags, n, par, ext, 2, llr, lsa=ls + 1, lsb=ls, warps=2, grid=2) == want).all()


def test_fast_kernel_emulated_on_other_codes():
    rng = np.random.default_rng(4)
    for name, n, k in (("frozen_n_4096_k_3072", 4096, 3072), ("FB_N512_K256", 512, 256), ("FB_N2048_K1024", 2048, 1024)):
        flags = scpd.packed_flags(name, n)
        llr = ol.test_llrs(rng, n, 9, k, 3.0)
        want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr)
        for g in (4, 8):
            assert (_wemu(g, flags, n, 16, 1, 2, llr, lsa=8, lsb=7) == want).all(), (name, g)
    for n in (128, 256):  # random flag tables
        for _ in range(6):
            flags = (rng.random(n) < rng.random()).astype(np.uint8)
            llr = rng.integers(-31, 32, size=(10, n)).astype(np.int8)
            want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr)
            for prune in (0, 2):
                assert (_wemu(8, flags, n, 16, 1, prune, llr) == want).all()


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "scpd.h")).read()
    declared = set(re.findall(r"\b(scpd_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"scpd_config", "scpd_decoder", "scpd_status", "scpd_format", "scpd_pruning"}
    assert declared == set(scpd.EXPORTS), declared ^ set(scpd.EXPORTS)
    raw = ctypes.CDLL(scpd.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name


def test_c_abi_argument_validation_without_gpu():
    """Configuration errors are reported before any CUDA call, with the reference's constraints."""
    flags = scpd.packed_flags("FB_N1024_K512", 1024)
    bad = [dict(n=1000), dict(par=1024), dict(par=3), dict(k=511), dict(llr_bits=4), dict(llr_bits=12),
           dict(fmt=7), dict(pruning=9), dict(extended=2)]
    for kw in bad:
        args = dict(n=1024, k=512, par=16, llr_bits=8, fmt=0, extended=1, pruning=2)
        args.update(kw)
        fl = flags if args["n"] == 1024 else np.zeros(args["n"], np.uint8)
        with pytest.raises(scpd.ScpdError) as e:
            scpd.Decoder(args["n"], args["k"], fl, par=args["par"], llr_bits=args["llr_bits"], fmt=args["fmt"],
                         extended=args["extended"], pruning=args["pruning"])
        assert e.value.status == scpd.E_CONFIG, kw
    assert abs(scpd.sigma(2.5, 0.5) - ol.sigma(2.5, 0.5)) < 1e-7
    assert scpd.lib.scpd_status_string(scpd.E_IO) == b"i/o error"
    assert scpd.lib.scpd_decode(None, None, 1, None, None) == scpd.E_ARG


def _remu(fmt, g, flags, n, par, q, ext, prune, llr, ls=-1, beta_sm=1, warps=1, grid=1):
    out = np.zeros((len(llr), max(n // 32, 1)), np.uint32)
    rc = _emu_lib('libwarp_emu.so').emu_raw_decode(fmt, g, int(np.log2(n)), int(np.log2(par)), q, ext, prune, ol.P(flags),
                                                   ol.P(llr), ctypes.c_size_t(len(llr)), ol.P(out), ls, beta_sm, warps, grid)
    assert rc == 0, rc
    return out


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("q", [5, 6, 8, 9])
def test_raw_kernel_source_under_warp_emulator(fmt, q):
    """decode_raw.cuh (the every-configuration kernel) executed lane by lane on the CPU against the oracle:
    both formats, LLR_BITS 5 (the +-31 alphabet wraps) ... 9, PAR 1 ... 256, EXTENDED 0/1, N 8 ... 1024, the whole int8
    input range (wrap-around on the sc_fifo<LLR> write), storage split between shared memory and workspace."""
    rng = np.random.default_rng(100 * fmt + q)
    for n, par, ext, g in [(8, 2, 1, 8), (8, 4, 1, 32), (16, 1, 1, 8), (64, 16, 0, 8), (64, 32, 1, 32), (256, 16, 1, 8),
                           (512, 256, 1, 32), (1024, 64, 0, 8)]:
        k = n // 2
        flags = np.zeros(n, np.uint8)
        flags[rng.permutation(n)[:k]] = 1
        if n == 1024:
            flags = scpd.packed_flags("FB_N1024_K512", n)
        nfr = 32 // g + 2
        llr = rng.integers(-31, 32, (nfr, n)).astype(np.int8)
        llr[0] = 0
        llr[1] = rng.integers(-128, 128, n).astype(np.int8)  # out of range for every LLR_BITS here
        want = ol.decode_packed(n, par, q, fmt, ext, flags, llr)
        for prune in ((0, 2) if n <= 64 else (2,)):
            got = _remu(fmt, g, flags, n, par, q, ext, prune, llr)
            assert (got == want).all(), (n, par, ext, g, prune)
        if n == 256:
            got = _remu(fmt, g, flags, n, par, q, ext, 1, llr, ls=4, beta_sm=0, warps=2, grid=2)
            assert (got == want).all(), (n, par, ext, g, "workspace")


@pytest.mark.parametrize("fmt", [0, 1])
def test_reference_pruning_mode_under_warp_emulator(fmt):
    """SCPD_PRUNE_REF_LEVEL2 (the reference built with PRUNING_LEVEL 2: R0 / R1 / REP / SPC / H0) on decode_raw.cuh,
    lane by lane on the CPU, against sco_decode_l2 -- itself pinned on eleven builds of the reference's own sources
    (test_oracle.py).  PAR 2 ... 256, LLR_BITS 6 ... 9, EXTENDED 0/1, 8 / 32 lanes per frame; noisy frames, the whole
    Q-bit range, near-zero LLRs (ties in the minimum search, zero sums); storage split over shared memory / workspace."""
    rng = np.random.default_rng(40 + fmt)
    differs = 0
    for tab, n, k in (("FB_N8_K4", 8, 4), ("FB_N512_K256", 512, 256), ("FB_N1024_K512", 1024, 512),
                      ("frozen_n_4096_k_3072", 4096, 3072)):
        flags = scpd.packed_flags(tab, n)
        for par, q, ext, g in ((16, 8, 1, 32), (4, 9, 1, 8), (64, 6, 0, 32), (2, 7, 1, 8), (256, 8, 1, 32)):
            if 2 * par > n or (n == 4096 and par not in (16, 64)):
                continue
            m = min(2 ** (q - 1) - 1, 127)
            llr = np.concatenate([ol.channel(n, 2, ol.sigma(1.0 if k / n < 0.6 else 3.0, k / n)),
                                  rng.integers(-m, m + 1, size=(2, n)).astype(np.int8),
                                  rng.integers(-2, 3, size=(2, n)).astype(np.int8), np.zeros((1, n), np.int8)])
            want = ol.pack_bits(ol.decode_l2(n, par, q, fmt, ext, flags, llr))
            assert (_remu(fmt, g, flags, n, par, q, ext, 3, llr) == want).all(), (n, par, q, ext, g)
            differs += int((ol.decode_packed(n, par, q, fmt, ext, flags, llr) != want).any(axis=1).sum())
            if n == 512 and par == 16:
                got = _remu(fmt, g, flags, n, par, q, ext, 3, llr, ls=5, beta_sm=0, warps=2, grid=2)
                assert (got == want).all(), "workspace"
    assert differs > 10  # the mode is a different decoder, and these inputs show it


def test_reference_pruning_mode_stage_matrix_and_validation():
    """scpd_stage_profile at SCPD_PRUNE_REF_LEVEL2 is the reference FSM at PRUNING_LEVEL 2: an independent recursion over
    the frozen table counts the same F / F_REP / G / G_R1 / G_SPC / H / R visits; the mode needs 2 <= PAR <= 256."""
    n, k, par = 1024, 512, 16
    flags = scpd.packed_flags("FB_N1024_K512", n)

    def wtype(o):
        w = flags[o:o + par]
        c = int(w.sum())
        return "R0" if c == 0 else "R1" if c == par else "REP" if (c == 1 and w[-1]) else "SPC" if (c == par - 1 and not w[0]) else "RN"

    def ctype(o, size):
        t = [wtype(o + i) for i in range(0, size, par)]
        if all(x == "R0" for x in t):
            return "R0"
        if all(x == "R1" for x in t):
            return "R1"
        if all(x == "R0" for x in t[:-1]) and t[-1] == "REP":
            return "REP"
        if t[0] == "SPC" and all(x == "R1" for x in t[1:]):
            return "SPC"
        return "RN"

    acc = dict.fromkeys(scpd.STAGE_FUNCS, 0)

    def rec(o, size, root=False):
        if size <= par:
            acc["R"] += 1
            return
        h = size // 2
        tl, tr = ("RN", "RN") if root else (ctype(o, h), ctype(o + h, h))
        if tl == "R0":
            acc["R0"] += 1
        elif tl == "REP":
            acc["REP"] += 1
        else:
            acc["F"] += 1
            rec(o, h)
        if tr == "R1":
            acc["R1"] += 1
        elif tr == "SPC":
            acc["SPC"] += 1
        else:
            acc["G"] += 1
            rec(o + h, h)
        acc["H"] += 1

    rec(0, n, True)
    vis, it, total = scpd.stage_profile(n, k, flags, par=par, pruning=scpd.PRUNE_REF_LEVEL2)
    assert {f: int(vis[i].sum()) for i, f in enumerate(scpd.STAGE_FUNCS)} == acc
    assert acc["REP"] > 0 and acc["SPC"] > 0 and acc["R1"] > 0
    plain = scpd.stage_profile(n, k, flags, par=par, pruning=scpd.PRUNE_NONE)[2]
    assert total < plain * 4 // 5  # why the reference ships with it: 484 instead of 640 loop trips per frame at c1
    for bad_par in (1, 512):
        with pytest.raises(scpd.ScpdError) as e:
            scpd.Decoder(n, k, flags, par=bad_par, pruning=scpd.PRUNE_REF_LEVEL2)
        assert e.value.status == scpd.E_CONFIG


def test_stage_profile_matches_reference_fsm_trip_counts():
    """scpd_stage_profile (host only): at PRUNE_NONE the matrix is the reference FSM at PRUNING_LEVEL 0 -- every node
    above the leaf runs f_loop, g_loop and h_loop for half its PAR-wide words (my_module.h:343,373,704,903) and every
    leaf is one R visit -- and with pruning it follows an independent recursion over the frozen table."""
    n, k, par = 1024, 512, 16
    flags = scpd.packed_flags("FB_N1024_K512", n)
    vis, it, total = scpd.stage_profile(n, k, flags, par=par, pruning=scpd.PRUNE_NONE)
    F, G, H, R, R0, R1 = range(6)
    for l in range(5, 11):
        nodes, words = n >> l, (1 << l) // par
        for fn in (F, G, H):
            assert vis[fn][l] == nodes and it[fn][l] == nodes * words // 2
    assert vis[R][4] == n // par and vis[R0].sum() == 0 and vis[R1].sum() == 0
    assert total == 3 * 6 * (n // par) // 2 + n // par

    def rec(l, o, prune, acc):
        size = 1 << l
        c = int(flags[o:o + size].sum())
        if prune >= 1 and c == 0:
            acc["R0"] += 1
            return
        if size <= par:
            acc["R"] += 1
            return
        if prune >= 2 and c == size:
            acc["R1"] += 1
            return
        h = size // 2
        if prune >= 1 and flags[o:o + h].sum() == 0:
            acc["R0"] += 1
        else:
            acc["F"] += 1
            rec(l - 1, o, prune, acc)
        if prune >= 1 and flags[o:o + h].sum() != 0 and flags[o + h:o + size].sum() == 0:
            acc["R0"] += 1
        else:
            acc["G"] += 1
            rec(l - 1, o + h, prune, acc)
        acc["H"] += 1

    for prune in (1, 2):
        acc = dict.fromkeys(scpd.STAGE_FUNCS, 0)
        rec(10, 0, prune, acc)
        vis, it, total2 = scpd.stage_profile(n, k, flags, par=par, pruning=prune)
        assert {f: int(vis[i].sum()) for i, f in enumerate(scpd.STAGE_FUNCS)} == acc
        assert total2 < total
    rnd = np.random.default_rng(0)
    bad = scpd.Config(48, 1, 16, 8, 0, 1, 0, 0)
    assert scpd.lib.scpd_stage_profile(ctypes.byref(bad), ol.P(flags), ctypes.byref(scpd.StageMatrix())) == scpd.E_CONFIG


def _cemu(w, flags, n, prune, llr, lsa=-1, lsb=-1, grid=1):
    out = np.zeros((len(llr), n // 32), np.uint32)
    rc = _emu_lib('libwarp_emu.so').emu_fast_coop_decode(w, int(np.log2(n)), 4, 8, 1, prune, ol.P(flags), ol.P(llr),
                                                         ctypes.c_size_t(len(llr)), ol.P(out), lsa, lsb, grid)
    assert rc == 0, rc
    return out


@pytest.mark.parametrize("w", [2, 4, 8])
def test_cooperative_fast_kernel_under_cta_emulator(w):
    """The CTA-cooperative variant of decode_fast.cuh (W warps walk one frame pair; small batches of large trees):
    all W warps emulated concurrently with a real CTA barrier, every pruning mode, zero-LLR frames (rate-1 fallback
    decided CTA-wide), levels split between shared memory and workspace, odd batch, several pairs per CTA."""
    rng = np.random.default_rng(w)
    sets = [("FB_N1024_K512", 1024, 512, (0, 1, 2))]
    if w == 4:
        sets.append(("frozen_n_4096_k_3072", 4096, 3072, (2,)))
    for name, n, k, prunes in sets:
        flags = scpd.packed_flags(name, n)
        llr = ol.test_llrs(rng, n, 5 if n == 1024 else 3, k)
        llr[0] = 0
        llr[1][::5] = 0
        want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr)
        for prune in prunes:
            assert (_cemu(w, flags, n, prune, llr, grid=2) == want).all(), (n, prune)
        ls = int(np.log2(n)) - 2
        assert (_cemu(w, flags, n, 2, llr, lsa=ls, lsb=ls - 1, grid=1) == want).all(), (n, "workspace")


# (words per lane of the register kernels, or 0 for the shared-memory kernel; frame sizes that instantiation serves)
_COUNT_VARIANTS = [(1, 32), (1, 256), (1, 1024), (2, 2048), (4, 4096), (8, 8192), (16, 16384), (32, 32768), (0, 1024), (0, 65536)]


@pytest.mark.parametrize("wpl,n", _COUNT_VARIANTS)
def test_counter_and_transform_kernels_under_warp_emulator(wpl, n):
    """count.cuh executed on the CPU (tests/emu/warp_emu.cpp, a fiber per lane, eight warps per CTA): the ten counters of
    the Monte-Carlo loop from one pass over x^ -- codeword bits as sc_error_counter.h:68-125 counts them (10-bit wrap
    included), information bits through (x^ ^ x) F^(x)n -- and the transform alone, against the oracle's counter and
    transform; all-zero, shared and per-frame references; frames with no, few and more than 1023 bit errors."""
    rng = np.random.default_rng(1000 * wpl + n)
    lib = _emu_lib('libwarp_emu.so')
    wpf = n // 32
    nfr = 21 if n <= 4096 else 5
    flags = (rng.random(n) < 0.5).astype(np.uint8)
    k = int(flags.sum())
    mask = ol.pack_bits(flags[None, :])[0]
    sent = rng.integers(0, 2, (nfr, n)).astype(np.uint8)
    err = (rng.random((nfr, n)) < rng.choice([0.0, 2.0 / n, 0.3], size=(nfr, 1))).astype(np.uint8)
    err[0] = 0
    err[1] = 1  # every bit wrong: n errors, n & 1023 under the wrap
    for mode in (0, 1, 2):  # all-zero codeword, one shared codeword, one per frame
        x = np.zeros((nfr, n), np.uint8) if mode == 0 else np.broadcast_to(sent[0], (nfr, n)) if mode == 1 else sent
        xhat = x ^ err
        xw = np.ascontiguousarray(ol.pack_bits(xhat))
        rw = None if mode == 0 else np.ascontiguousarray(ol.pack_bits(x[:1] if mode == 1 else x))
        cnt = np.zeros(10, np.uint64)
        rc = lib.emu_count_all(wpl, wpf, n, k, ctypes.c_ulonglong(nfr), ol.P(xw), ol.P(rw) if rw is not None else None,
                               int(mode == 2), ol.P(mask), ol.P(cnt), 3)
        assert rc == 0
        e = ol.polar_transform(err)[:, flags == 1].sum(axis=1)
        want = ol.count_errors(n, xhat, np.ascontiguousarray(x)) + [int(e.sum()), int((e != 0).sum()), nfr * k, nfr]
        assert [int(v) for v in cnt] == want, (wpl, n, mode)
    words = np.ascontiguousarray(ol.pack_bits(sent))
    out = np.zeros_like(words)
    assert lib.emu_polar_transform(wpl, wpf, ctypes.c_ulonglong(nfr), ol.P(words), ol.P(out), 2) == 0
    assert (out == ol.pack_bits(ol.polar_transform(sent))).all()
    assert lib.emu_polar_transform(wpl, wpf, ctypes.c_ulonglong(nfr), ol.P(out), ol.P(out), 2) == 0  # in place; an involution
    assert (out == words).all()
