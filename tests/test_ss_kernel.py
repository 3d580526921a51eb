"""CPU tier for the slot-sliced kernel (csrc/decode_ss.cuh, ss_plan.h): the kernel source executed lane by lane
(tests/emu/ss_emu.cpp; fp16x2 emulated with exact float pairs, tensor memory as a per-lane array) against the oracle."""
import ctypes
import os

import numpy as np
import pytest

import oracle_lib as ol
import sc_polar_decoder_hls_b200 as scpd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = {}


def ss_emu(flags, n, par, q, ext, prune, llr, smem=18 * 1024, lsa=-1, lwin=-1, fuse=1, ltm=0, pre=2, xf=0):
    if "l" not in _LIB:
        _LIB["l"] = ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libss_emu.so"))
    out = np.zeros((len(llr), n // 32), np.uint32)
    st = (ctypes.c_uint64 * 6)()
    rc = _LIB["l"].ss_emu_decode(int(np.log2(n)), q, int(np.log2(par)), ext, prune, ol.P(flags), ol.P(llr),
                                 ctypes.c_size_t(len(llr)), ol.P(out), ctypes.c_size_t(smem), lsa, lwin, fuse, st, ltm, pre, xf)
    assert rc == 0, rc
    return out, [int(v) for v in st]


def _llrs(rng, n, k, q, nfr=26):
    ma = (1 << (q - 1)) - 1
    llr = ol.test_llrs(rng, n, nfr, k, maxabs=min(31, ma))
    llr[-1][rng.random(n) < 0.5] = 0   # CA2 zeros: the all-information shortcut must fall back
    llr[-2] = 0
    wide = rng.integers(-ma, ma + 1, size=(8, n)).astype(np.int8)  # the whole range of the internal saturation
    return np.concatenate([llr, wide])


# (LLR_BITS, PAR, EXTENDED) instantiated in the emulator
VARIANTS = [(8, 16, 1), (8, 16, 0), (6, 16, 1), (7, 16, 1), (8, 4, 1), (8, 8, 1), (6, 4, 0)]


@pytest.mark.parametrize("name,n,k,q,par,ext", [("FB_N512_K256", 512, 256) + v for v in VARIANTS] +
                         [("FB_N1024_K512", 1024, 512, 8, 16, 1)])
def test_ss_kernel_source_against_oracle(name, n, k, q, par, ext):
    """Plane arithmetic above 64 LLRs, plane -> fp16x2 transposition, the pattern-specialised 8-LLR routines, static
    scale / saturation inside and outside the un-saturated leaf decoder, every pruning mode, ragged last task."""
    flags = scpd.packed_flags(name, n)
    llr = _llrs(np.random.default_rng(q * 100 + par + ext), n, k, q)
    want = ol.decode_packed(n, par, q, 0, ext, flags, llr)
    for prune in (0, 1, 2):
        got, st = ss_emu(flags, n, par, q, ext, prune, llr)
        assert (got == want).all(), prune
    assert st[1] + st[2] > 0


@pytest.mark.parametrize("lsa,lwin,ltm", [(-1, -1, 0), (7, 8, 8), (7, 9, 9), (6, 8, 8), (8, 10, 9), (7, 10, 0), (6, 8, 0),
                                          (7, 9, 10), (6, 8, 11)])
def test_ss_kernel_storage_plans(lsa, lwin, ltm):
    """LLR levels in shared memory / tensor memory / workspace in every combination the planner can produce, partial
    sums in the shared window or moved to the workspace (c2: N = 4096)."""
    name, n, k = "frozen_n_4096_k_3072", 4096, 3072
    flags = scpd.packed_flags(name, n)
    llr = _llrs(np.random.default_rng(abs(lsa) * 10 + ltm), n, k, 8, nfr=10)
    want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr)
    for prune in (0, 1, 2):
        got, _ = ss_emu(flags, n, 16, 8, 1, prune, llr, smem=64 * 1024, lsa=lsa, lwin=lwin, ltm=ltm)
        assert (got == want).all(), prune


def test_ss_kernel_arbitrary_flag_tables():
    """Flag tables no polar construction produces: the 8-LLR nodes run the run-time fallback of the walker."""
    rng = np.random.default_rng(7)
    for n in (128, 256, 1024):
        for trial in range(3):
            flags = (rng.random(n) < rng.random()).astype(np.uint8)
            llr = rng.integers(-127, 128, size=(33, n)).astype(np.int8)
            llr[rng.random(llr.shape) < 0.3] = 0
            want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr)
            for prune in (0, 1, 2):
                got, _ = ss_emu(flags, n, 16, 8, 1, prune, llr)
                assert (got == want).all(), (n, trial, prune)


def test_ss_known_patterns_cover_the_packaged_tables():
    """The nine specialised 8-bit patterns are all that the packaged frozen tables contain (ss_plan.h, SS_KNOWN8)."""
    known = {0x00, 0xFF, 0xFE, 0xE8, 0x80, 0xE0, 0xFC, 0xF8, 0xC0}
    for name, n in (("FB_N1024_K512", 1024), ("frozen_n_4096_k_3072", 4096), ("frozen_n_32768_k_29492_snr_4_5", 32768),
                    ("frozen_n_131072_k_117964", 131072), ("frozen_n_524288_k_262144", 524288)):
        f = scpd.packed_flags(name, n).reshape(-1, 8)
        pats = set(np.packbits(f, axis=1, bitorder="little").ravel().tolist())
        assert pats <= known, (name, pats - known)


@pytest.mark.parametrize("name,n,k", [("FB_N512_K256", 512, 256), ("FB_N1024_K512", 1024, 512), ("frozen_n_4096_k_3072", 4096, 3072)])
def test_ss_leading_f_levels_computed_with_the_planes(name, n, k):
    """The first f ops of the walk depend on the channel alone: ss_planes_kernel computes them (0 ... 3 levels) and the
    schedule starts below; the leftmost nodes are then read from the plane buffer, also when their level otherwise
    lives in tensor memory, and a table whose left half is all-frozen (no leading f) prefuses nothing."""
    flags = scpd.packed_flags(name, n)
    llr = _llrs(np.random.default_rng(n), n, k, 8, nfr=8)
    want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr)
    depths = set()
    for pre in (0, 1, 2, 3):
        for ltm in (0, int(np.log2(n)) - 1):
            for prune in (0, 2):
                got, st = ss_emu(flags, n, 16, 8, 1, prune, llr, smem=10 * 1024, ltm=ltm if ltm >= 8 else 0, pre=pre)
                assert (got == want).all(), (pre, ltm, prune)
                assert st[4] <= pre
                depths.add(st[4])
    assert max(depths) >= (1 if n == 512 else 2)
    half = flags.copy()
    half[:n // 2] = 0
    got, st = ss_emu(half, n, 16, 8, 1, 2, llr, pre=3)
    assert st[4] == 0 and (got == ol.decode_packed(n, 16, 8, 0, 1, half, llr)).all()
    got, st = ss_emu(np.ones(n, np.uint8), n, 16, 8, 1, 2, llr, pre=3)   # all-information root: hard decision first
    assert st[4] == 0 and (got == ol.decode_packed(n, 16, 8, 0, 1, np.ones(n, np.uint8), llr)).all()


@pytest.mark.parametrize("lsa,ltm,pre", [(6, 0, 0), (7, 0, 2), (7, 9, 2), (6, 8, 1), (8, 0, 3)])
def test_ss_fused_f_g_with_the_childs_opening_f(lsa, ltm, pre):
    """SS_XF_*: an f / g op fused with the f that opens its child (both levels written, the middle one not read back), used
    for large trees where those levels stream through DRAM; here forced onto c2 so that it runs wherever the three levels
    live in global memory, in every pruning mode, with zero LLRs (the fallback walk behind an R1 op is never fused) and
    together with the levels computed by the plane conversion."""
    name, n, k = "frozen_n_4096_k_3072", 4096, 3072
    flags = scpd.packed_flags(name, n)
    llr = _llrs(np.random.default_rng(lsa * 10 + ltm), n, k, 8, nfr=6)
    want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr)
    for prune in (0, 1, 2):
        got, st = ss_emu(flags, n, 16, 8, 1, prune, llr, smem=64 * 1024, lsa=lsa, lwin=8, ltm=ltm, pre=pre, xf=1)
        assert (got == want).all(), prune
        assert st[5] > 0
        # xf = 3: fused ops and the smallest workspace level in its own array (the one an L2 persisting window covers)
        got, _ = ss_emu(flags, n, 16, 8, 1, prune, llr, smem=64 * 1024, lsa=lsa, lwin=8, ltm=ltm, pre=pre, xf=3)
        assert (got == want).all(), (prune, "hot")
    got, st = ss_emu(flags, n, 16, 8, 1, 2, llr, smem=64 * 1024, lsa=lsa, lwin=8, ltm=ltm, pre=pre, xf=0)
    assert (got == want).all() and st[5] == 0
    got, st = ss_emu(flags, n, 16, 8, 1, 2, llr, smem=64 * 1024, lsa=lsa, lwin=8, ltm=ltm, pre=pre, xf=2)
    assert (got == want).all() and st[5] == 0
