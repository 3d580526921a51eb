"""CPU tier for the bit-sliced kernel (csrc/decode_bs.cuh, bs_arith.cuh): the kernel source executed lane
by lane under the warp emulator (tests/emu/warp_emu.cpp) against the oracle, in both number formats."""
import ctypes
import os

import numpy as np
import pytest

import oracle_lib as ol
import sc_polar_decoder_hls_b200 as scpd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = {}


def _emu():
    if "l" not in _LIB:
        _LIB["l"] = ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libwarp_emu.so"))
    return _LIB["l"]


def bs_emu(fmt, g, flags, n, par, q, ext, prune, llr, lsa=-1, lsb=-1, smem=56000, warps=1, grid=1, fuse=3):
    out = np.zeros((len(llr), n // 32), np.uint32)
    rc = _emu().emu_bs_decode(fmt, g, int(np.log2(n)), int(np.log2(par)), q, ext, prune, ol.P(flags), ol.P(llr),
                              ctypes.c_size_t(len(llr)), ol.P(out), lsa, lsb, smem, warps, grid, fuse)
    assert rc == 0, rc
    return out


# (format, LLR_BITS, PAR, EXTENDED, lanes per frame group) instantiated in the emulator
VARIANTS = [(0, 8, 16, 1, 32), (0, 8, 16, 1, 16), (0, 8, 16, 1, 8), (0, 8, 16, 0, 8), (0, 6, 16, 1, 16),
            (1, 6, 16, 1, 8), (1, 8, 16, 1, 32), (1, 6, 16, 0, 16), (0, 7, 64, 1, 8), (1, 7, 64, 1, 32),
            (0, 8, 2, 1, 8)]


@pytest.mark.parametrize("fmt,q,par,ext,g", VARIANTS)
def test_bs_kernel_source_under_warp_emulator(fmt, q, par, ext, g):
    """Plane arithmetic (f, saturating / un-saturated g, 2-bit terminals), width growth inside the leaf
    decoder, fused 16-LLR subtree walk, every pruning mode, ragged last group, CA2 zero fallback."""
    n, k = 512, 256
    flags = scpd.packed_flags("FB_N512_K256", n)
    rng = np.random.default_rng(100 * fmt + q + par + g)
    llr = ol.test_llrs(rng, n, 70, k, maxabs=min(31, (1 << (q - 1)) - 1))
    llr[-1][rng.random(n) < 0.5] = 0
    llr[-2] = 0
    want = ol.decode_packed(n, par, q, fmt, ext, flags, llr)
    for prune in (0, 1, 2):
        assert (bs_emu(fmt, g, flags, n, par, q, ext, prune, llr) == want).all(), prune


@pytest.mark.parametrize("name,n,k", [("FB_N1024_K512", 1024, 512), ("frozen_n_4096_k_3072", 4096, 3072)])
def test_bs_kernel_storage_split_and_grid(name, n, k):
    """Alpha levels / partial sums pushed out to the workspace, several warps and CTAs, full-range LLRs."""
    flags = scpd.packed_flags(name, n)
    rng = np.random.default_rng(n)
    for g, lsa, lsb, warps, grid, nfr in ((32, 6, 6, 1, 1, 33), (8, 7, 6, 2, 2, 250), (16, 8, 9, 2, 1, 97)):
        llr = ol.test_llrs(rng, n, nfr, k, maxabs=127)
        llr[-1][rng.random(n) < 0.5] = 0
        want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr)
        for fuse in (0, 2, 3):  # X(l) F(l-1) [F(l-2)] fused into one pass, or every op on its own
            got = bs_emu(0, g, flags, n, 16, 8, 1, 2, llr, lsa, lsb, 50000, warps, grid, fuse)
            assert (got == want).all(), (g, lsa, lsb, fuse)


def test_bs_kernel_random_flag_tables():
    rng = np.random.default_rng(12)
    for n in (128, 256):
        for _ in range(5):
            flags = (rng.random(n) < rng.random()).astype(np.uint8)
            llr = rng.integers(-31, 32, size=(40, n)).astype(np.int8)
            for fmt, q, g in ((0, 8, 16), (1, 6, 8)):
                want = ol.decode_packed(n, 16, q, fmt, 1, flags, llr)
                for prune in (0, 2):
                    assert (bs_emu(fmt, g, flags, n, 16, q, 1, prune, llr) == want).all(), (n, fmt, prune)


def test_bs_schedule_compiler_invariants():
    """Host logic of csrc/bs_plan.h: one word per word of the generic schedule (so the skip counts behind an
    all-information node stay valid), fused passes only on generic levels and followed by exactly the no-op
    words they subsume, and no skip ever lands on a no-op."""
    lib = _emu()
    lib.emu_bs_schedule.restype = ctypes.c_longlong
    LLOW, LSUB, NOP = 6, 4, 8 * 7
    for name, n in (("FB_N1024_K512", 1024), ("frozen_n_4096_k_3072", 4096), ("frozen_n_32768_k_29492_snr_4_5", 32768)):
        flags = scpd.packed_flags(name, n)
        for fmt in (0, 1):
            for fuse in (0, 2, 3):
                buf = np.zeros(1 << 18, np.uint32)
                gw = ctypes.c_uint32()
                m = lib.emu_bs_schedule(fmt, int(np.log2(n)), 4, 1, 2, ol.P(flags), fuse, ol.P(buf),
                                        ctypes.c_size_t(len(buf)), ctypes.byref(gw))
                assert m == gw.value and 0 < m <= len(buf)
                w = buf[:m]
                code, lev = w & 63, (w >> 6) & 31
                i, targets, nfused = 0, set(), 0
                while i < m:
                    c = int(code[i])
                    if c == 0:
                        i += 1
                    elif c == 8 * LSUB + 1:  # fused subtree + node types
                        i += 2
                    elif c % 8 == 7 and c >= 8 * 5:  # R1 of a specialised or generic level: skip count follows
                        skip = int(w[i + 1])
                        assert fmt == 0 or skip == 0
                        targets.add(i + 2 + skip)
                        i += 2
                    elif c in (9, 10, 11, 17, 18, 19):  # X(l) F(l-1) [F(l-2)]
                        depth = 2 if c < 16 else 3
                        assert fuse >= depth and lev[i] - depth + 1 > LLOW
                        for d in range(1, depth):
                            assert code[i + d] == NOP and lev[i + d] == lev[i] - d
                        nfused += 1
                        i += depth
                    else:
                        assert c != NOP, "a no-op that no fused pass owns"
                        i += 1
                assert (nfused > 0) == (fuse >= 2)
                for t in targets:
                    assert t < m and code[t] != NOP


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("q", [6, 8])
def test_bs_element_functions_exhaustive(fmt, q):
    """bs_arith.cuh f / saturating g / un-saturated g against the oracle's element functions
    (oracle/sc_oracle.c: sco_f, sco_g, sco_g_ext) on EVERY pair of Q-bit values and both partial-sum bits.
    CA2: same numeric value (a zero may carry either sign internally); SIGMAG: same (sign, magnitude) pattern."""
    L = ol.lib()
    P = q - 1
    if fmt == 0:
        vals = np.arange(-(2 ** P - 1), 2 ** P)
        sgn, mag = (vals < 0).astype(np.uint8), np.abs(vals).astype(np.uint32)
        pat = (vals & (2 ** q - 1)).astype(np.uint32)
    else:
        pat = np.arange(2 ** q, dtype=np.uint32)
        sgn, mag = (pat >> P).astype(np.uint8), (pat & (2 ** P - 1)).astype(np.uint32)
    ia, ib, iu = np.meshgrid(np.arange(len(pat)), np.arange(len(pat)), np.arange(2), indexing="ij")
    ia, ib, iu = ia.ravel(), ib.ravel(), iu.ravel().astype(np.uint8)
    sa, ma, sb, mb = (np.ascontiguousarray(x) for x in (sgn[ia], mag[ia], sgn[ib], mag[ib]))
    so, mo = np.zeros(len(ia), np.uint8), np.zeros(len(ia), np.uint32)
    for op, fn, wout in ((0, L.sco_f, q), (1, L.sco_g, q), (2, L.sco_g_ext, q + 1)):
        rc = _emu().emu_bs_prim(fmt, q, op, ctypes.c_size_t(len(ia)), ol.P(sa), ol.P(ma), ol.P(sb), ol.P(mb), ol.P(iu),
                                ol.P(so), ol.P(mo))
        assert rc == 0
        step = max(1, len(ia) // 60000)  # the oracle is called element by element from Python: sample evenly
        for i in range(0, len(ia), step):
            a, b, u = int(pat[ia[i]]), int(pat[ib[i]]), int(iu[i])
            want = fn(fmt, q, a, b) if op == 0 else fn(fmt, q, a, b, u)
            if fmt == 0:
                got = -int(mo[i]) if so[i] else int(mo[i])
                assert got == L.sco_value(0, wout, want), (op, a, b, u)
            else:
                assert (int(so[i]) << (wout - 1)) | int(mo[i]) == want, (op, a, b, u)


@pytest.mark.parametrize("sync_every", [0, 3])
def test_bs_kernel_whole_cta_with_schedule_in_shared_memory(sync_every):
    """The CTA-level paths of the headline kernel: all 4 warps of a CTA emulated concurrently with a real barrier,
    the schedule copied into shared memory by the CTA (the default below 8 KB of schedule), optional barrier ops
    (SCPD_BS_SYNC), several rounds per CTA and a ragged last round (warps without frames still reach every barrier)."""
    n, k = 1024, 512
    flags = scpd.packed_flags("FB_N1024_K512", n)
    rng = np.random.default_rng(5 + sync_every)
    llr = ol.test_llrs(rng, n, 32 * 9 + 5, k)  # 10 frame groups over 2 CTAs of 4 warps: 2 rounds, the second ragged
    llr[-1][::3] = 0
    want = ol.decode_packed(n, 16, 8, 0, 1, flags, llr)
    _emu().emu_bs_cta_mode(1, sync_every)
    try:
        for g, prune in ((32, 2), (16, 0)):
            got = bs_emu(0, g, flags, n, 16, 8, 1, prune, llr, smem=7168, warps=4, grid=2 if g == 32 else 1, fuse=0)
            assert (got == want).all(), (g, prune)
    finally:
        _emu().emu_bs_cta_mode(0, 0)
