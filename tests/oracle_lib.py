"""ctypes view of oracle/liboracle.so and oracle/_ref/*.so for the tests (checker only)."""
import ctypes
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")


class OCfg(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("n", "par", "llr_bits", "format", "extended")]


class OStats(ctypes.Structure):
    _fields_ = [("visits", ctypes.c_uint64 * 24), ("with_zero", ctypes.c_uint64 * 24)]


def build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("sc_oracle.c", "sc_oracle.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build_oracle())
        L.sco_sigma.restype = ctypes.c_float
        L.sco_sigma.argtypes = [ctypes.c_float, ctypes.c_float]
        L.sco_xs128_uniform.restype = ctypes.c_float
        L.sco_xs128_uniform.argtypes = [ctypes.c_uint32]
        L.sco_quantize.argtypes = [ctypes.c_float]
        for fn in ("sco_f", "sco_g", "sco_g_ext", "sco_input"):
            getattr(L, fn).restype = ctypes.c_uint32
        _lib = L
    return _lib


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def decode(n, par, q, fmt, ext, flags, llr):
    """Oracle decode -> uint8 [B, n] codeword estimate."""
    llr = np.ascontiguousarray(llr, np.int8)
    flags = np.ascontiguousarray(flags, np.uint8)
    out = np.zeros(llr.shape, np.uint8)
    cfg = OCfg(n, par, q, fmt, ext)
    rc = lib().sco_decode(ctypes.byref(cfg), P(flags), P(llr), ctypes.c_size_t(llr.shape[0]), P(out))
    assert rc == 0, rc
    return out


def decode_l2(n, par, q, fmt, ext, flags, llr):
    """Oracle restatement of the reference built with PRUNING_LEVEL 2 (REP / SPC shortcuts): NOT plain SC."""
    llr = np.ascontiguousarray(llr, np.int8)
    flags = np.ascontiguousarray(flags, np.uint8)
    out = np.zeros(llr.shape, np.uint8)
    cfg = OCfg(n, par, q, fmt, ext)
    rc = lib().sco_decode_l2(ctypes.byref(cfg), P(flags), P(llr), ctypes.c_size_t(llr.shape[0]), P(out))
    assert rc == 0, rc
    return out


def decode_packed(n, par, q, fmt, ext, flags, llr, threads=1):
    llr = np.ascontiguousarray(llr, np.int8)
    flags = np.ascontiguousarray(flags, np.uint8)
    out = np.zeros((llr.shape[0], max(1, n // 32)), np.uint32)
    cfg = OCfg(n, par, q, fmt, ext)
    rc = lib().sco_decode_packed_mt(ctypes.byref(cfg), P(flags), P(llr), ctypes.c_size_t(llr.shape[0]), P(out),
                                    threads)
    assert rc == 0, rc
    return out


def channel(n, nframes, sigma, first_frame=0, seed=0xF0, codeword=None):
    out = np.zeros((nframes, n), np.int8)
    cw, per = None, 0
    if codeword is not None:
        cwa = np.ascontiguousarray(codeword, np.uint8)
        cw, per = P(cwa), int(cwa.ndim == 2)
    lib().sco_channel(n, ctypes.c_size_t(first_frame), ctypes.c_size_t(nframes), ctypes.c_uint8(seed),
                      ctypes.c_float(sigma), cw, per, P(out))
    return out


def count_errors(n, xhat_bits, ref=None):
    xhat_bits = np.ascontiguousarray(xhat_bits, np.uint8)
    cnt = (ctypes.c_uint64 * 6)()
    r, per = None, 0
    if ref is not None:
        ra = np.ascontiguousarray(ref, np.uint8)
        r, per = P(ra), int(ra.ndim == 2)
    lib().sco_count_errors(n, ctypes.c_size_t(xhat_bits.shape[0]), P(xhat_bits), r, per, cnt)
    return [int(v) for v in cnt]


def polar_transform(bits):
    b = np.ascontiguousarray(bits, np.uint8).copy()
    flat = b.reshape(-1, b.shape[-1])
    for row in flat:
        lib().sco_polar_transform(P(row), row.shape[0])
    return flat.reshape(b.shape)


def sigma(ebn0_db, rate):
    return float(lib().sco_sigma(ebn0_db, rate))


def pack_bits(bits):
    """uint8 [B, n] -> uint32 [B, max(1, n/32)] LSB-first (layout of scpd_decode's output)."""
    bits = np.ascontiguousarray(bits, np.uint8)
    n = bits.shape[-1]
    if n < 32:
        pad = np.zeros(bits.shape[:-1] + (32 - n,), np.uint8)
        bits = np.concatenate([bits, pad], axis=-1)
    return np.packbits(bits, axis=-1, bitorder="little").view(np.uint32)


def unpack_bits(words, n):
    w = np.ascontiguousarray(words).view(np.uint8)
    return np.unpackbits(w, axis=-1, bitorder="little")[..., :n]


def golden_codewords():
    cw = json.load(open(os.path.join(GOLDEN, "codewords.json")))
    return {k: np.array([[int(c) for c in row] for row in v], np.uint8) for k, v in cw.items()}


def ref_lib(tag):
    """oracle/_ref/refdec_<tag>.so (reference sources compiled natively) or None if not built."""
    so = os.path.join(ORACLE_DIR, "_ref", f"refdec_{tag}.so")
    if not os.path.exists(so):
        return None
    return ctypes.CDLL(so)


def ref_decode(R, flags, llr):
    cfg = (ctypes.c_int32 * 6)()
    R.ref_config(cfg)
    n = cfg[0]
    llr = np.ascontiguousarray(llr, np.int8)
    flags = np.ascontiguousarray(flags, np.uint8)
    out = np.zeros(llr.shape, np.uint8)
    rc = R.ref_decode(P(flags), P(llr), ctypes.c_size_t(llr.shape[0]), P(out))
    assert rc == 0 and llr.shape[1] == n
    return out


def test_llrs(rng, n, nframes, k, ebn0=2.5, maxabs=31, seed=0xF0):
    """Half uniformly random over the quantiser alphabet, half channel output (all-zero codeword)."""
    llr = rng.integers(-maxabs, maxabs + 1, size=(nframes, n)).astype(np.int8)
    h = nframes // 2
    if h:
        llr[:h] = channel(n, h, sigma(ebn0, k / n), seed=seed)
    return llr
