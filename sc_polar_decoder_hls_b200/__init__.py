"""sc_polar_decoder_hls_b200 -- B200-native batched SC polar decoder (ctypes view of the C ABI).

The product is libscpd.so (include/scpd.h): hand-written sm_100a kernels behind a C ABI that
mirrors the reference's decoder module (src/module/my_module.h) and testbench harness.  This
module only binds it for the Python tests and bench.py; PyTorch is used by callers for device
memory and streams, never for the decode itself.  There is no CPU fallback: importing fails
loudly when the library is missing.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCPD_LIB_PATH") or os.path.join(_HERE, "libscpd.so")  # the override is a development aid
DATA_DIR = os.path.join(_HERE, "data")

FMT_CA2, FMT_SIGMAG = 0, 1
PRUNE_NONE, PRUNE_R0, PRUNE_R0_R1, PRUNE_REF_LEVEL2 = 0, 1, 2, 3
OK, E_ARG, E_CONFIG, E_UNSUPPORTED, E_IO, E_CUDA, E_NOMEM = range(7)


class Config(ctypes.Structure):
    """scpd_config: replaces config.h:2-16 and polar_parameters.h:4-11 of the reference."""
    _fields_ = [(k, ctypes.c_uint32) for k in
                ("n", "k", "par", "llr_bits", "format", "extended", "pruning", "reserved")]


class StageMatrix(ctypes.Structure):
    """scpd_stage_matrix: [function][level] node visits and loop iterations per frame."""
    _fields_ = [("visits", (ctypes.c_uint64 * 32) * 8), ("iterations", (ctypes.c_uint64 * 32) * 8),
                ("total_iterations", ctypes.c_uint64)]


STAGE_FUNCS = ("F", "G", "H", "R", "R0", "R1", "REP", "SPC")


class ScpdError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"scpd status {status}: {msg}")
        self.status = status


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, u8p = c.c_void_p, c.c_void_p
    sig = {
        "scpd_frozen_load_order": (c.c_int, [c.c_char_p, c.c_uint32, c.c_uint32, u8p]),
        "scpd_frozen_load_flags": (c.c_int, [c.c_char_p, c.c_uint32, u8p, c.POINTER(c.c_uint32)]),
        "scpd_frozen_write_order": (c.c_int, [c.c_char_p, c.c_uint32, vp]),
        "scpd_frozen_write_flags": (c.c_int, [c.c_char_p, c.c_uint32, u8p]),
        "scpd_write_polar_parameters": (c.c_int, [c.c_char_p, c.c_uint32, c.c_uint32, c.c_int, u8p]),
        "scpd_create": (c.c_int, [c.POINTER(Config), u8p, c.c_int, c.POINTER(vp)]),
        "scpd_destroy": (None, [vp]),
        "scpd_decode": (c.c_int, [vp, vp, c.c_size_t, vp, vp]),
        "scpd_decode_host": (c.c_int, [vp, vp, c.c_size_t, vp]),
        "scpd_validate_llr": (c.c_int, [vp, vp, c.c_size_t, c.POINTER(c.c_uint64), vp]),
        "scpd_extract_info": (c.c_int, [vp, vp, c.c_size_t, vp, vp]),
        "scpd_get_config": (c.c_int, [vp, c.POINTER(Config)]),
        "scpd_schedule_stats": (c.c_int, [vp, c.POINTER(c.c_uint64), c.POINTER(c.c_uint64)]),
        "scpd_launch_count": (c.c_uint64, [vp]),
        "scpd_kernel_timing": (c.c_int, [vp, c.c_int]),
        "scpd_last_kernel_ms": (c.c_int, [vp, c.POINTER(c.c_float)]),
        "scpd_kernel_name": (c.c_char_p, [vp]),
        "scpd_last_kernel_name": (c.c_char_p, [vp]),
        "scpd_stage_profile": (c.c_int, [c.POINTER(Config), u8p, c.POINTER(StageMatrix)]),
        "scpd_sigma": (c.c_float, [c.c_float, c.c_float]),
        "scpd_channel_mode": (c.c_int, [c.c_int]),
        "scpd_channel_generate": (c.c_int, [c.c_uint32, c.c_uint64, c.c_size_t, c.c_uint8, c.c_float, vp,
                                            c.c_int, vp, vp]),
        "scpd_count_errors": (c.c_int, [c.c_uint32, c.c_size_t, vp, vp, c.c_int, vp, vp]),
        "scpd_run_ber": (c.c_int, [vp, c.c_float, c.c_float, c.c_uint64, c.c_uint64, c.c_uint8, u8p,
                                   c.POINTER(c.c_uint64)]),
        "scpd_run_ber_ex": (c.c_int, [vp, c.c_float, c.c_float, c.c_uint64, c.c_uint64, c.c_uint8, c.c_int, u8p, c.c_uint32,
                                      c.c_uint64, c.POINTER(c.c_uint64)]),
        "scpd_stage_timing": (c.c_int, [vp, c.c_int]),
        "scpd_stage_time": (c.c_int, [vp, vp, vp]),
        "scpd_r1_votes": (c.c_int, [vp, vp, vp]),
        "scpd_last_error": (c.c_char_p, []),
        "scpd_status_string": (c.c_char_p, [c.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
EXPORTS = ["scpd_frozen_load_order", "scpd_frozen_load_flags", "scpd_frozen_write_order",
           "scpd_frozen_write_flags", "scpd_write_polar_parameters", "scpd_create", "scpd_destroy",
           "scpd_decode", "scpd_decode_host", "scpd_validate_llr", "scpd_extract_info", "scpd_get_config",
           "scpd_schedule_stats", "scpd_launch_count", "scpd_kernel_timing", "scpd_last_kernel_ms",
           "scpd_kernel_name", "scpd_last_kernel_name", "scpd_stage_profile", "scpd_sigma", "scpd_channel_generate", "scpd_channel_mode",
           "scpd_count_errors", "scpd_run_ber", "scpd_run_ber_ex", "scpd_stage_timing", "scpd_stage_time", "scpd_r1_votes", "scpd_last_error",
           "scpd_status_string"]
SRC_CODEWORDS, SRC_RANDOM = 0, 1


def check(status):
    if status != OK:
        raise ScpdError(status, lib.scpd_last_error().decode(errors="replace"))


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ------------------------------------------------------------------ frozen tables
def load_order(path, n, k):
    """Frozen_Bit_Tab/FB_N*_K*.txt -> n information flags (Writer.h:35-93)."""
    out = np.zeros(n, np.uint8)
    check(lib.scpd_frozen_load_order(os.fsencode(path), n, k, _np_ptr(out)))
    return out


def load_flags(path, n):
    """Generated_Frozen_Bit/frozen_n_*_k_*.txt -> n information flags (Writer.h:95-105)."""
    out = np.zeros(n, np.uint8)
    k = ctypes.c_uint32()
    check(lib.scpd_frozen_load_flags(os.fsencode(path), n, _np_ptr(out), ctypes.byref(k)))
    return out, int(k.value)


def write_flags(path, flags):
    flags = np.ascontiguousarray(flags, np.uint8)
    check(lib.scpd_frozen_write_flags(os.fsencode(path), len(flags), _np_ptr(flags)))


def write_order(path, order):
    order = np.ascontiguousarray(order, np.uint32)
    check(lib.scpd_frozen_write_order(os.fsencode(path), len(order), _np_ptr(order)))


def write_polar_parameters(path, par, flags, en=0):
    flags = np.ascontiguousarray(flags, np.uint8)
    check(lib.scpd_write_polar_parameters(os.fsencode(path), len(flags), par, en, _np_ptr(flags)))


def packed_flags(name, n):
    """Information flags of a packaged frozen set (data/<name>.bits, see tools/make_golden.py)."""
    raw = np.frombuffer(open(os.path.join(DATA_DIR, name + ".bits"), "rb").read(), np.uint8)
    return np.unpackbits(raw, bitorder="little")[:n].copy()


def sigma(ebn0_db, rate):
    return float(lib.scpd_sigma(ebn0_db, rate))


def stage_profile(n, k, flags, par=16, pruning=PRUNE_NONE):
    """The function x level matrix of the reference's sc_monitor (sc_monitor.h:50-441) from the frozen table:
    returns (visits, iterations, total) with visits / iterations numpy [8, 32] indexed [STAGE_FUNCS, log2 node size]."""
    flags = np.ascontiguousarray(flags, np.uint8)
    cfg = Config(n, k, par, 8, FMT_CA2, 1, pruning, 0)
    m = StageMatrix()
    check(lib.scpd_stage_profile(ctypes.byref(cfg), _np_ptr(flags), ctypes.byref(m)))
    return (np.array(m.visits, np.uint64).reshape(8, 32), np.array(m.iterations, np.uint64).reshape(8, 32),
            int(m.total_iterations))


# ------------------------------------------------------------------ decoder handle
class Decoder:
    """Host mirror of SC_MODULE(my_module): FB port at construction, e -> s per decode call."""

    def __init__(self, n, k, flags, par=16, llr_bits=8, fmt=FMT_CA2, extended=1, pruning=PRUNE_R0_R1,
                 device=0):
        flags = np.ascontiguousarray(flags, np.uint8)
        if flags.shape != (n,):
            raise ScpdError(E_ARG, "flags must have n entries")
        self.cfg = Config(n, k, par, llr_bits, fmt, extended, pruning, 0)
        self.n, self.k = n, k
        self.wpf = max(1, n // 32)
        self.device = device
        h = ctypes.c_void_p()
        check(lib.scpd_create(ctypes.byref(self.cfg), _np_ptr(flags), device, ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and lib is not None:
            lib.scpd_destroy(self._h)
        self._h = None

    __del__ = close

    def decode_ptr(self, d_llr, nframes, d_xhat, stream=0):
        check(lib.scpd_decode(self._h, d_llr, nframes, d_xhat, stream))

    def decode(self, llr, out=None, stream=None):
        """llr: torch int8 CUDA tensor [B, n] -> int32 CUDA tensor [B, n/32] of packed x^."""
        import torch
        assert llr.is_cuda and llr.dtype == torch.int8 and llr.is_contiguous() and llr.shape[-1] == self.n
        b = llr.shape[0]
        if out is None:
            out = torch.empty((b, self.wpf), dtype=torch.int32, device=llr.device)
        s = torch.cuda.current_stream(llr.device).cuda_stream if stream is None else stream
        self.decode_ptr(llr.data_ptr(), b, out.data_ptr(), s)
        return out

    def decode_host(self, llr):
        """llr: numpy int8 [B, n] (host) -> numpy uint32 [B, n/32]; copies inside the call."""
        llr = np.ascontiguousarray(llr, np.int8)
        out = np.zeros((llr.shape[0], self.wpf), np.uint32)
        check(lib.scpd_decode_host(self._h, _np_ptr(llr), llr.shape[0], _np_ptr(out)))
        return out

    def validate_llr(self, llr):
        """Number of LLRs of a CUDA int8 batch outside the input contract |llr| <= 2^(llr_bits-1) - 1."""
        import torch
        out = ctypes.c_uint64()
        s = torch.cuda.current_stream(llr.device).cuda_stream
        check(lib.scpd_validate_llr(self._h, llr.data_ptr(), llr.shape[0], ctypes.byref(out), s))
        return int(out.value)

    def extract_info(self, xhat, stream=None):
        import torch
        out = torch.empty_like(xhat)
        s = torch.cuda.current_stream(xhat.device).cuda_stream if stream is None else stream
        check(lib.scpd_extract_info(self._h, xhat.data_ptr(), xhat.shape[0], out.data_ptr(), s))
        return out

    def schedule_stats(self):
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        check(lib.scpd_schedule_stats(self._h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    @property
    def launches(self):
        return int(lib.scpd_launch_count(self._h))

    @property
    def kernel_name(self):
        return lib.scpd_kernel_name(self._h).decode()

    @property
    def last_kernel_name(self):
        """The kernel the last decode launched (the dispatch depends on the batch size)."""
        return lib.scpd_last_kernel_name(self._h).decode()

    def kernel_timing(self, enable=True):
        check(lib.scpd_kernel_timing(self._h, 1 if enable else 0))

    def last_kernel_ms(self):
        """Duration of the tree-walk kernel of the last decode (CUDA events on the decode's stream)."""
        ms = ctypes.c_float()
        check(lib.scpd_last_kernel_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def run_ber(self, ebn0_db, rate, nframes, first_frame=0, seed=0xF0, codeword=None):
        """src/testbench/main.cpp on the device: returns the six counters of scpd_run_ber."""
        cnt = (ctypes.c_uint64 * 6)()
        cw = None
        if codeword is not None:
            cwa = np.ascontiguousarray(codeword, np.uint8)
            cw = _np_ptr(cwa)
        check(lib.scpd_run_ber(self._h, ebn0_db, rate, first_frame, nframes, seed, cw, cnt))
        return [int(v) for v in cnt]

    def run_ber_ex(self, ebn0_db, rate, nframes, first_frame=0, seed=0xF0, codewords=None, random_payload=False,
                   payload_seed=1):
        """scpd_run_ber_ex: codewords [ncw, n] sent in turn (the reference's sc_encoder with ncw = 3), or random
        information words encoded on the device; ten counters (codeword bits, then information bits)."""
        cnt = (ctypes.c_uint64 * 10)()
        cw, ncw = None, 0
        if codewords is not None:
            cwa = np.ascontiguousarray(np.atleast_2d(codewords), np.uint8)
            cw, ncw = _np_ptr(cwa), cwa.shape[0]
        check(lib.scpd_run_ber_ex(self._h, ebn0_db, rate, first_frame, nframes, seed,
                                  SRC_RANDOM if random_payload else SRC_CODEWORDS, cw, ncw, payload_seed, cnt))
        return [int(v) for v in cnt]

    def stage_timing(self, enable=True):
        check(lib.scpd_stage_timing(self._h, 1 if enable else 0))

    def r1_votes(self):
        """(votes, fallbacks): uint64 [32] per log2 node size since the last call (scpd_r1_votes; stage timing on)."""
        v, f = np.zeros(32, np.uint64), np.zeros(32, np.uint64)
        check(lib.scpd_r1_votes(self._h, _np_ptr(v), _np_ptr(f)))
        return v, f

    def stage_time(self):
        """(cycles, visits): uint64 [6, 32] per (function, level) since the last call (scpd_stage_time)."""
        cyc, vis = np.zeros((6, 32), np.uint64), np.zeros((6, 32), np.uint64)
        check(lib.scpd_stage_time(self._h, _np_ptr(cyc), _np_ptr(vis)))
        return cyc, vis


def payload_words(seed, first_frame, nframes, info_flags):
    """Host restatement of the device payload source of scpd_run_ber_ex (harness.cuh: payload_word): uint32
    [nframes, n/32] information words, frozen positions 0."""
    n = len(info_flags)
    wpf = max(1, n // 32)
    m64 = (1 << 64) - 1
    f = (np.arange(nframes, dtype=np.uint64) + np.uint64(first_frame))[:, None]
    w = np.arange(wpf, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        z = (np.uint64((seed * 0x9E3779B97F4A7C15) & m64) + f * np.uint64(0xBF58476D1CE4E5B9) + w * np.uint64(0x94D049BB133111EB))
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    flags = np.zeros(wpf * 32, np.uint8)
    flags[:n] = info_flags
    mask = np.packbits(flags, bitorder="little").view(np.uint32)
    return (z & np.uint64(0xFFFFFFFF)).astype(np.uint32) & mask[None, :]


def channel_mode(mode=-1):
    """scpd_channel_mode: 0 libm-grade Box-Muller, 2 guarded SFU approximations (default, same LLRs), 1 approximations
    only; returns the previous mode (mode < 0: only reads it)."""
    return int(lib.scpd_channel_mode(mode))


def channel_generate(n, nframes, sigma_v, first_frame=0, seed=0xF0, codeword=None, device=0, stream=None):
    """Device channel (scpd_channel_generate) -> torch int8 [nframes, n]."""
    import torch
    dev = torch.device("cuda", device)
    out = torch.empty((nframes, n), dtype=torch.int8, device=dev)
    cw_ptr, per_frame = None, 0
    if codeword is not None:
        cw = torch.as_tensor(np.ascontiguousarray(codeword, np.uint8)).to(dev)
        per_frame = 1 if cw.dim() == 2 else 0  # (packed rows, mode 2, are an internal format of scpd_run_ber_ex)
        cw_ptr = cw.data_ptr()
    s = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
    with torch.cuda.device(dev):
        check(lib.scpd_channel_generate(n, first_frame, nframes, seed, sigma_v, cw_ptr, per_frame,
                                        out.data_ptr(), s))
        if codeword is not None:
            torch.cuda.current_stream(dev).synchronize()  # keep cw alive until the kernel has read it
    return out


def count_errors(n, xhat, ref_words=None, stream=None):
    """scpd_count_errors on a torch int32 [B, n/32] tensor -> list of six ints."""
    import torch
    dev = xhat.device
    cnt = torch.zeros(6, dtype=torch.int64, device=dev)
    ref_ptr, per_frame = None, 0
    if ref_words is not None:
        ref = torch.as_tensor(np.ascontiguousarray(ref_words).view(np.int32)).to(dev)
        per_frame = 1 if ref.dim() == 2 else 0
        ref_ptr = ref.data_ptr()
    s = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
    with torch.cuda.device(dev):
        check(lib.scpd_count_errors(n, xhat.shape[0], xhat.data_ptr(), ref_ptr, per_frame, cnt.data_ptr(), s))
        return [int(v) for v in cnt.cpu()]
