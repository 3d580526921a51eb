"""CPU tier: the channel-chain restatement against the known answers of SURVEY.md App. C."""
import ctypes

import numpy as np

import oracle_lib as ol


class Xs(ctypes.Structure):
    _fields_ = [(k, ctypes.c_uint32) for k in "xyzw"]


def test_xorshift128_known_answers():
    L = ol.lib()
    L.sco_xs128_next.restype = ctypes.c_uint32
    a, b = Xs(), Xs()
    L.sco_xs128_seed(ctypes.byref(a), ctypes.byref(b), ctypes.c_uint8(0xF0))
    assert (a.x, a.y, a.z, a.w) == (0x10301070, 0xF5F9F7F2, 0x0E2C5AF1, 0xE57A9269)
    assert (b.x, b.y, b.z, b.w) == (0x90705030, 0xF2F4F6F8, 0x0C4A2E0F, 0x03030302)
    assert [L.sco_xs128_next(ctypes.byref(a)) for _ in range(4)] == [0x7559AD26, 0x4F258218, 0x23B2A2DF, 0x126A951A]
    assert [L.sco_xs128_next(ctypes.byref(b)) for _ in range(4)] == [0x11E02282, 0x44F65570, 0x199131B7, 0x02913A94]


def test_sigma_and_quantiser():
    assert abs(ol.sigma(2.5, 0.5) - 0.7498942) < 1e-6  # main.cpp:91-98
    L = ol.lib()
    q = lambda y: L.sco_quantize(ctypes.c_float(y))
    assert [q(1.0), q(-1.0), q(0.24), q(-0.24), q(0.25), q(100.0), q(-100.0), q(7.76)] == [4, -4, 0, 0, 1, 31, -31, 31]


def test_channel_statistics_and_stream_positions():
    n = 1024
    sig = ol.sigma(2.5, 0.5)
    llr = ol.channel(n, 64, sig)
    assert llr.min() >= -31 and llr.max() <= 31
    assert abs((llr == 0).mean() - 0.11) < 0.01  # ~11 % zeros at 2.5 dB
    assert abs(llr.mean() - 4.0 * 0.93) < 0.5
    # frame f of a long run == a run started at frame f (draws [f n/2, (f+1) n/2) of both streams)
    assert (ol.channel(n, 3, sig, first_frame=61) == llr[61:64]).all()
    cw = (np.arange(n) % 2).astype(np.uint8)
    neg = ol.channel(n, 4, sig, codeword=cw)
    assert neg[:, 1::2].mean() < -3 and neg[:, 0::2].mean() > 3
    # first noise pair of App. C: (-0.4704, 1.0026) -> y = 1 + sigma*n -> LLR = trunc(4y)
    assert llr[0, 0] == int(4 * (1 + sig * -0.4704)) and llr[0, 1] == int(4 * (1 + sig * 1.0026))


def test_error_counter_wrap():
    n = 2048
    x = np.zeros((3, n), np.uint8)
    x[0, :1024] = 1   # 1024 errors: wraps to 0 in the reference's sc_uint<10>
    x[1, :5] = 1
    c = ol.count_errors(n, x)
    assert c == [1029, 2, 3 * n, 3, 5, 1]
