"""CPU tier: frozen-table loaders / writers of the C ABI (no GPU needed)."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
import sc_polar_decoder_hls_b200 as scpd

REF = "/root/reference"
INDEX = json.load(open(os.path.join(ol.GOLDEN, "frozen_index.json")))


def _packed_sha(flags):
    import hashlib
    return hashlib.sha256(np.packbits(flags, bitorder="little").tobytes()).hexdigest()


def test_flag_file_roundtrip(tmp_path):
    for name, n in (("frozen_n_4096_k_3072", 4096), ("frozen_n_131072_k_117964", 131072), ("FB_N8_K4", 8)):
        flags = scpd.packed_flags(name, n)
        p = tmp_path / (name + ".txt")
        scpd.write_flags(p, flags)
        got, k = scpd.load_flags(p, n)
        assert (got == flags).all() and k == flags.sum()
        assert os.path.getsize(p) == 2 * n  # "N tokens separated by one space, no newline" (App. B)


def test_order_file_roundtrip_and_subsetting(tmp_path):
    rng = np.random.default_rng(0)
    n = 1024
    order = rng.permutation(n).astype(np.uint32)
    p = tmp_path / "FB_N1024_K300.txt"
    scpd.write_order(p, order)
    txt = open(p, "rb").read()
    assert txt.startswith(b"1024\n0\n0\n") and txt.endswith(b"    ")
    flags = scpd.load_order(p, n, 300)
    assert flags.sum() == 300 and set(np.nonzero(flags)[0]) == set(order[:300].tolist())
    # a table for N=1024 also serves N=256: indices >= 256 are dropped first (Writer.h:64-71)
    sub = scpd.load_order(p, 256, 100)
    want = [v for v in order if v < 256][:100]
    assert set(np.nonzero(sub)[0]) == set(want)
    # CRLF files as shipped in Frozen_Bit_Tab/
    crlf = tmp_path / "crlf.txt"
    open(crlf, "wb").write(b"1024\r\n0\r\n0\r\n" + b"    ".join(str(int(v)).encode() for v in order) + b"    ")
    assert (scpd.load_order(crlf, n, 300) == flags).all()


def test_loader_errors(tmp_path):
    with pytest.raises(scpd.ScpdError) as e:
        scpd.load_flags(tmp_path / "missing.txt", 8)
    assert e.value.status == scpd.E_IO
    p = tmp_path / "short.txt"
    open(p, "w").write("0 1 1")
    with pytest.raises(scpd.ScpdError):
        scpd.load_flags(p, 8)
    open(p, "w").write("0 1 2 0 0 0 0 0")
    with pytest.raises(scpd.ScpdError):
        scpd.load_flags(p, 8)
    open(p, "w").write("8\n0\n0\n0 1 2 3 4 5 6")  # not a permutation of 0..7
    with pytest.raises(scpd.ScpdError):
        scpd.load_order(p, 8, 4)
    open(p, "w").write("8\n0\n0\n0 1 2 3 4 5 6 x")
    with pytest.raises(scpd.ScpdError):
        scpd.load_order(p, 8, 4)
    with pytest.raises(scpd.ScpdError) as e:
        scpd.load_order(p, 8, 9)
    assert e.value.status == scpd.E_CONFIG


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
def test_all_reference_tables_load():
    """Every file of Frozen_Bit_Tab/ and Generated_Frozen_Bit/ against the committed index."""
    for rel, meta in INDEX.items():
        n, k = meta["n"], meta["k"]
        if rel.startswith("Frozen_Bit_Tab"):
            flags = scpd.load_order(os.path.join(REF, rel), n, k)
        else:
            flags, kk = scpd.load_flags(os.path.join(REF, rel), n)
            assert kk == k
        assert flags.sum() == k and _packed_sha(flags) == meta["sha256_packed"], rel


def test_packaged_sets_match_index():
    for name, rel in (("FB_N1024_K512", "Frozen_Bit_Tab/FB_N1024_K512.txt"),
                      ("frozen_n_4096_k_3072", "Generated_Frozen_Bit/frozen_n_4096_k_3072.txt"),
                      ("frozen_n_524288_k_262144", "Generated_Frozen_Bit/frozen_n_524288_k_262144.txt")):
        meta = INDEX[rel]
        assert _packed_sha(scpd.packed_flags(name, meta["n"])) == meta["sha256_packed"]


FBG = os.path.join(ol.ORACLE_DIR, "_ref", "FB_Generator")


@pytest.mark.skipif(not (os.path.exists(FBG) and os.path.isdir(REF)), reason="needs oracle/_ref/FB_Generator")
@pytest.mark.parametrize("n,k,par,en,src,isflag", [
    (1024, 512, 16, 0, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0),
    (1024, 512, 64, 1, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0),
    (512, 256, 4, 0, "Frozen_Bit_Tab/FB_N512_K256.txt", 0),
    (4096, 3072, 16, 1, "Generated_Frozen_Bit/frozen_n_4096_k_3072.txt", 1),
    (256, 100, 8, 0, "Frozen_Bit_Tab/FB_N1024_K512.txt", 0),
])
def test_polar_parameters_header_identical_to_fb_generator(tmp_path, n, k, par, en, src, isflag):
    """scpd_write_polar_parameters / scpd_frozen_write_order byte-for-byte against the reference's
    own Frozen_Bit_Generator run on the same table."""
    work = tmp_path / "x" / "y"
    work.mkdir(parents=True)
    (tmp_path / "Frozen_Bit_Tab").mkdir()
    out = tmp_path / "out"
    out.mkdir()
    subprocess.check_call([FBG, str(n), str(k), str(par), str(en), os.path.join(REF, src), str(isflag), str(out) + "/"],
                          cwd=work, stdout=subprocess.DEVNULL)
    if isflag:
        flags, _ = scpd.load_flags(os.path.join(REF, src), n)
    else:
        flags = scpd.load_order(os.path.join(REF, src), n, k)
    mine = tmp_path / "mine.h"
    scpd.write_polar_parameters(mine, par, flags, en)
    assert open(mine, "rb").read() == open(out / "polar_parameters.h", "rb").read()
    if not isflag:
        tok = open(os.path.join(REF, src)).read().split()
        order = np.array([int(v) for v in tok[3:] if int(v) < n], np.uint32)
        aff = tmp_path / "aff.txt"
        scpd.write_order(aff, order)
        assert open(aff, "rb").read() == open(tmp_path / "Frozen_Bit_Tab" / f"FB_N{n}_K{k}.txt", "rb").read()
