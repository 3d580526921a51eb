"""Build libscpd.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libscpd.so")
SOURCES = ["scpd_api.cu", "frozen_io.cpp", "monitor.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newest_source_mtime():
    t = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for fn in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, fn)))
    return t


STAMP = os.path.join(HERE, "libscpd.flavour")


def _flavour():
    return "fast" if os.environ.get("SCPD_FAST_BUILD") else "full"


def _stamp():
    try:
        return open(STAMP).read().strip()
    except OSError:
        return ""


def build_lib(force=False, verbose=False):
    """Compile every CUDA source of the package into libscpd.so. Returns the library path.
    SCPD_FAST_BUILD=1 (development) leaves out most template instantiations; a library built that way
    is rebuilt in full the next time the variable is not set."""
    if (not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime()
            and _stamp() == _flavour()):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):
            return LIB  # GPU box without a toolkit: use the library that travelled with the snapshot
        raise RuntimeError("nvcc not found and no prebuilt libscpd.so")
    extra = ["-DSCPD_FAST_BUILD"] if os.environ.get("SCPD_FAST_BUILD") else []  # development: fewer instantiations
    extra += os.environ.get("SCPD_NVCC_EXTRA", "").split()  # development: e.g. -DSCPD_SS_THREADS=768
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout)
    with open(STAMP, "w") as f:
        f.write(_flavour() + "\n")
    if verbose:
        print(r.stdout)
    return LIB


def wait_for_lib(timeout=600.0):
    """Block until another process has finished building an up-to-date libscpd.so."""
    import time
    t0 = time.time()
    while time.time() - t0 < timeout:
        if os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime() and _stamp() == _flavour():
            time.sleep(0.5)
            return LIB
        time.sleep(0.5)
    raise RuntimeError("timed out waiting for libscpd.so")


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
