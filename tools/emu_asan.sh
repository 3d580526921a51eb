#!/bin/bash
# The CPU emulations of the kernels (tests/emu) built with AddressSanitizer: every buffer of the emulated launch (shared
# memory region, workspace, plane buffer, hot-level array, tensor memory columns) is an exact-size allocation, so an
# out-of-range index of the kernel source shows up here.  (compute-sanitizer is not available on the GPU pool.)
set -e
cd "$(dirname "$0")/.."
ASAN=$(g++ -print-file-name=libasan.so)
for lib in ss_emu warp_emu; do
  cp tests/emu/lib$lib.so /tmp/lib$lib.backup.so
  g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++17 -fPIC -shared -I tests/emu/fake_cuda -o tests/emu/lib$lib.so tests/emu/$lib.cpp 2>/dev/null
done
trap 'for lib in ss_emu warp_emu; do cp /tmp/lib$lib.backup.so tests/emu/lib$lib.so; done' EXIT
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 python -m pytest tests/test_ss_kernel.py tests/test_host_logic.py tests/test_bs_kernel.py -x -q
