#!/bin/bash
# development aid: build libscpd with extra nvcc flags into sc_polar_decoder_hls_b200/variants/libscpd_<name>.so
# usage: build_variant.sh <name> <flags...>      (select at run time with SCPD_LIB_PATH)
set -e
cd "$(dirname "$0")/../sc_polar_decoder_hls_b200/csrc"
NAME=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DSCPD_FAST_BUILD "$@" \
  -o ../variants/libscpd_$NAME.so scpd_api.cu frozen_io.cpp monitor.cpp
echo built ../variants/libscpd_$NAME.so
