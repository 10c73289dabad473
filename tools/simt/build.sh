#!/bin/bash
# Host build of the device sources against the SIMT interpreter (tools/simt/simt.h); the tests build it on demand too.
#   build.sh        plain build
#   build.sh asan   AddressSanitizer build: every load / store of the device code is checked against the exact bounds of the
#                   host buffers that stand in for global memory (same slack as the product allocates).  Run the tests with
#                     LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python -m pytest tests/test_kernel_sim.py tests/test_read_kernels_sim.py
set -e
cd "$(dirname "$0")"
EXTRA=""
if [ "$1" = asan ]; then EXTRA="-fsanitize=address -fno-omit-frame-pointer"; fi
g++ -std=c++17 -O1 -g -fPIC -shared -Wno-unknown-pragmas $EXTRA sim_compact.cpp -o libpgs_sim.so
