#!/bin/bash
# The host-compiled kernels (tests/host/kernels_host.cpp) under AddressSanitizer + UBSan: every out-of-range index of a
# step / face / particle kernel on the small cases of tests/test_kernels_host.py would be a heap-buffer-overflow here
# (each device array is its own heap block).  ~8 min on 8 cores (fiber-run particle cases); not part of
# the default suite.  compute-sanitizer on the device: tools/r02_sanitize.sh.
set -e
here=$(cd "$(dirname "$0")/.." && pwd)
cd "$here"
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -fPIC -shared -Wno-unknown-pragmas \
    -I tests/host/fake -I d3q19-single-phase_b200/csrc -o tests/host/libkernels_host.so tests/host/kernels_host.cpp
touch tests/host/libkernels_host.so
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 python -m pytest tests/test_kernels_host.py -q "$@"
rc=$?
rm -f tests/host/libkernels_host.so          # the next normal test run rebuilds the plain library
exit $rc
