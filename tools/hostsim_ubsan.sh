#!/bin/bash
# The C-ABI library itself (csrc/d3q19_api.cu + the kernels, host-sim build of tests/host/make_hostsim.py) under UBSan: the
# single-rank GPU test files, the 3-rank worker and the reference-driver worker; any "runtime error" line fails the script.
# TEST USE ONLY; ~6 min.  (ASan is not used here: the fibers that stand in for cooperating threads switch stacks.)
set -e
here=$(cd "$(dirname "$0")/.." && pwd)
cd "$here"
python tests/host/make_hostsim.py > /dev/null
out=gpurun_out/ubsan; mkdir -p $out
g++ -std=c++17 -O1 -g -fsanitize=undefined -ffp-contract=off -fPIC -shared -Wno-unknown-pragmas -DHS_FULL_RUNTIME \
    -I tests/host/fake -I d3q19-single-phase_b200/csrc -I include -o $out/libd3q19b200_hostsim.so tests/host/_gen/d3q19_api_hostsim.cpp -ldl -pthread
export LD_PRELOAD="$(gcc -print-file-name=libubsan.so)" UBSAN_OPTIONS=print_stacktrace=1 D3Q19_LIB=$here/$out/libd3q19b200_hostsim.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_particles.py tests/test_gpu_restart.py tests/test_golden.py -m gpu -q -x \
    -p no:cacheprovider -k "not 1000_steps and not faxen" > $out/single.log 2>&1 || { tail -20 $out/single.log; exit 1; }
HOSTSIM_SHORT=1 python tests/host/hostsim_mrank_worker.py 3 > $out/mrank3.log 2>&1 || { tail -20 $out/mrank3.log; exit 1; }
python tests/refdriver_worker.py --lib $D3Q19_LIB --ranks 2 --restart 5 > $out/refdrv.log 2>&1 || { tail -20 $out/refdrv.log; exit 1; }
n=$(cat $out/single.log $out/mrank3.log $out/refdrv.log | grep -c "runtime error" || true)
tail -1 $out/single.log; tail -1 $out/mrank3.log
echo "UBSan runtime errors: $n"
[ "$n" = "0" ]
