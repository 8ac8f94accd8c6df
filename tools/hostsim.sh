#!/bin/bash
# Runs GPU-marked tests (or any command) against the host-sim build of the library on a GPU-less box -- TEST USE ONLY,
# see tests/host/make_hostsim.py.  Everything is computed on the CPU; nothing printed is a GPU result or a timing.
#   tools/hostsim.sh                                   # the single-GPU test files (1000-step cases included: minutes)
#   tools/hostsim.sh python tests/host/hostsim_mrank_worker.py 4
#   tools/hostsim.sh python tests/host/hostsim_random_calls.py 100
set -e
here=$(cd "$(dirname "$0")/.." && pwd)
cd "$here"
python tests/host/make_hostsim.py > /dev/null
export D3Q19_LIB=$here/tests/host/_gen/libd3q19b200_hostsim.so D3Q19_DRIVER=$here/tests/host/_gen/channel_driver_hostsim
export D3Q19_TEST_FULLSIZE=${D3Q19_TEST_FULLSIZE:-64x16x16}
if [ $# -eq 0 ]; then
    exec python -m pytest tests -m gpu -q -p no:cacheprovider
fi
exec "$@"
