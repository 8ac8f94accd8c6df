#!/bin/bash
# compute-sanitizer over the particle tests of the FINAL particle kernels (half-warp link windows, early slab cull, split
# lubrication): memcheck over the whole file, racecheck over the cases with shared-memory / shuffle traffic.
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 75 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_particles.py -m gpu -q -x -p no:cacheprovider \
    > gpurun_out/${tag}_memcheck_particles.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${tag}_memcheck_particles.log
timeout 60 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_particles.py -m gpu -q -p no:cacheprovider \
    -k "mask_and_links or moving or near_contact or many" \
    > gpurun_out/${tag}_racecheck_particles.log 2>&1; echo "racecheck exit $?" >> gpurun_out/${tag}_racecheck_particles.log
tail -n 3 gpurun_out/${tag}_memcheck_particles.log gpurun_out/${tag}_racecheck_particles.log
