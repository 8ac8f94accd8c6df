#!/bin/bash
# What the first gpurun calls of the next round should run (DESIGN.md "Next").  Everything lands in gpurun_out/;
# copy what is to be judged into profiles/ afterwards.  A number printed under ncu is never a bench value.
#
#   gpurun --timeout 1500 -- 'bash tools/next_round_gpu.sh one'        # 1 GPU, ~12 min
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/next_round_gpu.sh two'   # 2 GPUs, ~8 min
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/next_round_gpu.sh eight'  # 8 GPUs, ~5 min
set -u
mkdir -p gpurun_out
tag=${TAG:-r02a}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"

case "${1:-one}" in
one)
    # 1. the whole GPU suite (particle kernels changed since the last GPU run)
    timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1
    tail -3 gpurun_out/${tag}_pytest_gpu.log
    # 2. headline line, particle line (round 1c: 2.44 ms per step), in-place line
    timeout 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
    timeout 300 python bench.py --particles 100 --no-cpu --steps 100 > gpurun_out/${tag}_bench_part.json 2>> gpurun_out/${tag}_bench.err
    timeout 300 python bench.py --scheme aa --no-cpu > gpurun_out/${tag}_bench_aa.json 2>> gpurun_out/${tag}_bench.err
    # the 128-bit / two-nodes-per-thread AB step (k_step_ab2, 168 registers, 3 CTAs per SM): bench line + full capture
    timeout 300 python bench.py --vec2 --no-cpu --no-e2e > gpurun_out/${tag}_bench_vec2.json 2>> gpurun_out/${tag}_bench.err
    timeout 300 python bench.py --vec2 --scheme aa --no-cpu --no-e2e > gpurun_out/${tag}_bench_vec2_aa.json 2>> gpurun_out/${tag}_bench.err
    D3Q19_VEC2=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step_ab2 -s 4 -c 2 \
        -o gpurun_out/prof_${tag}_vec2 python tools/prof_step.py --scheme ab --steps 8 > /dev/null 2>&1
    D3Q19_VEC2=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step_aa2 -s 4 -c 2 \
        -o gpurun_out/prof_${tag}_vec2_aa python tools/prof_step.py --scheme aa --steps 8 > /dev/null 2>&1
    # 3. launch list of the particle step (shares of the bookkeeping kernels) and of the plain step
    timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_particles.csv \
        python bench.py --particles 100 --no-cpu --no-e2e --steps 3 --warmup 1 > /dev/null 2>&1
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/${tag}_launches.csv \
        python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 > /dev/null 2>&1
    # 4. sanitizers on the small case (SURVEY.md section 5 "race detection"): memcheck over smoke(), racecheck over the
    #    particle tests (shared-memory scans and reductions)
    timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" \
        > gpurun_out/${tag}_memcheck_smoke.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${tag}_memcheck_smoke.log
    timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_particles.py -m gpu -q -k "mask_and_links" \
        > gpurun_out/${tag}_racecheck_particles.log 2>&1; echo "racecheck exit $?" >> gpurun_out/${tag}_racecheck_particles.log
    tail -2 gpurun_out/${tag}_memcheck_smoke.log gpurun_out/${tag}_racecheck_particles.log
    ;;
two)
    # the established multi-GPU suite first (its particle section now exchanges the refill source planes over NCCL)
    timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_zzzz_gpu_experimental.py -m gpu -q > gpurun_out/${tag}_pytest_gpu_2gpu.log 2>&1
    tail -3 gpurun_out/${tag}_pytest_gpu_2gpu.log
    # boundary stream: parity first, then its effect on thin and thick slabs, then the timelines
    MGPU_ONLY=bstream timeout 600 $TR --nproc-per-node 2 --master-port 29611 tests/mgpu_worker.py > gpurun_out/${tag}_mgpu_bstream_n2.log 2>&1
    tail -2 gpurun_out/${tag}_mgpu_bstream_n2.log
    for sc in strong weak; do
        for bs in "" "--boundary-stream" "--direct-faces" "--boundary-stream --direct-faces"; do
            name=${tag}_n2_${sc}$(echo "$bs" | tr -d ' ' | sed 's/--/_/g; s/-//g')
            timeout 300 $TR --nproc-per-node 2 --master-port 29612 bench.py --gpus 2 --scaling $sc --no-e2e $bs > gpurun_out/$name.json 2> gpurun_out/$name.err
        done
    done
    # 8-GPU strong scaling per-GPU block on 2 GPUs: 512x256x64 total = 32 planes each
    for bs in "" "--boundary-stream" "--direct-faces" "--boundary-stream --direct-faces"; do
        name=${tag}_n2_thin$(echo "$bs" | tr -d ' ' | sed 's/--/_/g; s/-//g')
        timeout 300 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 --scaling strong --workload 512x256x64 --no-e2e $bs > gpurun_out/$name.json 2> gpurun_out/$name.err
    done
    # the C++ driver with the ranks of the reference's job as threads (tests + one timed run)
    timeout 600 python -m pytest tests/test_zz_cpp_driver.py -m gpu -q > gpurun_out/${tag}_pytest_cpp_driver_2gpu.log 2>&1
    tail -2 gpurun_out/${tag}_pytest_cpp_driver_2gpu.log
    timeout 300 d3q19-single-phase_b200/host/channel_driver --ranks 2 --nx 512 --ny 256 --nz 512 --turbulent --nsteps 1000 \
        > gpurun_out/${tag}_cpp_driver_n2.log 2>&1
    grep "time loop" gpurun_out/${tag}_cpp_driver_n2.log
    grep -h '"value"' gpurun_out/${tag}_n2_*.json | python -c "
import json, sys
for l in sys.stdin:
    d = json.loads(l); print(d['config']['parallelism'][:70], d['config']['per_gpu'], d['scaling'], round(d['value']), 'MLUPS', round(d['ms_per_step'], 4), 'ms')"
    ;;
eight)
    for sc in strong weak; do
        for bs in "" "--boundary-stream --direct-faces"; do
            name=${tag}_n8_${sc}$(echo "$bs" | tr -d ' ' | sed 's/--/_/g; s/-//g')
            timeout 300 $TR --nproc-per-node 8 --master-port 29614 bench.py --gpus 8 --scaling $sc --no-e2e $bs > gpurun_out/$name.json 2> gpurun_out/$name.err
        done
    done
    MGPU_ONLY=bstream timeout 400 $TR --nproc-per-node 8 --master-port 29615 tests/mgpu_worker.py > gpurun_out/${tag}_mgpu_bstream_n8.log 2>&1
    tail -2 gpurun_out/${tag}_mgpu_bstream_n8.log
    # configs[4]: particle-laden channel on 8 GPUs (100 spheres per 512x256x256 slab, weak)
    timeout 300 $TR --nproc-per-node 8 --master-port 29616 bench.py --gpus 8 --particles 800 --no-e2e --steps 100 > gpurun_out/${tag}_n8_particles.json 2> gpurun_out/${tag}_n8_particles.err
    ;;
esac
