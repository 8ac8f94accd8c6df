#!/bin/bash
# Round 2, 1-GPU call: the whole GPU suite on the rewritten particle path / restart / full-size reference parity, bench lines,
# the kernel variants of this round (wall-split accessors, lean prefetch) burst and sustained, ncu launch list + full capture.
#   gpurun --timeout 1700 -- 'TAG=r02c bash tools/r02_one.sh'
set -u
mkdir -p gpurun_out
tag=${TAG:-r02c}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python bench.py --steps 1000 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --scheme aa --no-cpu --no-parity --steps 1000 > gpurun_out/${tag}_bench_aa.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --particles 100 --no-cpu --no-parity --steps 200 > gpurun_out/${tag}_bench_part.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --particles 100 --scheme aa --no-cpu --no-parity --steps 200 > gpurun_out/${tag}_bench_part_aa.json 2>> gpurun_out/${tag}_bench.err
# kernel variants: single launches (burst clocks) and 600 sustained steps under the power cap
for v in base ws pf ws3; do
    lib=$PWD/build/variants/libd3q19b200_$v.so
    [ -f "$lib" ] || continue
    D3Q19_LIB=$lib timeout 200 python tools/kernel_sweep.py 512x256x256 30 >> gpurun_out/${tag}_variants.jsonl 2>> gpurun_out/${tag}_bench.err
    for sc in aa ab; do
        D3Q19_LIB=$lib timeout 200 python bench.py --scheme $sc --no-cpu --no-e2e --no-parity --steps 600 2>> gpurun_out/${tag}_bench.err | sed "s/^{/{\"variant\": \"$v\", /" >> gpurun_out/${tag}_variants_sustained.jsonl
    done
done
timeout 200 python tools/kernel_sweep.py 512x256x256 30 | sed 's/^{/{"variant": "shipped", /' >> gpurun_out/${tag}_variants.jsonl
timeout 200 python tools/kernel_sweep.py 1024x1024x32 20 | sed 's/^{/{"variant": "shipped", /' >> gpurun_out/${tag}_variants.jsonl
D3Q19_LIB=$PWD/build/variants/libd3q19b200_base.so timeout 200 python tools/kernel_sweep.py 1024x1024x32 20 >> gpurun_out/${tag}_variants.jsonl 2>> gpurun_out/${tag}_bench.err
# launch lists (shares) and one full capture per scheme of the shipped kernels
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_particles.csv \
    python bench.py --particles 100 --no-cpu --no-e2e --no-parity --steps 3 --warmup 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --no-cpu --no-e2e --no-parity --steps 5 --warmup 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 \
    -o gpurun_out/prof_${tag}_ab python tools/prof_step.py --scheme ab --steps 8 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 2 \
    -o gpurun_out/prof_${tag}_aa python tools/prof_step.py --scheme aa --steps 8 > /dev/null 2>&1
sha256sum d3q19-single-phase_b200/libd3q19b200.so > gpurun_out/${tag}_lib.sha256
# sanitizer over the new particle kernels
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_particles.py -m gpu -q -x \
    > gpurun_out/${tag}_memcheck_particles.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${tag}_memcheck_particles.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_particles.py -m gpu -q -k "mask_and_links or moving" \
    > gpurun_out/${tag}_racecheck_particles.log 2>&1; echo "racecheck exit $?" >> gpurun_out/${tag}_racecheck_particles.log
tail -n 2 gpurun_out/${tag}_memcheck_particles.log gpurun_out/${tag}_racecheck_particles.log
cat gpurun_out/${tag}_variants.jsonl
grep -h '"value"' gpurun_out/${tag}_bench*.json gpurun_out/${tag}_variants_sustained.jsonl | python -c "
import json, sys
for l in sys.stdin:
    d = json.loads(l); print(d.get('variant', '-'), d['config']['scheme'], d['config']['particles'][:12], round(d['value']), 'MLUPS', round(d['ms_per_step'], 4), 'ms', round(d['roofline']['frac'], 4), (d.get('clocks') or {}).get('sm_mhz'))"
