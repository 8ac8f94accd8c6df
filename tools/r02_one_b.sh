#!/bin/bash
# Round 2, second 1-GPU call: direction-aware prefetch on 512- and 1024-wide rows (configs[3]'s planes), prefetch distance,
# configs[3] itself on one GPU, the re-segmented particle path, the new tests (Faxen anchor, strain rate, restart).
set -u
mkdir -p gpurun_out
tag=${TAG:-r02d}
timeout 900 python -m pytest tests/test_gpu_particles.py tests/test_golden.py tests/test_gpu_restart.py -m gpu -x -q -s > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/${tag}_pytest_gpu.log; grep -h "drag " gpurun_out/${tag}_pytest_gpu.log
for size in 512x256x256 1024x1024x32; do
    for pf in 0 -1 32 64 256 512; do
        timeout 200 python tools/kernel_sweep.py $size 20 $pf | sed 's/^{/{"variant": "shipped", /' >> gpurun_out/${tag}_variants.jsonl 2>> gpurun_out/${tag}_bench.err
    done
    for v in pfold nows; do
        D3Q19_LIB=$PWD/build/variants/libd3q19b200_$v.so timeout 200 python tools/kernel_sweep.py $size 20 0 >> gpurun_out/${tag}_variants.jsonl 2>> gpurun_out/${tag}_bench.err
    done
done
cat gpurun_out/${tag}_variants.jsonl
timeout 300 python bench.py --scheme aa --no-cpu --no-e2e --no-parity --steps 600 > gpurun_out/${tag}_bench_aa.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --no-cpu --no-e2e --no-parity --steps 600 > gpurun_out/${tag}_bench_ab.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --scheme aa --no-cpu --no-e2e --no-parity --steps 600 > gpurun_out/${tag}_bench_aa2.json 2>> gpurun_out/${tag}_bench.err
timeout 500 python bench.py --workload c4 --no-cpu --no-e2e --no-parity --steps 20 --warmup 3 > gpurun_out/${tag}_bench_c4.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --particles 100 --no-cpu --no-parity --steps 200 > gpurun_out/${tag}_bench_part.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --particles 100 --scheme aa --no-cpu --no-parity --steps 200 > gpurun_out/${tag}_bench_part_aa.json 2>> gpurun_out/${tag}_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_particles.csv \
    python bench.py --particles 100 --no-cpu --no-e2e --no-parity --steps 3 --warmup 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 2 \
    -o gpurun_out/prof_${tag}_aa_1024 python tools/prof_step.py --scheme aa --steps 8 --size 1024x1024x32 > /dev/null 2>&1
grep -h '"value"' gpurun_out/${tag}_bench*.json | python -c "
import json, sys
for l in sys.stdin:
    d = json.loads(l); print(d['implementation']['scheme'], d['config']['per_gpu'], d['config']['particles'][:12], round(d['value']), 'MLUPS', round(d['ms_per_step'], 4), 'ms', round(d['roofline']['frac'], 4), (d.get('clocks') or {}))"
