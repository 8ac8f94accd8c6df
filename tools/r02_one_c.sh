#!/bin/bash
# Round 2, third 1-GPU call: the whole GPU suite at HEAD, the headline lines, the plain kernel with / without its flag wait,
# the particle step after the sweeps' rewrite, full ncu captures for profiles/traffic.json.
set -u
mkdir -p gpurun_out
tag=${TAG:-r02f}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python bench.py --steps 1000 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_driver_args.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --scheme aa --no-cpu --no-parity --steps 1000 > gpurun_out/${tag}_bench_aa.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --particles 100 --no-cpu --no-parity --steps 200 > gpurun_out/${tag}_bench_part.json 2>> gpurun_out/${tag}_bench.err
for i in 1 2; do
    timeout 200 python tools/kernel_sweep.py 512x256x256 20 0 | sed 's/^{/{"variant": "shipped", /' >> gpurun_out/${tag}_variants.jsonl 2>> gpurun_out/${tag}_bench.err
    D3Q19_LIB=$PWD/build/variants/libd3q19b200_nowait.so timeout 200 python tools/kernel_sweep.py 512x256x256 20 0 >> gpurun_out/${tag}_variants.jsonl 2>> gpurun_out/${tag}_bench.err
done
cat gpurun_out/${tag}_variants.jsonl
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_particles.csv \
    python bench.py --particles 100 --no-cpu --no-e2e --no-parity --steps 3 --warmup 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --no-cpu --no-e2e --no-parity --steps 5 --warmup 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 \
    -o gpurun_out/prof_${tag}_ab python tools/prof_step.py --scheme ab --steps 8 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 2 \
    -o gpurun_out/prof_${tag}_aa python tools/prof_step.py --scheme aa --steps 8 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_beads_links -s 2 -c 1 \
    -o gpurun_out/prof_${tag}_links python bench.py --particles 100 --no-cpu --no-e2e --no-parity --steps 3 --warmup 3 > /dev/null 2>&1
python -c "
import importlib.util
s = importlib.util.spec_from_file_location('b', 'bench.py'); b = importlib.util.module_from_spec(s); s.loader.exec_module(b)
print(b.kernel_source_hash())" > gpurun_out/${tag}_kernel_source.sha256
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
grep -h '"value"' gpurun_out/${tag}_bench*.json | python -c "
import json, sys
for l in sys.stdin:
    d = json.loads(l); print(d.get('impl'), (d.get('implementation') or {}).get('scheme'), d['config']['per_gpu'], d['config']['particles'][:12], d['steps'], d['warmup'], round(d['value']), 'MLUPS', round(d['ms_per_step'], 4), 'ms', round((d.get('roofline') or {}).get('frac', 0), 4), 'e2e', d['e2e'] and round(d['e2e']['value']), (d.get('clocks') or {}))"
