#!/bin/bash
# Round 2, last call (2 GPUs): HEAD after the particle-path changes that followed the 8-GPU run -- multi-GPU suite, the weak line
# with its parity check, configs[4] at 2 GPUs and 1 GPU with launch lists.
set -u
mkdir -p gpurun_out
tag=${TAG:-r02h}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_particles.py -m gpu -q > gpurun_out/${tag}_pytest_gpu_2gpu.log 2>&1
tail -n 3 gpurun_out/${tag}_pytest_gpu_2gpu.log
run2() { name=$1; shift; timeout 300 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; }
run2 n2_weak --steps 400
run2 n2_particles --particles 200 --no-e2e --steps 200
run2 n2_particles_nccl --particles 200 --no-e2e --no-parity --steps 200 --halo nccl
run2 n2_particles_800 --particles 800 --no-e2e --no-parity --steps 200 --rad 7.5
run2 n2_thin --scaling strong --workload 512x256x64 --no-e2e --no-parity --no-cpu --steps 600
timeout 300 python bench.py --particles 100 --no-cpu --no-parity --steps 200 > gpurun_out/${tag}_n1_particles.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --particles 400 --rad 7.5 --no-cpu --no-parity --steps 200 > gpurun_out/${tag}_n1_particles_400.json 2>> gpurun_out/${tag}_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_particles.csv \
    python bench.py --particles 100 --no-cpu --no-e2e --no-parity --steps 3 --warmup 3 > /dev/null 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_driver_args.json 2>> gpurun_out/${tag}_bench.err
python - <<'PY'
import glob, json, os
def load(f):
    for l in open(f):
        if l.startswith('{'):
            return json.loads(l)
for f in sorted(glob.glob('gpurun_out/%s_*.json' % os.environ.get('TAG', 'r02h'))):
    d = load(f)
    if not d or 'value' not in d:
        print(os.path.basename(f), 'NO LINE'); continue
    print(os.path.basename(f)[5:-5].ljust(22), d['config']['per_gpu'].ljust(22), d['implementation']['scheme'], round(d['value']), 'MLUPS', round(d['ms_per_step'], 4), 'ms',
          'parity', (d.get('parity_check') or {}).get('bit_exact'), 'e2e', d['e2e'] and round(d['e2e']['value']), d['implementation']['parallelism'][:30])
PY
