#!/bin/bash
# Round 2, second 2-GPU call: the copy-engine face transport (two schedules) and NCCL's CTA budget on the 8-GPU strong-scaling
# slab thickness, two independent single-GPU jobs side by side (is the loss communication at all?), multi-GPU suite.
#   gpurun --gpus 2 --timeout 1300 -- 'TAG=r02e bash tools/r02_two_b.sh'
set -u
mkdir -p gpurun_out
tag=${TAG:-r02e}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--no-e2e --no-parity --no-cpu --steps 400"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_zz_cpp_driver.py -m gpu -q > gpurun_out/${tag}_pytest_gpu_2gpu.log 2>&1
tail -n 3 gpurun_out/${tag}_pytest_gpu_2gpu.log
# two INDEPENDENT single-GPU runs at the same time, then one alone
CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --workload 512x256x32 $B > gpurun_out/${tag}_pair_gpu0.json 2>> gpurun_out/${tag}_bench.err &
CUDA_VISIBLE_DEVICES=1 timeout 200 python bench.py --workload 512x256x32 $B > gpurun_out/${tag}_pair_gpu1.json 2>> gpurun_out/${tag}_bench.err
wait
timeout 200 python bench.py --workload 512x256x32 $B > gpurun_out/${tag}_n1_thin.json 2>> gpurun_out/${tag}_bench.err
timeout 200 python bench.py --workload 512x256x128 $B > gpurun_out/${tag}_n1_128.json 2>> gpurun_out/${tag}_bench.err
run2() { name=$1; shift; timeout 300 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; }
for c in 1 2 4 8 32; do run2 n2_thin_nccl_ctas$c --scaling strong --workload 512x256x64 $B --halo nccl --nccl-max-ctas $c; done
run2 n2_thin_put --scaling strong --workload 512x256x64 $B --halo put
run2 n2_thin_put_split --scaling strong --workload 512x256x64 $B --halo put --halo-split-min 3
run2 n2_thin_peer --scaling strong --workload 512x256x64 $B --halo peer
run2 n2_thin_put_aa --scaling strong --workload 512x256x64 $B --halo put --scheme aa
run2 n2_thin_nccl_aa --scaling strong --workload 512x256x64 $B --halo nccl --scheme aa
for v in "nccl" "put" "put --halo-split-min 3"; do
    timeout 200 $TR --nproc-per-node 2 --master-port 29614 tools/step_timeline.py --halo $v --scaling strong --size 512x256x64 --steps 100 >> gpurun_out/${tag}_timeline_thin.jsonl 2>> gpurun_out/${tag}_timeline.err
done
run2 n2_strong_nccl --scaling strong $B --halo nccl
run2 n2_strong_put --scaling strong $B --halo put
run2 n2_weak_nccl --steps 400
run2 n2_weak_put --steps 400 --halo put
run2 n2_c4_nccl --workload c4 --no-e2e --no-parity --steps 20 --warmup 3
run2 n2_c4_put --workload c4 --no-e2e --no-parity --steps 20 --warmup 3 --halo put
run2 n2_particles --particles 200 --no-e2e --steps 100
# one GPU: the particle step after this round's trims, its launch list and a full capture of the bounce-back kernel
timeout 300 python bench.py --particles 100 --no-cpu --no-parity --steps 200 > gpurun_out/${tag}_n1_particles.json 2>> gpurun_out/${tag}_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_particles.csv \
    python bench.py --particles 100 --no-cpu --no-e2e --no-parity --steps 3 --warmup 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_beads_ibb -s 2 -c 1 \
    -o gpurun_out/prof_${tag}_ibb python bench.py --particles 100 --no-cpu --no-e2e --no-parity --steps 3 --warmup 3 > /dev/null 2>&1
python - <<'PY'
import glob, json, os
def load(f):
    for l in open(f):
        if l.startswith('{'):
            return json.loads(l)
for f in sorted(glob.glob('gpurun_out/%s_*.json' % os.environ.get('TAG', 'r02e'))):
    d = load(f)
    if not d or 'value' not in d:
        print(os.path.basename(f), 'NO LINE'); continue
    print(os.path.basename(f)[5:-5].ljust(28), d['config']['per_gpu'].ljust(24), d['implementation']['scheme'], round(d['value']), 'MLUPS', round(d['ms_per_step'], 4), 'ms',
          'parity', (d.get('parity_check') or {}).get('bit_exact'), 'e2e', d['e2e'] and round(d['e2e']['value']), (d['clocks'] or {}).get('sm_mhz'), (d['clocks'] or {}).get('samples'))
PY
cat gpurun_out/${tag}_timeline_thin.jsonl
