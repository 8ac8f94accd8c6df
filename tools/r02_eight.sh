#!/bin/bash
# Round 2, the 8-GPU call (charged 8x: every line here is something the driver's own 1/2/4/8 weak run does not produce):
# parity on 8 ranks, configs[2] strong at 8 (and 4), configs[3] weak at 8 (and 4), configs[4] with particles at 8.
#   gpurun --gpus 8 --timeout 900 -- 'TAG=r02g bash tools/r02_eight.sh'
set -u
mkdir -p gpurun_out
tag=${TAG:-r02g}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { n=$1; name=$2; shift 2; timeout 400 $TR --nproc-per-node $n --master-port 29621 bench.py --gpus $n "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; }
# 1. weak default with the parity check (golden vector of the reference on 8 slabs, both schemes; 3 moving spheres, 8 slabs vs 1 domain) and e2e
run 8 n8_weak --steps 200
# 2. configs[2]: 512x256x256 over 8 GPUs (32 planes each)
run 8 n8_strong --scaling strong --no-e2e --no-parity --steps 600
run 8 n8_strong_nccl --scaling strong --no-e2e --no-parity --steps 600 --halo nccl
# 3. configs[4]: particle-laden channel, 100 spheres of radius 15 per 512x256x256 slab, avedensity every 100 steps
run 8 n8_particles --particles 800 --no-e2e --no-parity --steps 200
# 4. multi-rank parity worker on 8 ranks: every copy-engine case of tests/mgpu_worker.py + the halo watchdog
MGPU_ONLY=put timeout 300 $TR --nproc-per-node 8 --master-port 29622 tests/mgpu_worker.py > gpurun_out/${tag}_mgpu_put_n8.log 2>&1
tail -n 2 gpurun_out/${tag}_mgpu_put_n8.log
# 5. configs[3]: 1024x1024x944 per GPU (150 GB of populations each), in place
run 8 n8_c4 --workload c4 --no-e2e --no-parity --steps 20 --warmup 3
# 6. two 4-GPU jobs side by side on disjoint GPUs: configs[3] at 4, configs[2] at 4
CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 400 $TR --nproc-per-node 4 --master-port 29631 bench.py --gpus 4 --workload c4 --no-e2e --no-parity --steps 20 --warmup 3 \
    > gpurun_out/${tag}_n4_c4.json 2> gpurun_out/${tag}_n4_c4.err &
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 $TR --nproc-per-node 4 --master-port 29632 bench.py --gpus 4 --scaling strong --no-e2e --no-parity --steps 600 \
    > gpurun_out/${tag}_n4_strong.json 2> gpurun_out/${tag}_n4_strong.err
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 $TR --nproc-per-node 4 --master-port 29632 bench.py --gpus 4 --particles 400 --no-e2e --no-parity --steps 200 \
    > gpurun_out/${tag}_n4_particles.json 2> gpurun_out/${tag}_n4_particles.err
wait
python - <<'PY'
import glob, json, os
def load(f):
    for l in open(f):
        if l.startswith('{'):
            return json.loads(l)
for f in sorted(glob.glob('gpurun_out/%s_n*.json' % os.environ.get('TAG', 'r02g'))):
    d = load(f)
    if not d or 'value' not in d:
        print(os.path.basename(f), 'NO LINE', open(f.replace('.json', '.err')).read()[-600:]); continue
    print(os.path.basename(f)[5:-5].ljust(18), d['config']['per_gpu'].ljust(24), d['implementation']['scheme'], d['scaling'], round(d['value']), 'MLUPS', round(d['ms_per_step'], 4), 'ms',
          'parity', (d.get('parity_check') or {}).get('bit_exact'), 'e2e', d['e2e'] and round(d['e2e']['value']), d['implementation']['parallelism'][:40], (d['clocks'] or {}).get('sm_mhz'))
PY
