#!/bin/bash
# Round 2, 2-GPU call: multi-GPU suite, bench parity_check on real GPUs, per-step timelines and the opt-in schedules on the
# 8-GPU strong-scaling slab thickness (32 planes per GPU), C4 weak scaling at N=2.
#   gpurun --gpus 2 --timeout 1100 -- 'bash tools/r02_two.sh'
set -u
mkdir -p gpurun_out
tag=${TAG:-r02b}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_zzzz_gpu_experimental.py tests/test_zz_cpp_driver.py -m gpu -q > gpurun_out/${tag}_pytest_gpu_2gpu.log 2>&1
tail -n 3 gpurun_out/${tag}_pytest_gpu_2gpu.log
# default weak line with the parity check (fluid golden + particle slabs-vs-domain)
timeout 300 $TR --nproc-per-node 2 --master-port 29611 bench.py --gpus 2 --steps 200 > gpurun_out/${tag}_n2_weak.json 2> gpurun_out/${tag}_n2_weak.err
# one GPU alone on the thin block (the "ideal" of 8-GPU strong scaling) and on the 128-plane block
timeout 200 python bench.py --workload 512x256x32 --no-cpu --no-e2e --no-parity --steps 400 > gpurun_out/${tag}_n1_thin.json 2>> gpurun_out/${tag}_bench.err
timeout 200 python bench.py --workload 512x256x128 --no-cpu --no-e2e --no-parity --steps 400 > gpurun_out/${tag}_n1_128.json 2>> gpurun_out/${tag}_bench.err
# thin slabs (512x256x64 over 2 GPUs = 32 planes each): every transport / schedule, bench line + timeline
for v in "nccl" "nccl --no-overlap" "nccl --boundary-stream" "nccl --direct-faces" "nccl --boundary-stream --direct-faces" "peer" "put"; do
    name=${tag}_n2_thin_$(echo "$v" | tr -d ' ' | sed 's/--/_/g; s/-//g')
    timeout 200 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 --scaling strong --workload 512x256x64 --no-e2e --no-parity --steps 400 --halo $v > gpurun_out/$name.json 2> gpurun_out/$name.err
done
for v in "nccl" "peer" "put"; do
    timeout 200 $TR --nproc-per-node 2 --master-port 29614 tools/step_timeline.py --halo $v --scaling strong --size 512x256x64 --steps 100 >> gpurun_out/${tag}_timeline_thin.jsonl 2>> gpurun_out/${tag}_timeline.err
done
D3Q19_BOUNDARY_STREAM=1 timeout 200 $TR --nproc-per-node 2 --master-port 29614 tools/step_timeline.py --halo nccl --scaling strong --size 512x256x64 --steps 100 | sed 's/"halo": "nccl"/"halo": "nccl+bstream"/' >> gpurun_out/${tag}_timeline_thin.jsonl 2>> gpurun_out/${tag}_timeline.err
D3Q19_DIRECT_FACES=1 timeout 200 $TR --nproc-per-node 2 --master-port 29614 tools/step_timeline.py --halo nccl --scaling strong --size 512x256x64 --steps 100 | sed 's/"halo": "nccl"/"halo": "nccl+direct"/' >> gpurun_out/${tag}_timeline_thin.jsonl 2>> gpurun_out/${tag}_timeline.err
# strong scaling of configs[2] at N=2 (128 planes each) with the default and the candidates
for v in "nccl" "nccl --boundary-stream --direct-faces" "peer"; do
    name=${tag}_n2_strong_$(echo "$v" | tr -d ' ' | sed 's/--/_/g; s/-//g')
    timeout 200 $TR --nproc-per-node 2 --master-port 29615 bench.py --gpus 2 --scaling strong --no-e2e --no-parity --steps 400 --halo $v > gpurun_out/$name.json 2> gpurun_out/$name.err
done
# configs[3] at N=2: 1024x1024x944 per GPU, in place, device init
timeout 400 $TR --nproc-per-node 2 --master-port 29616 bench.py --gpus 2 --workload c4 --no-e2e --no-parity --steps 20 --warmup 3 > gpurun_out/${tag}_n2_c4.json 2> gpurun_out/${tag}_n2_c4.err
# particle-laden weak line at N=2 (100 spheres per slab)
timeout 300 $TR --nproc-per-node 2 --master-port 29617 bench.py --gpus 2 --particles 200 --no-e2e --no-parity --steps 100 > gpurun_out/${tag}_n2_particles.json 2> gpurun_out/${tag}_n2_particles.err
grep -h '"value"' gpurun_out/${tag}_n*.json | python -c "
import json, sys
for l in sys.stdin:
    d = json.loads(l); print(d['config']['parallelism'][:80], '|', d['config']['per_gpu'], d['scaling'], round(d['value']), 'MLUPS', round(d['ms_per_step'], 4), 'ms', (d.get('parity_check') or {}).get('bit_exact'))"
cat gpurun_out/${tag}_timeline_thin.jsonl
