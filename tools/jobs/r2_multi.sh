#!/bin/bash
# bench under torchrun on N GPUs of one box: bash tools/jobs/r2_multi.sh N
N=$1
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 > gpurun_out/r2_bench_N$N.json 2> gpurun_out/r2_bench_N$N.err; echo rc=$?
grep -v "^\[W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r2_bench_N$N.err | tail -6
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_N$N.json')); print({k:d[k] for k in ('value','ms_per_step','device_ms_per_step','gpu_launches')}); print(json.dumps(d['parity'])); print(json.dumps(d['multi_gpu'])[:1500]); e=d['e2e']; print({k:e[k] for k in e if k!='path'}); print(d['config']['entries_per_gpu'], d['config']['entries_total'])"
