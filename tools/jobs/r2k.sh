#!/bin/bash
python -m pytest tests/test_memdir_api_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in "FEI_HOST_THP=1" "FEI_HOST_THP=0" "FEI_HOST_THP=1"; do
  env $v python bench.py --entries 500000 --steps 1 --warmup 1 --e2e-entries 100000 --e2e-batches 1 --chain-blocks 10000 --cpu-sample 2000 > gpurun_out/r2k.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2k.json')); a=d['extra']['python_api_on_disk']; print('$v', round(a['search_memories_cold_s'],3), {k:(round(v,3) if isinstance(v,float) else v) for k,v in a['cold_stages'].items() if k!='cold_path'})"
done
