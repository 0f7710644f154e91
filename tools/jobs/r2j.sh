#!/bin/bash
# cold-pack variants on the 1 M-file tree (API leg of the bench only matters here)
for v in "FEI_PIN_COLD=0" "FEI_COLD_STAT=1" "FEI_PIN_COLD=1" "FEI_PIN_COLD=0"; do
  env $v python bench.py --entries 500000 --steps 1 --warmup 1 --e2e-entries 100000 --e2e-batches 1 --chain-blocks 10000 --cpu-sample 2000 > gpurun_out/r2j.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2j.json')); a=d['extra']['python_api_on_disk']; print('$v', round(a['search_memories_cold_s'],3), {k:(round(v,3) if isinstance(v,float) else v) for k,v in a['cold_stages'].items() if k!='cold_path'}, 'restore first query', round(a['snapshot']['first_query_after_restore_s'],2))"
done
