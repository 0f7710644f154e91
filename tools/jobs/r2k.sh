#!/bin/bash
# API leg of the bench (1 M files on disk) with the stage breakdown of the cold pack
python bench.py --entries 500000 --steps 1 --warmup 1 --e2e-entries 100000 --e2e-batches 1 --chain-blocks 10000 --cpu-sample 2000 > gpurun_out/r2k.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2k.json')); a=d['extra']['python_api_on_disk']; print(round(a['search_memories_cold_s'],3), {k:(round(v,3) if isinstance(v,float) else v) for k,v in a['cold_stages'].items() if k!='cold_path'}, 'restore first query', round(a['snapshot']['first_query_after_restore_s'],2))"
