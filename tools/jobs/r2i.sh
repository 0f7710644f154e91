python -m pytest tests/test_memdir_api_gpu.py tests/test_scan_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --chain-blocks 100000 --e2e-batches 2 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo rc=$?
grep -v "^\[W\|^$" gpurun_out/r2i_bench.err | tail -5
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench.json')); a=d['extra']['python_api_on_disk']; print({k:a[k] for k in a if k not in('query',)})"
FEI_COLD_STAT=1 python bench.py --steps 3 --chain-blocks 100000 --e2e-batches 2 > gpurun_out/r2i_bench_stat.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_stat.json')); a=d['extra']['python_api_on_disk']; print(a['search_memories_cold_s'], a['cold_stages'])"
