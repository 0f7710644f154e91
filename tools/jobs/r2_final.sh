#!/bin/bash
# final state of the round on one GPU: smoke, the whole GPU test suite, the default bench, the reference arm (short)
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench_default_N1.json 2> gpurun_out/r2_bench_default_N1.err; echo "bench rc=$?"
grep -v "^\[W\|^$" gpurun_out/r2_bench_default_N1.err | tail -4
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_default_N1.json'))
print({k:d[k] for k in ('value','ms_per_step','device_ms_per_step','gpu_launches','clocks')}); print(d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['traffic']); print(d['parity'])
e=d['e2e']; print({k:e[k] for k in e if k!='path'})
x=d['extra']; print({k:(x[k] if not isinstance(x[k],dict) else {kk:vv for kk,vv in x[k].items() if kk in ('value','ms','body_ms','metric')}) for k in x if k not in ('python_api_on_disk','cfg4_validate_chain')})
a=x['python_api_on_disk']; print({k:a[k] for k in a if k!='query'})
c=x['cfg4_validate_chain']; print({k:c[k] for k in c if k not in ('cpu_baseline',)})
"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-600 gpurun_out/r2_bench_reference_arm.json
