#!/bin/bash
# ncu captures of round 2 (one GPU).  Launch list of a reduced default bench, then --set full of the headline kernel, the raw
# ingest kernels and the chain kernels of the Python API path.  Numbers printed under ncu are never bench values.
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --entries 2000000 --steps 2 --warmup 1 --e2e-entries 200000 --e2e-batches 2 --api-files 2000 --chain-blocks 100000 > gpurun_out/r2_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_body -s 2 -c 1 -f -o gpurun_out/r2_k_body_batch32 python tools/prof_scan.py batch32 1000000 > gpurun_out/r2_ncu_body.log 2>&1
echo "k_body rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_raw_measure|k_raw_write|k_tile_copy' -s 3 -c 3 -f -o gpurun_out/r2_ingest python tools/prof_load.py 1000000 2 > gpurun_out/r2_ncu_ingest.log 2>&1
echo "ingest rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_json|k_sha256|k_prepare' -c 6 -f -o gpurun_out/r2_chain_api python tools/prof_validate.py 1000000 > gpurun_out/r2_ncu_chain.log 2>&1
echo "chain rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
