#!/bin/bash
# compute-sanitizer over the kernels that are new in round 2 (raw ingest, GPU canonical JSON, window counters + gate kernel,
# slot extraction, snapshot, incremental delta): memcheck on the small-size GPU tests, racecheck on the ingest kernels.
export FEI_SANITIZER=1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_scan_gpu.py -m gpu -x -q -k "raw_ingest or empty_and_tiny or hits_capacity or global_base or random_headers or adversarial" 2>&1 | tail -4
echo "memcheck scan rc=${PIPESTATUS[0]}"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_chain_gpu.py -m gpu -x -q -k "json or resident or single_block or reference_verdicts or chain_memory or mine" 2>&1 | tail -4
echo "memcheck chain rc=${PIPESTATUS[0]}"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_memdir_api_gpu.py -m gpu -x -q -k "matches_oracle or raising or sigma or in_place or incremental_sync or snapshot or cold_pack" 2>&1 | tail -4
echo "memcheck api rc=${PIPESTATUS[0]}"
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_scan_gpu.py -m gpu -x -q -k "raw_ingest" 2>&1 | tail -4
echo "racecheck ingest rc=${PIPESTATUS[0]}"
