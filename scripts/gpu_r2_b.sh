#!/bin/bash
# round 2, job B: driver with its state in shared memory; new session pipeline (e2e)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
B="python bench.py --no-cpu-baseline --steps 3 --warmup 3"
$B > gpurun_out/r2b_c2.json 2> gpurun_out/r2b.err
$B --no-e2e --band-rows 128 > gpurun_out/r2b_c2_b128.json 2>> gpurun_out/r2b.err
$B --no-e2e --exps 1.7 --band-rows 512 > gpurun_out/r2b_p17.json 2>> gpurun_out/r2b.err
$B --no-e2e --exps 1.0 --band-rows 512 > gpurun_out/r2b_p10.json 2>> gpurun_out/r2b.err
$B --no-e2e --exps 1.5 --band-rows 512 > gpurun_out/r2b_p15.json 2>> gpurun_out/r2b.err
$B --no-e2e --exps 1.1,1.3,1.5,1.7 --band-rows 512 > gpurun_out/r2b_gen4.json 2>> gpurun_out/r2b.err
$B --precision fp32 > gpurun_out/r2b_fp32.json 2>> gpurun_out/r2b.err
$B --no-e2e --workload c3 --band-rows 512 > gpurun_out/r2b_c3.json 2>> gpurun_out/r2b.err
$B --no-e2e --workload c5 --band-rows 256 > gpurun_out/r2b_c5.json 2>> gpurun_out/r2b.err
tail -5 gpurun_out/r2b.err
for f in gpurun_out/r2b_*.json; do python scripts/summ.py $f | cut -c12-330; done
