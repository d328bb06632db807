#!/bin/bash
# round 2, job C: lean whole-tile driver instantiation; new bench.py (c4 slab default + parity block)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
B="python bench.py --no-cpu-baseline --steps 3 --warmup 3 --workload c2 --no-parity"
$B > gpurun_out/r2c_c2.json 2> gpurun_out/r2c.err
$B --no-e2e --band-rows 128 > gpurun_out/r2c_c2_b128.json 2>> gpurun_out/r2c.err
V2R_IDW_FP64_PAIRED=0 $B --no-e2e > gpurun_out/r2c_c2_unpaired.json 2>> gpurun_out/r2c.err
$B --no-e2e --exps 1.7 --band-rows 512 > gpurun_out/r2c_p17.json 2>> gpurun_out/r2c.err
V2R_IDW_TUNE=g4x4 $B --no-e2e --exps 1.7 --band-rows 512 > gpurun_out/r2c_p17_g4x4.json 2>> gpurun_out/r2c.err
$B --no-e2e --exps 1.0 --band-rows 512 > gpurun_out/r2c_p10.json 2>> gpurun_out/r2c.err
$B --no-e2e --exps 1.5 --band-rows 512 > gpurun_out/r2c_p15.json 2>> gpurun_out/r2c.err
$B --no-e2e --exps 1.1,1.3,1.5,1.7 --band-rows 512 > gpurun_out/r2c_gen4.json 2>> gpurun_out/r2c.err
$B --precision fp32 > gpurun_out/r2c_fp32.json 2>> gpurun_out/r2c.err
$B --no-e2e --workload c3 --band-rows 512 > gpurun_out/r2c_c3.json 2>> gpurun_out/r2c.err
$B --no-e2e --workload c5 --band-rows 256 > gpurun_out/r2c_c5.json 2>> gpurun_out/r2c.err
# the new default: c4 slab, with parity block and cpu baseline (short)
python bench.py --steps 2 --warmup 3 --cpu-seconds 5 > gpurun_out/r2c_default.json 2>> gpurun_out/r2c.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --slab 7168:256 > gpurun_out/r2c_c4_256rows.json 2>> gpurun_out/r2c.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --slab 7168:128 > gpurun_out/r2c_c4_128rows.json 2>> gpurun_out/r2c.err
tail -5 gpurun_out/r2c.err
for f in gpurun_out/r2c_*.json; do python scripts/summ.py $f | cut -c12-330; done
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2c_default.json') if l.startswith('{')][-1])
print(json.dumps({k:d[k] for k in ('e2e','parity','cpu_baseline','wall_s','config')}, indent=1))
P
