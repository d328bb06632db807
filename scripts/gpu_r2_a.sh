#!/bin/bash
# round 2, job A: parity tests on the restructured kernels + first numbers
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3"
$B > gpurun_out/r2a_c2.json 2> gpurun_out/r2a_c2.err
V2R_IDW_TAIL=0 $B > gpurun_out/r2a_c2_notail.json 2>> gpurun_out/r2a_c2.err
$B --band-rows 512 > gpurun_out/r2a_c2_b512.json 2>> gpurun_out/r2a_c2.err
V2R_IDW_TAIL=0 $B --band-rows 512 > gpurun_out/r2a_c2_b512_notail.json 2>> gpurun_out/r2a_c2.err
$B --band-rows 128 > gpurun_out/r2a_c2_b128.json 2>> gpurun_out/r2a_c2.err
V2R_IDW_TAIL=0 $B --band-rows 128 > gpurun_out/r2a_c2_b128_notail.json 2>> gpurun_out/r2a_c2.err
$B --exps 1.7 --band-rows 512 > gpurun_out/r2a_p17.json 2>> gpurun_out/r2a_c2.err
$B --exps 2.2 --band-rows 512 > gpurun_out/r2a_p22.json 2>> gpurun_out/r2a_c2.err
$B --exps 4.2 --band-rows 512 > gpurun_out/r2a_p42.json 2>> gpurun_out/r2a_c2.err
$B --exps 1.1,1.3,1.5,1.7 --band-rows 512 > gpurun_out/r2a_gen4.json 2>> gpurun_out/r2a_c2.err
$B --precision fp32 > gpurun_out/r2a_fp32.json 2>> gpurun_out/r2a_c2.err
for t in f4x4 f2x8m2 f4x8 f2x16; do V2R_IDW_TUNE=$t $B --precision fp32 > gpurun_out/r2a_fp32_$t.json 2>> gpurun_out/r2a_c2.err; done
for t in g2x2 g4x4 g4x2; do V2R_IDW_TUNE=$t $B --exps 1.7 --band-rows 512 > gpurun_out/r2a_p17_$t.json 2>> gpurun_out/r2a_c2.err; done
$B --exps 1.5 --band-rows 512 > gpurun_out/r2a_p15.json 2>> gpurun_out/r2a_c2.err
$B --exps 1.0 --band-rows 512 > gpurun_out/r2a_p10.json 2>> gpurun_out/r2a_c2.err
tail -5 gpurun_out/r2a_c2.err
for f in gpurun_out/r2a_*.json; do echo -n "$f: "; python scripts/summ.py $f; done
