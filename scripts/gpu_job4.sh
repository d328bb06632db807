set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/u_p2_paired.json 2>&1
V2R_IDW_TUNE=u2 $B > gpurun_out/u_p2_paired_u2.json 2>&1
python bench.py --steps 2 --band-rows 1024 --exps 1.7 --no-e2e --no-cpu-baseline > gpurun_out/u_p17.json 2>&1
python bench.py --steps 2 --band-rows 1024 --exps 1.5 --no-e2e --no-cpu-baseline > gpurun_out/u_p15.json 2>&1
python bench.py --steps 2 --band-rows 1024 --exps 1.0 --no-e2e --no-cpu-baseline > gpurun_out/u_p10.json 2>&1
python bench.py --steps 2 --band-rows 1024 --exps 3.0 --no-e2e --no-cpu-baseline > gpurun_out/u_p30.json 2>&1
python bench.py --steps 2 --band-rows 1024 --exps 1.1,1.3,1.5,1.7 --no-e2e --no-cpu-baseline > gpurun_out/u_gen4.json 2>&1
python bench.py --workload c5 --steps 1 --band-rows 512 --no-e2e --no-cpu-baseline > gpurun_out/u_c5_band.json 2>&1
python scripts/summ.py gpurun_out/u_*.json
