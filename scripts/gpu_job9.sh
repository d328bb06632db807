set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
S="python bench.py --workload c3 --steps 2 --band-rows 1024 --no-e2e --no-cpu-baseline"
$S > gpurun_out/v_c3_def.json 2>&1
V2R_IDW_TUNE=b2 $S > gpurun_out/v_c3_b2.json 2>&1
python bench.py --workload c5 --steps 1 --band-rows 256 --no-e2e --no-cpu-baseline > gpurun_out/v_c5_band.json 2>&1
python bench.py --steps 2 --band-rows 1024 --exps 1.5 --no-e2e --no-cpu-baseline > gpurun_out/v_p15.json 2>&1
python bench.py --steps 2 --band-rows 1024 --exps 0.5,1,1.5,2 --no-e2e --no-cpu-baseline > gpurun_out/v_e4rt.json 2>&1
python scripts/summ.py gpurun_out/v_*.json
