set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
G="python bench.py --steps 2 --band-rows 1024 --exps 1.7 --no-e2e --no-cpu-baseline"
$G > gpurun_out/w_gen_def.json 2>&1
V2R_IDW_TUNE=g2x4 $G > gpurun_out/w_gen_2x4.json 2>&1
V2R_IDW_TUNE=g4x4 $G > gpurun_out/w_gen_4x4.json 2>&1
python bench.py --steps 2 --band-rows 1024 --exps 1.1,1.3,1.5,1.7 --no-e2e --no-cpu-baseline > gpurun_out/w_gen4.json 2>&1
python scripts/summ.py gpurun_out/w_*.json
