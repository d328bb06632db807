set -x
python -m pytest tests/test_gpu_parity.py -x -q -k "random_1k or c1_vs_oracle or sweeps_match or edge_semantics or randomized or golden or tile_boundaries" 2>&1 | tail -3
G="python bench.py --steps 2 --band-rows 1024 --exps 1.7 --no-e2e --no-cpu-baseline"
$G > gpurun_out/y_gen_def.json 2>&1
V2R_IDW_TUNE=g2x4 $G > gpurun_out/y_gen_2x4.json 2>&1
python bench.py --steps 2 --band-rows 1024 --exps 1.1,1.3,1.5,1.7 --no-e2e --no-cpu-baseline > gpurun_out/y_gen4.json 2>&1
python scripts/summ.py gpurun_out/y_*.json
