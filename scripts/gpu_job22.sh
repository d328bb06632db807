set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
G="python bench.py --steps 2 --band-rows 1024 --no-e2e --no-cpu-baseline"
$G --exps 1.7 > gpurun_out/q_gen_p17.json 2>&1
$G --exps 4.2 > gpurun_out/q_gen_p42.json 2>&1
$G --exps 1.1,1.3,1.5,1.7 > gpurun_out/q_gen4.json 2>&1
python scripts/summ.py gpurun_out/q_*.json
