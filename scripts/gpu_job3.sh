set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/t_p2_paired.json 2>&1
V2R_IDW_FP64_PAIRED=0 $B > gpurun_out/t_p2_unpaired.json 2>&1
V2R_IDW_FP64_PAIRED=0 V2R_IDW_TUNE=t384 $B > gpurun_out/t_p2_t384.json 2>&1
V2R_IDW_FP64_PAIRED=0 V2R_IDW_TUNE=b2 $B > gpurun_out/t_p2_b2.json 2>&1
P="python bench.py --steps 2 --band-rows 1024 --exps 1.5 --no-e2e --no-cpu-baseline"
$P > gpurun_out/t_p15_def.json 2>&1
V2R_IDW_TUNE=t384 $P > gpurun_out/t_p15_t384.json 2>&1
V2R_IDW_TUNE=b2 $P > gpurun_out/t_p15_b2.json 2>&1
S="python bench.py --workload c3 --steps 2 --band-rows 1024 --no-e2e --no-cpu-baseline"
$S > gpurun_out/t_c3_def.json 2>&1
V2R_IDW_TUNE=b2 $S > gpurun_out/t_c3_b2.json 2>&1
V2R_IDW_TUNE=c1x4 $S > gpurun_out/t_c3_c1x4.json 2>&1
python scripts/summ.py gpurun_out/t_*.json
