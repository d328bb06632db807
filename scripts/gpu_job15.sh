set -x
python scripts/cli_c2.py
CMD="python bench.py --steps 2 --band-rows 256 --exps 1.7 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_gen.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:idw_fp64 -s 3 -c 1 -o gpurun_out/prof_fp64_gen_band256 $CMD > gpurun_out/ncu_gen.log 2>&1
tail -2 gpurun_out/ncu_gen.log | cut -c1-200
