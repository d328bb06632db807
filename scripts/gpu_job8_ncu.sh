set -x
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_default_paired.csv $CMD > gpurun_out/ncu_a.log 2>&1
$CMD > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:idw_fp64 -s 3 -c 1 -o gpurun_out/prof_fp64_c2_paired $CMD > gpurun_out/ncu_b.log 2>&1
CMD32="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --precision fp32"
$CMD32 > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:idw_fp32 -s 3 -c 1 -o gpurun_out/prof_fp32_c2 $CMD32 > gpurun_out/ncu_c.log 2>&1
CMD3="python bench.py --workload c3 --band-rows 256 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD3 > gpurun_out/plain_d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:idw_fp64 -s 3 -c 1 -o gpurun_out/prof_fp64_c3_band256 $CMD3 > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
