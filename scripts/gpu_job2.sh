set -x
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain_default2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:idw_fp64 -s 3 -c 1 -o gpurun_out/prof_fp64_c2_full $CMD > gpurun_out/ncu_f.log 2>&1
python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 1200 gpurun_out/bench_c3.json
python bench.py --workload c4 --band-rows 2048 --steps 2 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c4_band.json 2> gpurun_out/bench_c4.err; tail -c 1200 gpurun_out/bench_c4_band.json
python bench.py --band-rows 512 --exps 1.7 --steps 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_p17.json 2>&1; tail -c 700 gpurun_out/bench_p17.json
python bench.py --band-rows 1024 --exps 1.5 --steps 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_p15.json 2>&1; tail -c 700 gpurun_out/bench_p15.json
python bench.py --precision fp32 --steps 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2>&1; tail -c 900 gpurun_out/bench_fp32.json
