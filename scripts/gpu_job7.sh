set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
G="python bench.py --steps 2 --band-rows 1024 --exps 1.7 --no-e2e --no-cpu-baseline"
$G > gpurun_out/g_def.json 2>&1
V2R_IDW_TUNE=g2x4 $G > gpurun_out/g_2x4.json 2>&1
V2R_IDW_TUNE=g4x4 $G > gpurun_out/g_4x4.json 2>&1
V2R_IDW_TUNE=g1x2 $G > gpurun_out/g_1x2.json 2>&1
python scripts/summ.py gpurun_out/g_*.json
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "tile_boundaries or edge_semantics or golden or fp32_tile" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/memcheck.log
