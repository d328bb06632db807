set -x
B="python bench.py --steps 3 --no-e2e --no-cpu-baseline"
for t in none c6x4 c4x6 c5x4; do V2R_IDW_TUNE=$t $B > gpurun_out/z_c2_$t.json 2>&1; done
C="python bench.py --workload c4 --band-rows 1536 --steps 1 --no-e2e --no-cpu-baseline"
for t in none c6x4 c4x6; do V2R_IDW_TUNE=$t $C > gpurun_out/z_c4_$t.json 2>&1; done
V2R_IDW_TUNE=c6x4 python -m pytest tests/test_gpu_parity.py -x -q -k "c2_full_size or tile_boundaries or c4_clustered" 2>&1 | tail -2
python scripts/summ.py gpurun_out/z_*.json
