set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; cut -c1-300 gpurun_out/final_ref.json
python bench.py > gpurun_out/final_default.json 2> gpurun_out/final_default.err; cut -c1-400 gpurun_out/final_default.json
python bench.py --precision fp32 --steps 3 > gpurun_out/final_fp32.json 2> gpurun_out/final_fp32.err
python bench.py --workload c4 --steps 1 --no-e2e --no-cpu-baseline > gpurun_out/final_c4_1gpu.json 2> gpurun_out/final_c4.err
python bench.py --workload c3 --steps 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/final_c3.json 2> gpurun_out/final_c3.err
python scripts/summ.py gpurun_out/final_*.json
