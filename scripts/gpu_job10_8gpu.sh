set -x
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --workload c5 --scaling strong --band-rows 512 --steps 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/m8_c5_strong_band512.json 2> gpurun_out/m8_c5.err; tail -3 gpurun_out/m8_c5.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --workload c4 --scaling strong --precision fp32 --steps 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/m8_c4_strong_fp32.json 2> gpurun_out/m8_c4f.err; tail -3 gpurun_out/m8_c4f.err
python scripts/summ.py gpurun_out/m8_c5_strong_band512.json gpurun_out/m8_c4_strong_fp32.json
