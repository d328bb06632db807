set -x
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/m8_c2_weak.json 2> gpurun_out/m8_c2_weak.err; tail -3 gpurun_out/m8_c2_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload c4 --scaling strong --steps 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/m8_c4_strong.json 2> gpurun_out/m8_c4.err; tail -3 gpurun_out/m8_c4.err
python scripts/summ.py gpurun_out/m8_*.json
