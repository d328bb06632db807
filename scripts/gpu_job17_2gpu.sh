set -x
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/m2_final.json 2> gpurun_out/m2_final.err; tail -2 gpurun_out/m2_final.err
python scripts/summ.py gpurun_out/m2_final.json
