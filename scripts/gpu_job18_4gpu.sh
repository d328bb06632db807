set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/m4_c2_weak.json 2> gpurun_out/m4_a.err; tail -2 gpurun_out/m4_a.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 4 --workload c4 --scaling strong --steps 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/m4_c4_strong.json 2> gpurun_out/m4_b.err; tail -2 gpurun_out/m4_b.err
python scripts/summ.py gpurun_out/m4_*.json
