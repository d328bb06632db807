set -x
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/m2_c2_weak.json 2> gpurun_out/m2_c2_weak.err; tail -3 gpurun_out/m2_c2_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --scaling strong --band-rows 1024 --steps 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/m2_c4_strong_band.json 2> gpurun_out/m2_c4.err; tail -3 gpurun_out/m2_c4.err
python - <<'PY'
import torch, numpy as np, time, sys
sys.path.insert(0, '.')
import v2r_b200 as V
from v2r_b200 import workloads
from v2r_b200.idw import make_grid
# distributed.py path sanity under 2 ranks is covered by bench; here: single-process ctx timing on c2
x,y,w,xi,yi,ex = workloads.make('c2'); ps = V.MakeCoordSpace(x,y,w,xi,yi); g = make_grid(xi,yi)
for n in (1,2):
    with V.IdwContext(n) as c:
        c.solve(ps,g,ex); t=time.perf_counter(); z=c.solve(ps,g,ex); dt=time.perf_counter()-t
        print('single-process ctx n_gpus',n,'c2 solve s',round(dt,4),'G evals/s',round(g.rows*g.cols*len(ps)/dt/1e9,1), c.last_timing())
PY
python scripts/summ.py gpurun_out/m2_*.json
