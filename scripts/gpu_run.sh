#!/bin/bash
# One parameterised runner for everything this repository runs on a B200 box:
#     gpurun --timeout 1800 -- 'bash scripts/gpu_run.sh <suite> [tag]'
# Outputs go to gpurun_out/<tag>_* (scratch); what is quoted in DESIGN.md / profiles/README.md is copied to profiles/.
#   tests     pytest -m gpu
#   classes   one bench line per exponent class / workload (diagnostic bands) + the default bench
#   tail      short c4 bands with and without the tail split (V2R_IDW_TAIL=0)
#   ncu       --set full capture per kernel class (summarised on the box)
#   ncu-more  rsqrt / root4 single-exponent kernels (full set) and the c5 4-exponent kernel (sections without source counters)
#   ncu-lite  metrics-only: FP64-pipe active vs elapsed + DRAM bytes on short bands and the default launch, launch list of the bench
#   multi     (gpurun --gpus N) multi-GPU tests, c2 weak (direct and NCCL-gather copy-out) and the default bench under
#             torchrun; `multi <tag> quick` runs the default bench only
#   tune      tuning builds (cell blocks, staged dy^2) against the default kernels
#   quick     a 90-second sanity pass for host-side changes: smoke(), the CLI tests, the small parity tests, one c2 bench line
#   quad      the p = 2 tuning build with four points per reciprocal (V2R_IDW_TUNE=quad): small parity tests, then c2 and a
#             32-row c4 band against the default build, parity block included
#   stream    streaming API against one whole-grid solve (scripts/stream_timing.py) and the CLI on c2 / c5-shaped inputs
suite=${1:-tests}
tag=${2:-r2}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out/${tag}
Q="--no-e2e --no-cpu-baseline --no-parity"
summ() { for f in "$@"; do python scripts/summ.py "$f" | cut -c12-260; done; }

case "$suite" in
tests)
    timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log
    tail -8 ${O}_pytest.log ;;
classes)
    B="python bench.py --steps 3 --warmup 3 $Q --workload c2"
    $B > ${O}_c2.json 2> ${O}.err
    V2R_IDW_FP64_PAIRED=0 $B > ${O}_c2_unpaired.json 2>> ${O}.err
    for p in 1.0 3.0 1.5 0.5 1.7 2.2 4.2 0.0 -1.0; do $B --exps=$p --band-rows 512 > ${O}_p${p}.json 2>> ${O}.err; done
    $B --exps 1.1,1.3,1.5,1.7 --band-rows 512 > ${O}_gen4.json 2>> ${O}.err
    $B --precision fp32 > ${O}_fp32.json 2>> ${O}.err
    python bench.py --steps 3 --warmup 3 $Q --workload c3 --band-rows 512 > ${O}_c3_band512.json 2>> ${O}.err
    python bench.py --steps 2 --warmup 3 $Q --workload c5 --band-rows 256 > ${O}_c5_band256.json 2>> ${O}.err
    python bench.py --steps 3 --warmup 3 $Q --precision fp32 > ${O}_c4slab_fp32.json 2>> ${O}.err
    python bench.py > ${O}_default.json 2>> ${O}.err
    python bench.py --impl reference --steps 3 --warmup 1 > ${O}_reference.json 2>> ${O}.err
    tail -3 ${O}.err; summ ${O}_*.json ;;
tail)
    for rows in 128 256 512 2048; do
        python bench.py --steps 3 --warmup 3 $Q --slab 7168:$rows > ${O}_c4_rows${rows}.json 2>> ${O}.err
        V2R_IDW_TAIL=0 python bench.py --steps 3 --warmup 3 $Q --slab 7168:$rows > ${O}_c4_rows${rows}_notail.json 2>> ${O}.err
    done
    summ ${O}_c4_rows*.json ;;
ncu)
    prof() {  # name, kernel regex, bench args...   (plain run first, `&&`, no pipe: B200_PROFILING.md)
        local name=$1 rx=$2; shift 2
        python bench.py --steps 1 --warmup 3 $Q "$@" > ${O}_plain_$name.log 2>&1 &&
        ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o ${O}_$name python bench.py --steps 1 --warmup 3 $Q "$@" > ${O}_ncu_$name.log 2>&1
        echo "$name rc=$?"
        # gpurun_out/ is capped at 64 MiB: keep the small summaries (metrics JSON + stall / opcode table), drop the report
        python profiles/summarize_ncu.py ${O}_$name.ncu-rep ${O}_$name > ${O}_summ_$name.log 2>&1 && rm -f ${O}_$name.ncu-rep
    }
    prof fp64_p2_c2_full idw_fp64 --workload c2
    prof fp64_p2_c4_band256 idw_fp64 --slab 7168:256
    prof fp64_general_p17_band512 idw_fp64 --workload c2 --exps 1.7 --band-rows 512
    prof fp32_packed_c2 idw_fp32 --workload c2 --precision fp32
    prof fp64_c3_band256 idw_fp64 --workload c3 --band-rows 256
    prof fp64_c5_e4_band16 idw_fp64 --workload c5 --band-rows 16
    prof fp64_rsqrt_p1_band512 idw_fp64 --workload c2 --exps 1.0 --band-rows 512
    prof fp64_root4_p15_band512 idw_fp64 --workload c2 --exps 1.5 --band-rows 512
    ls -la ${O}_*metrics.json ;;
ncu-more)   # the captures the `ncu` suite did not reach: single-exponent rsqrt / root4 kernels (full set), and the c5
            # 4-exponent kernel WITHOUT the source counters (the instrumented passes of that kernel take minutes each)
    prof() {
        local name=$1 rx=$2; shift 2
        python bench.py --steps 1 --warmup 3 $Q "$@" > ${O}_plain_$name.log 2>&1 &&
        ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o ${O}_$name python bench.py --steps 1 --warmup 3 $Q "$@" > ${O}_ncu_$name.log 2>&1
        echo "$name rc=$?"
        python profiles/summarize_ncu.py ${O}_$name.ncu-rep ${O}_$name > ${O}_summ_$name.log 2>&1 && rm -f ${O}_$name.ncu-rep
    }
    prof fp64_rsqrt_p1_band512 idw_fp64 --workload c2 --exps 1.0 --band-rows 512
    prof fp64_root4_p15_band512 idw_fp64 --workload c2 --exps 1.5 --band-rows 512
    name=fp64_c5_e4_band16
    python bench.py --steps 1 --warmup 3 $Q --workload c5 --band-rows 16 > ${O}_plain_$name.log 2>&1 &&
    timeout 600 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section WarpStateStats \
        --clock-control none -k regex:idw_fp64 -s 3 -c 1 -f -o ${O}_$name python bench.py --steps 1 --warmup 3 $Q --workload c5 --band-rows 16 > ${O}_ncu_$name.log 2>&1
    echo "$name rc=$?"
    python profiles/summarize_ncu.py ${O}_$name.ncu-rep ${O}_$name --metrics-only > ${O}_summ_$name.log 2>&1 && rm -f ${O}_$name.ncu-rep
    ls -la ${O}_*metrics.json ;;
ncu-lite)   # metrics-only passes: FP64 pipe active vs elapsed + DRAM bytes on short bands and the default launch, launch list
    M=sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
    for rows in 128 256 2048; do
        python bench.py --steps 1 --warmup 3 $Q --slab 7168:$rows > ${O}_plain_tail$rows.log 2>&1 &&
        ncu --metrics $M --clock-control none -k regex:idw_fp64 -s 3 -c 1 --csv --log-file ${O}_tail_c4_rows$rows.csv python bench.py --steps 1 --warmup 3 $Q --slab 7168:$rows > ${O}_ncu_tail$rows.log 2>&1
        echo "tail $rows rc=$?"
    done
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-cells 200 > ${O}_plain_launches.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-cells 200 > ${O}_ncu_launches.log 2>&1
    echo "launches rc=$?"
    python scripts/ncu_traffic.py ${O}_tail_c4_rows2048.csv 2048 16384 967142 > ${O}_traffic.log 2>&1; cat ${O}_traffic.log
    cp profiles/traffic.json ${O}_traffic.json
    ls -la ${O}_*.csv ;;
multi)
    n=$(nvidia-smi -L | wc -l)
    T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541"
    if [ "${3:-full}" = "full" ]; then
        timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > ${O}_pytest_multi.log 2>&1; tail -4 ${O}_pytest_multi.log
        $T bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline --workload c2 --scaling weak > ${O}_m${n}_c2_weak.json 2> ${O}_m${n}.err
        V2R_IDW_GATHER=nccl $T bench.py --gpus $n --steps 2 --warmup 3 --no-cpu-baseline --workload c2 --scaling weak > ${O}_m${n}_c2_weak_ncclgather.json 2>> ${O}_m${n}.err
    fi
    $T bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > ${O}_m${n}_default.json 2>> ${O}_m${n}.err
    tail -3 ${O}_m${n}.err; summ ${O}_m${n}_*.json
    python - <<P
import json, glob
for f in sorted(glob.glob("${O}_m${n}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, json.dumps({"e2e": d["e2e"], "parity": d["parity"]})[:1200])
    except Exception as e:
        print(f, "unreadable", e)
P
    ;;
tune)   # tuning builds against the defaults (V2R_IDW_TUNE=<tag>, see the variant tables in csrc/idw_launch*.cu)
    B="python bench.py --steps 3 --warmup 3 $Q"
    for t in none nys; do   # nys = every lane computes its own dy^2 (no staging in shared memory)
        V2R_IDW_TUNE=$t $B --workload c2 > ${O}_tune_c2_$t.json 2>> ${O}.err
        V2R_IDW_TUNE=$t $B --slab 7168:256 > ${O}_tune_c4r256_$t.json 2>> ${O}.err
        for p in 1.0 1.5 1.7; do V2R_IDW_TUNE=$t $B --workload c2 --exps=$p --band-rows 512 > ${O}_tune_p${p}_$t.json 2>> ${O}.err; done
        V2R_IDW_TUNE=$t $B --workload c2 --exps 1.1,1.3,1.5,1.7 --band-rows 512 > ${O}_tune_gen4_$t.json 2>> ${O}.err
        V2R_IDW_TUNE=$t $B --workload c3 --band-rows 512 > ${O}_tune_c3_$t.json 2>> ${O}.err
        V2R_IDW_TUNE=$t python bench.py --steps 2 --warmup 3 $Q --workload c5 --band-rows 128 > ${O}_tune_c5_$t.json 2>> ${O}.err
    done
    for t in c4x6 ys4x6; do V2R_IDW_TUNE=$t $B --workload c2 > ${O}_tune_c2_$t.json 2>> ${O}.err; done
    for t in none f2x8; do V2R_IDW_TUNE=$t $B --workload c2 --precision fp32 > ${O}_tune_fp32_$t.json 2>> ${O}.err; done
    tail -3 ${O}.err; summ ${O}_tune_*.json ;;
stream)
    python scripts/stream_timing.py > ${O}_stream_timing.log 2>&1; tail -5 ${O}_stream_timing.log
    python scripts/cli_c2.py > ${O}_cli_c2.log 2>&1; tail -5 ${O}_cli_c2.log ;;
quick)
    timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a ${O}_smoke.log
    timeout 60 python -m pytest tests/test_cli.py tests/test_gpu_parity.py -m gpu -x -q \
        -k "cli or golden or c1_vs or edge or stream or gpkg or error or tile_boundaries" > ${O}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a ${O}_pytest.log
    tail -3 ${O}_pytest.log
    timeout 40 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > ${O}_c2.json 2> ${O}.err; echo "bench rc=$?"
    summ ${O}_c2.json ;;
quad)
    V2R_IDW_TUNE=quad timeout 30 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
        -k "c1_vs or tile_boundaries or edge_semantics or paired_kernels or randomized or row_bands or golden" > ${O}_pytest.log 2>&1
    echo "pytest(quad) rc=$?" | tee -a ${O}_pytest.log; tail -2 ${O}_pytest.log
    timeout 12 python bench.py --workload c2 --steps 3 --warmup 3 $Q > ${O}_c2_default.json 2>> ${O}.err
    V2R_IDW_TUNE=quad timeout 14 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --parity-cells 500 > ${O}_c2_quad.json 2>> ${O}.err
    V2R_IDW_TUNE=quad timeout 16 python bench.py --slab 7168:32 --steps 3 --warmup 3 --no-cpu-baseline --parity-cells 150 > ${O}_c4rows32_quad.json 2>> ${O}.err
    tail -3 ${O}.err; summ ${O}_c2_default.json ${O}_c2_quad.json ${O}_c4rows32_quad.json ;;
*) echo "unknown suite $suite"; exit 2 ;;
esac
