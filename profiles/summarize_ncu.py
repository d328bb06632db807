#!/usr/bin/env python
"""Turns an Nsight Compute report (gpurun_out/*.ncu-rep, scratch) into the small tracked summaries under
profiles/: selected raw metrics per launch (JSON) and the warp-stall / instruction-mix table of the hot
kernel (markdown).  Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_name"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(r"^(gpu__time_duration\.sum|l1tex__m_xbar2l1tex_read_bytes\.sum|lts__t_sector_hit_rate\.pct|lts__t_sectors\.sum|dram__bytes_(read|write)\.sum|lts__t_bytes\.sum|lts__t_sectors_op_(read|write)\.sum|"
                  r"launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+|waves_per_multiprocessor|shared_mem_per_block_static)|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__inst_executed_pipe_(fp64|xu|fma|fmaheavy|fmalite|alu|lsu|uniform)\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                  r"sm__pipe_(fp64|fma|fmaheavy|xu)_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|sm__cycles_active\.avg|"
                  r"sm__cycles_elapsed\.avg\.per_second|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|"
                  r"smsp__sass_thread_inst_executed_op_(dfma|dmul|dadd|ffma|fmul|fadd)_pred_on\.sum|"
                  r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__inst_executed_pipe_\w+\.sum)$")


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, check=True).stdout


def main(rep, out):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if KEEP.match(h) and r[i] not in ("", "nan", "-nan"):
                d[f"{h} [{units[i]}]"] = r[i]
        launches.append(d)
    json.dump(launches, open(out + "_metrics.json", "w"), indent=1)
    if len(sys.argv) > 3 and sys.argv[3] == "--metrics-only":   # captures taken without the source counters
        print("wrote", out + "_metrics.json")
        return

    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    hi = [i for i, r in enumerate(src) if r and r[0] == "Address"][0]
    h = src[hi]
    end = [i for i, r in enumerate(src) if i > hi and r and r[0] == "Kernel Name"]
    body = src[hi + 1:(end[0] if end else len(src))]
    stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    S, X = h.index("# Samples"), h.index("Instructions Executed")
    total = sum(int(r[S] or 0) for r in body) or 1
    tot = {h[i]: sum(int(r[i] or 0) for r in body) for i in stall}
    mix = {}
    for r in body:
        op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
        op = op.split(".")[0] + ("." + r[1].split()[0].split(".")[1] if op == "MUFU" and "." in r[1].split()[0] else "")
        mix[op] = mix.get(op, 0) + int(r[X] or 0)
    n_inst = sum(mix.values()) or 1
    with open(out + "_stalls.md", "w") as f:
        f.write(f"# {src[0][1]}\n\nsource: `{rep}` (ncu --set full --clock-control none --import-source on)\n\n")
        f.write(f"## warp stall samples (all samples = {total})\n\n| reason | samples | share |\n|---|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            if v:
                f.write(f"| {k} | {v} | {100 * v / total:.1f}% |\n")
        f.write(f"\n## executed warp instructions by opcode (total {n_inst})\n\n| opcode | warp instructions | share |\n|---|---:|---:|\n")
        for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:16]:
            f.write(f"| {k} | {v} | {100 * v / n_inst:.2f}% |\n")
        f.write("\n## hottest instructions\n\n| samples | SASS |\n|---:|---|\n")
        for r in sorted(body, key=lambda r: -int(r[S] or 0))[:12]:
            f.write(f"| {r[S]} | `{r[1].strip()}` |\n")
    print("wrote", out + "_metrics.json", out + "_stalls.md")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
