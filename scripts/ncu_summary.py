#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) of the sweep kernel into a small text file
for profiles/: headline metrics, DRAM traffic per launch vs algorithmic bytes, pipe
utilisation, stall reasons, instruction mix and the hottest SASS lines.

    python scripts/ncu_summary.py gpurun_out/sweep.ncu-rep --cells 67108864 > profiles/sweep_rNN.txt
"""
import argparse
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
    "sm__cycles_active.avg", "smsp__cycles_active.avg",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--cells", type=int, default=0, help="cell-updates per launch (for bytes/instr per cell)")
    ap.add_argument("--json", default="", help="also write {dram_bytes_per_launch,...} here")
    a = ap.parse_args()

    rows = list(csv.reader(io.StringIO(ncu(["-i", a.rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# ncu summary of {a.rep}")
    traffic = []
    for n, d in enumerate(data):
        name = d[hdr.index("Kernel Name")]
        print(f"\n## launch {n}: {name[:100]}")
        for k in KEYS:
            if k in hdr:
                print(f"  {k:72s} {d[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        def val(k):
            v = float(d[hdr.index(k)].replace(",", ""))
            u = units[hdr.index(k)]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        try:
            tr = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            traffic.append(tr)
            ms = float(d[hdr.index("gpu__time_duration.sum")].replace(",", ""))
            if units[hdr.index("gpu__time_duration.sum")] == "us":
                ms /= 1e3
            elif units[hdr.index("gpu__time_duration.sum")] == "ns":
                ms /= 1e6
            print(f"  -> DRAM traffic per launch {tr/1e9:.3f} GB = {tr/ (ms*1e-3) / 1e9:.0f} GB/s (under the profiler)")
            if a.cells:
                inst = float(d[hdr.index("smsp__inst_executed.sum")].replace(",", ""))
                print(f"  -> {tr/a.cells:.1f} DRAM bytes per cell-update (algorithmic: 128); "
                      f"{inst*32/a.cells:.0f} thread-instructions per cell-update")
        except Exception as e:  # noqa
            print("  (traffic unavailable)", e)
        st = [(h, d[i]) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
        st = sorted(((h, float(v.replace(",", ""))) for h, v in st if v not in ("", "n/a")), key=lambda x: -x[1])
        print("  stall reasons (warps stalled per issue-active cycle):")
        for h, v in st[:8]:
            print(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")

    src = list(csv.reader(io.StringIO(ncu(["-i", a.rep, "--page", "source", "--csv"]))))
    # first kernel's table only
    try:
        h = next(i for i, r in enumerate(src) if r and r[0] == "Address")
        shdr = src[h]
        body = []
        for r in src[h + 1:]:
            if not r or r[0] == "Kernel Name":
                break
            body.append(r)
        iS, iA, iE = shdr.index("Source"), shdr.index("Warp Stall Sampling (All Samples)"), shdr.index("Instructions Executed")
        tot = sum(int(r[iA]) for r in body) or 1
        mix = collections.Counter()
        for r in body:
            toks = r[iS].split()
            op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
            mix[op.split(".")[0]] += int(r[iE])
        t = sum(mix.values()) or 1
        print("\n## instruction mix (warp-instructions executed)")
        for op, c in mix.most_common(16):
            print(f"  {op:10s} {100*c/t:5.1f} %")
        print("\n## hottest SASS lines (share of stall samples)")
        scols = [(i, x) for i, x in enumerate(shdr) if x.startswith("stall_") and "Not Issued" not in x]
        for k in sorted(range(len(body)), key=lambda k: -int(body[k][iA]))[:12]:
            r = body[k]
            rs = sorted(((x, int(r[i])) for i, x in scols if r[i] not in ("", "0")), key=lambda x: -x[1])[:2]
            print(f"  {100*int(r[iA])/tot:5.1f} %  {r[iS].strip()[:56]:56s} {rs}")
    except StopIteration:
        print("(no source page)")
    if a.json and traffic:
        json.dump({"dram_bytes_per_launch": sum(traffic) / len(traffic)}, open(a.json, "w"))


if __name__ == "__main__":
    main()
