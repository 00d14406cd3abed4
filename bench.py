#!/usr/bin/env python
"""bench.py — headline benchmark of the fv2d hot path on B200.

Metric (BASELINE.json): Mcell-updates/s, fp64, Kelvin-Helmholtz 8192^2, HLLC + PLM (config
C3 of BASELINE.md), one "step" = one full time step of the reference loop body
(main.cpp:66-83: dt + update + consToPrim + checkNegatives), IO excluded.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code

N > 1 is launched by torchrun (one rank per GPU); the grid is split into y-slabs
(strong scaling: the total grid is fixed) and ghost rows are exchanged over peer memory.

Prints ONE JSON line (rank 0).  See DESIGN.md §6 for how every field is measured.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BYTES_PER_CELL_UPDATE = 128  # read Q,U + write U,Q, fp64 x 4 fields
WORKLOADS = {
    # name: (ini, overrides)
    "kelvin_helmholtz_8192_plm_hllc": ("kelvin_helmholtz.ini", {"mesh.Nx": 8192, "mesh.Ny": 8192,
                                                                 "solvers.reconstruction": "plm"}),
    "blast_4096_pcm_hllc": ("blast.ini", {"mesh.Nx": 4096, "mesh.Ny": 4096}),
    "rayleigh_taylor_16384_plm_hllc": ("rayleigh_taylor.ini", {"mesh.Nx": 16384, "mesh.Ny": 16384}),
    "c91_8192_pcm_hllc_tc_visc": ("C91.ini", {"mesh.Nx": 8192, "mesh.Ny": 8192}),
}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def write_ini(workload: str, extra=None) -> str:
    """Materialise the workload's .ini (settings/<base> + overrides) for the reference binary."""
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    from make_goldens import apply_overrides

    base, ov = WORKLOADS[workload]
    ov = dict(ov)
    ov.update(extra or {})
    text = apply_overrides((ROOT / "settings" / base).read_text(), ov)
    f = tempfile.NamedTemporaryFile("w", suffix=f"_{workload}.ini", delete=False)
    f.write(text)
    f.close()
    return f.name


def run_reference_binary(workload: str, extra, steps: int, warmup: int):
    """Times oracle/_ref/fv2d_ref (the unmodified reference, Kokkos-OpenMP) on the host cores."""
    ref = ROOT / "oracle" / "_ref" / "fv2d_ref"
    ncores = os.cpu_count() or 1
    ini = write_ini(workload, extra)
    if ref.exists():
        env = dict(os.environ, OMP_NUM_THREADS=str(ncores), OMP_PROC_BIND="spread", OMP_PLACES="threads")
        out = subprocess.run([str(ref), ini, "--steps", str(steps), "--warmup", str(warmup), "--bench", "--quiet"],
                             env=env, capture_output=True, text=True, cwd=tempfile.gettempdir())
        m = re.search(r"bench: steps=(\d+) seconds=([\d.eE+-]+) mcell_updates_per_s=([\d.eE+-]+) threads=(\d+)", out.stdout)
        if out.returncode == 0 and m:
            return {"value": float(m.group(3)), "seconds": float(m.group(2)), "steps": int(m.group(1)),
                    "cores": int(m.group(4)), "kind": "reference"}
    # fall back to the C restatement (oracle port)
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as O
    from fv2d_b200 import capi

    dev, run = capi.params_from_ini(ini)
    Q = capi.init_problem(dev, run)
    U = O.prim_to_cons(dev, Q)
    if warmup:
        O.run(dev, run.time_stepping, run.epsilon_reset_negative, 1e30, Q, U, warmup)
    t0 = time.perf_counter()
    n, *_ = O.run(dev, run.time_stepping, run.epsilon_reset_negative, 1e30, Q, U, steps)
    secs = time.perf_counter() - t0
    return {"value": dev.Nx * dev.Ny * n / secs / 1e6, "seconds": secs, "steps": n, "cores": ncores, "kind": "port"}


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    base, ov = WORKLOADS[args.workload]
    # bounded sample of the workload: same configuration and row length, 1/16 of the rows
    ny = max(64, int(ov.get("mesh.Ny", 256)) // args.ref_row_fraction)
    extra = {"mesh.Ny": ny}
    r = run_reference_binary(args.workload, extra, args.steps, args.warmup)
    sample = (f"{args.workload} with Ny={ny} (1/{args.ref_row_fraction} of the rows, same Nx), {r['steps']} timed steps "
              f"after {args.warmup} warm-up, {r['cores']} OpenMP threads")
    line = {
        "impl": "reference", "metric": "Mcell-updates/s", "value": r["value"], "unit": "Mcell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["seconds"] / max(r["steps"], 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "sample": sample},
        "cpu_baseline": {"value": r["value"], "unit": "Mcell-updates/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": sample},
        "e2e": {"value": r["value"], "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def native_arm(args):
    import torch
    import torch.distributed as dist

    from fv2d_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    base, ov = WORKLOADS[args.workload]
    ov = dict(ov)
    if args.nx:
        ov["mesh.Nx"] = args.nx
    if args.ny:
        ov["mesh.Ny"] = args.ny
    dev, run = capi.params_from_ini(ROOT / "settings" / base, ov)
    Nx, Ny = dev.Nx, dev.Ny

    ctx = capi.Context(dev, run.time_stepping, run.epsilon_reset_negative, device=local_rank, rank=rank, nranks=world)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    Nyl, joff = ctx.Ny, ctx.j_offset
    # synthetic initial condition from the reference's own init function, on the host; each
    # rank evaluates only the rows of its own y-slab (+ ghost rows)
    os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 1) // world)))
    t_init = time.perf_counter()
    Qloc = capi.init_problem_rows(dev, run, joff, Nyl + 2 * dev.Ng)
    t_init = time.perf_counter() - t_init
    if world > 1:
        from fv2d_b200 import multigpu

        multigpu.connect(ctx, dist)
    ctx.upload_Q(Qloc)
    ctx.prim_to_cons()
    ctx.compute_dt()
    ctx.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    ctx.run_steps(args.warmup)
    barrier()

    # ---- timed region: K fused steps, state resident in HBM, dt resident on the device
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    ctx.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        ctx.run_steps(args.steps)
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    sweep_ms, sweep_launches, total_launches = ctx.profile_read()
    ctx.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tmax = torch.tensor([ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    value = Nx * Ny * args.steps / (ms * 1e-3) / 1e6
    neg = ctx.negative_counts()

    # ---- roofline of the dominant kernel (the fused sweep), per launch, local slab
    peak, peak_src = measured_peak_gbs()
    per_launch_s = (sweep_ms / max(sweep_launches, 1)) * 1e-3
    achieved = BYTES_PER_CELL_UPDATE * Nx * Nyl / per_launch_s / 1e9
    traffic = None
    tj = ROOT / "profiles" / "sweep_traffic.json"
    if tj.exists():
        try:
            tinfo = json.loads(tj.read_text())
            if tinfo.get("workload") == args.workload and tinfo.get("Nx") == Nx and tinfo.get("Ny_local") == Nyl:
                traffic = tinfo["dram_bytes_per_launch"]
        except Exception:
            pass

    # ---- end to end through the host-buffer C ABI call: every step uploads the state from
    #      pinned host memory, advances one step and reads the new state + dt back
    #      (N > 1: every rank does so for its own y-slab, ghost rows included; wall clock between
    #      two barriers, maximum over the ranks)
    e2e = None
    if args.e2e_steps > 0:
        hin = torch.from_numpy(Qloc).pin_memory()
        hout = torch.empty_like(hin).pin_memory()
        a_in, a_out = hin.numpy(), hout.numpy()
        dts = np.zeros(1)
        ctx.advance_host(a_in, a_out, 1, dts)  # warm-up
        a_in, a_out = a_out, a_in
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ctx.advance_host(a_in, a_out, 1, dts)
            a_in, a_out = a_out, a_in
        barrier()
        secs = time.perf_counter() - t0
        if world > 1:
            tsec = torch.tensor([secs], device="cuda", dtype=torch.float64)
            dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
            secs = float(tsec.item())
        nbytes = int(Qloc.nbytes) * world
        e2e = {"value": Nx * Ny * args.e2e_steps / secs / 1e6, "unit": "Mcell-updates/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes + 8 * world, "steps": args.e2e_steps,
               "api": "fv2d_advance_host(ctx, hostQ_in, hostQ_out, 1, &dt)" + (" on every rank's y-slab" if world > 1 else "")}

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference itself on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_n = args.cpu_sample
        r = run_reference_binary(args.workload, {"mesh.Nx": sample_n, "mesh.Ny": sample_n}, args.cpu_steps, 2)
        cpu = {"value": r["value"], "unit": "Mcell-updates/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"{args.workload} scaled to {sample_n}x{sample_n}, {r['steps']} timed steps after 2 warm-up "
                         f"({r['seconds']:.1f} s of CPU work)"}

    if rank == 0:
        line = {
            "metric": "Mcell-updates/s", "value": value, "unit": "Mcell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "Nx": Nx, "Ny": Ny, "riemann_solver": "hllc",
                       "reconstruction": {0: "pcm", 1: "pcm_wb", 2: "plm"}[dev.reconstruction],
                       "time_stepping": "euler" if run.time_stepping == 0 else "rk2",
                       "decomposition": f"{world} y-slab(s)", "l2_policy": "working set (3 arrays x %.2f GB) >> 126 MB L2"
                       % (Qloc.nbytes * world / 1e9), "host_init_s": round(t_init, 2)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_sweep (fused RK stage)", "peak_source": peak_src,
                         "bytes_per_cell_update": BYTES_PER_CELL_UPDATE, "ms_per_launch": per_launch_s * 1e3,
                         "share_of_step": sweep_ms / ms if ms > 0 else None},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(total_launches), "clocks": clocks,
            "sanity": {"negative_density": neg[0], "negative_pressure": neg[1], "nan": neg[2]},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="kelvin_helmholtz_8192_plm_hllc", choices=sorted(WORKLOADS))
    ap.add_argument("--nx", type=int, default=0, help="override Nx (development only)")
    ap.add_argument("--ny", type=int, default=0, help="override Ny (development only)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=2048, help="edge of the CPU-baseline sample grid")
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--ref-row-fraction", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    return reference_arm(args) if args.impl == "reference" else native_arm(args)


if __name__ == "__main__":
    sys.exit(main())
