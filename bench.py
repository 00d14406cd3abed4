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
# fp64-pipe instructions (DFMA/DMUL/DADD/DSETP) the shipped sweep executes per cell-update, from the
# SASS of the row loop (scripts/sass_lines.py; profiles/README.md): numerator of the fp64 roofline
FP64_INSTR_PER_CELL_UPDATE = {"kelvin_helmholtz_8192_plm_hllc": 177, "rayleigh_taylor_16384_plm_hllc": 183}
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
         "clocks_event_reasons.sw_power_cap,enforced.power.limit")

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
        sm, mx, reasons, pw, limit = [], [], set(), [], None
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
            if len(f) > 8:
                try:
                    limit = float(f[8])
                except ValueError:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "power_limit_w": limit, "samples": len(sm), "reasons": sorted(reasons)}


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
    """--impl reference: the reference's own CPU implementation (oracle/_ref/fv2d_ref = the
    unmodified reference headers + Kokkos-OpenMP) on the box's host cores, on the SAME grid as the
    native arm.  The number of timed steps is capped so that the run ends within a few minutes
    (one 8192^2 step is ~2.5 s of CPU time on 16 cores); throughput per step does not depend on it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = min(args.steps, args.ref_max_steps), min(args.warmup, 3)
    r = run_reference_binary(args.workload, {}, steps, warmup)
    base, ov = WORKLOADS[args.workload]
    sample = (f"{args.workload} at its full size ({ov.get('mesh.Nx')}x{ov.get('mesh.Ny')}), {r['steps']} timed steps after "
              f"{warmup} warm-up, {r['cores']} OpenMP threads")
    line = {
        "impl": "reference", "metric": "Mcell-updates/s", "value": r["value"], "unit": "Mcell-updates/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": warmup,
        "ms_per_step": 1e3 * r["seconds"] / max(r["steps"], 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "Nx": ov.get("mesh.Nx"), "Ny": ov.get("mesh.Ny"), "sample": sample},
        "cpu_baseline": {"value": r["value"], "unit": "Mcell-updates/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": sample},
        "e2e": {"value": r["value"], "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


class Job:
    """One workload resident on this rank's GPU (its y-slab of the grid)."""

    def __init__(self, env, workload, extra=None):
        from fv2d_b200 import capi

        self.env = env
        base, ov = WORKLOADS[workload]
        ov = dict(ov)
        ov.update(extra or {})
        self.workload = workload
        self.dev, self.run = capi.params_from_ini(ROOT / "settings" / base, ov)
        self.Nx, self.Ny = self.dev.Nx, self.dev.Ny
        self.ctx = capi.Context(self.dev, self.run.time_stepping, self.run.epsilon_reset_negative, device=env.local_rank,
                                rank=env.rank, nranks=env.world)
        self.ctx.set_stream(env.stream.cuda_stream)
        self.Nyl = self.ctx.Ny
        # synthetic initial condition from the reference's own init function, on the host; each
        # rank evaluates only the rows of its own y-slab (+ ghost rows)
        t0 = time.perf_counter()
        self.Qloc = capi.init_problem_rows(self.dev, self.run, self.ctx.j_offset, self.Nyl + 2 * self.dev.Ng)
        self.host_init_s = time.perf_counter() - t0
        if env.world > 1:
            from fv2d_b200 import multigpu

            multigpu.connect(self.ctx, env.dist)
        self.ctx.upload_Q(self.Qloc)
        self.ctx.prim_to_cons()
        self.ctx.compute_dt()
        self.ctx.sync()

    def timed(self, steps, bracket_launches=False):
        """`steps` fused steps between barrier + synchronize, CUDA events on the context's stream,
        maximum over the ranks.  Returns (ms, sweep_ms, sweep_launches, all_launches) of this region.
        sweep_ms: by default the region itself (the sweep is the only kernel of a step; consecutive
        launches overlap prologue and tail, so a launch's share of the region is region / launches);
        with bracket_launches every launch gets its own event pair (and runs on its own)."""
        import torch

        env, ctx = self.env, self.ctx
        ctx.sync_wait(reset=True)
        ctx.profile_enable(1 if bracket_launches else 2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        env.barrier()
        with torch.cuda.stream(env.stream):
            e0.record(env.stream)
            ctx.run_steps(steps)
            e1.record(env.stream)
        env.barrier()
        ms_local = e0.elapsed_time(e1)
        ms = env.max_over_ranks(ms_local)
        sweep_ms, sweep_launches, total_launches = ctx.profile_read()
        if not bracket_launches:
            sweep_ms = ms_local
        ctx.profile_enable(0)
        # per-step wait of this rank's first CTA for the other ranks' CFL mails (the decomposition's
        # synchronisation cost; 0 on one GPU), maximum over the ranks
        self.cfl_wait_us = env.max_over_ranks(ctx.sync_wait()[1] / max(sweep_launches, 1))
        return ms, sweep_ms, sweep_launches, total_launches

    def global_hash(self):
        """Hash of the conserved state of the WHOLE grid: slab hashes added modulo 2^64."""
        h = self.ctx.state_hash()
        if self.env.world > 1:
            import torch

            mine = torch.tensor([h - (1 << 64) if h >= (1 << 63) else h], dtype=torch.int64, device="cuda")
            out = [torch.empty_like(mine) for _ in range(self.env.world)]
            self.env.dist.all_gather(out, mine)
            h = sum(int(t.item()) for t in out) % (1 << 64)
        return h

    def roofline(self, sweep_ms, sweep_launches, peak):
        per_launch_s = (sweep_ms / max(sweep_launches, 1)) * 1e-3
        achieved = BYTES_PER_CELL_UPDATE * self.Nx * self.Nyl / per_launch_s / 1e9
        return per_launch_s, achieved, achieved / peak

    def close(self):
        self.ctx.sync()
        self.env.barrier()
        self.ctx.close()
        self.Qloc = None


class Env:
    def __init__(self):
        import torch
        import torch.distributed as dist

        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
        torch.cuda.set_device(self.local_rank)
        self.dist = dist
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.stream = torch.cuda.Stream()
        self.torch = torch
        # one process per GPU: keep the host side (threads and the pages they touch first) on the NUMA
        # node the GPU's PCIe root port belongs to
        self.numa = None
        if self.world > 1 and not os.environ.get("FV2D_NO_NUMA_BIND"):
            from fv2d_b200 import multigpu

            self.numa = multigpu.bind_to_gpu_numa_node(self.local_rank)
        os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 1) // self.world)))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world > 1:
            tmax = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
            self.dist.all_reduce(tmax, op=self.dist.ReduceOp.MAX)
            return float(tmax.item())
        return float(x)


HASH_STEPS = 16  # the state hash is taken after this many steps from the initial condition, whatever --steps / --warmup


def native_arm(args):
    import torch

    from fv2d_b200 import capi

    env = Env()
    world, rank = env.world, env.rank
    peak, peak_src = measured_peak_gbs()

    extra = {}
    if args.nx:
        extra["mesh.Nx"] = args.nx
    if args.ny:
        extra["mesh.Ny"] = args.ny
    job = Job(env, args.workload, extra)
    ctx, Nx, Ny, Nyl = job.ctx, job.Nx, job.Ny, job.Nyl

    # ---- state hash after a FIXED number of steps from the initial condition: the N-GPU result
    #      is bitwise the 1-GPU result, so this value must be the same in every --gpus N line
    ctx.run_steps(HASH_STEPS)
    state_hash = job.global_hash()

    # ---- warm-up
    ctx.run_steps(args.warmup)
    env.barrier()

    # ---- timed region: K fused steps, state resident in HBM, dt resident on the device;
    #      repeated `reps` times, the MEDIAN repetition is reported
    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    reps = []
    for k in range(max(1, args.reps)):
        if k:
            # repetitions are independent measurements: let the board's power average settle, or the
            # later ones run into the power cap the earlier ones built up (that regime is what the
            # `sustained` block below measures, on purpose)
            time.sleep(args.rep_pause)
        reps.append(job.timed(args.steps))
    clocks = sampler.stop() if rank == 0 else None
    cfl_wait_main = job.cfl_wait_us
    order = sorted(range(len(reps)), key=lambda k: reps[k][0])
    ms, sweep_ms, sweep_launches, total_launches = reps[order[len(order) // 2]]
    value = Nx * Ny * args.steps / (ms * 1e-3) / 1e6
    per_launch_s, achieved, frac = job.roofline(sweep_ms, sweep_launches, peak)

    # ---- sustained regime: one region of >= 200 back-to-back steps whatever --steps says (a short
    #      region runs at boost clocks; a production run is thousands of steps under the power cap)
    sustained = None
    if args.sustained_steps > 0:
        n_sus = max(args.sustained_steps, args.steps)
        sampler2 = ClockSampler(env.local_rank)
        if rank == 0:
            sampler2.start()
            time.sleep(0.2)
        s_ms, s_sweep_ms, s_launches, _ = job.timed(n_sus)
        s_clocks = sampler2.stop() if rank == 0 else None
        s_per_launch, s_achieved, s_frac = job.roofline(s_sweep_ms, s_launches, peak)
        sustained = {"steps": n_sus, "ms_per_step": s_ms / n_sus, "value": Nx * Ny * n_sus / (s_ms * 1e-3) / 1e6,
                     "unit": "Mcell-updates/s", "frac": s_frac, "ms_per_launch": s_per_launch * 1e3, "clocks": s_clocks}
    # ---- the same kernel with every launch bracketed by its own event pair (launches then run one
    #      after the other, no overlap of a launch's prologue with the previous one's tail)
    b_ms, b_sweep_ms, b_launches, _ = job.timed(min(args.steps, 20), bracket_launches=True)
    bracketed = {"steps": min(args.steps, 20), "ms_per_launch": b_sweep_ms / max(b_launches, 1),
                 "ms_per_step": b_ms / min(args.steps, 20), "share_of_step": b_sweep_ms / b_ms if b_ms > 0 else None}
    neg = ctx.negative_counts()

    traffic = None
    tj = ROOT / "profiles" / "sweep_traffic.json"
    if tj.exists():
        try:
            tinfo = json.loads(tj.read_text())
            if tinfo.get("workload") == args.workload and tinfo.get("Nx") == Nx and tinfo.get("Ny_local") == Nyl:
                traffic = tinfo["dram_bytes_per_launch"]
        except Exception:
            pass

    # ---- second roofline: the fp64 pipe.  Peak = measured DFMA rate of this device (dependent-DFMA
    #      chains); achieved = fp64-pipe instructions the sweep executes per cell-update (counted in the
    #      SASS of the shipped kernel, profiles/README.md) x cell-updates per second of the sweep.
    fp64 = None
    try:
        peak_dfma = capi.fp64_peak(env.local_rank)
        ach = FP64_INSTR_PER_CELL_UPDATE.get(args.workload)
        if ach:
            rate = ach * Nx * Nyl / per_launch_s
            fp64 = {"peak_dfma_per_s": peak_dfma, "peak_tflops": 2 * peak_dfma / 1e12, "fp64_instr_per_cell_update": ach,
                    "achieved_instr_per_s": rate, "frac": rate / peak_dfma,
                    "peak_source": "measured on this device: fv2d_debug_fp64_peak (dependent DFMA chains, 8 per thread)"}
    except Exception as e:  # noqa
        fp64 = {"error": str(e)}

    # ---- end to end through the host-buffer C ABI call: every step uploads the state from
    #      pinned host memory, advances one step and reads the new state + dt back
    #      (N > 1: every rank does so for its own y-slab, ghost rows included; wall clock between
    #      two barriers, maximum over the ranks)
    e2e = None
    if args.e2e_steps > 0:
        hin = torch.from_numpy(job.Qloc).pin_memory()
        hout = torch.empty_like(hin).pin_memory()
        a_in, a_out = hin.numpy(), hout.numpy()
        nbytes = int(job.Qloc.nbytes) * world

        def serial(n):
            """fv2d_advance_host: upload, computeDt, step, download - one transfer after the other"""
            nonlocal a_in, a_out
            dts = np.zeros(1)
            env.barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                ctx.advance_host(a_in, a_out, 1, dts)
                a_in, a_out = a_out, a_in
            env.barrier()
            return env.max_over_ranks(time.perf_counter() - t0)

        serial(1)  # warm-up
        if world == 1:
            # the streamed call: the state moves in row blocks, upload | sweep | download overlapped; the
            # dt each call announces for its output is the next call's hint and is verified on the way
            _, hint, _ = ctx.advance_host_stream(a_in, a_out, 0.0)  # first call of a chain: no hint
            a_in, a_out = a_out, a_in
            _, hint, _ = ctx.advance_host_stream(a_in, a_out, hint)  # warm-up of the streamed path
            a_in, a_out = a_out, a_in
            env.barrier()
            n_streamed = 0
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                _, hint, st = ctx.advance_host_stream(a_in, a_out, hint)
                n_streamed += int(st)
                a_in, a_out = a_out, a_in
            env.barrier()
            secs = time.perf_counter() - t0
            secs_serial = serial(min(args.e2e_steps, 3))
            row_bytes = 4 * 8 * (Nx + 2 * job.dev.Ng)
            up_rows = Ny + (job.dev.Ng if job.dev.boundary_y == capi.BC_PERIODIC else 0)
            e2e = {"value": Nx * Ny * args.e2e_steps / secs / 1e6, "unit": "Mcell-updates/s",
                   "h2d_bytes_per_step": row_bytes * up_rows, "d2h_bytes_per_step": nbytes + 8 * 64,
                   "steps": args.e2e_steps, "steps_streamed": n_streamed, "ms_per_step": 1e3 * secs / args.e2e_steps,
                   "api": "fv2d_advance_host_stream(ctx, hostQ_in, hostQ_out, dt_hint = dt_next of the previous call, "
                          "&dt_used, &dt_next, &streamed): pinned host arrays, one step per call, hint verified per call",
                   "serial": {"value": Nx * Ny * min(args.e2e_steps, 3) / secs_serial / 1e6, "unit": "Mcell-updates/s",
                              "api": "fv2d_advance_host(ctx, hostQ_in, hostQ_out, 1, &dt): upload, step, download in turn"}}
        else:
            secs = serial(args.e2e_steps)
            e2e = {"value": Nx * Ny * args.e2e_steps / secs / 1e6, "unit": "Mcell-updates/s",
                   "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes + 8 * world, "steps": args.e2e_steps,
                   "api": "fv2d_advance_host(ctx, hostQ_in, hostQ_out, 1, &dt) on every rank's y-slab"}
        del hin, hout, a_in, a_out
    host_init_s, slab_gb = job.host_init_s, job.Qloc.nbytes * world / 1e9
    job.close()

    # ---- north_star's scaling targets, in every line: strong scaling of Rayleigh-Taylor 16384^2
    #      (the same grid on N GPUs) and weak scaling of Kelvin-Helmholtz (8192 rows per GPU)
    def side_block(workload, extra_ov, steps, scaling):
        j = Job(env, workload, extra_ov)
        j.ctx.run_steps(max(3, args.warmup // 2))
        env.barrier()
        b_ms, b_sweep_ms, b_launches, _ = j.timed(steps)
        pl, ach, fr = j.roofline(b_sweep_ms, b_launches, peak)
        out = {"workload": workload, "Nx": j.Nx, "Ny": j.Ny, "scaling": scaling, "steps": steps,
               "value": j.Nx * j.Ny * steps / (b_ms * 1e-3) / 1e6, "unit": "Mcell-updates/s", "ms_per_step": b_ms / steps,
               "frac": fr, "ms_per_launch": pl * 1e3, "cfl_mail_wait_us_per_step": j.cfl_wait_us}
        j.close()
        return out

    strong_16384 = weak = None
    if not args.no_scaling_blocks:
        strong_16384 = side_block("rayleigh_taylor_16384_plm_hllc", {}, args.side_steps, "strong")
        if world == 1 and args.workload == "kelvin_helmholtz_8192_plm_hllc" and not extra:
            weak = {"workload": args.workload, "Nx": Nx, "Ny": Ny, "rows_per_gpu": Ny, "scaling": "weak", "steps": args.steps,
                    "value": value, "unit": "Mcell-updates/s", "ms_per_step": ms / args.steps, "frac": frac,
                    "ms_per_launch": per_launch_s * 1e3, "note": "N=1: the headline run itself"}
        else:
            weak = side_block("kelvin_helmholtz_8192_plm_hllc", {"mesh.Ny": 8192 * world}, max(args.side_steps, 40), "weak")
            weak["rows_per_gpu"] = 8192

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference itself, same grid, fewer steps
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = run_reference_binary(args.workload, extra, args.cpu_steps, 1)
        cpu = {"value": r["value"], "unit": "Mcell-updates/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"{args.workload} at its full size ({Nx}x{Ny}), {r['steps']} timed steps after 1 warm-up "
                         f"({r['seconds']:.1f} s of CPU work)"}

    if rank == 0:
        line = {
            "metric": "Mcell-updates/s", "value": value, "unit": "Mcell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "Nx": Nx, "Ny": Ny, "riemann_solver": "hllc",
                       "reconstruction": {0: "pcm", 1: "pcm_wb", 2: "plm"}[job.dev.reconstruction],
                       "time_stepping": "euler" if job.run.time_stepping == 0 else "rk2",
                       "decomposition": f"{world} y-slab(s)", "host_numa_binding": env.numa, "l2_policy": "working set (3 arrays x %.2f GB) >> 126 MB L2"
                       % slab_gb, "host_init_s": round(host_init_s, 2),
                       "repetitions": len(reps), "value_is": "median repetition", "idle_between_repetitions_s": args.rep_pause,
                       "ms_per_step_all_repetitions": [r[0] / args.steps for r in reps]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": frac,
                         "traffic": traffic, "kernel": "k_sweep (fused RK stage, persistent)", "peak_source": peak_src,
                         "bytes_per_cell_update": BYTES_PER_CELL_UPDATE, "ms_per_launch": per_launch_s * 1e3,
                         "launch_duration_is": "timed region / launches in it (one kernel per step, launched back to back "
                                               "with programmatic stream serialization)",
                         "event_bracketed_launches": bracketed, "share_of_step": bracketed["share_of_step"], "fp64": fp64,
                         "cfl_mail_wait_us_per_step": cfl_wait_main},
            "sustained": sustained, "state_hash": {"after_steps": HASH_STEPS, "u64": f"0x{state_hash:016x}"},
            "strong_16384": strong_16384, "weak": weak,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(total_launches), "clocks": clocks,
            "sanity": {"negative_density": neg[0], "negative_pressure": neg[1], "nan": neg[2]},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        env.dist.destroy_process_group()
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
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--reps", type=int, default=3, help="repetitions of the timed region; the median is reported")
    ap.add_argument("--rep-pause", type=float, default=1.0, help="idle seconds between repetitions of the timed region")
    ap.add_argument("--sustained-steps", type=int, default=200, help="length of the sustained-regime region (0: skip)")
    ap.add_argument("--side-steps", type=int, default=20, help="timed steps of the strong_16384 / weak blocks")
    ap.add_argument("--no-scaling-blocks", action="store_true", help="skip the strong_16384 / weak blocks")
    ap.add_argument("--cpu-steps", type=int, default=6, help="timed steps of the in-line CPU baseline (full grid)")
    ap.add_argument("--ref-max-steps", type=int, default=40, help="cap on the timed steps of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    return reference_arm(args) if args.impl == "reference" else native_arm(args)


if __name__ == "__main__":
    sys.exit(main())
