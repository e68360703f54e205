#!/usr/bin/env python
"""Benchmark of the Nyles LES time step (RHS + multigrid projection) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Metric (BASELINE.json): LES cell-updates/s, fp64, one "step" = one steady-state LFAM3 step
(2 right-hand sides + 2 pressure projections) of the whole domain.  N > 1 is launched by
torchrun, one rank per GPU; the domain is cut into z slabs (weak scaling: 512^3 cells per GPU).

The JSON line follows the driver contract; extra keys:
  roofline      dominant kernel family: algorithmic bytes per launch group / event-timed duration
  roofline_step whole step against the operator-level byte model of SURVEY.md 8(d)
  cpu_baseline  the CPU restatement of the reference (oracle/, OpenMP build) on the host cores
  e2e           the same metric through Nyles.step_host(): state in pinned HOST buffers, uploaded
                and downloaded inside every timed step
`--impl reference` times the CPU restatement alone (the reference's Fortran cannot be built in
this image: no gfortran / MPI).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "LES cell-updates/s (fp64, RHS+projection)"
UNIT = "cell-updates/s"

# ---- workloads ---------------------------------------------------------------
# weak scaling, 512^3 cells per GPU, global (nx, ny, nz).  Default: the near-cubic boxes of SURVEY.md 8(d);
# the slab interfaces grow from 512^2 to 1024^2 at 8 GPUs.  "weakz": the box grows along z only (constant
# 512^2 interfaces) -- measured and rejected: mgfor stops coarsening when nx or ny reaches 2, so a 512x512x4096
# box is left with a 2x2x16 coarsest grid that is only smoothed and the solves run into maxite (39 V-cycles
# per step instead of 8, 177 ms per step at 8 GPUs).
WEAK_SHAPES = {1: (512, 512, 512), 2: (512, 512, 1024), 4: (512, 1024, 1024), 8: (1024, 1024, 1024)}
TALL_SHAPES = {1: (512, 512, 512), 2: (512, 512, 1024), 4: (512, 512, 2048), 8: (512, 512, 4096)}


def workload(name, n_gpus):
    if name in ("weak512", "weakz"):      # configs[4] (= configs[2] at N=1): RT instability, closed box, LES
        nx, ny, nz = (WEAK_SHAPES if name == "weak512" else TALL_SHAPES)[n_gpus]
        return dict(name="rayleigh-taylor 512^3 per GPU (weak-scaling sweep%s), closed, LES, LFAM3"
                         % ("" if name == "weak512" else ", box extended along z"),
                    nx=nx, ny=ny, nz=nz, dx=0.25, geometry="closed", modelname="LES", ic="rt",
                    cfl=0.8, dt_max=0.1)
    if name.startswith("tgv"):     # tgv256 = configs[1]
        n = int(name[3:])
        return dict(name="taylor-green vortex %d^3, perio_xyz, Euler3d, LFAM3" % n, nx=n, ny=n, nz=n * n_gpus,
                    dx=2 * np.pi / n, geometry="perio_xyz", modelname="Euler3d", ic="tgv", cfl=0.8, dt_max=0.4 * 32 / n)
    if name == "lock":             # configs[0]
        return dict(name="lock-exchange 128x32x32, closed, LES, LFAM3", nx=128, ny=32, nz=32 * n_gpus, dx=0.25,
                    geometry="closed", modelname="LES", ic="lock", cfl=0.8, dt_max=0.1)
    if name == "rt512strong":      # configs[2] on 1 and 2 B200: ONE 512^3 box cut into z slabs (strong scaling)
        return dict(name="rayleigh-taylor 512^3 (one box, %d z-slabs), closed, LES, LFAM3" % n_gpus, nx=512, ny=512,
                    nz=512, dx=0.25, geometry="closed", modelname="LES", ic="rt", cfl=0.8, dt_max=0.1,
                    scaling="strong")
    if name in ("plume", "plume256"):
        # configs[3]: experiments/forced_convection/forced_plume.py:14-80 scaled to 1024x1024x512 (L = 16x16x8),
        # closed, LES, rotating f = 1, Gaussian-column heat source, linear stratification, cfl 0.8, dt_max 0.8;
        # the box is fixed and cut into N z-slabs.  plume256: the same set-up at 256x256x128 (tests, one GPU)
        f = 1 if name == "plume" else 4
        return dict(name="turbulent plume %dx%dx%d (%d z-slabs), closed, LES, rotating + forced, LFAM3"
                         % (1024 // f, 1024 // f, 512 // f, n_gpus), nx=1024 // f, ny=1024 // f, nz=512 // f,
                    dx=16.0 / (1024 // f), geometry="closed", modelname="LES", ic="plume", cfl=0.8, dt_max=0.8,
                    rotating=True, coriolis=1.0, forced=True, scaling="strong")
    if name.startswith("rt"):      # rtN: cubic RT box of N^3 per GPU (used for the CPU sample)
        n = int(name[2:])
        return dict(name="rayleigh-taylor %d^3, closed, LES, LFAM3" % n, nx=n, ny=n, nz=n * n_gpus, dx=0.25,
                    geometry="closed", modelname="LES", ic="rt", cfl=0.8, dt_max=0.1)
    raise SystemExit("unknown workload %r" % name)


def initial_condition(w, x, y, z, rank):
    """b, u, v (canonical (k,j,i) arrays or None) from 1-D cell-centre coordinates of this rank."""
    shape = (len(z), len(y), len(x))
    Lz = w["nz"] * w["dx"]
    dx = w["dx"]
    rng = np.random.default_rng(1234 + rank)
    if w["ic"] == "rt":            # experiments/rayleightaylor/RT.py:69-71
        noise = 0.01 * rng.standard_normal(shape)
        noise += (0.5 * Lz - z)[:, None, None]
        noise /= dx
        return np.tanh(noise, out=noise), None, None
    if w["ic"] == "lock":          # experiments/lockechange/lockexchange.py:65-72
        Lx = w["nx"] * dx
        noise = 0.1 * rng.standard_normal(shape)
        return np.tanh((x[None, None, :] + noise - 0.25 * Lx) / (2 * dx)), None, None
    if w["ic"] == "tgv":           # experiments/taylorgreen/tgv.py:68-84 (covariant: times dx)
        X, Y, Z = x[None, None, :], y[None, :, None], z[:, None, None]     # cell centres, as tgv.py does
        u = np.sin(X + 1.2) * np.cos(Y + 1.8) * np.cos(Z + 0.5) * dx
        v = -np.cos(X + 1.2) * np.sin(Y + 1.8) * np.cos(Z + 0.5) * dx
        return None, u, v
    if w["ic"] == "plume":         # forced_plume.py:94-95: linear stratification b = 0.1 (z/Lz - 0.5), fluid at rest
        return np.broadcast_to((0.1 * (z / Lz - 0.5))[:, None, None], shape).copy(), None, None
    raise ValueError(w["ic"])


class PlumeForcing(object):
    """The user forcing object of experiments/forced_convection/forced_plume.py:68-80: a heat source
    Q = 0.1 exp(-z/0.02) (1 - tanh(d/0.1))/2 in a column around the axis of the box (x, y, z scaled by the box
    lengths), added to db after the right-hand side.  `add` is the reference's protocol; `device_tendencies`
    (nyles_b200) hands the same array to the fused RHS + time-scheme launches."""

    def __init__(self, w, x, y, z, device=None):
        X = x[None, None, :] / (w["nx"] * w["dx"]) - 0.5
        Y = y[None, :, None] / (w["ny"] * w["dx"]) - 0.5
        Z = z[:, None, None] / (w["nz"] * w["dx"])
        msk = 0.5 * (1. - np.tanh(np.sqrt(X ** 2 + Y ** 2) / 0.1))
        self.Q = 1e-1 * np.exp(-Z / 0.02) * msk
        if device is not None:
            import torch
            self.Q = torch.as_tensor(self.Q, dtype=torch.float64).to(device)

    def add(self, state, dstate, time):
        db = dstate.b.view("i")
        db += self.Q

    def device_tendencies(self, state, time):
        return {"b": self.Q}


def bytes_per_cell_step(w, n_vc):
    """Operator-level compulsory HBM bytes per cell per LFAM3 step."""
    return (984.0 if w["modelname"] == "Euler3d" else 1080.0) + 199.0 * n_vc


# ---- clocks --------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU restatement of the reference (oracle) ------------------------------------------------
def cpu_reference_run(w, steps, warmup):
    """Time `steps` LFAM3 steps of the OpenMP build of the oracle on this host.  Returns
    (cell-updates/s, ms per step, threads, v-cycles per step)."""
    from oracle import model as M
    # all host cores (torchrun exports OMP_NUM_THREADS=1 to its workers; the baseline is not a worker)
    threads = int(os.environ.get("NY_CPU_THREADS", os.cpu_count() or 1))
    os.environ["OMP_NUM_THREADS"] = str(threads)
    p = M.make_param(nx=w["nx"], ny=w["ny"], nz=w["nz"], geometry=w["geometry"], Lx=w["nx"] * w["dx"],
                     Ly=w["ny"] * w["dx"], Lz=w["nz"] * w["dx"], modelname=w["modelname"], cfl=w["cfl"],
                     dt_max=w["dt_max"], **{k: w[k] for k in ("rotating", "coriolis", "forced") if k in w})
    o = M.LES(p, flavour="fast")
    g = o.grid
    if w.get("forced"):
        o.forcing = PlumeForcing(w, g.x_b_1D, g.y_b_1D, g.z_b_1D)
    b, u, v = initial_condition(w, g.x_b_1D, g.y_b_1D, g.z_b_1D, 0)
    if b is not None:
        o.state.b.view("i")[:] = b
    if u is not None:
        o.state.u["i"].view("i")[:] = u
        o.state.u["j"].view("i")[:] = v
    o.diagnose_var(o.state)
    t = 0.0
    for _ in range(1 + warmup):                    # Euler start-up step + warm-up
        dt = o.compute_dt(); o.forward(t, dt); t += dt
    nlog = len(o.mg_log)
    t0 = time.perf_counter()
    for _ in range(steps):
        dt = o.compute_dt(); o.forward(t, dt); t += dt
    wall = time.perf_counter() - t0
    n_vc = sum(m[0] for m in o.mg_log[nlog:]) / float(steps)
    cells = w["nx"] * w["ny"] * w["nz"]
    return cells * steps / wall, 1e3 * wall / steps, threads, n_vc


def cpu_model_string():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_sample(args):
    """The bounded sample of the workload that the CPU restatement is timed on (same IC, geometry and model on a
    smaller box, rate quoted per cell), or -- `--cpu-sample same` -- the workload itself."""
    name = args.cpu_sample
    if name == "same":
        return workload(args.workload, args.gpus), True
    if name == "auto":
        wl = args.workload
        name = "plume256" if wl.startswith("plume") else "lock" if wl == "lock" else \
            "tgv128" if wl.startswith("tgv") else "rt256"
    return workload(name, 1), name == args.workload and args.gpus == 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample, is_workload = cpu_sample(args)
    w = workload(args.workload, args.gpus)
    value, ms, threads, n_vc = cpu_reference_run(sample, args.steps, args.warmup)
    desc = ("%s: %s, %d LFAM3 steps after 1 Euler + %d warm-up steps; rate is per cell"
            % (sample["name"], "the workload itself" if is_workload else
               "same IC/geometry/model as the workload on a %dx%dx%d sub-box" % (sample["nx"], sample["ny"], sample["nz"]),
               args.steps, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "sample": desc, "sample_is_the_workload": is_workload,
                       "vcycles_per_step": n_vc, "cpu": cpu_model_string()},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---- the B200 arm ---------------------------------------------------------------------------
# algorithmic bytes per cell of one timed launch group of each kernel family ("Kernels")
def family_bytes_per_cell(euler):
    return {
        # the RHS launches apply the time-scheme update themselves and write each field once (ny_rhs_step):
        # tracer: b, U x3 read + the other time level (bb in the predictor, bn in the corrector) -> new b
        "rhs_tracer": 48.0,
        # momentum: U x3, vor x3, ke (, b) read; predictor reads u, ub (6 arrays), corrector reads un (3): mean 4.5
        # arrays per launch; writes the new u (3 arrays)
        "rhs_momentum": (56.0 if euler else 64.0) + 36.0 + 24.0,
        "vorticity_ke": 40.0,                     # vorticity: u x3 -> vor x3 (48); ke: u x3 -> ke (32); mean per group
        "div": 32.0, "gradp": 56.0, "U_from_u": 48.0,
        "timescheme": 36.0,                       # predictor 48, corrector 24 per field; mean per group
        "maxspeed": 24.0,
        "mg_smooth_fine": 24.0,                   # x, b -> x (two Jacobi sweeps fused in one pass)
        "mg_residual_fine": 24.0,                 # x, b -> r
        "mg_restrict_fine": 9.0, "mg_prolong_fine": 17.0, "mg_norm": 8.0,
        "mg_down_fine": 25.0,                     # smooth + residual + restriction fused: x, b -> x, b_coarse
        "mg_up_fine": 25.0,                       # prolongation + smooth + residual norm fused: x, b, x_coarse -> x
        "mg_embed_extract": 16.0,
    }


# dram bytes and executed fp64 instructions per launch group, from the `ncu --set full` capture of this same command
# (tools/ncu_traffic.py writes the map; the capture itself is never a bench value).  The map carries a hash of the
# kernel sources it was taken from: numbers of another build are reported as stale, not silently quoted.
def ncu_facts():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        from ncu_traffic import source_stamp
        stamp = source_stamp()
    except Exception:                     # noqa: BLE001
        stamp = None
    for path in (os.path.join(ROOT, "gpurun_out", "ncu_traffic.json"), os.path.join(ROOT, "profiles", "ncu_traffic.json")):
        try:
            with open(path) as f:
                d = json.load(f)
        except (OSError, ValueError):
            continue
        d["stale"] = not (stamp is not None and d.get("stamp") == stamp)
        d["path"] = os.path.relpath(path, ROOT)
        return d
    return {"families": {}, "fp64": {}, "stale": True, "path": None}


def comm_stats(steps):
    """Slab exchanges and bytes pushed to the neighbours over NVLink by this rank, per step (ny_comm_stats)."""
    from nyles_b200 import comm, lib
    if not comm.active():
        return None
    import ctypes as C
    n, b = C.c_longlong(), C.c_longlong()
    lib.check(lib.load().ny_comm_stats(comm.get(), C.byref(n), C.byref(b), 1))
    return {"exchanges_per_step": n.value / float(steps), "bytes_sent_per_step_per_rank": b.value / float(steps)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import nyles_b200
    from nyles_b200 import lib, nyles, parameters
    nyles_b200.FAST_ARITH = bool(args.fast_arith)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py --gpus %d must run under torchrun with %d ranks (found WORLD_SIZE=%d)"
                         % (args.gpus, args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device is visible (there is no CPU fallback)")
    torch.cuda.set_device(local)
    parity = None
    if world > 1:
        # N-rank correctness travels with the scaling record: before anything is timed, three small problems
        # (closed LES, perio_xyz Euler3d, perio_xy rotating LES; 64 x 64 x 64N) are stepped on ONE GPU by every
        # rank and then on N z-slabs -- fields, dt and V-cycle counts must be bit-equal (nyles_b200/selfcheck.py)
        from nyles_b200 import selfcheck
        ref_runs = None if args.no_selfcheck else selfcheck.single_gpu_runs(world)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if ref_runs is not None:
            ok, report = selfcheck.slab_runs(ref_runs)
            parity = ("ok" if ok else "FAILED", report)
            del ref_runs
            if not ok and rank == 0:
                print("bench.py: slab self-check FAILED: %r" % (report,), file=sys.stderr)

    w = workload(args.workload, args.gpus)
    parameters.InextensibleDict.unfreeze()
    up = parameters.UserParameters()
    up.model["modelname"] = w["modelname"]
    up.model["geometry"] = w["geometry"]
    up.model["Lx"], up.model["Ly"], up.model["Lz"] = w["nx"] * w["dx"], w["ny"] * w["dx"], w["nz"] * w["dx"]
    up.discretization["global_nx"], up.discretization["global_ny"], up.discretization["global_nz"] = \
        w["nx"], w["ny"], w["nz"]
    up.MPI["npz"] = args.gpus
    up.time["cfl"], up.time["dt_max"] = w["cfl"], w["dt_max"]
    for k in ("rotating", "coriolis", "forced"):
        if k in w:
            up.physics[k] = w[k]
    up.IO["datadir"] = ""                                   # no history output inside the benchmark
    ny = nyles.Nyles(up)
    model, g = ny.model, ny.grid
    if w.get("forced"):                                     # "the user must attach the forcing to the model"
        model.forcing = PlumeForcing(w, g.x_b_1D, g.y_b_1D, g.z_b_1D, device=model.state.b.tensor.device)
        if args.unfused_forcing:                            # the reference's protocol only: RHS -> dstate -> add -> ts
            model.forcing.device_tendencies = None
    b, u, v = initial_condition(w, g.x_b_1D, g.y_b_1D, g.z_b_1D, rank)
    if b is not None:
        model.state.b.view("i")[:] = b
    if u is not None:
        model.state.u["i"].view("i")[:] = u
        model.state.u["j"].view("i")[:] = v
    del b, u, v
    model.diagnose_var(model.state)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(t):
        dt = ny.compute_dt()
        model.forward(t, dt)
        return t + dt

    t = 0.0
    t = step(t)                                            # Euler start-up step of LFAM3 (not steady state)
    for _ in range(args.warmup):
        t = step(t)

    # ---- timed region 1: state resident in HBM ----------------------------------------------
    cells = w["nx"] * w["ny"] * w["nz"]
    euler = w["modelname"] == "Euler3d"
    barrier()
    lib.launch_count_reset()
    comm_stats(1)                                          # reads and resets the exchange counters
    lib.prof_start()
    vc0 = model.mg.nvcycles
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        t = step(t)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.launch_count()
    prof = lib.prof_collect()
    lib.prof_start(0)
    n_vc = (model.mg.nvcycles - vc0) / float(args.steps)
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()
    value = cells * args.steps / (ms * 1e-3)
    link = comm_stats(args.steps)

    # ---- timed region 2: through the host-buffer API (e2e) -----------------------------------
    host = ny.allocate_host_state()
    for h, d in zip(host, ny.prognostic_tensors()):
        h.copy_(d)
    barrier()
    e2_steps = max(1, args.e2e_steps)                       # its own count: 10 by default, whatever --steps says
    t = ny.step_host(t, host) + t                          # warm the pinned path once
    barrier()
    e0.record()
    for _ in range(e2_steps):
        t += ny.step_host(t, host)
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms2], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms2 = tt.item()
    state_bytes = sum(h.numel() * 8 for h in host) * world
    e2e = {"value": cells * e2_steps / (ms2 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": state_bytes,
           "d2h_bytes_per_step": state_bytes + 8, "steps": e2_steps,
           "api": "nyles_b200.nyles.Nyles.step_host (prognostic state in pinned host memory)"}
    del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family ------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except (OSError, ValueError):
        pass
    peak_gbs, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy)") if "hbm_gbs" in peaks \
        else (6650.0, "fallback of B200_PROFILING.md")
    fam = family_bytes_per_cell(euler)
    facts = ncu_facts()
    fp64_peak = None
    try:                                   # ~0.5 s of DFMA-only launches: what the fp64 pipe sustains on this board
        out = lib.C.c_double()
        lib.check(lib.load().ny_debug_fp64_peak(lib.context(), 0.5, lib.C.byref(out), lib.stream()))
        fp64_peak = out.value * 32.0        # thread-level instructions per second
    except Exception as exc:               # noqa: BLE001
        print("fp64 peak measurement failed: %r" % (exc,), file=sys.stderr)
    local_cells = cells / world
    timed = {k: v for k, v in prof.items() if v[1] > 0}
    total_ms = sum(v[0] for v in timed.values())
    top = max((k for k in timed if k in fam), key=lambda k: timed[k][0], default=None)
    shares = {k: round(v[0] / (ms), 4) for k, v in sorted(timed.items(), key=lambda kv: -kv[1][0])}
    if top is not None:
        t_ms, n = timed[top]
        achieved = fam[top] * local_cells / (t_ms / n * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                    "frac": achieved / peak_gbs, "traffic": None if facts["stale"] else facts.get("families", {}).get(top),
                    "traffic_source": ("%s%s" % (facts["path"], " (STALE: taken from other kernel sources)" if facts["stale"] else ""))
                    if facts["path"] else None,
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_cell": fam[top], "launch_groups": n, "avg_ms": t_ms / n,
                    "share_of_step": t_ms / ms}
        f64 = facts.get("fp64", {}).get(top)
        if f64 and not facts["stale"] and not args.fast_arith:
            # the WENO kernels are bound by the fp64 pipe (and by the power it draws), not by HBM: executed fp64
            # instructions per cell from the ncu capture against the DFMA-only rate this board sustains under its
            # power cap (ny_debug_fp64_peak, measured below) and against 64 lanes per SM per clock at the maximum clock
            rate = f64["dp_instr_per_cell"] * local_cells / (t_ms / n * 1e-3)
            roofline["fp64_pipe"] = {"dp_instr_per_cell": f64["dp_instr_per_cell"], "achieved_instr_per_s": rate,
                                     "peak_instr_per_s_sustained_measured": fp64_peak,
                                     "frac_of_sustained_measured": rate / fp64_peak if fp64_peak else None,
                                     "peak_instr_per_s_nominal": 148 * 64 * 1.965e9, "frac_of_nominal": rate / (148 * 64 * 1.965e9),
                                     "ncu_pipe_fp64_pct_isolated_launch": f64["ncu_pipe_fp64_pct"],
                                     "note": "actual limiter of this kernel (with the 1 kW power cap); the hbm fraction "
                                             "above is low by construction"}
    else:
        roofline = None
    B = bytes_per_cell_step(w, n_vc)
    step_gbs = B * value / world / 1e9
    roofline_step = {"bytes_per_cell_step": B, "vcycles_per_step": n_vc, "achieved": step_gbs, "unit": "GB/s per GPU",
                     "frac_of_measured": step_gbs / peak_gbs, "frac_of_nominal_8TBs": step_gbs / 8000.0,
                     "kernel_time_share": shares, "event_timed_ms_per_step": total_ms / args.steps}

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        sample, _ = cpu_sample(args)
        v, cms, threads, cvc = cpu_reference_run(sample, args.cpu_steps, 1)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%s, %d LFAM3 steps timed after 1 Euler + 1 warm-up step (%.0f ms/step, %.1f V-cycles/step); "
                         "OpenMP -O3 -march=native build of the C restatement of the Fortran path, %s"
                         % (sample["name"], args.cpu_steps, cms, cvc, cpu_model_string())}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": w.get("scaling", "weak"),
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "grid": [w["nx"], w["ny"], w["nz"]], "cells": cells,
                       "parallelism": "z-slabs x%d" % world, "timestepping": "LFAM3",
                       "l2": "inputs larger than L2 (every field is %.0f MB per GPU)" % (local_cells * 8 / 1e6),
                       "vcycles_per_step": n_vc,
                       "arithmetic": "fast (re-associated weno5, <=1e-12 per RHS)" if args.fast_arith
                       else "strict (source order, no FMA: bit-identical to the reference restatement)"},
            "roofline": roofline, "roofline_step": roofline_step, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks}
    if parity is not None:
        line["parity_check"] = parity[0]
        line["parity_check_detail"] = parity[1]
    if link is not None:
        line["nvlink"] = link
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    # the contract is ONE JSON line on stdout; libraries (NCCL's version banner) also write there, so the
    # real stdout is kept aside for that line and fd 1 points to stderr for everything else
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="weak512")
    # 256^3: 10-30 s of work for the 16 host threads of the GPU box, and (unlike 128^3) far out of the CPU's caches
    ap.add_argument("--cpu-sample", default="auto",
                    help="CPU sample: auto (a smaller box of the workload), a workload name, or `same`")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-selfcheck", action="store_true", help="N > 1: skip the slabs-vs-one-GPU bit-equality check")
    ap.add_argument("--unfused-forcing", action="store_true",
                    help="plume: hide device_tendencies, i.e. run the reference's forcing.add protocol (unfused RHS)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fast-arith", action="store_true", help="re-associated weno5 (see include/nyles_b200.h)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print("note: fewer than 3 warm-up steps requested", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
