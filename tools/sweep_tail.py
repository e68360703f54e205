#!/usr/bin/env python
"""One process, several settings of the V-cycle tail (narrow / wide thresholds), several workloads: ms per step with
CUDA events, V-cycles per step, and a checksum of the prognostic state after the same number of steps -- the settings
change which kernels run the small levels, never the arithmetic, so the checksums of a workload must all be equal.

    python tools/sweep_tail.py [--workloads lock,tgv256,weak512] [--combos 4096:0,2048:150000,2048:300000] [--steps 10]

A combo is tail_cells:wide_cells (ny_mg_set_tail_cells / ny_mg_set_wide_cells; wide 0 = levels above tail_cells run
the fused plane-marching legs).  Prints one JSON line per (workload, combo)."""
import argparse
import gc
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402


def build(wname):
    from nyles_b200 import nyles, parameters
    w = bench.workload(wname, 1)
    parameters.InextensibleDict.unfreeze()
    up = parameters.UserParameters()
    up.model["modelname"] = w["modelname"]
    up.model["geometry"] = w["geometry"]
    up.model["Lx"], up.model["Ly"], up.model["Lz"] = w["nx"] * w["dx"], w["ny"] * w["dx"], w["nz"] * w["dx"]
    up.discretization["global_nx"], up.discretization["global_ny"], up.discretization["global_nz"] = \
        w["nx"], w["ny"], w["nz"]
    up.time["cfl"], up.time["dt_max"] = w["cfl"], w["dt_max"]
    for k in ("rotating", "coriolis", "forced"):
        if k in w:
            up.physics[k] = w[k]
    up.IO["datadir"] = ""
    ny = nyles.Nyles(up)
    model, g = ny.model, ny.grid
    if w.get("forced"):
        model.forcing = bench.PlumeForcing(w, g.x_b_1D, g.y_b_1D, g.z_b_1D, device=model.state.b.tensor.device)
    b, u, v = bench.initial_condition(w, g.x_b_1D, g.y_b_1D, g.z_b_1D, 0)
    if b is not None:
        model.state.b.view("i")[:] = b
    if u is not None:
        model.state.u["i"].view("i")[:] = u
        model.state.u["j"].view("i")[:] = v
    return w, ny


def main():
    import torch
    from nyles_b200 import lib
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="lock,tgv256,weak512")
    ap.add_argument("--combos", default="4096:0,2048:150000,2048:300000")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    L = lib.load()
    for wname in args.workloads.split(","):
        sums = []
        for combo in args.combos.split(","):
            tail, wide = (int(v) for v in combo.split(":"))
            L.ny_mg_set_tail_cells(tail)
            L.ny_mg_set_wide_cells(wide)
            try:
                w, ny = build(wname)
            finally:
                L.ny_mg_set_tail_cells(2048)
                L.ny_mg_set_wide_cells(300000)
            model = ny.model
            model.diagnose_var(model.state)
            steps = args.steps * (20 if w["nx"] * w["ny"] * w["nz"] < (1 << 20) else 1)
            t = 0.0
            for _ in range(1 + args.warmup):
                dt = ny.compute_dt()
                model.forward(t, dt)
                t += dt
            torch.cuda.synchronize()
            lib.launch_count_reset()
            vc0 = model.mg.nvcycles
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                dt = ny.compute_dt()
                model.forward(t, dt)
                t += dt
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            h = hashlib.sha256()
            for tn in ny.prognostic_tensors():
                h.update(tn.cpu().numpy().tobytes())
            sums.append(h.hexdigest()[:16])
            print(json.dumps({"workload": wname, "tail_cells": tail, "wide_cells": wide, "ms_per_step": round(ms, 4),
                              "steps": steps, "vcycles_per_step": (model.mg.nvcycles - vc0) / steps,
                              "launches_per_step": lib.launch_count() / steps, "t_end": t, "state_sha256": sums[-1]}),
                  flush=True)
            del ny, model
            gc.collect()
            torch.cuda.empty_cache()
        print(json.dumps({"workload": wname, "all_settings_bit_identical": len(set(sums)) == 1}), flush=True)


if __name__ == "__main__":
    main()
