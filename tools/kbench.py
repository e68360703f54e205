#!/usr/bin/env python
"""Kernel micro-benchmarks on one GPU (CUDA events, inputs larger than L2): RHS kernels on smooth
random fields and on the still/sharp fields of the benchmark's early steps, and every multigrid
level-1 operation.  Development tool; bench.py is the judged measurement."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nyles_b200 import lib  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


class Clocks(object):
    """SM clock / power sampled through NVML while a kernel loop runs."""
    def __init__(self):
        import threading
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        self.rows, self.stop = [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            self.rows.append((self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM),
                              self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            self.stop.wait(0.02)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join()

    def summary(self):
        r = self.rows[len(self.rows) // 3:] or self.rows
        return "sm %.0f MHz (min %.0f), %.0f W" % (np.median([x[0] for x in r]), min(x[0] for x in r), max(x[1] for x in r))


def sustained(fn, seconds=1.5):
    """ms per call and clocks over a loop of about `seconds`."""
    t1 = timeit(fn, n=3, warm=1)
    n = max(5, int(seconds * 1e3 / t1))
    with Clocks() as c:
        t = timeit(fn, n=n, warm=0)
    return t, c.summary()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--fast", type=int, default=0)
    ap.add_argument("--what", default="rhs,mg")
    ap.add_argument("--sustained", type=int, default=0)
    ap.add_argument("--mom", type=int, default=0, help="momentum kernel variant (ny_set_momentum_variant)")
    a = ap.parse_args()
    n = a.n
    L = lib.load()
    ctx = lib.context()
    lib.set_arith(bool(a.fast))
    lib.check(L.ny_set_momentum_variant(ctx, a.mom))
    dev = "cuda"
    cells = n ** 3
    st = lib.stream()
    gen = torch.Generator(device=dev).manual_seed(1)

    def smooth_field(amp=1.0):
        z = torch.linspace(0, 6.28, n, device=dev, dtype=torch.float64)
        f = (torch.sin(3 * z)[:, None, None] * torch.cos(2 * z)[None, :, None] * torch.sin(5 * z + 1)[None, None, :])
        return (amp * (f + 0.05 * torch.randn((n, n, n), device=dev, dtype=torch.float64, generator=gen))).contiguous()

    if "rhs" in a.what:
        for label in ("turbulent", "still"):
            if label == "turbulent":
                F = [smooth_field() for _ in range(8)]
            else:
                F = [torch.zeros((n, n, n), device=dev, dtype=torch.float64) for _ in range(8)]
                zz = torch.arange(n, device=dev, dtype=torch.float64)
                F[0] = torch.tanh((0.5 * n - zz - 0.5)[:, None, None] + 0.01 * torch.randn((n, n, n), device=dev, dtype=torch.float64, generator=gen)).contiguous()
                for q in range(1, 7):
                    F[q] = 1e-7 * torch.randn((n, n, n), device=dev, dtype=torch.float64, generator=gen)
            b, Ux, Uy, Uz, wx, wy, wz, ke = F
            out = [torch.empty_like(b) for _ in range(4)]
            e = lib.ext(b)
            t_tr = timeit(lambda: lib.check(L.ny_upwind(ctx, lib.ptr(b), lib.ptr(Ux), lib.ptr(Uy), lib.ptr(Uz), lib.ptr(out[0]), e, st)))
            t_rhs = timeit(lambda: lib.check(L.ny_rhs(ctx, lib.ptr(b), lib.ptr(Ux), lib.ptr(Uy), lib.ptr(Uz), lib.ptr(wx), lib.ptr(wy),
                                                      lib.ptr(wz), lib.ptr(ke), lib.ptr(out[0]), lib.ptr(out[1]), lib.ptr(out[2]),
                                                      lib.ptr(out[3]), 0.25, 0, e, st)))
            print("rhs %-9s fast=%d mom=%d n=%d: tracer %.3f ms (%.0f GB/s alg), momentum %.3f ms (%.0f GB/s alg)"
                  % (label, a.fast, a.mom, n, t_tr, 40 * cells / t_tr / 1e6, t_rhs - t_tr, 88 * cells / (t_rhs - t_tr) / 1e6), flush=True)
            if a.sustained:
                t, c = sustained(lambda: lib.check(L.ny_rhs(ctx, lib.ptr(b), lib.ptr(Ux), lib.ptr(Uy), lib.ptr(Uz), lib.ptr(wx), lib.ptr(wy),
                                                            lib.ptr(wz), lib.ptr(ke), lib.ptr(out[0]), lib.ptr(out[1]), lib.ptr(out[2]),
                                                            lib.ptr(out[3]), 0.25, 1, e, st)))
                print("   sustained momentum alone (Euler flag: no tracer launch) %.3f ms, %s" % (t, c), flush=True)
                lib.prof_start()
                for _ in range(5):
                    lib.check(L.ny_rhs(ctx, lib.ptr(b), lib.ptr(Ux), lib.ptr(Uy), lib.ptr(Uz), lib.ptr(wx), lib.ptr(wy),
                                       lib.ptr(wz), lib.ptr(ke), lib.ptr(out[0]), lib.ptr(out[1]), lib.ptr(out[2]),
                                       lib.ptr(out[3]), 0.25, 0, e, st))
                prof = lib.prof_collect()
                lib.prof_start(0)
                print("   event-timed families: " + ", ".join("%s %.3f ms" % (k, v[0] / 5) for k, v in prof.items() if v[1]), flush=True)
                t, c = sustained(lambda: lib.check(L.ny_upwind(ctx, lib.ptr(b), lib.ptr(Ux), lib.ptr(Uy), lib.ptr(Uz), lib.ptr(out[0]), e, st)))
                print("   sustained tracer %.3f ms, %s" % (t, c), flush=True)
                t, c = sustained(lambda: lib.check(L.ny_vortex_force(ctx, lib.ptr(Ux), lib.ptr(Uy), lib.ptr(Uz), lib.ptr(wx), lib.ptr(wy),
                                                                     lib.ptr(wz), lib.ptr(out[1]), lib.ptr(out[2]), lib.ptr(out[3]), e, st)))
                print("   sustained vortex force %.3f ms, %s" % (t, c), flush=True)
            del F, out, b, Ux, Uy, Uz, wx, wy, wz, ke
    if "tail" in a.what:                       # one-launch V-cycle tail: which levels should it take?
        from nyles_b200.mgfordriver import MG
        for cells in (0, 512, 4096, 32768):
            L.ny_mg_set_tail_cells(cells)
            mg = MG(1, 1, n, n, n, 3, 1)
            shape = mg.get_arrayshape(1)
            x = torch.randn(shape, device=dev, dtype=torch.float64, generator=gen)
            mg.set_array(x, ivar=1)
            mg.set_array(x * 0.1, ivar=2)
            del x
            mg.op("fill", 1)
            mg.op("vcycle", 1)
            lib.prof_start()
            for _ in range(8):
                mg.op("vcycle", 1)
            prof = lib.prof_collect()
            lib.prof_start(0)
            print("tail_cells=%d: " % cells + ", ".join("%s %.3f ms/%d" % (k, v[0] / 8, v[1] // 8) for k, v in prof.items() if v[1]), flush=True)
            del mg
            torch.cuda.empty_cache()
        L.ny_mg_set_tail_cells(4096)
    if "mg" in a.what:
        from nyles_b200.mgfordriver import MG
        for topo in (1, 6):
            mg = MG(1, 1, n, n, n, 3, topo)
            shape = mg.get_arrayshape(1)
            x = torch.randn(shape, device=dev, dtype=torch.float64, generator=gen)
            mg.set_array(x, ivar=1)
            mg.set_array(x * 0.1, ivar=2)
            del x
            for name, bpc in (("smooth", 24), ("residual", 24), ("restriction", 9), ("prolongation", 17), ("vcycle", 0)):
                t = timeit(lambda: mg.op(name, 1), n=4, warm=1)
                print("mg topo=%d n=%d %-12s %.3f ms%s" % (topo, n, name, t, (" (%.0f GB/s alg)" % (bpc * cells / t / 1e6)) if bpc else ""), flush=True)
            for fused in (1, 0):
                mg.set_fused_legs(bool(fused))
                mg.op("fill", 1)
                mg.op("vcycle", 1)
                lib.prof_start()
                for _ in range(4):
                    mg.op("vcycle", 1)
                prof = lib.prof_collect()
                lib.prof_start(0)
                print("   vcycle fused=%d: " % fused + ", ".join("%s %.3f ms/%d" % (k, v[0] / 4, v[1] // 4) for k, v in prof.items() if v[1]), flush=True)
            del mg
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
