import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nyles_b200 import lib
from oracle.kernels import Kernels
K = Kernels("strict")
L = lib.load(); ctx = lib.context()
n = (16, 8, 128)
g = torch.Generator(device="cuda").manual_seed(1)
T = torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
Ux = -torch.randn(n, device="cuda", dtype=torch.float64, generator=g).abs() - 0.1
Z = torch.zeros(n, device="cuda", dtype=torch.float64)
res = []
for v in (1, 2):
    lib.check(L.ny_set_momentum_variant(ctx, v))
    out = torch.full(n, 3.0, device="cuda", dtype=torch.float64)
    lib.check(L.ny_upwind(ctx, lib.ptr(T), lib.ptr(Ux), lib.ptr(Z), lib.ptr(Z), lib.ptr(out), lib.ext(out), lib.stream()))
    torch.cuda.synchronize()
    res.append(out.cpu().numpy())
t, u = T.cpu().numpy(), Ux.cpu().numpy()
k, j = 5, 3
def F(i, sh):
    return u[k, j, i] * K.weno5(t[k, j, i + 3 + sh], t[k, j, i + 2 + sh], t[k, j, i + 1 + sh], t[k, j, i + sh], t[k, j, i - 1 + sh])
for i in (40, 41, 50):
    print("i", i, "old", res[0][k, j, i], "new", res[1][k, j, i])
    for sh in (-2, -1, 0, 1, 2):
        for shm in (-2, -1, 0, 1, 2):
            val = (0.0 + F(i - 1, shm)) - F(i, sh)
            if val == res[1][k, j, i]: print("   new matches Fxm shift", shm, "Fx shift", sh)
            if val == res[0][k, j, i]: print("   old matches Fxm shift", shm, "Fx shift", sh)
