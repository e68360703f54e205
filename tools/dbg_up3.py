import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nyles_b200 import lib
L = lib.load(); ctx = lib.context()
n = tuple(int(a) for a in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "rand"
g = torch.Generator(device="cuda").manual_seed(1)
F = [torch.randn(n, device="cuda", dtype=torch.float64, generator=g) for _ in range(4)]
if mode == "ypos": F[2] = F[2].abs() + 0.1; F[1] *= 0; F[3] *= 0
if mode == "yneg": F[2] = -F[2].abs() - 0.1; F[1] *= 0; F[3] *= 0
if mode == "xpos": F[1] = F[1].abs() + 0.1; F[2] *= 0; F[3] *= 0
if mode == "xneg": F[1] = -F[1].abs() - 0.1; F[2] *= 0; F[3] *= 0
res = []
for v in (1, 2):
    lib.check(L.ny_set_momentum_variant(ctx, v))
    out = torch.full(n, 3.0, device="cuda", dtype=torch.float64)
    lib.check(L.ny_upwind(ctx, *[lib.ptr(t) for t in F], lib.ptr(out), lib.ext(out), lib.stream()))
    torch.cuda.synchronize()
    res.append(out.cpu().numpy())
bad = np.argwhere(res[0] != res[1])
print("shape", n, mode, "mismatches", len(bad), "of", res[0].size)
if len(bad):
    print(" k:", [int(x) for x in sorted(set(bad[:, 0]))][:12], " j:", [int(x) for x in sorted(set(bad[:, 1]))][:40], " i:", [int(x) for x in sorted(set(bad[:, 2]))][:70])
