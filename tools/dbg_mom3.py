import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nyles_b200 import lib
L = lib.load(); ctx = lib.context()
lib.check(L.ny_set_momentum_variant(ctx, 2))
n = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (24, 20, 40)
g = torch.Generator(device="cuda").manual_seed(1)
F = [torch.randn(n, device="cuda", dtype=torch.float64, generator=g) for _ in range(8)]
out = [torch.zeros(n, device="cuda", dtype=torch.float64) for _ in range(4)]
b, Ux, Uy, Uz, wx, wy, wz, ke = F
e = lib.ext(b)
lib.check(L.ny_rhs(ctx, lib.ptr(b), lib.ptr(Ux), lib.ptr(Uy), lib.ptr(Uz), lib.ptr(wx), lib.ptr(wy), lib.ptr(wz), lib.ptr(ke),
                   lib.ptr(out[0]), lib.ptr(out[1]), lib.ptr(out[2]), lib.ptr(out[3]), 0.25, 0, e, lib.stream()))
torch.cuda.synchronize()
ref = [o.clone() for o in out]
lib.check(L.ny_set_momentum_variant(ctx, 1))
lib.check(L.ny_rhs(ctx, lib.ptr(b), lib.ptr(Ux), lib.ptr(Uy), lib.ptr(Uz), lib.ptr(wx), lib.ptr(wy), lib.ptr(wz), lib.ptr(ke),
                   lib.ptr(out[0]), lib.ptr(out[1]), lib.ptr(out[2]), lib.ptr(out[3]), 0.25, 0, e, lib.stream()))
torch.cuda.synchronize()
print("shape", n, "equal:", [bool(torch.equal(a, b_)) for a, b_ in zip(ref, out)], "maxdiff", [float((a - b_).abs().max()) for a, b_ in zip(ref, out)])
