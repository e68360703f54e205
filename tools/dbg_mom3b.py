import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nyles_b200 import lib
L = lib.load(); ctx = lib.context()
lib.check(L.ny_set_momentum_variant(ctx, 2))
n = int(sys.argv[1]); reps = int(sys.argv[2]); smooth = int(sys.argv[3])
dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(1)
def smooth_field(amp=1.0):
    z = torch.linspace(0, 6.28, n, device=dev, dtype=torch.float64)
    f = (torch.sin(3 * z)[:, None, None] * torch.cos(2 * z)[None, :, None] * torch.sin(5 * z + 1)[None, None, :])
    return (amp * (f + 0.05 * torch.randn((n, n, n), device=dev, dtype=torch.float64, generator=gen))).contiguous()
F = [smooth_field() if smooth else torch.randn((n, n, n), device=dev, dtype=torch.float64, generator=gen) for _ in range(8)]
out = [torch.empty((n, n, n), device=dev, dtype=torch.float64) for _ in range(4)]
b, Ux, Uy, Uz, wx, wy, wz, ke = F
e = lib.ext(b)
for r in range(reps):
    lib.check(L.ny_rhs(ctx, lib.ptr(b), lib.ptr(Ux), lib.ptr(Uy), lib.ptr(Uz), lib.ptr(wx), lib.ptr(wy), lib.ptr(wz), lib.ptr(ke),
                       lib.ptr(out[0]), lib.ptr(out[1]), lib.ptr(out[2]), lib.ptr(out[3]), 0.25, 0, e, lib.stream()))
    if len(sys.argv) > 4: torch.cuda.synchronize()
torch.cuda.synchronize()
print("ok", n, reps, smooth, float(out[1].abs().max()))
