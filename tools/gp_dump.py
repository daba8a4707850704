"""Bitwise fingerprint of the GP sampler's outputs on fixed seeded inputs (for A/B builds: run under each PFN_B200_LIB and
compare the printed hashes), plus the kernel time at the cfg-2 shape.  usage: python tools/gp_dump.py"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformerscandobayesianinference_b200 import _lib as L
dev = torch.device("cuda:0")
h = hashlib.sha256()
for (Bn, T, F, kt, noise) in [(320, 1000, 1, 0, 1e-4), (24, 1000, 1, 0, 1e-4), (40, 333, 3, 3, 1e-2), (300, 64, 2, 1, 1e-2)]:
    torch.manual_seed(Bn + T)
    x = torch.rand(Bn, T, F, device=dev); z = torch.randn(Bn, T, device=dev)
    ls = torch.rand(Bn, F, device=dev) * 0.5 + 0.35; os_ = torch.rand(Bn, device=dev) + 0.5
    nz = torch.full((Bn,), noise, device=dev)
    ldw = (T + 3) // 4 * 4
    y = torch.empty(Bn, T, device=dev); work = torch.zeros(Bn, T, ldw, device=dev); info = torch.zeros(Bn, device=dev, dtype=torch.int32)
    L.gp_sample(x, z, ls, os_, nz, 0.0, kt, y, work, info)
    torch.cuda.synchronize()
    fac = work[:, :, :T].transpose(1, 2).tril().contiguous()
    h.update(y.cpu().numpy().tobytes()); h.update(fac.cpu().numpy().tobytes()); h.update(info.cpu().numpy().tobytes())
    print(f"B={Bn} T={T} F={F} kernel={kt}: bad pivots {int((info != 0).sum())}, |y| mean {y.abs().mean().item():.6f}")
print("sha256", h.hexdigest())
