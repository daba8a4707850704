"""One launch of the GP sampler kernel at the cfg-2 shape (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformerscandobayesianinference_b200 import _lib as L
dev = torch.device("cuda:0")
Bn, T = int(sys.argv[1]) if len(sys.argv) > 1 else 296, 1000
x = torch.rand(Bn, T, 1, device=dev); z = torch.randn(Bn, T, device=dev)
ls = torch.full((Bn, 1), .6, device=dev); os_ = torch.ones(Bn, device=dev); nz = torch.full((Bn,), 1e-4, device=dev)
y = torch.empty(Bn, T, device=dev); work = torch.empty(Bn, T, T, device=dev); info = torch.zeros(Bn, device=dev, dtype=torch.int32)
for _ in range(3):
    L.gp_sample(x, z, ls, os_, nz, 0.0, 0, y, work, info)
torch.cuda.synchronize()
print("info", int(info.max()))
