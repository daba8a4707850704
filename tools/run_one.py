"""Launch one kernel a few times (for ncu): python tools/run_one.py {attn_fwd|attn_bwd|gemm|gp} """
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformerscandobayesianinference_b200 import _lib as L
dev = torch.device("cuda:0")
what = sys.argv[1]
B, H, dh, T, sep = int(os.environ.get("B", 128)), 4, 128, 1000, 500
E = H * dh
if what.startswith("attn"):
    qkv = torch.randn(T * B, 3 * E, device=dev).to(torch.bfloat16)
    out = torch.empty(T * B, E, device=dev, dtype=torch.bfloat16); lse = torch.empty(B * H, T, device=dev)
    for _ in range(3):
        L.attention_fwd(qkv, out, lse, T, B, H, dh, sep, use_tc=True)
    if what == "attn_bwd":
        dout = torch.randn(T * B, E, device=dev).to(torch.bfloat16); dqkv = torch.empty_like(qkv); delta = torch.empty_like(lse)
        for _ in range(3):
            L.attention_bwd(qkv, out, lse, dout, dqkv, delta, T, B, H, dh, sep, use_tc=True)
elif what == "gemm":
    N = T * B
    x = torch.randn(N, 512, device=dev).to(torch.bfloat16); w = torch.randn(1024, 512, device=dev).to(torch.bfloat16)
    y = torch.empty(N, 1024, device=dev, dtype=torch.bfloat16); u = torch.empty_like(y); bias = torch.randn(1024, device=dev)
    for _ in range(3):
        L.gemm(x, w, y, bias=bias, C2=u, epilogue=L.EPI_GELU, use_tc=True)
elif what == "gp":
    Bn = 296
    x = torch.rand(Bn, T, 1, device=dev); z = torch.randn(Bn, T, device=dev)
    ls = torch.full((Bn, 1), .6, device=dev); os_ = torch.ones(Bn, device=dev); nz = torch.full((Bn,), 1e-4, device=dev)
    y = torch.empty(Bn, T, device=dev); work = torch.empty(Bn, T, T, device=dev); info = torch.zeros(Bn, device=dev, dtype=torch.int32)
    for _ in range(2):
        L.gp_sample(x, z, ls, os_, nz, 0.0, 0, y, work, info)
torch.cuda.synchronize()
