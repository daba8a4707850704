"""Stress the tcgen05 attention kernels: repeat each case many times and report which (t, b, h) rows go wrong."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformerscandobayesianinference_b200 import _lib as L
from oracle import pfn_oracle as O

dev = torch.device("cuda:0")
CASES = [(200, 2, 4, 100), (256, 2, 2, 128), (1000, 2, 4, 500), (513, 1, 2, 257), (300, 3, 1, 299), (384, 1, 2, 320)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
mode = sys.argv[2] if len(sys.argv) > 2 else "fwd"
for (T, B, H, sep) in CASES:
    dh, E = 128, H * 128
    torch.manual_seed(T + sep)
    qkv = (torch.randn(T * B, 3 * E, device=dev) * 1.5).to(torch.bfloat16)
    ref, ref_lse = O.attention_ref(qkv.float().cpu().double(), T, B, H, dh, sep)
    scale = ref.abs().max().item()
    bad = 0
    first = None
    outs = []
    for r in range(reps):
        out = torch.full((T * B, E), float("nan"), device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B * H, T, device=dev)
        L.attention_fwd(qkv, out, lse, T, B, H, dh, sep, use_tc=True)
        torch.cuda.synchronize()
        o = out.float().cpu().double()
        err = (o - ref).abs()
        if not torch.isfinite(o).all() or err.max().item() > 2e-2 * scale:
            bad += 1
            if first is None:
                e = err.reshape(T, B, H, dh).amax(-1)
                e[~torch.isfinite(e)] = 1e9
                idx = (e > 2e-2 * scale).nonzero()
                first = (r, idx[:12].tolist(), len(idx), float(err[torch.isfinite(err)].max()))
    print(f"case T={T} B={B} H={H} sep={sep}: {bad}/{reps} bad; first: {first}", flush=True)
