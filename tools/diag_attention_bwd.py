import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformerscandobayesianinference_b200 import _lib as L
from oracle import pfn_oracle as O
dev = torch.device("cuda:0")
CASES = [(128, 1, 1, 64), (128, 1, 1, 128), (256, 1, 1, 64), (256, 1, 1, 128), (256, 1, 1, 192), (128, 2, 1, 64), (128, 1, 2, 64), (200, 1, 1, 100)]
for (T, B, H, sep) in CASES:
    dh, E = 128, H * 128
    torch.manual_seed(T * 3 + sep)
    qkv = (torch.randn(T * B, 3 * E, device=dev) * 1.2).to(torch.bfloat16)
    out = torch.empty(T * B, E, device=dev, dtype=torch.bfloat16); lse = torch.empty(B * H, T, device=dev)
    L.attention_fwd(qkv, out, lse, T, B, H, dh, sep, use_tc=True)
    dout = torch.randn(T * B, E, device=dev).to(torch.bfloat16)
    qr = qkv.float().cpu().double().requires_grad_(True)
    ref, _ = O.attention_ref(qr, T, B, H, dh, sep)
    (ref * dout.float().cpu().double()).sum().backward()
    for rep in range(2):
        dqkv = torch.full_like(qkv, float("nan")); delta = torch.empty(B * H, T, device=dev)
        L.attention_bwd(qkv, out, lse, dout, dqkv, delta, T, B, H, dh, sep, use_tc=True)
        torch.cuda.synchronize()
        got = dqkv.float().cpu().double()
        msg = []
        for name, k in (("dq", 0), ("dk", 1), ("dv", 2)):
            want = qr.grad[:, k * E:(k + 1) * E]; g = got[:, k * E:(k + 1) * E]
            err = (g - want).abs(); err[~torch.isfinite(g)] = 1e9
            e = err.reshape(T, B, H, dh).amax(-1)
            bad = (e > 3e-2 * (want.abs().max().item() + 1e-6)).nonzero()
            ts = sorted(set(int(x[0]) for x in bad))
            rng = f"{ts[0]}..{ts[-1]} ({len(ts)} rows)" if ts else "ok"
            msg.append(f"{name}: {rng} max {float(e.max()):.3g}/{float(want.abs().max()):.3g}")
        print(f"T={T} B={B} H={H} sep={sep} rep{rep}: " + " | ".join(msg), flush=True)
