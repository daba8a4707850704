"""Ablation timing of the attention backward kernels: variant libraries (tools/build_variants.sh attention_bwd_tc.cu ...)
run in subprocesses on the same GPU; results of ablated variants are numerically wrong, only their time matters."""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
CHILD = r'''
import sys, os
sys.path.insert(0, os.path.dirname(HERE))
import torch
from transformerscandobayesianinference_b200 import _lib as L
dev = torch.device("cuda:0")
B, H, dh, T, sep = 512, 4, 128, 1000, 500
E = H * dh
qkv = torch.randn(T * B, 3 * E, device=dev).to(torch.bfloat16)
out = torch.empty(T * B, E, device=dev, dtype=torch.bfloat16); lse = torch.empty(B * H, T, device=dev)
L.attention_fwd(qkv, out, lse, T, B, H, dh, sep, use_tc=True)
dout = torch.randn(T * B, E, device=dev).to(torch.bfloat16); dqkv = torch.empty_like(qkv); delta = torch.empty_like(lse)
def t(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = []
for name, sel in (("dq", 22), ("dkv", 21)):
    L.load().pfn_debug_attention_trace(None, 0, sel)
    res.append(f"{name} {t(lambda: L.attention_bwd(qkv, out, lse, dout, dqkv, delta, T, B, H, dh, sep, use_tc=True)):.3f} ms")
print(" | ".join(res))
'''.replace("HERE", repr(HERE))
for v in sys.argv[1:]:
    env = dict(os.environ, PFN_B200_LIB=os.path.join(HERE, "ubench", "_bin", f"libpfn_{v}.so"))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    print(f"{v:10s}: {r.stdout.strip() or r.stderr.strip()[-400:]}", flush=True)
