"""A/B timing of GEMM epilogue variants on ONE GPU in ONE session: each variant library (tools/build_variants.sh) runs in
its own subprocess (PFN_B200_LIB), interleaved twice so clock/power drift shows up as the spread between rounds."""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
CHILD = r'''
import sys, os
sys.path.insert(0, os.path.dirname(HERE))
import torch
from transformerscandobayesianinference_b200 import _lib as L
dev = torch.device("cuda:0"); N = 512000
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = []
for (Nn, K, bmn, name) in [(1536, 512, 0, "plain"), (512, 1536, 1, "plain K1536"), (512, 1024, 1, "plain K1024"), (1024, 512, 0, "gelu+c2g"), (1024, 512, 1, "mul"), (1024, 512, 0, "gelu"), (512, 512, 0, "+aux"), (512, 1024, 0, "+aux K1024"), (512, 512, 1, "rowdot")]:
    A = torch.randn(N, K, device=dev).to(torch.bfloat16)
    Bm = (torch.randn((K, Nn) if bmn else (Nn, K), device=dev) * K ** -0.5).to(torch.bfloat16)
    C = torch.zeros(N, Nn, device=dev, dtype=torch.bfloat16)
    kw = {}
    if name == "gelu+c2g": kw = dict(bias=torch.randn(Nn, device=dev), C2=torch.empty_like(C), epilogue=1, c2_gelu_grad=True)
    elif name == "gelu": kw = dict(bias=torch.randn(Nn, device=dev), epilogue=1)
    elif name == "mul": kw = dict(aux=torch.randn(N, Nn, device=dev).to(torch.bfloat16), epilogue=4)
    elif name == "rowdot": kw = dict(aux=torch.randn(N, Nn, device=dev).to(torch.bfloat16), epilogue=3, rowdot=(torch.zeros(N, Nn // 128, device=dev), 128))
    elif name.startswith("+aux"): kw = dict(bias=torch.randn(Nn, device=dev), aux=torch.randn(N, Nn, device=dev).to(torch.bfloat16))
    ms = t(lambda: L.gemm(A, Bm, C, b_mn_major=bool(bmn), M=N, N=Nn, K=K, use_tc=True, **kw))
    R = 4096
    acc = A[:R].float() @ (Bm.float() if bmn else Bm.float().t())
    if "bias" in kw: acc = acc + kw["bias"]
    ref2 = None
    if name in ("gelu+c2g", "gelu"):
        ref = torch.nn.functional.gelu(acc)
        if name == "gelu+c2g":
            ref2 = 0.5 * (1 + torch.erf(acc / 2 ** 0.5)) + acc * torch.exp(-0.5 * acc * acc) / (2 * 3.141592653589793) ** 0.5
    elif name == "mul": ref = acc * kw["aux"][:R].float()
    elif name == "rowdot": ref = acc
    elif name.startswith("+aux"): ref = acc + kw["aux"][:R].float()
    else: ref = acc
    err = (C[:R].float() - ref).abs().max().item() / ref.abs().max().item()
    extra = f" (err {err:.1e}"
    if ref2 is not None: extra += f", c2 {(kw['C2'][:R].float() - ref2).abs().max().item():.1e}"
    out.append(f"{name} {ms:.3f}{extra})")
print(" | ".join(out))
'''.replace("HERE", repr(HERE))
variants = sys.argv[1:] or ["tanh1", "logi1", "tanh2", "logi1nopf"]
for rnd in range(2):
    for v in variants:
        if v.startswith("env:"):       # in-tree library with an environment switch, e.g. env:PFN_GEMM_AUX_TMA=0
            k, val = v[4:].split("=", 1)
            env = dict(os.environ, **{k: val})
        else:
            env = dict(os.environ, PFN_B200_LIB=os.path.join(HERE, "ubench", "_bin", f"libpfn_{v}.so"))
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=100)
        print(f"[{rnd}] {v:24s}: {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
