import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformerscandobayesianinference_b200 import _lib as L
dev = torch.device("cuda:0")
N = 512000
CASES = [(N, 1536, 512, 0, 0, "qkv fwd"), (N, 512, 512, 0, 0, "out fwd"), (N, 512, 1536, 0, 1, "qkv dgrad"), (1536, 512, N, 1, 1, "qkv wgrad"),
         (N, 1024, 512, 0, 0, "mlp1 plain"), (N, 1024, 512, 0, 0, "mlp1 gelu+c2"), (N, 1024, 512, 0, 0, "mlp1 gelu+c2grad"), (N, 1024, 512, 0, 1, "mlp2 dgrad gelu'"), (N, 1024, 512, 0, 1, "mlp2 dgrad mul"), (N, 512, 512, 0, 0, "out fwd +aux")]
for (M, Nn, K, amn, bmn, name) in CASES:
    A = torch.randn((K, M) if amn else (M, K), device=dev).to(torch.bfloat16)
    Bm = torch.randn((K, Nn) if bmn else (Nn, K), device=dev).to(torch.bfloat16)
    wg = amn and bmn
    C = torch.zeros(M, Nn, device=dev, dtype=torch.float32 if wg else torch.bfloat16)
    kw = {}
    if "gelu+c2" in name:
        kw = dict(bias=torch.randn(Nn, device=dev), C2=torch.empty_like(C), epilogue=1, c2_gelu_grad="grad" in name)
    elif "mul" in name:
        kw = dict(aux=torch.randn(M, Nn, device=dev).to(torch.bfloat16), epilogue=4)
    elif "gelu'" in name:
        kw = dict(aux=torch.randn(M, Nn, device=dev).to(torch.bfloat16), epilogue=2)
    elif "+aux" in name:
        kw = dict(bias=torch.randn(Nn, device=dev), aux=torch.randn(M, Nn, device=dev).to(torch.bfloat16))
    run = lambda: L.gemm(A, Bm, C, a_mn_major=bool(amn), b_mn_major=bool(bmn), M=M, N=Nn, K=K, accumulate=wg, k_splits=37 if wg else 1, use_tc=True, **kw)
    for _ in range(3): run()
    buf = torch.zeros(148 * 8, device=dev, dtype=torch.int64)
    L.load().pfn_debug_attention_trace(buf.data_ptr(), 0, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    L.load().pfn_debug_attention_trace(None, 0, 0)
    st = buf.view(148, 8).double().mean(0).tolist()
    ms = e0.elapsed_time(e1)
    print(f"{name}: {ms:.3f} ms | mean clocks per CTA waiting: producer(empty) {st[0]:.0f}, MMA(full) {st[1]:.0f}, MMA(tempty) {st[2]:.0f}, epilogue(tfull) {st[3]:.0f}, epilogue(staging free) {st[4]:.0f}, epilogue(column loop) {st[5]:.0f} of which tmem ld wait {st[6]:.0f}")
