"""Micro-benchmarks (CUDA events) of individual kernels at cfg-2 shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformerscandobayesianinference_b200 import _lib as L
dev = torch.device("cuda:0")

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

what = sys.argv[1] if len(sys.argv) > 1 else "attn"
B, H, dh = 512, 4, 128
E = H * dh
if what in ("attnparts",):
    T, sep = 1000, 500
    qkv = (torch.randn(T * B, 3 * E, device=dev)).to(torch.bfloat16)
    out = torch.empty(T * B, E, device=dev, dtype=torch.bfloat16); lse = torch.empty(B * H, T, device=dev)
    dout = torch.randn(T * B, E, device=dev).to(torch.bfloat16); dqkv = torch.empty_like(qkv); delta = torch.empty_like(lse)
    L.attention_fwd(qkv, out, lse, T, B, H, dh, sep, use_tc=True)
    for name, sel in (("all", 0), ("dkv", 21), ("dq", 22), ("delta", 23)):
        L.load().pfn_debug_attention_trace(None, 0, sel)
        t = timeit(lambda: L.attention_bwd(qkv, out, lse, dout, dqkv, delta, T, B, H, dh, sep, use_tc=True))
        print(f"attn bwd part {name}: {t:.3f} ms")
    L.load().pfn_debug_attention_trace(None, 0, 0)
if what in ("attn", "all"):
    for (T, sep) in [(1000, 500), (500, 500), (1000, 1000), (1000, 0), (1000, 64), (2000, 1000)]:
        qkv = (torch.randn(T * B, 3 * E, device=dev)).to(torch.bfloat16)
        out = torch.empty(T * B, E, device=dev, dtype=torch.bfloat16); lse = torch.empty(B * H, T, device=dev)
        dout = torch.randn(T * B, E, device=dev).to(torch.bfloat16); dqkv = torch.empty_like(qkv); delta = torch.empty_like(lse)
        fl = 4 * E * B * (T * sep + (T - sep))
        for bm in (False, True):
            tf = timeit(lambda: L.attention_fwd(qkv, out, lse, T, B, H, dh, sep, use_tc=True, batch_major=bm))
            tb = timeit(lambda: L.attention_bwd(qkv, out, lse, dout, dqkv, delta, T, B, H, dh, sep, use_tc=True, batch_major=bm))
            print(f"attn T={T} sep={sep} batch_major={int(bm)}: fwd {tf:.3f} ms ({fl / tf / 1e9:.0f} TF/s)  bwd {tb:.3f} ms ({2 * fl / tb / 1e9:.0f} TF/s)", flush=True)
        del qkv, out, dout, dqkv
if what in ("gemm", "all"):
    N = 512000
    for (M, Nn, K, amn, bmn, name) in [(N, 1536, 512, 0, 0, "qkv fwd"), (N, 512, 512, 0, 0, "out fwd"), (N, 1024, 512, 0, 0, "mlp1 fwd"),
                                        (N, 512, 1024, 0, 0, "mlp2 fwd"), (N, 512, 1536, 0, 1, "qkv dgrad"), (N, 1024, 512, 0, 1, "mlp2 dgrad"),
                                        (1536, 512, N, 1, 1, "qkv wgrad"), (1024, 512, N, 1, 1, "mlp1 wgrad"), (512, 1024, N, 1, 1, "mlp2 wgrad")]:
        A = torch.randn((K, M) if amn else (M, K), device=dev).to(torch.bfloat16)
        Bm = torch.randn((K, Nn) if bmn else (Nn, K), device=dev).to(torch.bfloat16)
        wg = amn and bmn
        C = torch.zeros(M, Nn, device=dev, dtype=torch.float32 if wg else torch.bfloat16)
        for splits in ([16, 37, 74] if wg else [1]):
            t = timeit(lambda: L.gemm(A, Bm, C, a_mn_major=bool(amn), b_mn_major=bool(bmn), M=M, N=Nn, K=K, accumulate=wg, k_splits=splits, use_tc=True))
            print(f"gemm {name} {M}x{Nn}x{K} splits={splits}: {t:.3f} ms ({2 * M * Nn * K / t / 1e9:.0f} TF/s)", flush=True)
        del A, Bm, C
if what in ("gemmepi", "all"):
    N = 512000
    x = torch.randn(N, 512, device=dev).to(torch.bfloat16); w1 = (torch.randn(1024, 512, device=dev) * .05).to(torch.bfloat16)
    g = torch.empty(N, 1024, device=dev, dtype=torch.bfloat16); u = torch.empty_like(g); b1 = torch.randn(1024, device=dev)
    t = timeit(lambda: L.gemm(x, w1, g, bias=b1, C2=u, epilogue=L.EPI_GELU, use_tc=True)); print(f"mlp1 fwd bias+GELU+C2: {t:.3f} ms ({2*N*1024*512/t/1e9:.0f} TF/s)")
    w2 = (torch.randn(512, 1024, device=dev) * .05).to(torch.bfloat16); z = torch.empty(N, 512, device=dev, dtype=torch.bfloat16); b2 = torch.randn(512, device=dev)
    t = timeit(lambda: L.gemm(g, w2, z, bias=b2, aux=x, use_tc=True)); print(f"mlp2 fwd bias+residual: {t:.3f} ms ({2*N*1024*512/t/1e9:.0f} TF/s)")
    du = torch.empty_like(u)
    t = timeit(lambda: L.gemm(z, w2, du, b_mn_major=True, aux=u, epilogue=L.EPI_GELU_BWD, M=N, N=1024, K=512, use_tc=True)); print(f"mlp2 dgrad * GELU'(u): {t:.3f} ms ({2*N*1024*512/t/1e9:.0f} TF/s)")
    dh = torch.empty_like(x)
    t = timeit(lambda: L.gemm(du, w1, dh, b_mn_major=True, aux=x, M=N, N=512, K=1024, use_tc=True)); print(f"mlp1 dgrad + residual: {t:.3f} ms ({2*N*1024*512/t/1e9:.0f} TF/s)")
if what in ("gp", "all"):
    for (Bn, T) in [(512, 1000), (512, 500), (148, 1000), (296, 1000)]:
        x = torch.rand(Bn, T, 1, device=dev); z = torch.randn(Bn, T, device=dev)
        ls = torch.full((Bn, 1), .6, device=dev); os_ = torch.ones(Bn, device=dev); nz = torch.full((Bn,), 1e-4, device=dev)
        y = torch.empty(Bn, T, device=dev); work = torch.empty(Bn, T, (T + 3) // 4 * 4, device=dev); info = torch.zeros(Bn, device=dev, dtype=torch.int32)
        t = timeit(lambda: L.gp_sample(x, z, ls, os_, nz, 0.0, 0, y, work, info))
        print(f"gp_sample B={Bn} T={T}: {t:.3f} ms ({Bn * T ** 3 / 3 / t / 1e9:.1f} TFLOP/s fp32)", flush=True)
if what in ("row", "all"):
    N = 512000
    z = torch.randn(N, 512, device=dev).to(torch.bfloat16); h = torch.empty_like(z)
    g = torch.ones(512, device=dev); b = torch.zeros(512, device=dev); mean = torch.empty(N, device=dev); rstd = torch.empty(N, device=dev)
    t = timeit(lambda: L.layernorm_fwd(z, g, b, h, mean, rstd)); print(f"ln fwd: {t:.3f} ms ({N * 512 * 4 / t / 1e6:.0f} GB/s)")
    dz = torch.empty_like(z); dg = torch.zeros(512, device=dev); db = torch.zeros(512, device=dev); cs = torch.zeros(512, device=dev)
    t = timeit(lambda: L.layernorm_bwd(h, z, mean, rstd, g, dz, dg, db, cs)); print(f"ln bwd: {t:.3f} ms ({N * 512 * 6 / t / 1e6:.0f} GB/s)")
    for cols in (512, 1024, 1536):
        X = torch.randn(N, cols, device=dev).to(torch.bfloat16); o = torch.zeros(cols, device=dev)
        t = timeit(lambda: L.colsum(X, o)); print(f"colsum {cols}: {t:.3f} ms ({N * cols * 2 / t / 1e6:.0f} GB/s)")
