import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformerscandobayesianinference_b200 import _lib as L
dev = torch.device("cuda:0")
B, H, dh, T, sep = 512, 4, 128, 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 500
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
E = H * dh
qkv = torch.randn(T * B, 3 * E, device=dev).to(torch.bfloat16)
out = torch.empty(T * B, E, device=dev, dtype=torch.bfloat16); lse = torch.empty(B * H, T, device=dev)
for _ in range(2): L.attention_fwd(qkv, out, lse, T, B, H, dh, sep, use_tc=True)
dout = torch.randn(T * B, E, device=dev).to(torch.bfloat16); dqkv = torch.empty_like(qkv); delta = torch.empty_like(lse)
if which: L.attention_bwd(qkv, out, lse, dout, dqkv, delta, T, B, H, dh, sep, use_tc=True)
cap = 1500
NR = 11
buf = torch.zeros(NR * 4 * cap, device=dev, dtype=torch.int64)
L.load().pfn_debug_attention_trace(buf.data_ptr(), cap, which)
if which == 0: L.attention_fwd(qkv, out, lse, T, B, H, dh, sep, use_tc=True)
else: L.attention_bwd(qkv, out, lse, dout, dqkv, delta, T, B, H, dh, sep, use_tc=True)
torch.cuda.synchronize()
L.load().pfn_debug_attention_trace(None, 0, 0)
ev = [e for e in buf.view(NR * cap, 4).cpu().tolist() if e[3] != 0]
ev.sort(key=lambda e: e[3])
t0 = ev[0][3]
names = {10: "M:QTready", 24: "T:copied", 1: "P:Qload", 2: "P:KVload", 3: "P:Vload", 11: "M:KVok+QK", 12: "M:Pready", 13: "M:PVissued", 20: "T:S", 21: "T:Ppub", 22: "T:epi0", 23: "T:epi1", 14: "M:QKissued", 24: "T:barsync", 25: "T:loaded", 26: "T:computed", 27: "T:pdfree"}
# print the events of the first ~3 tiles after a warm start (skip first 2 tiles)
for e in ev[:int(os.environ.get('TRACE_N', '1200'))]:
    code, wid = e[0] % 100, e[0] // 100
    nm = names.get(code, str(code)) + (f"/w{wid}" if wid else "")
    print(f"{e[3] - t0:8d} {nm:14s} tile={e[1]} blk={e[2]}")
