import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformerscandobayesianinference_b200 import _lib as L
dev = torch.device("cuda:0")
B, H, dh, T, sep = 512, 4, 128, 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 500
E = H * dh
qkv = torch.randn(T * B, 3 * E, device=dev).to(torch.bfloat16)
out = torch.empty(T * B, E, device=dev, dtype=torch.bfloat16); lse = torch.empty(B * H, T, device=dev)
for _ in range(2): L.attention_fwd(qkv, out, lse, T, B, H, dh, sep, use_tc=True)
cap = 1500
buf = torch.zeros(3 * 4 * cap, device=dev, dtype=torch.int64)
L.load().pfn_debug_attention_trace(buf.data_ptr(), cap)
L.attention_fwd(qkv, out, lse, T, B, H, dh, sep, use_tc=True)
torch.cuda.synchronize()
L.load().pfn_debug_attention_trace(None, 0)
ev = [e for e in buf.view(3 * cap, 4).cpu().tolist() if e[3] != 0]
ev.sort(key=lambda e: e[3])
t0 = ev[0][3]
names = {1: "P:Qload", 2: "P:KVload", 10: "M:Qok", 11: "M:KVok+QK", 12: "M:Pready", 13: "M:PVissued", 20: "T:S", 21: "T:Ppub", 22: "T:epi0", 23: "T:epi1"}
# print the events of the first ~3 tiles after a warm start (skip first 2 tiles)
for e in ev[:400]:
    print(f"{e[3] - t0:8d} {names.get(e[0], e[0]):12s} tile={e[1]} blk={e[2]}")
