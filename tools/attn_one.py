"""One attention shape, a few launches: target for `ncu -k regex:attn_tc_kernel -s 2 -c 1 --set full`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minimagen_b200.ops import NativeOps
ops = NativeOps()
B, n, m, heads, d = int(os.environ.get("AB", 16)), int(os.environ.get("AN", 4096)), int(os.environ.get("AM", 4096)), 8, 64
inner = heads * d
q = torch.randn(B, n, inner, device="cuda", dtype=torch.float16) * 0.125
out = torch.empty_like(q)
null_kv = torch.randn(2, d, device="cuda")
kv = torch.randn(B, m, 2 * d, device="cuda", dtype=torch.float16)
for _ in range(4):
    ops.attention(q, n * inner, inner, kv, kv[..., d:], m * 2 * d, 2 * d, 0, null_kv, None, B, heads, n, m, out, n * inner, inner)
torch.cuda.synchronize()
print("done")
