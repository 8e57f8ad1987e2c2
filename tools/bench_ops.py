"""Micro-benchmarks of the non-GEMM kernels at the cfg-3 (SR 64->256, dim 128, batch 32) shapes.  GPU only.

Diagnostics for kernel tuning, not a bench value: CUDA events around `reps` back-to-back launches, rotating over enough
buffer sets that the working set exceeds the 126 MB L2 ("cold") or re-using one set ("warm").
Usage: python tools/bench_ops.py [gn ln linear quantile attn cast final stem]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minimagen_b200 import _native, ops as ops_mod   # noqa: E402

F16, F32, F64 = torch.float16, torch.float32, torch.float64
dev = torch.device("cuda", 0)
HBM = 6486.5   # GB/s, MEASURED_PEAKS.json


def timeit(fn_list, reps=20):
    """fn_list: callables doing the same work on different buffers; returns ms per call."""
    for f in fn_list:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn_list[i % len(fn_list)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms_cold, ms_warm, nbytes):
    print(f"{name:44s} cold {ms_cold * 1e3:8.1f} us ({nbytes / ms_cold / 1e6:7.0f} GB/s, {nbytes / ms_cold / 1e6 / HBM * 100:5.1f}% HBM)"
          f"   warm {ms_warm * 1e3:8.1f} us   [{nbytes / 1e6:.0f} MB]", flush=True)


def bench_gn(ops):
    B = 32
    for (hw, c0, c1) in [(128 * 128, 128, 0), (128 * 128, 128, 128), (64 * 64, 256, 0), (64 * 64, 256, 256),
                         (32 * 32, 512, 0), (32 * 32, 512, 512), (16 * 16, 1024, 0), (16 * 16, 1024, 1024),
                         (256 * 256, 128, 0)]:
        C = c0 + c1
        n = B * hw * C
        nbytes = n * 6
        nset = max(1, int(300e6 // nbytes) + 1)
        sets = []
        for _ in range(nset):
            s0 = torch.randn(B, hw, c0, device=dev)
            s1 = torch.randn(B, hw, c1, device=dev) if c1 else None
            st0 = torch.rand(B, c0 // 16, 2, device=dev, dtype=F64) * hw * 16
            st0[..., 1] += hw * 16
            st1 = None
            if c1:
                st1 = torch.rand(B, c1 // 16, 2, device=dev, dtype=F64) * hw * 16
                st1[..., 1] += hw * 16
            out = torch.empty(B, hw, C, device=dev, dtype=F16)
            sets.append((s0, s1, st0, st1, out))
        gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
        ss = torch.randn(B, 2 * C, device=dev)

        def mk(t):
            s0, s1, st0, st1, out = t
            return lambda: ops.gn_apply_silu(s0, c0, s1, c1, 0.7071, B, hw, 8, st0, 16, st1, 16 if c1 else 0, gamma, beta,
                                             ss, 2 * C, 1e-5, out)
        fns = [mk(t) for t in sets]
        report(f"gn_apply hw={hw} C={c0}+{c1}", timeit(fns), timeit(fns[:1]), nbytes)


def bench_ln(ops):
    for (rows, C) in [(8192, 1024), (32 * 59, 512)]:
        nbytes = rows * C * (4 + 2)
        nset = max(1, int(300e6 // nbytes) + 1)
        g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
        sets = [(torch.randn(rows, C, device=dev), torch.empty(rows, C, device=dev, dtype=F16)) for _ in range(nset)]
        fns = [(lambda t=t: ops.ln_rows(t[0], rows, C, g, b, 1e-5, 0, None, None, t[1])) for t in sets]
        report(f"ln_rows rows={rows} C={C} (f16 out)", timeit(fns), timeit(fns[:1]), nbytes)


def bench_linear(ops):
    for (M, K, Nn) in [(32, 512, 518 * 128), (32, 768, 1024), (32 * 59, 768, 128), (32, 512, 256), (32, 128, 512)]:
        x = torch.randn(M, K, device=dev)
        nbytes = Nn * K * 4 + M * K * 4 + M * Nn * 4
        nset = max(1, int(300e6 // nbytes) + 1)
        nset = min(nset, 8)
        sets = [(torch.randn(Nn, K, device=dev), torch.randn(Nn, device=dev), torch.empty(M, Nn, device=dev)) for _ in range(nset)]
        fns = [(lambda t=t: ops.linear_f32(x, M, K, t[0], t[1], Nn, 0, 0, None, t[2], None)) for t in sets]
        report(f"linear_f32 M={M} K={K} N={Nn}", timeit(fns), timeit(fns[:1]), nbytes)


def bench_quantile(ops):
    B, n = 32, 3 * 256 * 256
    x = torch.randn(B, n, device=dev)
    s = torch.empty(B, device=dev)
    pos = 0.95 * (n - 1)
    lo = int(pos)
    f = lambda: ops.step_quantile(x, B, n, lo, lo + 1, pos - lo, 1.0, s)
    ms = timeit([f])
    report("step_quantile B=32 n=196608", ms, ms, B * n * 4)


def bench_attn(ops):
    heads, d = 8, 64
    inner = heads * d
    for (B, n, m, shared) in [(32, 256, 256, True), (32, 256, 59, False), (64, 1024, 1024, True), (64, 4096, 4096, True),
                              (64, 1024, 258, False), (64, 4096, 258, False)]:
        q = torch.randn(B, n, inner, device=dev, dtype=F16) * 0.125
        out = torch.empty(B, n, inner, device=dev, dtype=F16)
        null_kv = torch.randn(2, d, device=dev)
        if shared:      # self-attention (layers.py:14-104): one shared k/v head (multi-query)
            kv = torch.randn(B, m, 2 * d, device=dev, dtype=F16)
            args = (kv, kv[..., d:], m * 2 * d, 2 * d, 0)
        else:           # cross-attention (layers.py:180-251): per-head k/v over m text tokens
            kv = torch.randn(B, m, 2 * inner, device=dev, dtype=F16)
            args = (kv, kv[..., inner:], m * 2 * inner, 2 * inner, d)
        fl = 4.0 * B * heads * n * (m + 1) * d
        res = []
        for tc in (True, False):
            ops.attention_tc = tc
            f = lambda: ops.attention(q, n * inner, inner, *args, null_kv, None, B, heads, n, m, out, n * inner, inner)
            ms = timeit([f], reps=5)
            res.append(f"{'tcgen05' if tc else 'mma.sync'} {ms * 1e3:9.1f} us ({fl / ms / 1e9:6.1f} TFLOP/s)")
        ops.attention_tc = True
        print(f"attention B={B} n={n} m={m} {'multi-query' if shared else 'per-head kv'}:  " + "   ".join(res), flush=True)


def bench_cast(ops):
    B = 32
    for (H, c0, c1, mode) in [(128, 128, 0, 2), (64, 256, 0, 2), (64, 256, 0, 1), (128, 128, 0, 1), (32, 512, 0, 1), (16, 1024, 0, 1)]:
        C = c0 + c1
        n = B * H * H * C
        mult = 4 if mode == 1 else 1
        nbytes = n * 2 + n * 2 * mult
        src = torch.randn(B, H, H, c0, device=dev, dtype=F16)
        out = torch.empty(B * mult, H, H, C, device=dev, dtype=F16)
        f = lambda: ops.cast_act(src, c0, None, 0, 1.0, B, H, H, mode, out)
        ms = timeit([f])
        report(f"cast_act f16->f16 H={H} C={C} mode={mode}", ms, ms, nbytes)


def bench_final(ops):
    B, H, C = 32, 256, 128
    act = torch.randn(B, H, H, C, device=dev, dtype=F16)
    w = torch.randn(16, C, 3, 3, device=dev)
    w[3:] = 0
    wp = ops.pack_conv_weight(w)
    bias = torch.zeros(16, device=dev)
    out = torch.empty(B, 3, H, H, device=dev)
    f = lambda: ops.conv_igemm(act, B, H, H, C, 0, C, wp, 16, 3, 3, 0, bias, None, out, None, (3 * H * H, H, 1),
                               out_sc=H * H, n_valid=3)
    ms = timeit([f])
    report("final_conv 3x3 128->3 (N=16) 256x256 b32", ms, ms, B * H * H * C * 2 + B * 3 * H * H * 4)


def bench_stem(ops):
    B, H = 32, 256
    a, b = torch.randn(B, 3, H, H, device=dev), torch.randn(B, 3, H, H, device=dev)
    out = torch.empty(B, H, H, 128, device=dev, dtype=F16)
    f = lambda: ops.stem_unroll(a, 3, b, 3, B, H, H, out)
    ms = timeit([f])
    report("stem_unroll 6ch -> 128-wide f16, 256x256 b32", ms, ms, B * H * H * (6 * 4 + 128 * 2))


def main():
    _native.load()
    ops = ops_mod.get_ops()
    which = sys.argv[1:] or ["gn", "ln", "linear", "quantile", "attn", "cast", "final", "stem"]
    table = {"gn": bench_gn, "ln": bench_ln, "linear": bench_linear, "quantile": bench_quantile, "attn": bench_attn,
             "cast": bench_cast, "final": bench_final, "stem": bench_stem}
    with torch.no_grad():
        for w in which:
            table[w](ops)


if __name__ == "__main__":
    main()
