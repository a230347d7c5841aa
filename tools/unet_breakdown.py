"""Per-op timing of one full-size UNet forward (B2 = 8) or VAE decode/encode (B = 4): every C-ABI call is bracketed by
CUDA events on the launching stream and grouped by (entry point, shape signature). torch-side copies (skip concat) show up
as `torch.*`. Eager launches: tiny kernels carry a few microseconds of host gap, so read this for SHARES and for the
per-shape tensor throughput of the big GEMMs/convs; the ncu launch list gives the cold, serialised durations.

    python tools/unet_breakdown.py [unet|vae_decode|vae_encode] [out.json]
"""
import ctypes
import json
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200 import _lib  # noqa: E402
from coma_b200._lib import GemmArgs  # noqa: E402
from coma_b200.inpaint import nn  # noqa: E402
from coma_b200.inpaint.unet import UNet  # noqa: E402
from coma_b200.inpaint.vae import VAE  # noqa: E402
from coma_b200.inpaint import synthetic as so  # noqa: E402  (seeded random state dicts)

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "unet"
B = 4
records = []
orig_call = _lib.call


def signature(name, args):
    if name == "coma_gemm_f16_ex":
        g = ctypes.cast(args[0], ctypes.POINTER(GemmArgs)).contents
        nb = max(g.nb1, 1) * max(g.nb2, 1)
        return f"M={g.M} N={g.N} K={g.K} nb={nb}", 2.0 * g.M * g.N * g.K * nb
    if name == "coma_conv3x3_strided_f16":
        _, Bn, H, W, C, _, st, _, _, _, N = args[:11]
        return f"B={Bn} HWin={H}x{W} Cin={C} Cout={N} stride={st}", 2.0 * Bn * (H // st) * (W // st) * 9 * C * N
    if name in ("coma_conv3x3_f16", "coma_conv3x3_f16_ws"):
        _, Bn, H, W, C, _, _, _, N = args[:9]
        return f"B={Bn} HW={H}x{W} Cin={C} Cout={N}", 2.0 * Bn * H * W * 9 * C * N
    if name == "coma_attention_fwd_f16":
        Bn, heads, S, L, d = args[3:8]
        return f"B={Bn} heads={heads} S={S} L={L} d={d}", 4.0 * Bn * heads * S * L * d
    ints = [a for a in args if isinstance(a, int) and 0 < a < (1 << 24)]
    return " ".join(str(a) for a in ints[:5]), 0.0


def timed_call(name, *args):
    sig, flop = signature(name, args)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig_call(name, *args)
    b.record()
    records.append((name, sig, flop, a, b))


def run():
    g = torch.Generator(device=dev).manual_seed(0)
    if what == "unet":
        net = UNet(so.make_unet_state_dict(0), device=dev)
        x = nn.new_act(2 * B, 64, 64, 9, dev)
        x.t.copy_(torch.randn((2 * B * 4096, 9), device=dev, generator=g).half())
        ctx = (torch.randn((2 * B * 77, 768), device=dev, generator=g) * 0.02).half()
        tt = torch.full((2 * B,), 961.0, device=dev)
        kv = net.context_kv(ctx, 77, 2 * B)     # as in the pipeline: cross-attention K / V once per call, not per step
        return lambda: net.forward(x, tt, ctx, 77, ctx_kv=kv)
    vae = VAE(so.make_vae_state_dict(1), device=dev)
    if what == "vae_decode":
        z = nn.new_act(B, 64, 64, 4, dev)
        z.t.copy_(torch.randn((B * 4096, 4), device=dev, generator=g).half())
        return lambda: vae.decode(z)
    img = nn.new_act(B, 512, 512, 3, dev)
    img.t.copy_(torch.tanh(torch.randn((B * 512 * 512, 3), device=dev, generator=g)).half())
    return lambda: vae.encode_moments(img)


fn = run()
fn()
fn()
torch.cuda.synchronize()
nn.call = timed_call          # nn.py binds `call` at import time
_lib.call = timed_call
reps = 3
for _ in range(reps):
    torch.cuda._sleep(int(80e6))   # ~40 ms of GPU spin: the host enqueues the whole forward meanwhile, so the events
    fn()                           # bracket back-to-back kernels and not the host-side launch preparation
    torch.cuda.synchronize()
nn.call = orig_call
_lib.call = orig_call
gr = nn.Graphed(fn)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gr()
t0.record()
for _ in range(5):
    gr()
t1.record()
torch.cuda.synchronize()
total_ms = t0.elapsed_time(t1) / 5
agg = OrderedDict()
for name, sig, flop, a, b in records:
    k = (name.replace("coma_", ""), sig)
    e = agg.setdefault(k, [0, 0.0, 0.0])
    e[0] += 1
    e[1] += a.elapsed_time(b)
    e[2] += flop
rows = []
for (name, sig), (n, ms, flop) in agg.items():
    rows.append(dict(op=name, shape=sig, calls=n // reps, ms=ms / reps, tflops=(flop / reps) / (ms / reps) / 1e9 if flop else None))
rows.sort(key=lambda r: -r["ms"])
in_calls = sum(r["ms"] for r in rows)
by_op = {}
for r in rows:
    by_op[r["op"]] = by_op.get(r["op"], 0.0) + r["ms"]
print(f"{what}: {total_ms:.3f} ms per forward (CUDA graph replay), {in_calls:.3f} ms inside C-ABI calls, {len(records) // reps} calls")
for k, v in sorted(by_op.items(), key=lambda kv: -kv[1]):
    print(f"  {k:32s} {v:8.3f} ms  {100 * v / total_ms:5.1f}%")
print("| op | shape | calls | ms | TFLOP/s |\n|---|---|---|---|---|")
for r in rows[:70]:
    tf = f"{r['tflops']:.0f}" if r["tflops"] else ""
    print(f"| {r['op']} | {r['shape']} | {r['calls']} | {r['ms']:.3f} | {tf} |")
if len(sys.argv) > 2:
    json.dump(dict(what=what, total_ms=total_ms, by_op=by_op, rows=rows), open(sys.argv[2], "w"), indent=1)
