"""UNet forward (B2 = 8) as ONE chain of launches against TWO half-batch chains on two streams inside one CUDA graph: do the chains fill
each other's dependency bubbles?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200.inpaint import nn  # noqa: E402
from coma_b200.inpaint.unet import UNet  # noqa: E402
from coma_b200.inpaint import synthetic as so  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
net = UNet(so.make_unet_state_dict(0), device=dev)
NB = int(os.environ.get("NB", 8))
L = 77


def inputs(b):
    x = nn.new_act(b, 64, 64, 9, dev)
    x.t.copy_(torch.randn((b * 4096, 9), device=dev, generator=g).half())
    ctx = (torch.randn((b * L, 768), device=dev, generator=g) * 0.02).half()
    tt = torch.full((b,), 961.0, device=dev)
    return x, tt, ctx


def time_graph(gr, n=20):
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            gr.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    return best


# one chain
x, tt, ctx = inputs(NB)
kv = net.context_kv(ctx, L, NB)
net.forward(x, tt, ctx, L, ctx_kv=kv)
torch.cuda.synchronize()
g1 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g1):
    out_full = net.forward(x, tt, ctx, L, ctx_kv=kv)
t_one = time_graph(g1)

# two half-batch chains on two streams
halves = [inputs(NB // 2) for _ in range(2)]
for h, (xa, ta, ca) in enumerate(halves):   # same data as the full batch, for a parity check
    xa.t.copy_(x.t[h * (NB // 2) * 4096:(h + 1) * (NB // 2) * 4096])
    ca.copy_(ctx[h * (NB // 2) * L:(h + 1) * (NB // 2) * L])
kvs = [net.context_kv(c, L, NB // 2) for (_, _, c) in halves]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for s, (xa, ta, ca), k in zip(streams, halves, kvs):   # eager warm-up on the side streams (workspaces are per stream)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        net.forward(xa, ta, ca, L, ctx_kv=k)
torch.cuda.synchronize()
g2 = torch.cuda.CUDAGraph()
outs = []
with torch.cuda.graph(g2):
    cur = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(cur)
    for s, (xa, ta, ca), k in zip(streams, halves, kvs):
        with torch.cuda.stream(s):
            outs.append(net.forward(xa, ta, ca, L, ctx_kv=k))
    for s in streams:
        cur.wait_stream(s)
t_two = time_graph(g2)
g1.replay()
g2.replay()
torch.cuda.synchronize()
both = torch.cat(outs, 0)
err = (both.float() - out_full.float()).abs().max().item() / out_full.float().abs().max().item()
print(f"B2={NB}: one chain {t_one:.3f} ms, two half-batch chains {t_two:.3f} ms, max diff / scale {err:.2e}")
