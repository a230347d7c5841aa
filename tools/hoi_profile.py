"""One UNet forward (B2=8) + one VAE decode/encode (B=4) at full size, for `ncu --metrics gpu__time_duration.sum`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200.inpaint import nn  # noqa: E402
from coma_b200.inpaint.unet import UNet  # noqa: E402
from coma_b200.inpaint.vae import VAE  # noqa: E402
from coma_b200.inpaint import synthetic as so  # noqa: E402

dev = torch.device("cuda:0")
B = 4
what = sys.argv[1] if len(sys.argv) > 1 else "unet"
g = torch.Generator(device=dev).manual_seed(0)
if what == "unet":
    net = UNet(so.make_unet_state_dict(0), device=dev)
    x = nn.new_act(2 * B, 64, 64, 9, dev)
    x.t.copy_(torch.randn((2 * B * 4096, 9), device=dev, generator=g).half())
    ctx = (torch.randn((2 * B * 77, 768), device=dev, generator=g) * 0.02).half()
    tt = torch.full((2 * B,), 961.0, device=dev)
    kv = net.context_kv(ctx, 77, 2 * B)
    net.forward(x, tt, ctx, 77, ctx_kv=kv)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("fwd")
    net.forward(x, tt, ctx, 77, ctx_kv=kv)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
else:
    vae = VAE(so.make_vae_state_dict(1), device=dev)
    z = nn.new_act(B, 64, 64, 4, dev)
    z.t.copy_(torch.randn((B * 4096, 4), device=dev, generator=g).half())
    img = nn.new_act(B, 512, 512, 3, dev)
    img.t.copy_(torch.tanh(torch.randn((B * 512 * 512, 3), device=dev, generator=g)).half())
    vae.decode(z)
    vae.encode_moments(img)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("fwd")
    vae.decode(z)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
