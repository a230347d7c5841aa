import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200.inpaint import nn
from coma_b200.inpaint.unet import UNet
from coma_b200.inpaint.vae import VAE
from coma_b200.inpaint import synthetic as so  # noqa: E402  (seeded random state dicts)
dev = torch.device("cuda:0"); B = 4
net = UNet(so.make_unet_state_dict(0), device=dev); vae = VAE(so.make_vae_state_dict(1), device=dev)
g = torch.Generator(device=dev).manual_seed(0)
x = nn.new_act(2 * B, 64, 64, 9, dev); x.t.copy_(torch.randn((2 * B * 4096, 9), device=dev, generator=g).half())
ctx = (torch.randn((2 * B * 77, 768), device=dev, generator=g) * 0.02).half(); tt = torch.full((2 * B,), 961.0, device=dev)
kv = net.context_kv(ctx, 77, 2 * B)
img = nn.new_act(B, 512, 512, 3, dev); img.t.copy_(torch.tanh(torch.randn((B * 512 * 512, 3), device=dev, generator=g)).half())
def ev(fn, n=10):
    fn(); torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
for flag in (False, True, False, True):
    nn.FUSED_GN_STATS = flag
    gr = nn.Graphed(lambda a, b, c: net.forward(a, b, c, 77, ctx_kv=kv), x, tt, ctx)
    ge = nn.Graphed(lambda a: vae.encode_moments(a), img)
    print("fused stats", flag, "unet", round(ev(gr), 3), "ms; vae encode", round(ev(ge), 3), "ms", flush=True)
    del gr, ge
