import os, sys, torch
sys.path.insert(0, "/root/repo")
from coma_b200.inpaint import nn
from coma_b200.inpaint.vae import VAE
from coma_b200.inpaint import synthetic as so  # noqa: E402  (seeded random state dicts)
dev = torch.device("cuda:0"); B = 4
vae = VAE(so.make_vae_state_dict(1), device=dev)
g = torch.Generator(device=dev).manual_seed(0)
z = nn.new_act(B, 64, 64, 4, dev); z.t.copy_(torch.randn((B * 4096, 4), device=dev, generator=g).half())
def ev(fn, n=5):
    fn(); torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
for flag in (False, True, False, True):
    nn.FUSED_GN_STATS = flag
    gr = nn.Graphed(lambda a: vae.decode(a), z)
    print("fused stats", flag, "decode graph", round(ev(gr), 3), "ms; eager", round(ev(lambda: vae.decode(z), 3), 3), "ms; mem", torch.cuda.max_memory_allocated() >> 20, "MB", flush=True)
    del gr
