"""Where does the pipeline's wall time go beyond the three CUDA graphs? Sustained replays (power-capped clocks) vs the pipeline
call, with SM clocks sampled by nvidia-smi."""
import os, sys, time, subprocess, threading
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200.inpaint import nn
from coma_b200.inpaint.pipeline import AdaptiveMaskInpaintPipeline, default_adaptive_mask_settings
from coma_b200.inpaint.segmenter import LuminanceSegmenter
from coma_b200.inpaint.unet import UNet
from coma_b200.inpaint.vae import VAE
from coma_b200.inpaint import synthetic as so  # noqa: E402  (seeded random state dicts)

dev = torch.device("cuda:0"); B = 4
clk = []; stop = threading.Event()
def sampler():
    while not stop.is_set():
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
        clk.append((time.time(), o)); stop.wait(0.1)
th = threading.Thread(target=sampler, daemon=True); th.start()
unet = UNet(so.make_unet_state_dict(0), device=dev); vae = VAE(so.make_vae_state_dict(1), device=dev)
g = torch.Generator(device=dev).manual_seed(0)
x = nn.new_act(2 * B, 64, 64, 9, dev); x.t.copy_(torch.randn((2 * B * 4096, 9), device=dev, generator=g).half())
ctx = (torch.randn((2 * B * 77, 768), device=dev, generator=g) * 0.02).half(); tt = torch.full((2 * B,), 961.0, device=dev)
kv = unet.context_kv(ctx, 77, 2 * B)
gr = nn.Graphed(lambda a, b, c: unet.forward(a, b, c, 77, ctx_kv=kv), x, tt, ctx)
def ev(fn, n):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
print("unet graph, 5 replays  :", ev(gr, 5), "ms")
t0 = time.time(); print("unet graph, 150 replays:", ev(gr, 150), "ms (sustained)"); t1 = time.time()
print("clocks during sustained:", [c[1] for c in clk if t0 + 0.5 < c[0] < t1][:12])
pipe = AdaptiveMaskInpaintPipeline(unet, vae); pipe.register_adaptive_mask_model(LuminanceSegmenter(128)); pipe.register_adaptive_mask_settings(default_adaptive_mask_settings(50))
rng = np.random.default_rng(0); image = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8); default = np.zeros((512, 512), np.uint8); default[64:448, 128:384] = 255
pe, ne = torch.randn((77, 768), generator=torch.Generator().manual_seed(1)) * 0.02, torch.zeros((77, 768))
def run():
    gens = [torch.Generator(device=dev).manual_seed(i) for i in range(B)]
    return pipe(image=image, default_mask_image=default, prompt_embeds=pe, negative_prompt_embeds=ne, guidance_scale=11.0, strength=0.98, num_inference_steps=50,
                generator=gens, enforce_full_mask_ratio=0.0, human_detection_thres=0.015, batch_size=B, output_type="np")
run(); torch.cuda.synchronize()
for it in range(3):
    t0 = time.time(); run(); torch.cuda.synchronize(); t1 = time.time()
    print(f"pipeline {it}: {t1 - t0:.3f} s; clocks:", [c[1] for c in clk if t0 < c[0] < t1][::2])
# host-side time of one pipeline call with the GPU work removed from the critical path: profile CPU time
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable(); run(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])
stop.set()
