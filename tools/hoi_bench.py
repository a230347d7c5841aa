"""Full-size HP-A timing on one GPU: SD-1.5 inpainting UNet / VAE with seeded random weights (no checkpoints are
available offline), synthetic 512x512 render + rectangular default mask, deterministic stub segmenter."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_b200 import _lib  # noqa: E402
from coma_b200.inpaint import nn  # noqa: E402
from coma_b200.inpaint.pipeline import AdaptiveMaskInpaintPipeline, default_adaptive_mask_settings  # noqa: E402
from coma_b200.inpaint.segmenter import LuminanceSegmenter  # noqa: E402
from coma_b200.inpaint.unet import UNet  # noqa: E402
from coma_b200.inpaint.vae import VAE  # noqa: E402
from coma_b200.inpaint import synthetic as so  # noqa: E402  (seeded random state dicts)

dev = torch.device("cuda:0")
B = int(os.environ.get("B", 4))
steps = int(os.environ.get("STEPS", 50))


def ev_time(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


t0 = time.time()
unet = UNet(so.make_unet_state_dict(0), device=dev)
vae = VAE(so.make_vae_state_dict(1), device=dev)
print(f"weights ready in {time.time() - t0:.1f}s", flush=True)
g = torch.Generator(device=dev).manual_seed(0)
x = nn.new_act(2 * B, 64, 64, 9, dev)
x.t.copy_(torch.randn((2 * B * 4096, 9), device=dev, generator=g).half())
ctx = (torch.randn((2 * B * 77, 768), device=dev, generator=g) * 0.02).half()
tt = torch.full((2 * B,), 961.0, device=dev)
out = {}
l0 = _lib.launch_count()
ms = ev_time(lambda: unet.forward(x, tt, ctx, 77))
out["unet_ms"] = ms
out["unet_tflops"] = 0.803 * 2 * B / ms * 1e3 / 1e3
out["unet_launches"] = (_lib.launch_count() - l0) // 4
gr = nn.Graphed(lambda a, b, c: unet.forward(a, b, c, 77), x, tt, ctx)
out["unet_graph_ms"] = ev_time(gr, n=5)
z = nn.new_act(B, 64, 64, 4, dev)
z.t.copy_(torch.randn((B * 4096, 4), device=dev, generator=g).half())
ms = ev_time(lambda: vae.decode(z), n=2)
out["vae_decode_ms"] = ms
out["vae_decode_tflops"] = 2.515 * B / ms
gd = nn.Graphed(lambda a: vae.decode(a), z)
out["vae_decode_graph_ms"] = ev_time(gd, n=3)
del gd
img = nn.new_act(B, 512, 512, 3, dev)
img.t.copy_(torch.tanh(torch.randn((B * 512 * 512, 3), device=dev, generator=g)).half())
ms = ev_time(lambda: vae.encode_moments(img), n=2)
out["vae_encode_ms"] = ms
out["vae_encode_tflops"] = 1.117 * B / ms
print(json.dumps(out), flush=True)

pipe = AdaptiveMaskInpaintPipeline(unet, vae)
pipe.register_adaptive_mask_model(LuminanceSegmenter(128))
pipe.register_adaptive_mask_settings(default_adaptive_mask_settings(steps))
rng = np.random.default_rng(0)
image = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
default = np.zeros((512, 512), np.uint8)
default[64:448, 128:384] = 255
pe, ne = torch.randn((77, 768), generator=torch.Generator().manual_seed(1)) * 0.02, torch.zeros((77, 768))
for it in range(2):
    gens = [torch.Generator(device=dev).manual_seed(i) for i in range(B)]
    torch.cuda.synchronize()
    t0 = time.time()
    res = pipe(image=image, default_mask_image=default, prompt_embeds=pe, negative_prompt_embeds=ne, guidance_scale=11.0, strength=0.98,
               num_inference_steps=steps, generator=gens, enforce_full_mask_ratio=0.0, human_detection_thres=0.015, batch_size=B,
               output_type="np")
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"pipeline run {it}: {dt:.2f} s for {B} images -> {B / dt:.3f} images/s", flush=True)
out["pipeline_s"] = dt
out["images_per_s"] = B / dt
out["max_mem_gb"] = torch.cuda.max_memory_allocated() / 2**30
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/hoi_bench.json", "w"), indent=1)
