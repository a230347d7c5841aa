"""The real-weights path of the 2D HOI stage, end to end on a synthetic asset tree and a synthetic checkpoint directory in the
diffusers layout (unet/ vae/ text_encoder/ tokenizer/ with config.json + safetensors — random weights at toy width, real file formats):
`src/generation/inpaint.py` flags -> enumerate_work -> set_pipeline (segmenter resolved first, configs read from config.json, safetensors
loader) -> CLIP text encoder on the B200 kernels -> batched adaptive-mask loop -> PNG files at the reference's paths
(results/generation/inpaintings/<sc>/<c>/<asset>/<view>/<mask>/<prompt>/<id:06>.png, src/generation/inpaint.py:235-236), the
contiguous work-list slice (:272-278) and --skip_done / --no_skip_done."""
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs, n = bs[:], 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return [chr(c) for c in cs]


def _make_checkpoint(d):
    from safetensors.torch import save_file
    from transformers import CLIPTextConfig, CLIPTextModel
    from oracle import sd_oracle as so
    ucfg, vcfg = so.tiny_unet_cfg(), so.tiny_vae_cfg()
    for sub in ("unet", "vae", "text_encoder", "tokenizer"):
        os.makedirs(os.path.join(d, sub))
    save_file({k: v.half().contiguous() for k, v in so.make_unet_state_dict(0, ucfg).items()}, os.path.join(d, "unet", "diffusion_pytorch_model.safetensors"))
    json.dump(dict(in_channels=9, out_channels=4, block_out_channels=list(ucfg["block_out_channels"]), layers_per_block=2,
                   attention_head_dim=ucfg["heads"], cross_attention_dim=ucfg["cross_attention_dim"], norm_num_groups=ucfg["groups"],
                   down_block_types=["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"]), open(os.path.join(d, "unet", "config.json"), "w"))
    save_file({k: v.half().contiguous() for k, v in so.make_vae_state_dict(1, vcfg).items()}, os.path.join(d, "vae", "diffusion_pytorch_model.safetensors"))
    json.dump(dict(in_channels=3, latent_channels=4, block_out_channels=list(vcfg["block_out_channels"]), layers_per_block=2,
                   norm_num_groups=vcfg["groups"], scaling_factor=0.18215), open(os.path.join(d, "vae", "config.json"), "w"))
    chars = _bytes_to_unicode()
    vocab = {}
    for c in chars:
        vocab[c] = len(vocab)
    for c in chars:
        vocab[c + "</w>"] = len(vocab)
    merges = ["#version: 0.2"]
    for a, b in (("t", "h"), ("th", "e</w>"), ("b", "a"), ("c", "k"), ("ba", "ck")):
        merges.append(f"{a} {b}")
        vocab[a + b] = len(vocab)
    vocab["<|startoftext|>"] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    json.dump(vocab, open(os.path.join(d, "tokenizer", "vocab.json"), "w"))
    open(os.path.join(d, "tokenizer", "merges.txt"), "w").write("\n".join(merges) + "\n")
    json.dump({"model_max_length": 77, "bos_token": "<|startoftext|>", "eos_token": "<|endoftext|>", "pad_token": "<|endoftext|>",
               "unk_token": "<|endoftext|>"}, open(os.path.join(d, "tokenizer", "tokenizer_config.json"), "w"))
    tcfg = dict(vocab_size=len(vocab), hidden_size=ucfg["cross_attention_dim"], intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                max_position_embeddings=77, layer_norm_eps=1e-5, hidden_act="quick_gelu")
    torch.manual_seed(0)
    save_file({k: v.contiguous() for k, v in CLIPTextModel(CLIPTextConfig(**tcfg)).state_dict().items() if v.is_floating_point()},
              os.path.join(d, "text_encoder", "model.safetensors"))
    json.dump(tcfg, open(os.path.join(d, "text_encoder", "config.json"), "w"))


def _make_assets(root, views=("view:00000", "view:00001"), size=64):
    from PIL import Image
    rng = np.random.default_rng(0)
    for sub in ("asset_renders", "asset_masks", "asset_segs"):
        os.makedirs(f"{root}/{sub}/BEHAVE/backpack/asset0", exist_ok=True)
    os.makedirs(f"{root}/prompts/BEHAVE/backpack/asset0")
    pickle.dump(dict(prompts=["the backpack"]), open(f"{root}/prompts/BEHAVE/backpack/asset0/prompts.pickle", "wb"))
    for v in views:
        Image.fromarray(rng.integers(0, 256, (size, size, 3), dtype=np.uint8)).save(f"{root}/asset_renders/BEHAVE/backpack/asset0/{v}.png")
        Image.fromarray(rng.integers(0, 256, (size, size, 3), dtype=np.uint8)).save(f"{root}/asset_segs/BEHAVE/backpack/asset0/{v}.png")
        os.makedirs(f"{root}/asset_masks/BEHAVE/backpack/asset0/{v}")
        m = np.zeros((size, size), np.uint8)
        m[8:56, 16:48] = 255
        Image.fromarray(m).save(f"{root}/asset_masks/BEHAVE/backpack/asset0/{v}/mask:00000.png")
        pickle.dump(dict(valid_mask_ids=["mask:00000"]), open(f"{root}/asset_masks/BEHAVE/backpack/asset0/{v}.pickle", "wb"))


def _cli(root, ckpt, *extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "src", "generation", "inpaint.py"), "--supercategories", "BEHAVE", "--categories", "backpack",
           "--asset_render_dir", f"{root}/asset_renders", "--asset_mask_dir", f"{root}/asset_masks", "--asset_seg_dir", f"{root}/asset_segs",
           "--prompts_dir", f"{root}/prompts", "--save_dir", f"{root}/inpaintings", "--model_dir", ckpt, "--num_img_per_combination", "3",
           "--default_ddim_steps", "50", "--batch_size", "3"] + list(extra)   # 50 steps: the provoke schedule is hard-coded up to step 45 (:125-129)
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, **(env or {})), timeout=900)


def test_inpaint_cli_end_to_end_with_synthetic_checkpoint(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from PIL import Image
    root, ckpt = str(tmp_path / "results"), str(tmp_path / "ckpt")
    os.makedirs(root)
    _make_checkpoint(ckpt)
    _make_assets(root)
    # default segmenter type "p" (PointRend) is a plug-in: must fail fast and say how to provide one — before any weights load
    r = _cli(root, ckpt)
    assert r.returncode != 0 and "--segmenter module:factory" in r.stderr and "NotImplementedError" in r.stderr
    # rank 0 of 2: the reference's contiguous slice (6 items: 2 views x 1 mask x 1 prompt x 1 augmentation x 3 ids -> sub = 6 // 2 + 1 = 4)
    r = _cli(root, ckpt, "--adaptive_mask_model_type", "stub", "--parallel_num", "2", "--parallel_idx", "0")
    assert r.returncode == 0, r.stderr[-3000:]
    pngs = sorted(p for p in (os.path.join(dp, f) for dp, _, fs in os.walk(f"{root}/inpaintings") for f in fs) if p.endswith(".png"))
    assert len(pngs) == 4
    assert pngs[0].endswith("inpaintings/BEHAVE/backpack/asset0/view:00000/mask:00000/the backpack/000000.png")
    img = np.asarray(Image.open(pngs[0]))
    assert img.shape == (64, 64, 3) and img.dtype == np.uint8 and img.std() > 0
    # rank 1 finishes the list; --skip_done leaves rank 0's files alone, a plug-in factory is accepted
    t0 = os.path.getmtime(pngs[0])
    r = _cli(root, ckpt, "--adaptive_mask_model_type", "p", "--segmenter", "tests.test_gpu_inpaint_cli:luminance_factory", "--parallel_num", "2",
             "--parallel_idx", "1", env={"PYTHONPATH": ROOT})
    assert r.returncode == 0, r.stderr[-3000:]
    pngs = sorted(p for p in (os.path.join(dp, f) for dp, _, fs in os.walk(f"{root}/inpaintings") for f in fs) if p.endswith(".png"))
    assert len(pngs) == 6 and os.path.getmtime(pngs[0]) == t0
    assert pngs[-1].endswith("view:00001/mask:00000/the backpack/000002.png")
    # same seeds, same inputs -> --no_skip_done regenerates bit-identical images
    before = np.asarray(Image.open(pngs[0])).copy()
    r = _cli(root, ckpt, "--adaptive_mask_model_type", "stub", "--no_skip_done", "--parallel_num", "2", "--parallel_idx", "0")
    assert r.returncode == 0, r.stderr[-3000:]
    assert os.path.getmtime(pngs[0]) > t0 and np.array_equal(np.asarray(Image.open(pngs[0])), before)


def luminance_factory(adaptive_mask_model_type, pointrend_threshold):
    from coma_b200.inpaint.segmenter import LuminanceSegmenter
    return LuminanceSegmenter(128)
