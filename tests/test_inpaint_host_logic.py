"""Host-side logic of the inpainting path that needs no GPU: the DDIM / dilate / provoke schedules, the work
list of `src/generation/inpaint.py` (output paths, sort order, per-rank contiguous slices) and the batching of work items."""
import os
import pickle

import numpy as np
import pytest


def test_schedules_match_reference_constants():
    from coma_b200.inpaint.pipeline import DDIMSchedule, default_adaptive_mask_settings
    s = DDIMSchedule()
    ts, ratio = s.timesteps(50, 0.98)
    assert ts == [961 - 20 * j for j in range(49)] and ratio == 20              # 981 - 20j with j = 0 dropped at strength 0.98
    assert s.timesteps(50, 1.0)[0][0] == 981 and len(s.timesteps(50, 1.0)[0]) == 50
    st = default_adaptive_mask_settings(50)
    # src/generation/inpaint.py:112-132: dilate iterations 20,10,5,4,3,2,1 for five steps each, then 0
    assert [st.dilate_scheduler(i) for i in range(49)] == sum(([k] * 5 for k in (20, 10, 5, 4, 3, 2, 1)), []) + [0] * 14
    # provoke on the 1-indexed steps {2,4,...,40,45}: 21 adapt calls in 49 steps
    assert [i + 1 for i in range(49) if st.provoke_scheduler(i)] == list(range(2, 41, 2)) + [45]


def test_ddim_alphas_follow_scaled_linear_betas():
    from coma_b200.inpaint.pipeline import DDIMSchedule
    s = DDIMSchedule()
    betas = np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=np.float64) ** 2
    ac = np.cumprod(1.0 - betas)
    for t in (961, 501, 21, 1):
        a_t, a_prev = s.alphas(t, 20)
        assert abs(a_t - ac[t]) < 2e-6 * ac[t]                                   # fp32 cumprod like diffusers vs fp64
        assert abs(a_prev - (ac[t - 20] if t >= 20 else ac[0])) < 2e-6           # set_alpha_to_one=False: final alpha = alphas_cumprod[0]
    assert s.alphas(1, 20)[1] == s.alphas_cumprod[0]


def _make_tree(root, n_views=3, n_masks=2, prompts=("a person carries the backpack", "a person holds the backpack")):
    from PIL import Image
    render, mask, seg, prm = (os.path.join(root, d) for d in ("render", "mask", "seg", "prompts"))
    for v in range(n_views):
        d = f"{render}/behave/backpack/asset0"
        os.makedirs(d, exist_ok=True)
        Image.fromarray(np.zeros((8, 8, 3), np.uint8)).save(f"{d}/view:{v:05}.png")
        md = f"{mask}/behave/backpack/asset0/view:{v:05}"
        os.makedirs(md, exist_ok=True)
        for m in range(n_masks):
            Image.fromarray(np.zeros((8, 8), np.uint8)).save(f"{md}/{m:03}.png")
        with open(f"{mask}/behave/backpack/asset0/view:{v:05}.pickle", "wb") as fh:
            pickle.dump(dict(valid_mask_ids=[f"{m:03}" for m in range(n_masks)]), fh)
    os.makedirs(f"{prm}/behave/backpack/asset0", exist_ok=True)
    with open(f"{prm}/behave/backpack/asset0/prompts.pickle", "wb") as fh:
        pickle.dump(dict(prompts=list(prompts)), fh)
    return render, mask, seg, prm


def test_work_list_paths_order_batches_and_rank_slices(tmp_path):
    from coma_b200 import dist as cdist
    from coma_b200.cli.inpaint import enumerate_work, group_batches
    render, mask, seg, prm = _make_tree(str(tmp_path))
    defaults = dict(ddim_steps=50, cfg_scale=11.0, strength=0.98, enforce_full_mask_ratio=0.0, human_detection_thres=0.015)
    items = enumerate_work(4, ["behave"], ["backpack"], render, mask, seg, prm, str(tmp_path / "out"), "ugly", defaults)
    assert len(items) == 3 * 2 * 2 * 4                                           # views x masks x prompts x images per combination
    paths = [it["result_save_pth"] for it in items]
    assert paths == sorted(paths)                                                 # the reference sorts the work list by output path
    first = items[0]
    # results/generation/inpaintings/<sc>/<c>/<asset>/<view>/<mask_id>/<prompt>/<id:06>.png (src/generation/inpaint.py:235-236)
    assert first["result_save_pth"].endswith("behave/backpack/asset0/view:00000/000/a person carries the backpack/000000.png")
    assert first["cfg_scale"] == 11.0 and first["strength"] == 0.98 and first["input_negprompt"] == "ugly"
    # batching: items that differ only in inpaint_id (the RNG seed) run as one pipeline call
    batches = group_batches(items, 8)
    assert [len(b) for b in batches] == [4] * 12
    assert all(len({(b["asset_render_pth"], b["asset_mask_pth"], b["input_prompt"]) for b in batch}) == 1 for batch in batches)
    assert [len(b) for b in group_batches(items, 3)][:2] == [3, 1]
    # per-rank slices: contiguous, [idx*(len//N+1), (idx+1)*(len//N+1)) like src/generation/inpaint.py:272-278, covering every item once
    for world in (1, 2, 3, 8, 50):
        seen, idx = [], list(range(len(items)))
        for r in range(world):
            lo, hi = cdist.work_item_slice(len(items), r, world)
            step = len(items) // world + 1
            assert idx[lo:hi] == idx[r * step:(r + 1) * step]                    # the reference's python slice, element for element
            seen += idx[lo:hi]
        assert seen == idx


def test_inpaint_human_writes_pngs_in_the_background(tmp_path):
    """`inpaint_human` with a stand-in pipeline (no GPU): every work item of the rank's slice ends up as a complete PNG at the reference's
    path, written by the background writer; `--skip_done` skips them on a second run; a failing write surfaces as an exception and leaves no
    truncated file behind."""
    import torch
    from PIL import Image
    from types import SimpleNamespace
    from coma_b200.cli import inpaint as cli
    render, mask, seg, prm = _make_tree(str(tmp_path), n_views=2, n_masks=1)
    defaults = dict(ddim_steps=50, cfg_scale=11.0, strength=0.98, enforce_full_mask_ratio=0.0, human_detection_thres=0.015)
    calls = []

    class FakePipeline:
        dev = torch.device("cpu")

        def __call__(self, generator, batch_size, **kw):
            calls.append(batch_size)
            imgs = [Image.fromarray(np.full((16, 16, 3), int(g.initial_seed()) % 251, np.uint8)) for g in generator]
            return SimpleNamespace(images=imgs)

    args = (3, ["behave"], ["backpack"], render, mask, seg, prm, str(tmp_path / "out"), "ugly", defaults)
    n = cli.inpaint_human(FakePipeline(), lambda text: None, *args, skip_done=True, verbose=False, parallel_num=1, parallel_idx=0, batch_size=8)
    assert n == 2 * 1 * 2 * 3 and calls == [3] * 4
    items = cli.enumerate_work(*args)
    for it in items:
        img = np.asarray(Image.open(it["result_save_pth"]))
        assert img.shape == (16, 16, 3) and (img == it["inpaint_id"] % 251).all()
        assert not [f for f in os.listdir(it["result_save_dir"]) if ".tmp" in f]
    assert cli.inpaint_human(FakePipeline(), lambda text: None, *args, skip_done=True, verbose=False, parallel_num=1, parallel_idx=0) == 0

    class Broken:
        def save(self, path, format=None):
            with open(path, "wb") as fh:
                fh.write(b"half a file")
            raise OSError("disk full")
    w = cli._PngWriter(workers=2)
    target = str(tmp_path / "broken.png")
    w.submit(Broken(), target)
    with pytest.raises(OSError):
        w.close()
    assert not os.path.exists(target) and not [f for f in os.listdir(tmp_path) if f.startswith("broken.png")]
