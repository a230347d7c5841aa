#!/usr/bin/env python
"""bench.py — ComA vertex-pairs/s on B200 (BASELINE.json metric, contact-extraction leg) + the occupancy and HOI legs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Headline workload (BASELINE.json configs[3], "ComA extraction: 2048 synthetic 3D HOI samples, SMPL-X 10475 x object 1500
verts, 8xB200"): the job aggregates 256 x N samples of H=10475 x O=1500 vertex pairs through ALL of
aggregate_single_sample_for_contact — pair distance / contact count / proximity (K2) and both 250-bin orientation
histograms (K3). Multi-GPU = ROW sharding (SURVEY 8e "zero-collective" form): rank r owns the human-vertex rows
human_slice(H, r, N) of every accumulator and sees ALL 256 x N samples, so the per-GPU work (256 x H x O vertex pairs) is
fixed as N grows (weak scaling; 8 ranks = the 2048 samples of the config) and no accumulator ever crosses GPUs; the only
exchange is the all-gather of the staged samples and the per-vertex maps of the read-out, both inside `e2e`.
A *vertex-pair* is one (sample, human vertex, object vertex) triple.

`value`   : device-resident inputs, CUDA-event timed, max over ranks.
`e2e`     : the same job through the reference-facing class API with HOST (numpy fp64) samples, each rank loading 1/N of them:
            ComA(human_slice) -> register_sample_to_cache x 256 -> aggregate_all_samples(exchange) (H2D + NVLink all-gather
            inside) -> get_aggregated_contact() -> numpy on the host.
`roofline`: the dominant kernel (K3) timed live on its stream; `roofline_sfu` its binding (MUFU) roofline; `roofline_k2_stream`
            the HBM-bound streaming form of K2 the north star sets its 90 % target on.
`occupancy`: BASELINE configs[4] (128^3 voxels, 4096 samples, H-sharded: 1310 rows per GPU): K4 scatter + K5c read-out + MAX
            all-reduce, device-timed and end to end.
`hoi`     : BASELINE configs[1] (N < 8: one viewpoint x batch 4 per rank) / configs[2] (N = 8: 36 viewpoints x batch 8, the
            reference's contiguous slice rule) through the public pipeline call.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference classes (oracle/_ref, staged by oracle/make_ref.py) on the host
            cores over a bounded sample of the same workload; `reference_cuda` = the same classes with device="cuda" on one
            B200 (the reference's production setting) — the denominator of the north star's ">= 20x" target.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, O, N = 10475, 1500, 250
S_PER_RANK = 256
PRESET = dict(spatial_grid_size=0.15, spatial_grid_thres=0.05, normal_gaussian_sigma=0.25, eps=1e-10,
              significant_contact_ratio=0.1)  # constants/coma/qual.py "qual:backpack_object_contact"
# DRAM traffic of one launch from `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum): the grids are read and written
# once per launch whatever the sample count; inputs are 0.3 MB / sample. Updated by the round-2 capture (profiles/r02_k3_full.md).
K3_NCU_DRAM_BYTES = 31.467172e9 + 31.378264e9
# K2 streaming kernel (one sample per launch), profiles/r01_k2_full.md: 125.9 MB read + 70.2 MB written while the capture ran (the
# rest of the 125.7 MB of accumulator writes is still dirty in the 126 MB L2 when the single profiled launch ends)
K2_NCU_DRAM_BYTES = 125.852928e6 + 70.227456e6
K3_MUFU_PER_EVAL = 2.0     # DENSE kernel: MUFU.SQRT + MUFU.EX2 per bin evaluation (cuobjdump)
MUFU_CLK_PER_WARP_INSTR = 8.05  # measured on B200 with tools/ubench_pipes.cu (4 lanes/clk per SM sub-partition)
# CONE-LIMITED kernel (the default): what binds it is the FP32 "heavy" pipe, the only one that executes the packed FFMA2 / FMUL2 /
# FADD2 (ncu: sm__pipe_fmaheavy_cycles_active 64 % vs 36 % for the whole FMA pipe, profiles/r02_k3cone_full.md).
K3C_EXEC_FRACTION = 0.748       # executed / algorithmic bin evaluations on this seeded workload: MUFU.EX2 warp-instructions x 32 lanes
                                # / (2 x 250 x pair-samples) from the ncu capture of the same generator (profiles/r02_k3cone_full.md)
K3C_PACKED_PER_STEP = 10.0      # FFMA2-class instructions per step of 64 evaluations at degree 5: 3 (dot) + 5 (Horner) + FMUL2 + FADD2
FFMA2_CLK_PER_WARP_INSTR = 2.1  # tools/ubench_pipes.cu
OCC = dict(Sg=128, S=4096, H_per_rank=1310, tol=3.0)   # BASELINE configs[4]


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.stop, self.th = gpu_index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if "bf16_tflops_sustained" in d:
            return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained: the loop runs for ~1 s)"
    return 1410.0, "fallback (B200_PROFILING.md sustained dense bf16)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------------------------ reference arms
def _reference_contact_once(ref, device, hs, n_samples, seed=123):
    """The UNMODIFIED reference ComA on `hs` human rows x O object vertices x N bins, `n_samples` samples -> seconds for
    register + aggregate_all_samples (utils/coma.py:253-323), i.e. the work of hs * O * n_samples vertex pairs."""
    import torch
    from coma_b200 import synth
    samples = synth.make_samples(n_samples, hs, O, seed=seed)
    c = ref.ComA(human_res=hs, obj_res=O, normal_res=N, spatial_res=0,
                 proximity_settings=dict(spatial_grid_size=PRESET["spatial_grid_size"], spatial_grid_thres=PRESET["spatial_grid_thres"]),
                 normal_gaussian_sigma=PRESET["normal_gaussian_sigma"], eps=PRESET["eps"], device=device)
    if device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in samples:
        c.register_sample_to_cache(**s)
    c.aggregate_all_samples()
    if device != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    del c
    return dt


def _quiet(fn, *a, **kw):
    """The reference prints tqdm bars to stderr; keep the single JSON line on stdout clean (stdout is untouched anyway)."""
    devnull = open(os.devnull, "w")
    old = sys.stderr
    sys.stderr = devnull
    try:
        return fn(*a, **kw)
    finally:
        sys.stderr = old
        devnull.close()


def cpu_reference_rate(target_seconds=12.0):
    """`cpu_baseline`: the reference's own CPU implementation of the path on all host cores, bounded sample.
    kind "reference": the unmodified classes from oracle/_ref (torch CPU, intra-op threads = host cores), one cfg-4 sample
    restricted to `hs` human rows; kind "port" (only if oracle/_ref was not staged): the C/OpenMP oracle."""
    cores = host_cores()
    from oracle import ref_loader
    if ref_loader.available():
        import torch
        torch.set_num_threads(cores)
        ref = ref_loader.load()
        hs = 16
        probe = _quiet(_reference_contact_once, ref, "cpu", hs, 1)
        hs = int(min(512, max(16, hs * target_seconds / max(probe, 1e-3)))) // 8 * 8   # its [hs,O,N,3] fp64 temporary is 9 MB per row
        dt = _quiet(_reference_contact_once, ref, "cpu", hs, 1)
        return dict(value=hs * O / dt, unit="vertex-pairs/s", cores=cores, kind="reference", step_ms=dt * 1e3,
                    sample=f"unmodified reference ComA(device='cpu'), 1 sample x {hs} of {H} human rows x {O} object verts x {N} bins "
                           f"(register + aggregate_all_samples), {dt:.1f} s, torch {torch.__version__} with {cores} intra-op threads")
    from coma_b200 import synth
    from oracle import oracle
    oracle.set_num_threads(cores)
    cores = oracle.num_threads()
    hv, hn, ov, on = synth.make_sample_arrays(1, H, O, seed=123)
    grid = oracle.fibonacci_sphere(N)

    def run(hs):
        t0 = time.perf_counter()
        oracle.pair_accumulate(hv[:, :hs], ov, PRESET["spatial_grid_thres"], PRESET["spatial_grid_size"])
        oracle.orient_accumulate(hn[:, :hs], on, grid, PRESET["normal_gaussian_sigma"], PRESET["eps"])
        return time.perf_counter() - t0
    hs = int(min(H, max(64, 64 * target_seconds / max(run(64), 1e-3)))) // 8 * 8
    dt = run(hs)
    return dict(value=hs * O / dt, unit="vertex-pairs/s", cores=cores, kind="port", step_ms=dt * 1e3,
                sample=f"1 sample x {hs} of {H} human rows x {O} object verts x {N} bins (K2+K3), {dt:.1f} s on {cores} OpenMP threads")


def reference_cuda_rate(hs=1024, n_samples=3):
    """The unmodified reference with device="cuda" on ONE B200 (src/coma/extract_coma.py:329): `hs` of the 10475 human rows
    (its [hs,O,N,3] fp64 temporary is 9.2 GB at hs = 1024; the full 10475 rows would need 94 GB), per-sample loop as shipped."""
    import torch
    from oracle import ref_loader
    if not (ref_loader.available() and torch.cuda.is_available()):
        return None
    ref = ref_loader.load()
    try:
        _quiet(_reference_contact_once, ref, "cuda", hs, 1)                 # warm-up (allocator, kernels)
        dt = _quiet(_reference_contact_once, ref, "cuda", hs, n_samples)
    except Exception as ex:  # informative leg only
        return {"error": repr(ex)[:200]}
    finally:
        torch.cuda.empty_cache()
    return dict(value=hs * O * n_samples / dt, unit="vertex-pairs/s", kind="reference", device=torch.cuda.get_device_name(0),
                sample=f"unmodified reference ComA(device='cuda'), {n_samples} samples x {hs} of {H} human rows x {O} x {N} bins, "
                       f"{dt * 1e3 / n_samples:.0f} ms per sample, torch {torch.__version__} eager")


def hoi_reference(dev):
    """Reference-shaped inpainting loop in torch eager fp16 on the same GPU: batch 1, decode every step,
    cv2 on the host; once with F.scaled_dot_product_attention (what a current torch would dispatch) and once unfused
    (baddbmm + softmax + bmm, what the reference's pinned torch 1.13 without xformers runs)."""
    out = {}
    try:
        from oracle.inpaint_loop_oracle import time_reference_loop
        for name, sdpa in (("sdpa", True), ("unfused", False)):
            s = time_reference_loop(dev, sdpa=sdpa)
            out[name] = {"images_per_s": 1.0 / s, "s_per_image": s}
        out["what"] = ("reference-shaped loop (batch 1, VAE decode every step, cv2.dilate on the host) with the restated SD-1.5 models in "
                       "torch eager fp16 on the same GPU; 'sdpa' = F.scaled_dot_product_attention, 'unfused' = the reference's torch-1.13 attention")
    except Exception as ex:
        out["error"] = repr(ex)[:200]
    return out


def run_reference(args, rank):
    """`--impl reference`: the reference's own CPU implementation on the host cores (rank 0 only), K steps of a bounded
    sample each; plus (informative keys) the same classes on one B200 and the reference-shaped HOI loop."""
    if rank != 0:
        return
    for _ in range(max(args.warmup, 1) - 1):
        cpu_reference_rate(target_seconds=2.0)
    rates = [cpu_reference_rate(target_seconds=max(4.0, 60.0 / max(args.steps, 1))) for _ in range(args.steps)]
    best = rates[int(np.argsort([r["value"] for r in rates])[len(rates) // 2])]
    line = {
        "impl": "reference", "metric": "ComA vertex-pairs/s", "value": best["value"], "unit": "vertex-pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": best.get("step_ms"), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"coma_contact cfg4-shape H={H} O={O} N={N}, K2+K3 (reference CPU path, kind={best['kind']})"},
        "cpu_baseline": best,
        "e2e": {"value": best["value"], "unit": "vertex-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        import torch
        if torch.cuda.is_available() and not args.no_reference_cuda:
            torch.cuda.set_device(0)
            line["reference_cuda"] = reference_cuda_rate()
            if not args.no_hoi:
                line["hoi"] = hoi_reference(torch.device("cuda", 0))
    except Exception as ex:
        line["reference_cuda"] = {"error": repr(ex)[:200]}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------ HOI leg
def hoi_leg(args, dev, rank, world, barrier):
    """BASELINE.json configs[1] (world < 8): adaptive-mask SD inpainting, one viewpoint x batch 4 per rank; configs[2] (world = 8):
    36 viewpoints x batch 8 sharded over the ranks with the reference's contiguous slice rule (src/generation/inpaint.py:272-278,
    sub = 36 // 8 + 1 = 5 views on ranks 0-6, one on rank 7). 512x512, 50 DDIM steps (strength 0.98 -> 49 UNet x2 evaluations,
    CFG 11, 21 adapt calls), the batch = seeds of one (render, mask, prompt). No collective: outputs are images. Seeded random
    weights with the real architecture (no checkpoints offline), synthetic renders, rectangular default mask, deterministic stub
    segmenter. images/s = finished 512x512 outputs per second, whole job, wall clock around the public pipeline calls (host image
    in, host images out), max over ranks."""
    import torch
    import torch.distributed as dist
    from coma_b200 import _lib
    from coma_b200 import dist as cdist
    from coma_b200.inpaint.pipeline import AdaptiveMaskInpaintPipeline, default_adaptive_mask_settings
    from coma_b200.inpaint.segmenter import LuminanceSegmenter
    from coma_b200.inpaint.unet import UNet
    from coma_b200.inpaint.vae import VAE
    from coma_b200.inpaint import synthetic as so   # seeded random state dicts with diffusers key names (no checkpoints offline)
    cfg3 = world == 8 and not args.hoi_cfg2
    B = 8 if cfg3 else args.hoi_batch
    n_views = 36 if cfg3 else world
    v0, v1 = cdist.work_item_slice(n_views, rank, world) if cfg3 else (rank, rank + 1)
    torch.cuda.empty_cache()
    pipe = AdaptiveMaskInpaintPipeline(UNet(so.make_unet_state_dict(0), device=dev), VAE(so.make_vae_state_dict(1), device=dev))
    pipe.register_adaptive_mask_model(LuminanceSegmenter(128))
    pipe.register_adaptive_mask_settings(default_adaptive_mask_settings(50))
    default = np.zeros((512, 512), np.uint8)
    default[64:448, 128:384] = 255
    pe = torch.randn((77, 768), generator=torch.Generator().manual_seed(1)) * 0.02
    ne = torch.zeros((77, 768))

    def run(view):
        image = np.random.default_rng(100 + view).integers(0, 256, (512, 512, 3), dtype=np.uint8)
        gens = [torch.Generator(device=dev).manual_seed(i) for i in range(B)]
        return image, pipe(image=image, default_mask_image=default, prompt_embeds=pe, negative_prompt_embeds=ne, guidance_scale=11.0,
                           strength=0.98, num_inference_steps=50, generator=gens, enforce_full_mask_ratio=0.0, human_detection_thres=0.015,
                           batch_size=B, output_type="np")
    image, out = run(v0)                    # warm-up: captures the CUDA graphs
    times = []
    l0 = _lib.launch_count()
    reps = 1 if cfg3 else 2
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        for v in range(v0, v1):
            image, out = run(v)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    launches = (_lib.launch_count() - l0) // (reps * max(v1 - v0, 1))
    t = torch.tensor([min(times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n_images = n_views * B
    per_gpu_tflops = B * (v1 - v0) * 159.7 / min(times)
    res = {"metric": "HOI images/s (512x512, 50-step DDIM, adaptive mask)", "value": n_images / t.item(), "unit": "images/s",
           "config": (f"BASELINE configs[2]: 36 viewpoints x batch 8 over {world} GPUs, contiguous slice rule (5,5,5,5,5,5,5,1 views)" if cfg3 else
                      f"BASELINE configs[1]: 1 viewpoint x batch {B} per GPU"),
           "s_job": t.item(), "batch_per_gpu": B, "views": n_views, "n_gpus": world,
           "data": "synthetic render + random weights (SD-1.5 inpainting UNet, SD VAE architectures)",
           "flop_per_image_T": 159.7, "tensor_tflops_achieved": n_images * 159.7 / t.item(),
           "roofline": {"bound": "tensor", "achieved": per_gpu_tflops, "peak": tensor_peak()[0], "unit": "TFLOP/s",
                        "frac": per_gpu_tflops / tensor_peak()[0], "peak_source": tensor_peak()[1],
                        "note": "rank 0's GPU; necessary work = 49 x 2 UNet evaluations + 22 VAE decodes + 23 VAE encodes per image (SURVEY 8d), whole "
                                "pipeline call incl. GroupNorm / softmax / mask logic / host copies, sustained clocks"},
           "e2e": {"api": "AdaptiveMaskInpaintPipeline.__call__ (host uint8 render + mask in, host uint8 images out)",
                   "h2d_bytes_per_step": int(image.nbytes + default.nbytes), "d2h_bytes_per_step": int(out.images.nbytes) if hasattr(out.images, "nbytes") else None},
           "launches_outside_graphs_per_batch": int(launches)}
    del pipe
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res["reference_torch_eager_fp16"] = hoi_reference(dev)
    return res


# ------------------------------------------------------------------------------------------------------------------ occupancy leg
def occupancy_leg(args, dev, rank, world, barrier):
    """BASELINE.json configs[4]: occupancy affordance, 128^3 voxels, 4096 synthetic samples, H-sharded (1310 of the 10475 human
    rows per GPU -> 11 GB of grids per GPU; 8 ranks = the whole config). K4 scatter of all samples + K5c read-out (per-vertex
    normalise + max over vertices) + ONE MAX all-reduce of the [128^3] field. Device-resident timing, then end to end through
    ComA_Occupancy with host fp64 samples (each rank loads 4096 / N of them; exchange over NVLink)."""
    import torch
    import torch.distributed as dist
    from coma_b200 import _lib, ops, synth
    from coma_b200 import dist as cdist
    from utils.coma_occupancy import ComA_Occupancy
    Sg, S, Hr = OCC["Sg"], args.occ_samples, OCC["H_per_rank"]
    Hj = Hr * world                                  # rows of the job (10480 at 8 ranks; the config's 10475 rounded up to 8 x 1310)
    h0, h1 = rank * Hr, (rank + 1) * Hr
    torch.cuda.empty_cache()
    occ = ComA_Occupancy(scale_tolerance=OCC["tol"], human_res=Hj, obj_res=4, normal_res=0, spatial_res=Sg, device=f"cuda:{dev.index}",
                         human_slice=(h0, h1))
    # device-resident canonical vertices of ALL samples for this rank's rows (fp32, already minus object vertex 0)
    chunks, obj0 = [], None
    for c0 in range(0, S, 512):
        hv, _, ov, _ = synth.make_sample_arrays(min(512, S - c0), Hj, 4, seed=900 + c0 // 512, dtype=np.float64)
        obj0 = ov[0, 0] if obj0 is None else obj0        # ONE object for the whole job (the reference asserts it never moves)
        chunks.append(torch.from_numpy((hv[:, h0:h1] - obj0[None, None]).astype(np.float32)))
    hvc = torch.cat(chunks).to(dev)
    del chunks

    def step():
        occ.spatial_occupancy_grids.zero_()
        ops.occupancy_accumulate(hvc, occ._centers, occ.rel_dist_thres, occ.spatial_occupancy_grids)
        return occ.return_aggregated_spatial_grids()       # K5c + MAX all-reduce

    step()
    hits = float(occ.spatial_occupancy_grids.nan_to_num(0).sum().item())   # grids are normalised now: recount below
    occ.spatial_occupancy_grids.zero_()
    ops.occupancy_accumulate(hvc, occ._centers, occ.rel_dist_thres, occ.spatial_occupancy_grids)
    hits = float(occ.spatial_occupancy_grids.sum(dtype=torch.float64).item())
    # occupied 512-byte granules (what K5c's second pass has to read and rewrite; the first pass reads everything once)
    occupied = 0
    if (Sg ** 3) % 128 == 0:
        for r0 in range(0, Hr, 64):
            occupied += int((occ.spatial_occupancy_grids[r0:r0 + 64].reshape(-1, 128) != 0).any(-1).sum().item())
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    barrier()
    l0 = _lib.launch_count()
    ev[0].record()
    occ.spatial_occupancy_grids.zero_()
    ops.occupancy_accumulate(hvc, occ._centers, occ.rel_dist_thres, occ.spatial_occupancy_grids)
    ev[1].record()
    field = ops.occupancy_readout(occ.spatial_occupancy_grids, None)
    ev[2].record()
    field = cdist.all_reduce_max_nan(field)
    ev[3].record()
    barrier()
    launches = _lib.launch_count() - l0
    k4_ms, k5_ms, tot_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[0].elapsed_time(ev[3])
    t = torch.tensor([tot_ms, k4_ms, k5_ms, hits], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        hs = t[3:].clone()
        dist.all_reduce(hs, op=dist.ReduceOp.SUM)
        tot_ms, k4_ms, k5_ms, hits = tm[0].item(), tm[1].item(), tm[2].item(), hs[0].item()
    peak, peak_src = measured_peaks()
    k5_dense_bytes = 12.0 * Hr * Sg ** 3                      # round 1: two reads + one write of every voxel
    k5_bytes = 4.0 * Hr * Sg ** 3 + 2 * 512.0 * occupied if occupied else k5_dense_bytes   # one full read + (read + write) of the occupied granules
    del hvc

    # end to end: host fp64 samples through the class API, 1/world of the samples loaded per rank
    S_e2e = min(S, args.occ_e2e_samples)
    mine_set = set(cdist.sample_shard(S_e2e, rank, world))
    host, obj = [], None
    for c0 in range(0, S_e2e, 256):
        ss = synth.make_samples(min(256, S_e2e - c0), Hj, 4, seed=900 + c0 // 256)
        obj = (ss[0]["obj_verts"], ss[0]["obj_normals"]) if obj is None else obj
        for j, s in enumerate(ss):
            if (c0 + j) in mine_set:
                s["obj_verts"], s["obj_normals"] = obj
                host.append(s)
    e2e_s = None
    for _ in range(2):   # one untimed pass (pinned staging buffers, first-touch pages, NCCL channels of the exchange), then the timed one
        occ.spatial_occupancy_grids.zero_()
        occ.debug_obj_vert = occ.debug_obj_normal = None
        occ.used, occ.used_count = {}, 0
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for s in host:
            occ.register_sample_to_cache(**s)
        occ.aggregate_all_samples(exchange=world > 1)
        f = occ.return_aggregated_spatial_grids().cpu().numpy()
        e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    h2d = occ.last_h2d_bytes
    del occ, host
    torch.cuda.empty_cache()
    return {"metric": "occupancy vertex-samples/s", "value": world * Hr * S / (tot_ms * 1e-3), "unit": "vertex-samples/s",
            "config": f"BASELINE configs[4]: {Sg}^3 voxels, {S} samples, H-sharded {Hr} rows/GPU x {world} GPUs, scale_tolerance {OCC['tol']}",
            "hits_per_s": hits / (k4_ms * 1e-3), "hits_per_vertex_sample": hits / (world * Hr * S),
            "ms": {"total": tot_ms, "k4_scatter": k4_ms, "k5c_readout": k5_ms, "max_all_reduce": tot_ms - k4_ms - k5_ms},
            "roofline_k5c": {"bound": "hbm", "kernel": "occupancy_rowsum_kernel + occupancy_norm_max_sparse_kernel (K5c)", "achieved": k5_bytes / (k5_ms * 1e-3) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": k5_bytes / (k5_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": k5_bytes, "peak_source": peak_src,
                             "occupied_granule_fraction": occupied * 128.0 / (Hr * Sg ** 3),
                             "dense_equivalent_gbs": k5_dense_bytes / (k5_ms * 1e-3) / 1e9,
                             "note": "algorithmic bytes = one read of every voxel (pass 1: row sums + occupied-granule flags) + read and write of the occupied 512-byte "
                                     "granules only (pass 2: 0 / sum rewrites an untouched zero as itself); dense_equivalent_gbs = round 1's 12 B per voxel over the same time"},
            "e2e": {"value": world * Hr * S_e2e / te.item(), "unit": "vertex-samples/s", "samples": S_e2e, "s": te.item(), "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(f.nbytes), "api": "ComA_Occupancy.register_sample_to_cache/aggregate_all_samples/return_aggregated_spatial_grids"},
            "gpu_launches": int(launches), "n_gpus": world}


def cfg1_leg(cpu_samples=4):
    """BASELINE.json configs[0] IN FULL: 32 samples, SMPL-X downsampled to 1000 vertices x 180 object vertices x 250 bins, preset
    qual:backpack_human_contact — the reference's own CPU-runnable case — end to end through the class API (register ->
    aggregate_all_samples -> get_aggregated_contact, host fp64 in, host maps out): the drop-in class on the B200, the UNMODIFIED
    reference with device="cuda" on the same B200 (all 32 samples) and on the host cores (`cpu_samples` of the 32: its loop is
    strictly per sample). The two aggregated human maps are compared on the spot (1e-4, the north-star tolerance)."""
    import torch
    from coma_b200 import synth
    from oracle import ref_loader
    from utils.coma import ComA, get_aggregated_contact
    Hc, Oc, Sc = 1000, 180, 32
    kw = dict(human_res=Hc, obj_res=Oc, normal_res=N, spatial_res=0, proximity_settings=dict(spatial_grid_size=0.07, spatial_grid_thres=0.03),
              normal_gaussian_sigma=0.25, eps=1e-10)
    base = synth.make_samples(Sc, Hc, Oc, seed=1)

    class _Fresh:   # every run gets its own copies of the host arrays (made outside the timed region)
        def __getitem__(self, sl):
            return [{k: v.copy() for k, v in x.items()} for x in base[sl]]
    samples = _Fresh()

    def run(cls, gac, device, smp):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        c = cls(device=device, **kw)
        for x in smp:
            c.register_sample_to_cache(**x)
        c.aggregate_all_samples()
        m, idx = gac(c, "human", 0.03)   # 0.03 x 32 samples: every pair that was hit at least once counts as significant
        m = np.asarray(m.cpu() if hasattr(m, "cpu") else m)
        return time.perf_counter() - t0, m, idx

    out = {"config": f"BASELINE configs[0] in full: {Sc} samples, H={Hc}, O={Oc}, N={N}, preset qual:backpack_human_contact; class API end to end",
           "unit": "vertex-pairs/s"}
    run(ComA, get_aggregated_contact, "cuda", samples[:2])                      # warm-up: allocator, pinned staging
    dt, mine, mine_idx = min((run(ComA, get_aggregated_contact, "cuda", samples[:]) for _ in range(3)), key=lambda r: r[0])
    out["value"] = Sc * Hc * Oc / dt
    out["s"] = dt
    if ref_loader.available():
        ref = ref_loader.load()
        try:
            _quiet(run, ref.ComA, ref.get_aggregated_contact, "cuda", samples[:2])
            dt_r, theirs, their_idx = _quiet(run, ref.ComA, ref.get_aggregated_contact, "cuda", samples[:])
            out["reference_cuda"] = {"value": Sc * Hc * Oc / dt_r, "s": dt_r, "kind": "reference", "sample": f"all {Sc} samples, device='cuda'"}
            out["agrees_with_reference_cuda"] = bool(np.array_equal(np.asarray(mine_idx), np.asarray(their_idx))
                                                     and np.allclose(mine, theirs, rtol=1e-4, atol=1e-12))
            torch.set_num_threads(host_cores())
            dt_c, _, _ = _quiet(run, ref.ComA, ref.get_aggregated_contact, "cpu", samples[:cpu_samples])
            out["reference_cpu"] = {"value": cpu_samples * Hc * Oc / dt_c, "s": dt_c, "cores": host_cores(), "kind": "reference",
                                    "sample": f"{cpu_samples} of the {Sc} samples, device='cpu' (per-sample loop)"}
        except Exception as ex:  # informative leg only
            out["reference_error"] = repr(ex)[:200]
        finally:
            torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples-per-rank", type=int, default=S_PER_RANK)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--no-hoi", action="store_true", help="skip the HOI images/s leg (adaptive-mask inpainting loop)")
    ap.add_argument("--no-occupancy", action="store_true", help="skip the occupancy leg (BASELINE configs[4])")
    ap.add_argument("--hoi-batch", type=int, default=4)
    ap.add_argument("--hoi-cfg2", action="store_true", help="at 8 GPUs run configs[1] per rank instead of configs[2] (36 views x batch 8)")
    ap.add_argument("--occ-samples", type=int, default=OCC["S"])
    ap.add_argument("--occ-e2e-samples", type=int, default=OCC["S"], help="samples of the end-to-end occupancy pass (default: the whole config)")
    ap.add_argument("--sample-sharded", action="store_true", help="round-1 form: shard SAMPLES + all-reduce(SUM) of the accumulators")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from coma_b200 import _lib, ops, synth
    from coma_b200 import dist as cdist
    from utils.coma import ComA, get_aggregated_contact

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    Sr = args.samples_per_rank
    row_sharded = world > 1 and not args.sample_sharded
    S = Sr * world if row_sharded else Sr                 # samples this rank aggregates per step
    hsl = cdist.human_slice(H, rank, world) if row_sharded else (0, H)
    Hl = hsl[1] - hsl[0]

    def make_coma():
        return ComA(human_res=H, obj_res=O, normal_res=N, spatial_res=0,
                    proximity_settings=dict(spatial_grid_size=PRESET["spatial_grid_size"], spatial_grid_thres=PRESET["spatial_grid_thres"]),
                    normal_gaussian_sigma=PRESET["normal_gaussian_sigma"], eps=PRESET["eps"], device=f"cuda:{local_rank}",
                    human_slice=hsl if row_sharded else None)

    # ---- synthetic samples (seeded per 256-sample block), resident in HBM for `value`: all S samples, this rank's rows
    blocks = range(world) if row_sharded else [rank]
    parts, host_block = [], None
    for b in blocks:
        arrs = synth.make_sample_arrays(Sr, H, O, seed=42 + b, dtype=np.float64)
        if b == rank:
            host_block = arrs                                                     # the samples THIS rank "loads" in the e2e leg
        parts.append([torch.from_numpy(np.ascontiguousarray(a[:, hsl[0]:hsl[1]] if i < 2 else a).astype(np.float32)) for i, a in enumerate(arrs)])
    hv, hn, ov, on = (torch.cat([p[i] for p in parts]).to(dev) for i in range(4))
    del parts
    coma = make_coma()
    grid = coma.canon_normal_grid.contiguous()
    k3_events = []

    def step(timed):
        # K2 + K3 over all samples of the step, accumulators in registers, one launch each (== ComA.aggregate_batch_for_contact)
        ops.pair_accumulate(hv, ov, PRESET["spatial_grid_thres"], PRESET["spatial_grid_size"], coma.significant_contact_count,
                            coma.contact_dist_expectation_grid_nom, sum_order=coma.reference_sum_order)
        coma.contact_dist_expectation_grid_denom += float(S)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.orient_accumulate(hn, on, grid, PRESET["normal_gaussian_sigma"], PRESET["eps"], [0, 0, 1], [0, 1, 0],
                              coma.prob_grid_canon_human_wrt_obj, coma.prob_grid_canon_obj_wrt_human,
                              bin_perm=coma._bin_perm, drop_bits=coma.orient_drop_bits, sum_order=coma.reference_sum_order)
        b.record()
        if timed:
            k3_events.append((a, b))
        coma.used_count += S
        if world > 1 and not row_sharded:
            coma.all_reduce()   # round-1 form: SUM of the sample-sharded accumulators over NVLink

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    k3_kernel = _lib.last_kernel()
    launches0 = _lib.launch_count()
    with ClockSampler(local_rank) as clocks:
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record()
        for _ in range(args.steps):
            step(True)
        t1.record()
        barrier()
    launches = _lib.launch_count() - launches0
    ms_total = t0.elapsed_time(t1)
    k3_ms = float(np.mean([a.elapsed_time(b) for a, b in k3_events]))
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = tmax.item() / args.steps
    pairs_per_step = world * Sr * H * O                   # whole job: (256 x world) samples x H x O, split by rows or by samples
    value = pairs_per_step / (ms_step * 1e-3)
    del hv, hn, ov, on

    # ---- K2 in its HBM-bound streaming form (one sample per launch, 16 B per vertex-pair), rotating over accumulator
    #      sets larger than L2 so every launch streams from HBM (full H x O, the reference's per-sample call)
    peak, peak_src = measured_peaks()
    k2 = None
    if rank == 0:
        nsets = 6  # 6 x 126 MB of accumulators > 126 MB L2
        cs = [torch.zeros((H, O), device=dev) for _ in range(nsets)]
        ns = [torch.zeros((H, O), device=dev) for _ in range(nsets)]
        hv1 = torch.from_numpy(host_block[0][:1].astype(np.float32)).to(dev)
        ov1 = torch.from_numpy(host_block[2][:1].astype(np.float32)).to(dev)
        for i in range(nsets):
            ops.pair_accumulate(hv1, ov1, 0.05, 0.15, cs[i], ns[i], sum_order="cuda")
        assert _lib.last_kernel() == "pair_accumulate_stream_kernel"
        torch.cuda.synchronize()
        n_k2 = 8 * nsets
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n_k2):   # back-to-back launches over rotating accumulator sets: average launch duration
            ops.pair_accumulate(hv1, ov1, 0.05, 0.15, cs[i % nsets], ns[i % nsets], sum_order="cuda")
        b.record()
        torch.cuda.synchronize()
        k2_ms = a.elapsed_time(b) / n_k2
        k2_bytes = 16.0 * H * O + 12.0 * (H + O)
        k2 = {"bound": "hbm", "kernel": "pair_accumulate_stream_kernel (K2, S=1)", "achieved": k2_bytes / (k2_ms * 1e-3) / 1e9,
              "peak": peak, "unit": "GB/s", "frac": k2_bytes / (k2_ms * 1e-3) / 1e9 / peak, "traffic": K2_NCU_DRAM_BYTES,
              "algorithmic_bytes": k2_bytes, "ms": k2_ms, "peak_source": peak_src,
              "timing": "48 back-to-back launches rotating over 6 accumulator sets (756 MB > 126 MB L2), one CUDA-event pair"}
        del cs, ns

    # ---- end to end through the class API with host samples (fresh instance per step, read-out to host)
    del coma
    torch.cuda.empty_cache()
    samples = [dict(human_verts=host_block[0][i], human_normals=host_block[1][i], obj_verts=host_block[2][i], obj_normals=host_block[3][i])
               for i in range(Sr)]
    e2e_ms, h2d, d2h = [], 0, 0
    for it in range(2 + args.steps):
        barrier()
        t = time.perf_counter()
        c = make_coma()
        for s in samples:
            c.register_sample_to_cache(**s)
        c.aggregate_all_samples(exchange=row_sharded)
        if world > 1 and not row_sharded:
            c.all_reduce()
        agg, idx = get_aggregated_contact(c, "human", PRESET["significant_contact_ratio"])
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) * 1e3
        h2d, d2h = c.last_h2d_bytes, agg.nbytes + H * O  # fp32 map + the bool significant-pair matrix
        del c
        if it >= 2:
            e2e_ms.append(dt)
    e2e_t = torch.tensor([float(np.mean(e2e_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = pairs_per_step / (e2e_t.item() * 1e-3)
    del samples, host_block

    occupancy = None if args.no_occupancy else occupancy_leg(args, dev, rank, world, barrier)
    hoi = None if args.no_hoi else hoi_leg(args, dev, rank, world, barrier)

    if rank == 0:
        ck = clocks.summary()
        k3_bytes = 24.0 * S * (Hl + O) + 16.0 * Hl * O * N
        sm_hz = (ck["sm_mhz"] or 1965.0) * 1e6
        evals_per_s = 2.0 * N * S * Hl * O / (k3_ms * 1e-3)
        line = {
            "metric": "ComA vertex-pairs/s", "value": value, "unit": "vertex-pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"coma_contact cfg4-shape: H={H} O={O} N={N}, {Sr * world} samples/step ({Sr} per GPU), K2+K3"
                                   + (" + NCCL all-reduce(SUM) of count/nom/PH/PO" if world > 1 and not row_sharded else ""),
                       "preset": "qual:backpack_object_contact",
                       "sharding": ("human-vertex rows (each rank: all samples x H/N rows, no accumulator collective)" if row_sharded
                                    else "samples" if world > 1 else "none"),
                       "l2_policy": f"working set ({16.0 * Hl * O * N / 1e9:.1f} GB of accumulators per step) >> 126 MB L2; K2-stream rotates 6 accumulator sets"},
            "e2e": {"value": e2e_value, "unit": "vertex-pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_t.item(),
                    "api": "ComA.register_sample_to_cache/aggregate_all_samples/get_aggregated_contact"
                           + (" (row-sharded: staged samples all-gathered over NVLink, per-vertex maps exchanged in the read-out)" if row_sharded else "")},
            "gpu_launches": int(launches),
            "clocks": ck,
            "roofline": {"bound": "hbm", "kernel": f"{k3_kernel} (K3)", "achieved": k3_bytes / (k3_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": k3_bytes / (k3_ms * 1e-3) / 1e9 / peak,
                         "traffic": K3_NCU_DRAM_BYTES if world == 1 else None,
                         "algorithmic_bytes": k3_bytes, "peak_source": peak_src, "ms": k3_ms,
                         "note": "K3 is FP32-pipe bound once samples are fused (see roofline_pipe / roofline_sfu); the HBM fraction is reported "
                                 "because the schema asks for it"},
        }
        if k3_kernel == "orient_accumulate_cone_kernel":
            steps_per_s = evals_per_s * K3C_EXEC_FRACTION / 64.0
            peak_instr = 148 * 4 * sm_hz / FFMA2_CLK_PER_WARP_INSTR
            line["roofline_pipe"] = {
                "bound": "fp32-heavy pipe (packed FFMA2)", "kernel": f"{k3_kernel} (K3)", "achieved": steps_per_s * K3C_PACKED_PER_STEP / 1e9,
                "peak": peak_instr / 1e9, "unit": "G packed-FP32 warp-instructions/s", "frac": steps_per_s * K3C_PACKED_PER_STEP / peak_instr,
                "algorithmic_bin_evals_per_s": evals_per_s, "executed_fraction": K3C_EXEC_FRACTION, "packed_instr_per_64_evals": K3C_PACKED_PER_STEP,
                "speedup_vs_dense_evaluation": "the dense kernel (all 2 x 250 bins, 2 MUFU each) ran this workload at 3.40e9 vertex-pairs/s (BENCH_r01)",
                "peak_source": "148 SMs x 4 sub-partitions / 2.1 clk per FFMA2 warp-instr (tools/ubench_pipes.cu) x median SM clock under load; "
                               "executed fraction from the ncu capture profiles/r02_k3cone_full.md"}
        else:
            line["roofline_sfu"] = {"bound": "sfu", "kernel": f"{k3_kernel} (K3)", "achieved": evals_per_s * K3_MUFU_PER_EVAL / 1e9,
                                    "peak": 148 * 4 * 32 * sm_hz / MUFU_CLK_PER_WARP_INSTR / 1e9, "unit": "G MUFU lane-ops/s",
                                    "frac": evals_per_s * K3_MUFU_PER_EVAL / (148 * 4 * 32 * sm_hz / MUFU_CLK_PER_WARP_INSTR),
                                    "bin_evals_per_s": evals_per_s, "mufu_per_eval": K3_MUFU_PER_EVAL,
                                    "peak_source": "148 SMs x 4 sub-partitions x 32 lanes / 8.05 clk per MUFU warp-instr (tools/ubench_pipes.cu) x median SM clock under load"}
        if k2 is not None:
            line["roofline_k2_stream"] = k2
        if occupancy is not None:
            line["occupancy"] = occupancy
        if hoi is not None:
            line["hoi"] = hoi
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_rate()
            if not args.no_reference_cuda:
                line["reference_cuda"] = reference_cuda_rate()
                try:
                    line["cfg1"] = cfg1_leg()
                except Exception as ex:   # informative leg: never lose the headline line over it
                    line["cfg1"] = {"error": repr(ex)[:300]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
