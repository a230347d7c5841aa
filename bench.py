#!/usr/bin/env python
"""bench.py — ComA vertex-pairs/s on B200 (BASELINE.json metric, contact-extraction leg).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[3], "ComA extraction: 2048 synthetic 3D HOI samples, SMPL-X 10475 x object 1500 verts,
8xB200 + NCCL histogram all-reduce"): every rank aggregates its 256-sample shard (weak scaling: 8 ranks = the 2048
samples of the config) of H=10475 x O=1500 vertex pairs through ALL of aggregate_single_sample_for_contact — pair
distance / contact count / proximity (K2) and both 250-bin orientation histograms (K3) — and, for N > 1, sums the
accumulators with one NCCL all-reduce per tensor. A *vertex-pair* is one (sample, human vertex, object vertex) triple.

`value`  : device-resident inputs, CUDA-event timed, max over ranks.
`e2e`    : the same job through the reference-facing class API with HOST (numpy fp64) samples:
           ComA(...) -> register_sample_to_cache x S -> aggregate_all_samples() (H2D inside) -> get_aggregated_contact()
           -> numpy on the host.
`roofline`: the dominant kernel (K3, orient_accumulate_kernel) timed live on its stream; plus the HBM-bound streaming
           form of K2 the north star sets its 90 % target on (`roofline_k2_stream`).
`cpu_baseline` / `--impl reference`: the CPU restatement of the reference (oracle/, OpenMP over all host cores) on a
           bounded sample of the same workload. The reference is pure Python and cannot travel to the GPU box.
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
# DRAM traffic of one launch from `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01_k3_full.md, same
# H x O x N, 32 samples: the grids are read and written once per launch whatever the sample count; inputs are 0.3 MB / sample)
K3_NCU_DRAM_BYTES = 31.467172e9 + 31.378264e9
# K2 streaming kernel (one sample per launch), profiles/r01_k2_full.md: 125.9 MB read + 70.2 MB written while the capture ran (the
# rest of the 125.7 MB of accumulator writes is still dirty in the 126 MB L2 when the single profiled launch ends)
K2_NCU_DRAM_BYTES = 125.852928e6 + 70.227456e6
K3_MUFU_PER_EVAL = 2.0     # MUFU.SQRT + MUFU.EX2 per bin evaluation in the K3 inner loop (cuobjdump)
MUFU_CLK_PER_WARP_INSTR = 8.05  # measured on B200 with tools/ubench_pipes.cu (4 lanes/clk per SM sub-partition)


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


def cpu_reference_rate(target_seconds=12.0):
    """Oracle port (C + OpenMP, all host cores) on a bounded sample of the workload: rows [0, Hs) of ONE cfg-4 sample."""
    from coma_b200 import synth
    from oracle import oracle
    # all host cores this process may run on — torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    oracle.set_num_threads(avail)
    cores = oracle.num_threads()
    hv, hn, ov, on = synth.make_sample_arrays(1, H, O, seed=123)
    grid = oracle.fibonacci_sphere(N)
    Hs = 64
    t0 = time.perf_counter()
    oracle.pair_accumulate(hv[:, :Hs], ov, PRESET["spatial_grid_thres"], PRESET["spatial_grid_size"])
    oracle.orient_accumulate(hn[:, :Hs], on, grid, PRESET["normal_gaussian_sigma"], PRESET["eps"])
    probe = time.perf_counter() - t0
    Hs = int(min(H, max(64, Hs * target_seconds / max(probe, 1e-3)))) // 8 * 8
    t0 = time.perf_counter()
    oracle.pair_accumulate(hv[:, :Hs], ov, PRESET["spatial_grid_thres"], PRESET["spatial_grid_size"])
    oracle.orient_accumulate(hn[:, :Hs], on, grid, PRESET["normal_gaussian_sigma"], PRESET["eps"])
    dt = time.perf_counter() - t0
    return dict(value=Hs * O / dt, unit="vertex-pairs/s", cores=cores, kind="port",
                sample=f"1 sample x {Hs} of {H} human rows x {O} object verts x {N} bins (K2+K3), {dt:.1f} s on {cores} OpenMP threads")


def run_reference(args, rank):
    """`--impl reference`: times the CPU restatement of the reference on the host cores (rank 0 only)."""
    if rank != 0:
        return
    for _ in range(max(args.warmup, 1) - 1):
        cpu_reference_rate(target_seconds=2.0)
    rates = [cpu_reference_rate(target_seconds=max(4.0, 60.0 / max(args.steps, 1))) for _ in range(args.steps)]
    best = rates[int(np.argsort([r["value"] for r in rates])[len(rates) // 2])]
    line = {
        "impl": "reference", "metric": "ComA vertex-pairs/s", "value": best["value"], "unit": "vertex-pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"coma_contact cfg4-shape H={H} O={O} N={N}, K2+K3 (reference CPU path, oracle port)"},
        "cpu_baseline": best,
        "e2e": {"value": best["value"], "unit": "vertex-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def hoi_leg(args, dev, rank, world, barrier):
    """BASELINE.json configs[1]/[2]: adaptive-mask SD inpainting, 512x512, 50 DDIM steps (strength 0.98 -> 49 UNet x2
    evaluations, CFG 11, 21 adapt calls), `--hoi-batch` work items per rank sharing (render, mask, prompt); one rank = one
    viewpoint shard (no collective: outputs are images). Seeded random weights with the real architecture (no checkpoints
    offline), synthetic render, rectangular default mask, deterministic stub segmenter. images/s = finished 512x512
    outputs per second, whole job (all ranks), wall clock around the public pipeline call (host image in, host images out)."""
    import torch
    import torch.distributed as dist
    from coma_b200 import _lib
    from coma_b200.inpaint.pipeline import AdaptiveMaskInpaintPipeline, default_adaptive_mask_settings
    from coma_b200.inpaint.segmenter import LuminanceSegmenter
    from coma_b200.inpaint.unet import UNet
    from coma_b200.inpaint.vae import VAE
    from oracle import sd_oracle as so   # weight generator only (seeded random state dicts with diffusers key names)
    B = args.hoi_batch
    torch.cuda.empty_cache()
    pipe = AdaptiveMaskInpaintPipeline(UNet(so.make_unet_state_dict(0), device=dev), VAE(so.make_vae_state_dict(1), device=dev))
    pipe.register_adaptive_mask_model(LuminanceSegmenter(128))
    pipe.register_adaptive_mask_settings(default_adaptive_mask_settings(50))
    rng = np.random.default_rng(100 + rank)
    image = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
    default = np.zeros((512, 512), np.uint8)
    default[64:448, 128:384] = 255
    pe = torch.randn((77, 768), generator=torch.Generator().manual_seed(1)) * 0.02
    ne = torch.zeros((77, 768))

    def run():
        gens = [torch.Generator(device=dev).manual_seed(i) for i in range(B)]
        return pipe(image=image, default_mask_image=default, prompt_embeds=pe, negative_prompt_embeds=ne, guidance_scale=11.0,
                    strength=0.98, num_inference_steps=50, generator=gens, enforce_full_mask_ratio=0.0, human_detection_thres=0.015,
                    batch_size=B, output_type="np")
    run()                                   # warm-up: captures the CUDA graphs
    times = []
    l0 = _lib.launch_count()
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        out = run()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    launches = (_lib.launch_count() - l0) // 2
    t = torch.tensor([min(times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = {"metric": "HOI images/s (512x512, 50-step DDIM, adaptive mask)", "value": world * B / t.item(), "unit": "images/s",
           "s_per_batch": t.item(), "batch_per_gpu": B, "n_gpus": world, "data": "synthetic render + random weights (SD-1.5 inpainting UNet, SD VAE architectures)",
           "flop_per_image_T": 159.7, "tensor_tflops_achieved": world * B * 159.7 / t.item(),
           "roofline": {"bound": "tensor", "achieved": B * 159.7 / t.item(), "peak": tensor_peak()[0], "unit": "TFLOP/s",
                        "frac": B * 159.7 / t.item() / tensor_peak()[0], "peak_source": tensor_peak()[1],
                        "note": "per GPU; necessary work = 49 x 2 UNet evaluations + 22 VAE decodes + 23 VAE encodes per image (SURVEY 8d), whole "
                                "pipeline call incl. GroupNorm / softmax / mask logic / host copies, sustained clocks"},
           "e2e": {"api": "AdaptiveMaskInpaintPipeline.__call__ (host uint8 render + mask in, host uint8 images out)",
                   "h2d_bytes_per_step": int(image.nbytes + default.nbytes), "d2h_bytes_per_step": int(out.images.nbytes) if hasattr(out.images, "nbytes") else None},
           "d2h_bytes": int(out.images.nbytes) if hasattr(out.images, "nbytes") else None,
           "launches_outside_graphs_per_batch": int(launches)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle.inpaint_loop_oracle import time_reference_loop
            s = time_reference_loop(dev)
            res["reference_torch_eager_fp16"] = {"images_per_s": 1.0 / s, "s_per_image": s,
                                                 "what": "reference-shaped loop (batch 1, decode every step, unfused attention, cv2 on host) in torch eager fp16 on the same GPU"}
        except Exception as ex:  # the baseline is informative only
            res["reference_torch_eager_fp16"] = {"error": repr(ex)[:200]}
    del pipe
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples-per-rank", type=int, default=S_PER_RANK)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hoi", action="store_true", help="skip the HOI images/s leg (adaptive-mask inpainting loop)")
    ap.add_argument("--hoi-batch", type=int, default=4)
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
    from utils.coma import ComA, get_aggregated_contact

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    S = args.samples_per_rank

    def make_coma():
        return ComA(human_res=H, obj_res=O, normal_res=N, spatial_res=0,
                    proximity_settings=dict(spatial_grid_size=PRESET["spatial_grid_size"], spatial_grid_thres=PRESET["spatial_grid_thres"]),
                    normal_gaussian_sigma=PRESET["normal_gaussian_sigma"], eps=PRESET["eps"], device=f"cuda:{local_rank}")

    # ---- synthetic shard of this rank (seeded per rank), resident in HBM for `value`
    hv_h, hn_h, ov_h, on_h = synth.make_sample_arrays(S, H, O, seed=42 + rank, dtype=np.float64)
    hv, hn, ov, on = (torch.from_numpy(a.astype(np.float32)).to(dev) for a in (hv_h, hn_h, ov_h, on_h))
    coma = make_coma()
    grid = coma.canon_normal_grid.contiguous()
    k3_events = []

    def step(timed):
        # K2 + K3 over the whole shard, accumulators in registers, one launch each (== ComA.aggregate_batch_for_contact)
        ops.pair_accumulate(hv, ov, PRESET["spatial_grid_thres"], PRESET["spatial_grid_size"], coma.significant_contact_count,
                            coma.contact_dist_expectation_grid_nom)
        coma.contact_dist_expectation_grid_denom += float(S)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.orient_accumulate(hn, on, grid, PRESET["normal_gaussian_sigma"], PRESET["eps"], [0, 0, 1], [0, 1, 0],
                              coma.prob_grid_canon_human_wrt_obj, coma.prob_grid_canon_obj_wrt_human)
        b.record()
        if timed:
            k3_events.append((a, b))
        coma.used_count += S
        if world > 1:
            coma.all_reduce()   # the job's single exchange step: SUM of the accumulators over NVLink

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
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
    value = world * S * H * O / (ms_step * 1e-3)

    # ---- K2 in its HBM-bound streaming form (one sample per launch, 16 B per vertex-pair), rotating over accumulator
    #      sets larger than L2 so every launch streams from HBM
    peak, peak_src = measured_peaks()
    nsets = 6  # 6 x 126 MB of accumulators > 126 MB L2
    cs = [torch.zeros((H, O), device=dev) for _ in range(nsets)]
    ns = [torch.zeros((H, O), device=dev) for _ in range(nsets)]
    hv1, ov1 = hv[:1].contiguous(), ov[:1].contiguous()
    for i in range(nsets):
        ops.pair_accumulate(hv1, ov1, 0.05, 0.15, cs[i], ns[i])
    torch.cuda.synchronize()
    n_k2 = 8 * nsets
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n_k2):   # back-to-back launches over rotating accumulator sets: average launch duration
        ops.pair_accumulate(hv1, ov1, 0.05, 0.15, cs[i % nsets], ns[i % nsets])
    b.record()
    torch.cuda.synchronize()
    k2_ms = a.elapsed_time(b) / n_k2
    k2_bytes = 16.0 * H * O + 12.0 * (H + O)
    del cs, ns

    # ---- end to end through the class API with host samples (fresh instance per step, read-out to host)
    del coma
    torch.cuda.empty_cache()
    samples = [dict(human_verts=hv_h[i], human_normals=hn_h[i], obj_verts=ov_h[i], obj_normals=on_h[i]) for i in range(S)]
    e2e_ms, h2d, d2h = [], 0, 0
    for it in range(2 + args.steps):
        barrier()
        t = time.perf_counter()
        c = make_coma()
        for s in samples:
            c.register_sample_to_cache(**s)
        c.aggregate_all_samples()
        if world > 1:
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
    e2e_value = world * S * H * O / (e2e_t.item() * 1e-3)

    hoi = None
    if not args.no_hoi:
        hoi = hoi_leg(args, dev, rank, world, barrier)

    if rank == 0:
        ck = clocks.summary()
        k3_bytes = 24.0 * S * (H + O) + 16.0 * H * O * N
        sm_hz = (ck["sm_mhz"] or 1965.0) * 1e6
        evals_per_s = 2.0 * N * S * H * O / (k3_ms * 1e-3)
        line = {
            "metric": "ComA vertex-pairs/s", "value": value, "unit": "vertex-pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"coma_contact cfg4-shape: H={H} O={O} N={N}, {S} samples/GPU/step, K2+K3"
                                   + (" + NCCL all-reduce(SUM) of count/nom/PH/PO" if world > 1 else ""),
                       "preset": "qual:backpack_object_contact", "sharding": "samples",
                       "l2_policy": "working set (31.4 GB of accumulators per step) >> 126 MB L2; K2-stream rotates 6 accumulator sets"},
            "e2e": {"value": e2e_value, "unit": "vertex-pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_t.item(), "api": "ComA.register_sample_to_cache/aggregate_all_samples/get_aggregated_contact"},
            "gpu_launches": int(launches),
            "clocks": ck,
            "roofline": {"bound": "hbm", "kernel": "orient_accumulate_kernel_x2 (K3)", "achieved": k3_bytes / (k3_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": k3_bytes / (k3_ms * 1e-3) / 1e9 / peak, "traffic": K3_NCU_DRAM_BYTES,
                         "algorithmic_bytes": k3_bytes, "peak_source": peak_src, "ms": k3_ms,
                         "note": "K3 is SFU (MUFU) / FP32-pipe bound once samples are fused (500 bin evaluations per pair-sample, "
                                 "2 MUFU each): see roofline_sfu; the HBM fraction is reported because the schema asks for it"},
            "roofline_sfu": {"bound": "sfu", "kernel": "orient_accumulate_kernel_x2 (K3)", "achieved": evals_per_s * K3_MUFU_PER_EVAL / 1e9,
                             "peak": 148 * 4 * 32 * sm_hz / MUFU_CLK_PER_WARP_INSTR / 1e9, "unit": "G MUFU lane-ops/s",
                             "frac": evals_per_s * K3_MUFU_PER_EVAL / (148 * 4 * 32 * sm_hz / MUFU_CLK_PER_WARP_INSTR),
                             "bin_evals_per_s": evals_per_s, "mufu_per_eval": K3_MUFU_PER_EVAL,
                             "peak_source": "148 SMs x 4 sub-partitions x 32 lanes / 8.05 clk per MUFU warp-instr (tools/ubench_pipes.cu) x median SM clock under load"},
            "roofline_k2_stream": {"bound": "hbm", "kernel": "pair_accumulate_stream_kernel (K2, S=1)", "achieved": k2_bytes / (k2_ms * 1e-3) / 1e9,
                                   "peak": peak, "unit": "GB/s", "frac": k2_bytes / (k2_ms * 1e-3) / 1e9 / peak, "traffic": K2_NCU_DRAM_BYTES,
                                   "algorithmic_bytes": k2_bytes,
                                   "ms": k2_ms, "peak_source": peak_src},
        }
        if hoi is not None:
            line["hoi"] = hoi
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_rate()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
