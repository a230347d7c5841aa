"""Segmenter plug-ins for the adaptive-mask loop. The reference's default is detectron2 PointRend
(utils/adaptive_mask_inpainting.py:1182-1236), whose weights and runtime are not available here; the interface is kept
(`callable(np.uint8[H,W,3]) -> {"mask": np.uint8[H,W], "asset_mask": ..., "vis": ...}`, attribute `use_visualizer`) and a
deterministic stand-in is provided for tests and benchmarks (used on BOTH sides of every comparison)."""
import numpy as np
import torch


class LuminanceSegmenter:
    """'Human' = pixels whose luminance exceeds `thres` (0..255). Works on numpy images (reference interface) and, as a
    B200 extension (`accepts_cuda`), on uint8 CUDA batches [B,H,W,3] without leaving the device."""
    use_visualizer = False
    accepts_cuda = True

    def __init__(self, thres=140):
        self.thres = float(thres)
        self._w = {}     # per-device luminance weights (built once: torch.tensor(..., device=cuda) is a synchronising copy)

    def __call__(self, image):
        if isinstance(image, torch.Tensor):
            w = self._w.get(image.device)
            if w is None:
                w = self._w[image.device] = torch.tensor([0.299, 0.587, 0.114], device=image.device)
            lum = image.float().mul(w).sum(-1)
            return {"mask": (lum > self.thres).to(torch.uint8), "asset_mask": None, "vis": None}
        lum = image.astype(np.float32) @ np.array([0.299, 0.587, 0.114], dtype=np.float32)
        return {"mask": (lum > self.thres).astype(np.uint8), "asset_mask": None, "vis": None}
