"""Host-side helpers mirroring the reference's boundary semantics (utils/misc.py:14-63, :66-83)."""
import numpy as np
import torch

_CONTAINERS = (dict, list, np.ndarray, torch.Tensor)


def to_np_torch_recursive(X, use_torch=True, device="cuda", np_float_type=np.float32, np_int_type=np.int64,
                          torch_float_type=torch.float32, torch_int_type=torch.int64):
    """Recursive numpy<->torch converter that FORCES fp32 / int64 — this is what defines the precision every sample
    enters the accumulation with (reference: utils/misc.py:14-63). Dicts / lists are converted in place."""
    if isinstance(X, dict):
        for k in X.keys():
            if isinstance(X[k], _CONTAINERS):
                X[k] = to_np_torch_recursive(X[k], use_torch, device)
    elif isinstance(X, list):
        for i in range(len(X)):
            if isinstance(X[i], _CONTAINERS):
                X[i] = to_np_torch_recursive(X[i], use_torch, device)
    elif isinstance(X, np.ndarray):
        if use_torch:
            X = torch.tensor(X, device=device)
    elif isinstance(X, torch.Tensor):
        X = X.to(device) if use_torch else X.detach().cpu().numpy()

    if isinstance(X, torch.Tensor):
        if X.dtype in (torch.float64, torch.float32, torch.float16, torch.bfloat16):
            X = X.type(torch_float_type)
        elif X.dtype in (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64):
            X = X.type(torch_int_type)
    elif isinstance(X, np.ndarray):
        if X.dtype in (np.float32, np.float16, np.float64):
            X = X.astype(np_float_type)
        elif X.dtype in (np.int64, np.int32, np.int16):
            X = X.astype(np_int_type)
    return X


def get_3d_indexgrid_ijk(N_x, N_y, N_z):
    """utils/misc.py:66-83 (non-raveled form): indices[:, i, j, k] == (i, j, k)."""
    return np.mgrid[0:N_x, 0:N_y, 0:N_z]


def normalize_vectors_np(vecs, eps=1e-8):
    """utils/transformations.py:8-11."""
    assert vecs.ndim == 2 and vecs.shape[-1] == 3
    return vecs / (np.sqrt(np.sum(np.square(vecs), axis=-1, keepdims=True)) + eps)


def get_uniform_points_on_sphere(num_points=1000):
    """Fibonacci-sphere bin centres, utils/coma.py:18-26 (float64)."""
    indices = np.arange(0, num_points, dtype=float) + 0.5
    phi = np.arccos(1 - 2 * indices / num_points)
    theta = np.pi * (1 + 5**0.5) * indices
    return np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)
