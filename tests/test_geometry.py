"""K7 / coma_b200.geometry: nearest-neighbour distances between point sets with a backward pass, and the reference's
`chamfer_distance` (src/application/optimize.py:155-165) / `minimum_distance` (src/generation/optimize_depth.py:29-44) on top.
CPU: the oracle restatement against the reference's literal torch.cdist expressions. GPU: kernels against the oracle (bit-exact
distances and indices) and against torch autograd through cdist (gradients)."""
import numpy as np
import pytest
import torch


def _ref_chamfer(A, B):   # src/application/optimize.py:155-165, verbatim
    dist_A_to_B = torch.cdist(A, B)
    dist_B_to_A = torch.cdist(B, A)
    min_dist_A_to_B, _ = torch.min(dist_A_to_B, dim=1)
    min_dist_B_to_A, _ = torch.min(dist_B_to_A, dim=1)
    return torch.mean(min_dist_A_to_B) + torch.mean(min_dist_B_to_A)


def _ref_minimum_distance(vertsA, vertsB, num_vertices=100):   # src/generation/optimize_depth.py:29-44 (one cdist batch)
    d = torch.cdist(vertsA.float(), vertsB.float())
    m = torch.min(d, dim=1).values
    s, _ = torch.sort(m)
    return torch.mean(s[:num_vertices])


def _clouds(na, nb, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((na, 3)).astype(np.float32) * 0.3
    b = (rng.standard_normal((nb, 3)) * 0.3 + 0.05).astype(np.float32)
    b[nb // 2] = a[min(3, na - 1)]         # an exact coincidence (zero distance -> zero gradient)
    if nb > 8:
        b[7] = b[2]                        # duplicate point: arg-min tie -> lowest index
    return a, b


def test_oracle_matches_reference_expressions():
    from oracle import oracle
    a, b = _clouds(700, 333, 0)
    d, idx = oracle.nearest_distance(a, b)
    full = torch.cdist(torch.from_numpy(a), torch.from_numpy(b), compute_mode="donot_use_mm_for_euclid_dist")
    m, j = torch.min(full, dim=1)
    np.testing.assert_allclose(d, m.numpy(), rtol=2e-7, atol=1e-12)
    same = idx == j.numpy()
    assert same.mean() > 0.999 and np.allclose(full.numpy()[np.arange(700), idx], m.numpy(), rtol=1e-6)
    # the reference's default cdist (matmul form for > 25 rows) agrees to its own cancellation error
    np.testing.assert_allclose(oracle.chamfer_distance(a, b), _ref_chamfer(torch.from_numpy(a), torch.from_numpy(b)).item(), rtol=2e-4)
    np.testing.assert_allclose(oracle.minimum_distance(a, b, 50), _ref_minimum_distance(torch.from_numpy(a), torch.from_numpy(b), 50).item(),
                               rtol=1e-3, atol=1e-4)
    assert d[3] == 0.0 and idx[3] == 333 // 2


@pytest.mark.gpu
@pytest.mark.parametrize("na,nb", [(700, 333), (10475, 1500), (1500, 10475), (5, 3), (257, 1025), (1, 2049), (3000, 40000)])
def test_nearest_distance_forward_backward(na, nb):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from coma_b200 import geometry
    from oracle import oracle
    dev = torch.device("cuda:0")
    a, b = _clouds(na, nb, na + nb)
    rd, ri = oracle.nearest_distance(a, b)
    A = torch.tensor(a, device=dev, requires_grad=True)
    B = torch.tensor(b, device=dev, requires_grad=True)
    d, idx = geometry.nearest_distance(A, B)
    np.testing.assert_array_equal(d.detach().cpu().numpy(), rd)            # IEEE sqrt of exactly accumulated squares: bit-exact
    np.testing.assert_array_equal(idx.cpu().numpy(), ri)
    w = torch.linspace(0.5, 1.5, na, device=dev)
    (d * w).sum().backward()
    A2 = torch.tensor(a, device=dev, requires_grad=True)
    B2 = torch.tensor(b, device=dev, requires_grad=True)
    full = torch.cdist(A2, B2, compute_mode="donot_use_mm_for_euclid_dist")
    (full.gather(1, idx.long()[:, None])[:, 0] * w).sum().backward()     # same arg-min choice: the gradient of the selected entries
    torch.testing.assert_close(A.grad, A2.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(B.grad, B2.grad, rtol=1e-4, atol=1e-5)
    assert torch.isfinite(A.grad).all() and torch.isfinite(B.grad).all()


@pytest.mark.gpu
def test_chamfer_and_minimum_distance_match_reference():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from coma_b200 import geometry
    from oracle import oracle
    dev = torch.device("cuda:0")
    a, b = _clouds(1000, 180, 5)
    A = torch.tensor(a, device=dev, requires_grad=True)
    B = torch.tensor(b, device=dev)
    c = geometry.chamfer_distance(A, B)
    np.testing.assert_allclose(c.item(), float(oracle.chamfer_distance(a, b)), rtol=1e-6)
    A2 = torch.tensor(a, device=dev, requires_grad=True)
    cr = _ref_chamfer(A2, B)                                             # the reference's literal code on the same GPU
    np.testing.assert_allclose(c.item(), cr.item(), rtol=2e-4)
    c.backward()
    cr.backward()
    torch.testing.assert_close(A.grad, A2.grad, rtol=5e-3, atol=2e-6)    # cdist's matmul form: cancellation in the reference itself
    m = geometry.minimum_distance(A.detach()[None], B[None], 100)
    np.testing.assert_allclose(m.item(), float(oracle.minimum_distance(a, b, 100)), rtol=1e-6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        geometry.nearest_distance(torch.zeros(4, 3), torch.zeros(4, 3))
