"""Tensor digests for golden fixtures.  TEST INFRASTRUCTURE ONLY.

A digest is small enough to commit and still pins a tensor tightly: float64
sum, abs-sum, L2 norm and 64 strided samples of the flattened tensor."""
import numpy as np
import torch

N_SAMPLES = 64


def digest(t: torch.Tensor) -> np.ndarray:
    """-> float64[3 + N_SAMPLES] = [sum, abs_sum, l2, samples...]."""
    f = t.detach().reshape(-1).double().cpu()
    n = f.numel()
    idx = (torch.arange(N_SAMPLES, dtype=torch.int64) * max(n // N_SAMPLES, 1)) % max(n, 1)
    s = f[idx] if n else torch.zeros(N_SAMPLES, dtype=torch.float64)
    head = torch.tensor([float(f.sum()), float(f.abs().sum()), float(f.pow(2).sum().sqrt())],
                        dtype=torch.float64)
    return torch.cat((head, s)).numpy()


def digest_close(d_ref: np.ndarray, t: torch.Tensor, rtol: float, what: str = '') -> None:
    """Assert tensor ``t`` matches a stored digest within ``rtol`` (relative to
    the tensor's scale: |diff| <= rtol * max|sample|, norms within rtol)."""
    d = digest(t)
    l2 = max(abs(d_ref[2]), 1e-30)
    assert abs(d[2] - d_ref[2]) <= rtol * l2, '%s: l2 %.9g vs golden %.9g' % (what, d[2], d_ref[2])
    assert abs(d[1] - d_ref[1]) <= rtol * max(abs(d_ref[1]), 1e-30), \
        '%s: abs_sum %.9g vs golden %.9g' % (what, d[1], d_ref[1])
    scale = max(float(np.abs(d_ref[3:]).max()), 1e-30)
    err = float(np.abs(d[3:] - d_ref[3:]).max())
    assert err <= rtol * scale, '%s: sample err %.3g (scale %.3g, rtol %g)' % (what, err, scale, rtol)
