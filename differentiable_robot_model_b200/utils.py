"""
Small tensor helpers
====================================
Host-side helpers with the names and semantics of the reference's ``utils.py`` (``utils.py:21-87``), for callers
that import them directly.  None of them is on the compute path of this package (the kernels do their own 3x3
algebra); unlike the reference they allocate on the device of their argument, so they work on CUDA tensors.
"""
import numpy as np
import torch


def convert_into_pytorch_tensor(variable):
    """``utils.py:72-78``"""
    if isinstance(variable, torch.Tensor):
        return variable
    return torch.as_tensor(np.asarray(variable), dtype=torch.float32)


def convert_into_at_least_2d_pytorch_tensor(variable):
    """``utils.py:81-87``"""
    t = convert_into_pytorch_tensor(variable)
    return t.unsqueeze(0) if t.dim() == 1 else t


def vector3_to_skew_symm_matrix(vec3):
    """``[B,3]`` (or ``[3]``) -> ``[B,3,3]`` with ``S(a) b = a x b`` (``utils.py:40-50``)."""
    v = convert_into_at_least_2d_pytorch_tensor(vec3)
    zero = torch.zeros_like(v[:, 0])
    return torch.stack([torch.stack([zero, -v[:, 2], v[:, 1]], dim=1),
                        torch.stack([v[:, 2], zero, -v[:, 0]], dim=1),
                        torch.stack([-v[:, 1], v[:, 0], zero], dim=1)], dim=1)


def cross_product(vec3a, vec3b):
    """Row-wise ``a x b`` through the skew matrix, like ``utils.py:21-25``."""
    a = convert_into_at_least_2d_pytorch_tensor(vec3a)
    b = convert_into_at_least_2d_pytorch_tensor(vec3b)
    return (vector3_to_skew_symm_matrix(a) @ b.unsqueeze(2)).squeeze(2)


def bfill_lowertriangle(A: torch.Tensor, vec: torch.Tensor):
    """Write ``vec`` into the strict lower triangle of the trailing two dimensions, in place (``utils.py:28-31``)."""
    ii, jj = np.tril_indices(A.size(-2), k=-1, m=A.size(-1))
    A[..., ii, jj] = vec
    return A


def bfill_diagonal(A: torch.Tensor, vec: torch.Tensor):
    """Write ``vec`` onto the diagonal of the trailing two dimensions, in place (``utils.py:34-37``)."""
    ii, jj = np.diag_indices(min(A.size(-2), A.size(-1)))
    A[..., ii, jj] = vec
    return A


def torch_square(x):
    return x * x


def exp_map_so3(omega, epsilon=1.0e-14):
    """Rodrigues formula with the reference's regulariser (``utils.py:57-69``)."""
    from .rigid_body_params import exp_map_so3 as _exp
    return _exp(omega.reshape(3), epsilon)
