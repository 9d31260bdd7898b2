"""
Link-parameter parametrisations
====================================
The small ``torch.nn.Module``s callers plug into ``make_link_param_learnable``.  They are host-side
producers of scalars / ``[1,3]`` / ``[3,3]`` tensors that ``link_table.build_link_table`` consumes; the
engine does not care which module produced a value, so the reference's own parametrisation classes
(``rigid_body_params.py:14-403``) work unchanged against this package as well.  Only the three
unconstrained / positive ones the kinematics and inverse-dynamics examples use are provided here
(``examples/learn_dynamics_iiwa.py:57-65``, ``examples/learn_kinematics_of_iiwa.py:33-38``).
"""
import torch


class UnconstrainedScalar(torch.nn.Module):
    """A free scalar (reference: rigid_body_params.py:14-23)."""

    def __init__(self, init_val=None):
        super().__init__()
        self.param = torch.nn.Parameter(torch.rand(1) if init_val is None else init_val)

    def forward(self):
        return self.param


class PositiveScalar(torch.nn.Module):
    """``l^2 + min_val`` (reference: rigid_body_params.py:26-43)."""

    def __init__(self, min_val=0.0, init_param_std=1.0, init_param=None):
        super().__init__()
        self._min_val = min_val
        if init_param is None:
            start = torch.empty(1, 1).normal_(mean=0.0, std=init_param_std)
        else:
            start = torch.sqrt(init_param - self._min_val)
        self.l = torch.nn.Parameter(start.squeeze())

    def forward(self):
        return (self.l * self.l + self._min_val).squeeze()


class UnconstrainedTensor(torch.nn.Module):
    """A free ``[dim1, dim2]`` tensor (reference: rigid_body_params.py:46-56)."""

    def __init__(self, dim1, dim2, init_tensor=None, init_std=0.1):
        super().__init__()
        self._dim1, self._dim2 = dim1, dim2
        if init_tensor is None:
            init_tensor = torch.empty(dim1, dim2).normal_(mean=0.0, std=init_std)
        self.param = torch.nn.Parameter(init_tensor)

    def forward(self):
        return self.param
