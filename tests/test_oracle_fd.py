"""CPU: pin the oracle's articulated-body restatement (oracle/drm_oracle.py: forward_dynamics) against golden
vectors generated from the reference's compute_forward_dynamics (tests/golden/make_golden_fd.py;
reference: robot_model.py:488-624), forward values and autograd gradients, symmetric and non-symmetric inertia."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, assert_close, urdf_path
from oracle import drm_oracle as O

PARAM_OF = {"trans": "trans", "rot_angles": "rpy", "mass": "mass", "com": "com", "inertia_mat": "inertia",
            "joint_damping": "damping"}


# With the perturbed (non-symmetric) inertias the Kinova hands are so ill-conditioned that the reference's own fp32
# evaluation is ~1e-2 away from the fp64 evaluation of the same formulas; those vectors pin nothing.
ILL_CONDITIONED_NONSYM = {"jaco", "jaco_clean"}


def load_fd(stem):
    return np.load(os.path.join(GOLDEN_DIR, stem + ".fd.npz"), allow_pickle=False)


def grad_robot(stem, dtype, g=None, tag="sym"):
    robot = O.load_robot(urdf_path(stem), dtype)
    if tag == "nonsym":
        inertia = torch.tensor(g["nonsym.inertia"], dtype=dtype)
        inertia[0] = robot.inertia[0]
        robot.inertia = inertia
    for name in ("trans", "rpy", "mass", "com", "inertia", "damping"):
        setattr(robot, name, getattr(robot, name).detach().clone().requires_grad_(True))
    return robot


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_forward_dynamics_matches_reference(robot_stem, dtype):
    g = load_fd(robot_stem)
    robot = O.load_robot(urdf_path(robot_stem), dtype)
    q, qd, f = (torch.tensor(g[k], dtype=dtype) for k in ("q", "qd", "f"))
    for grav in (0, 1):
        for damp in (0, 1):
            want = g[f"qdd.g{grav}d{damp}"]
            got = O.forward_dynamics(robot, q, qd, f, bool(grav), bool(damp))
            # fp32 evaluation noise of the reference itself: compare normwise per configuration
            scale = np.abs(want).max(axis=1, keepdims=True)
            assert np.all(np.abs(got.numpy() - want) <= 2e-4 * scale + 1e-6), (grav, damp)


@pytest.mark.parametrize("tag", ["sym", "nonsym"])
def test_forward_dynamics_gradients_match_reference_autograd(robot_stem, tag):
    if tag == "nonsym" and robot_stem in ILL_CONDITIONED_NONSYM:
        pytest.skip("fp32 reference vectors are not reproducible to better than 1e-2 for this model")
    g = load_fd(robot_stem)
    dt = torch.float64
    robot = grad_robot(robot_stem, dt, g, tag)
    q, qd, f = (torch.tensor(g[k], dtype=dt).requires_grad_(True) for k in ("q", "qd", "f"))
    qdd = O.forward_dynamics(robot, q, qd, f, True, True)
    want = g[f"{tag}.qdd"]
    assert np.all(np.abs(qdd.detach().numpy() - want) <= 2e-4 * np.abs(want).max(axis=1, keepdims=True) + 1e-6)
    loss = (torch.tensor(g["G_qdd"], dtype=dt) * qdd).sum()
    params = [robot.trans, robot.rpy, robot.mass, robot.com, robot.inertia, robot.damping]
    grads = torch.autograd.grad(loss, [q, qd, f] + params, allow_unused=True)
    by_name = dict(zip(["trans", "rpy", "mass", "com", "inertia", "damping"], grads[3:]))
    for t, key in zip(grads[:3], ("q", "qd", "f")):
        ref = g[f"{tag}.grad.{key}"]
        assert_close(t.numpy(), ref, rtol=2e-3, atol=2e-4 * max(np.abs(ref).max(), 1e-3), what=f"{tag}.{key}")
    checked = 0
    prefix = f"{tag}.grad."
    for key in g.files:
        if not key.startswith(prefix) or key[len(prefix):] in ("q", "qd", "f"):
            continue
        pname, idx = key[len(prefix):].rsplit(".", 1)
        mine = by_name[PARAM_OF[pname]]
        mine = torch.zeros_like(getattr(robot, PARAM_OF[pname])) if mine is None else mine
        ref = g[key]
        fam = max(np.abs(g[k]).max() for k in g.files if k.startswith(prefix + pname + "."))
        assert_close(mine[int(idx)].reshape(ref.shape).numpy(), ref, rtol=2e-3, atol=2e-4 * max(fam, 1e-6), what=key)
        checked += 1
    assert checked > 0


@pytest.mark.parametrize("stem", ["2link_robot", "iiwa7", "panda", "trifinger_edu", "iiwa7_allegro"])
def test_forward_dynamics_adjoint_recursions_match_autograd(stem):
    """The hand-derived adjoint of the articulated-body algorithm (oracle/adjoint_proto.py -- the recursions the
    CUDA backward kernel evaluates) against torch.autograd of the oracle, fp64, non-symmetric inertias."""
    from oracle import adjoint_proto as AP
    dt = torch.float64
    robot = O.load_robot(urdf_path(stem), dt)
    gen = torch.Generator().manual_seed(11)
    scale = robot.inertia.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-6)
    robot.inertia = robot.inertia + 0.05 * scale * torch.randn(robot.inertia.shape, generator=gen, dtype=dt)
    q, qd, _ = O.sample_inputs(robot, 4, seed=3, dtype=dt)
    f = torch.randn(4, robot.n_dofs, generator=gen, dtype=dt)
    G = torch.randn(4, robot.n_dofs, generator=gen, dtype=dt)
    table = O.link_table(robot).to(dt)
    qg, qdg, fg = (t.clone().requires_grad_(True) for t in (q, qd, f))
    names = ("trans", "rpy", "mass", "com", "inertia", "damping")
    for name in names:
        setattr(robot, name, getattr(robot, name).detach().clone().requires_grad_(True))
    params = [getattr(robot, name) for name in names]
    for grav, damp in ((True, True), (False, False)):
        qdd_o = O.forward_dynamics(robot, qg, qdg, fg, grav, damp)
        want = torch.autograd.grad((G * qdd_o).sum(), [qg, qdg, fg] + params, allow_unused=True)
        qdd, dq, dqd, df, tg = AP.forward_dynamics_with_backward(table, robot.parent, O.axis_codes(robot), robot.dof,
                                                                 q, qd, f, G, grav, damp)
        got_params = torch.autograd.grad((O.link_table(robot) * tg).sum(), params, allow_unused=True)
        rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-300))  # noqa: E731
        assert rel(qdd, qdd_o.detach()) < 1e-12
        for a, b in zip((dq, dqd, df) + tuple(got_params), want):
            if b is None:
                continue
            a = torch.zeros_like(b) if a is None else a
            assert rel(a, b) < 1e-9, (stem, grav, damp)


def test_mass_matrix_construction_matches_reference(robot_stem):
    """Column j = ID(q, 0, e_j) - ID(q, 0, 0) through the oracle's inverse dynamics vs the reference's
    compute_lagrangian_inertia_matrix (robot_model.py:403-450); the reference's fp32 subtraction of the gravity terms
    leaves noise of ~1e-6 x the gravity torque, hence the absolute floor."""
    g = load_fd(robot_stem)
    robot = O.load_robot(urdf_path(robot_stem), torch.float64)
    q = torch.tensor(g["q"], dtype=torch.float64)
    z = torch.zeros_like(q)
    cols = []
    for j in range(robot.n_dofs):
        e = z.clone()
        e[:, j] = 1
        cols.append(O.inverse_dynamics(robot, q, z, e, False, False))
    H = torch.stack(cols, dim=2).numpy()
    grav = np.abs(O.inverse_dynamics(robot, q, z, z, True, False).numpy()).max()
    for key in ("H.g1d1", "H.g0d0"):
        floor = 2e-5 * np.abs(g[key]).max() + (4e-6 * grav if key == "H.g1d1" else 0.0)
        assert np.all(np.abs(H - g[key]) <= 2e-4 * np.abs(g[key]) + floor), key
