"""GPU: the analytic adjoint of the articulated-body kernel (csrc/backward_aba.cu) through torch.autograd, against

  * the gradients of the reference's own autograd graph of compute_forward_dynamics (tests/golden/*.fd.npz), at the
    URDF parameters and with every inertia_mat perturbed to a NON-symmetric matrix;
  * torch.autograd of the fp64 oracle on larger seeded batches;
and that table gradients are bitwise reproducible.

Tolerance: forward dynamics amplifies fp32 rounding by the condition number of the articulated inertias, so each
gradient family (q, qd, f, and each link-parameter kind) is compared relative to the largest entry of that family:
arms 2e-3 (the reference's fp32 autograd is itself only reproducible to ~1e-3 there), hand models 2e-2.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, urdf_path
from test_backward_gpu import _ORACLE_PARAM, cuda, learnable_model
from oracle import drm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ARMS = {"iiwa7", "panda_no_gripper", "panda", "fetch_arm_no_gripper", "fetch_arm_no_gripper_small_damping", "2link_robot"}
ILL_CONDITIONED_NONSYM = {"jaco", "jaco_clean"}       # see tests/test_oracle_fd.py


def load_fd(stem):
    return np.load(os.path.join(GOLDEN_DIR, stem + ".fd.npz"), allow_pickle=False)


def family_close(got, want, fam_scale, tol, what):
    got, want = np.asarray(got, dtype=np.float64).reshape(-1), np.asarray(want, dtype=np.float64).reshape(-1)
    err = np.abs(got - want).max() if got.size else 0.0
    assert err <= tol * max(fam_scale, 1e-30), f"{what}: |err| {err:.3e} vs family scale {fam_scale:.3e} (tol {tol})"


@pytest.mark.parametrize("tag", ["sym", "nonsym"])
def test_forward_dynamics_gradients_match_reference_autograd(robot_stem, tag):
    if tag == "nonsym" and robot_stem in ILL_CONDITIONED_NONSYM:
        pytest.skip("fp32 reference vectors are not reproducible to better than 1e-2 for this model")
    g = load_fd(robot_stem)
    m, params = learnable_model(robot_stem)
    if tag == "nonsym":
        with torch.no_grad():
            for (i, pname), p in params.items():
                if pname == "inertia_mat":
                    p.copy_(cuda(g["nonsym.inertia"][i]))
    q, qd, f = cuda(g["q"], True), cuda(g["qd"], True), cuda(g["f"], True)
    qdd = m.compute_forward_dynamics(q, qd, f, include_gravity=True, use_damping=True)
    want = g[f"{tag}.qdd"]
    rel = np.abs(qdd.detach().cpu().numpy() - want) / np.abs(want).max(axis=1, keepdims=True)
    assert rel.max() < (2e-4 if robot_stem in ARMS else 2e-3)
    (cuda(g["G_qdd"]) * qdd).sum().backward()
    tol = 2e-3 if robot_stem in ARMS else 2e-2
    prefix = f"{tag}.grad."
    for key, t in (("q", q), ("qd", qd), ("f", f)):
        ref = g[prefix + key]
        family_close(t.grad.cpu().numpy(), ref, np.abs(ref).max(), tol, f"{tag}.{key}")
    checked = 0
    for key in g.files:
        if not key.startswith(prefix) or key[len(prefix):] in ("q", "qd", "f"):
            continue
        pname, idx = key[len(prefix):].rsplit(".", 1)
        p = params[(int(idx), pname)]
        got = torch.zeros_like(p) if p.grad is None else p.grad
        fam = max(np.abs(g[k]).max() for k in g.files if k.startswith(prefix + pname + "."))
        family_close(got.cpu().numpy(), g[key], fam, tol, key)
        checked += 1
    assert checked > 0


@pytest.mark.parametrize("stem,batch,grav,damp,nonsym", [("iiwa7", 1000, True, True, True), ("iiwa7", 333, False, False, False),
                                                         ("panda", 130, True, False, True), ("2link_robot", 65, True, True, True),
                                                         ("trifinger_edu", 97, True, True, False),
                                                         ("iiwa7_allegro", 50, True, True, False)])
def test_forward_dynamics_gradients_match_fp64_oracle(stem, batch, grav, damp, nonsym):
    robot = O.load_robot(urdf_path(stem), torch.float64)
    m, params = learnable_model(stem)
    if nonsym:
        gen = torch.Generator().manual_seed(17)
        scale = robot.inertia.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-6)
        robot.inertia = (robot.inertia + 0.05 * scale * torch.randn(robot.inertia.shape, generator=gen, dtype=torch.float64)).float().double()
        with torch.no_grad():
            for (i, pname), p in params.items():
                if pname == "inertia_mat":
                    p.copy_(robot.inertia[i].float().to(DEV))
    q, qd, _ = (t.float() for t in O.sample_inputs(robot, batch, seed=batch + 1, dtype=torch.float64))
    gen = torch.Generator().manual_seed(batch)
    f = torch.randn(batch, robot.n_dofs, generator=gen)
    G = torch.randn(batch, robot.n_dofs, generator=gen)
    qg, qdg, fg = (t.to(DEV).requires_grad_(True) for t in (q, qd, f))
    qdd = m.compute_forward_dynamics(qg, qdg, fg, include_gravity=grav, use_damping=damp)
    (G.to(DEV) * qdd).sum().backward()

    names = ("trans", "rpy", "mass", "com", "inertia", "damping")
    for name in names:
        setattr(robot, name, getattr(robot, name).detach().clone().requires_grad_(True))
    ins = [t.double().requires_grad_(True) for t in (q, qd, f)]
    qdd_o = O.forward_dynamics(robot, *ins, grav, damp)
    grads = torch.autograd.grad((G.double() * qdd_o).sum(), ins + [getattr(robot, nm) for nm in names], allow_unused=True)
    by = dict(zip(names, grads[3:]))
    tol = 2e-3 if stem in ARMS else 2e-2
    rel = (qdd.detach().cpu().double() - qdd_o.detach()).abs() / qdd_o.detach().abs().amax(dim=1, keepdim=True)
    assert rel.max() < (2e-4 if stem in ARMS else 5e-3)
    for t, w, what in zip((qg, qdg, fg), grads[:3], ("q", "qd", "f")):
        family_close(t.grad.cpu().numpy(), w.numpy(), float(w.abs().max()), tol, what)
    for (i, pname), p in params.items():
        want = by[_ORACLE_PARAM[pname]]
        if want is None:
            want = torch.zeros_like(getattr(robot, _ORACLE_PARAM[pname]))
        got = torch.zeros_like(p) if p.grad is None else p.grad
        family_close(got.cpu().numpy(), want[i].numpy(), float(want.abs().max()), tol, f"{pname}.{i}")


def test_table_gradient_is_bitwise_reproducible_and_input_only_path_works():
    import differentiable_robot_model_b200 as drm
    robot = O.load_robot(urdf_path("iiwa7"), torch.float32)
    q, qd, _ = O.sample_inputs(robot, 4099, seed=2)
    f = torch.randn(4099, 7, generator=torch.Generator().manual_seed(0))
    G = torch.randn(4099, 7, generator=torch.Generator().manual_seed(1)).to(DEV)

    def run():
        m, params = learnable_model("iiwa7")
        qdd = m.compute_forward_dynamics(q.to(DEV), qd.to(DEV), f.to(DEV), use_damping=True)
        (G * qdd).sum().backward()
        return torch.cat([p.grad.reshape(-1) for p in params.values()])

    a, b = run(), run()
    assert torch.equal(a, b)
    # constant model: only input gradients, no table-gradient reduction launched
    m = drm.DifferentiableKUKAiiwa(device=DEV)
    qg, qdg, fg = (t.to(DEV).requires_grad_(True) for t in (q, qd, f))
    (G * m.compute_forward_dynamics(qg, qdg, fg)).sum().backward()
    m2, _ = learnable_model("iiwa7")
    q2, qd2, f2 = (t.to(DEV).requires_grad_(True) for t in (q, qd, f))
    (G * m2.compute_forward_dynamics(q2, qd2, f2)).sum().backward()
    for x, y in ((qg, q2), (qdg, qd2), (fg, f2)):
        assert torch.equal(x.grad, y.grad)


def test_learning_link_inertia_from_accelerations():
    """The reference's forward-dynamics example in miniature (examples/learn_forward_dynamics_iiwa.py:54-99)."""
    import differentiable_robot_model_b200 as drm
    from differentiable_robot_model_b200.rigid_body_params import PositiveScalar, UnconstrainedTensor
    torch.manual_seed(0)
    gt = drm.DifferentiableKUKAiiwa(device=DEV)
    m = drm.DifferentiableRobotModel(gt.urdf_path, "learn", device=DEV)
    m.make_link_param_learnable("iiwa_link_1", "mass", PositiveScalar())
    m.make_link_param_learnable("iiwa_link_1", "com", UnconstrainedTensor(dim1=1, dim2=3))
    m.make_link_param_learnable("iiwa_link_1", "inertia_mat", UnconstrainedTensor(dim1=3, dim2=3))
    robot = O.load_robot(gt.urdf_path, torch.float32)
    q, qd, _ = (t.to(DEV) for t in O.sample_inputs(robot, 2048, seed=4))
    tau = 2.0 * torch.randn(2048, 7, device=DEV)
    with torch.no_grad():
        target = gt.compute_forward_dynamics(q, qd, tau, use_damping=True)
    var = target.var(dim=0)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    losses = []
    for _ in range(150):
        opt.zero_grad()
        pred = m.compute_forward_dynamics(q, qd, tau, use_damping=True)
        loss = (((pred - target) ** 2) / var).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert np.isfinite(losses).all()
    assert losses[-1] < 0.2 * losses[0], (losses[0], losses[-1])
