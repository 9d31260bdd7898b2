"""CPU: pin oracle/drm_oracle.py against the golden vectors generated from the reference itself
(tests/golden/make_golden.py) and against the known answers of SURVEY.md section 8(c)."""
import numpy as np
import pytest
import torch

from conftest import assert_close, canon_quat, load_golden, urdf_path
from oracle import drm_oracle as O


def _inputs(g, dtype):
    return tuple(torch.tensor(g[k], dtype=dtype) for k in ("q", "qd", "qdd"))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_matches_reference_forward(robot_stem, dtype):
    g = load_golden(robot_stem)
    robot = O.load_robot(urdf_path(robot_stem), dtype)
    q, qd, qdd = _inputs(g, dtype)

    # parsed link parameters are bit-identical to what the reference parsed
    for key, mine in (("trans", robot.trans), ("rpy", robot.rpy), ("axis", robot.axis), ("mass", robot.mass),
                      ("com", robot.com), ("inertia", robot.inertia), ("damping", robot.damping)):
        np.testing.assert_array_equal(mine.to(torch.float32).numpy(), g[key], err_msg=key)
    assert robot.parent == g["parent"].tolist() and robot.dof == g["dof"].tolist()
    assert robot.names == g["link_names"].tolist()

    R, p, _, _, _ = O.kinematic_state(robot, q, qd)
    assert_close(torch.stack(R).numpy(), g["all_R"], what="all-link rotations")
    assert_close(torch.stack(p).numpy(), g["all_p"], what="all-link positions")
    for link in g["fk_links"].tolist():
        pos, quat = O.forward_kinematics(robot, q, link)
        assert_close(pos.numpy(), g[f"pos.{link}"], what=f"pos {link}")
        assert_close(canon_quat(quat.numpy()), canon_quat(g[f"quat.{link}"]), what=f"quat {link}")
        jl, ja = O.jacobian(robot, q, link)
        assert_close(jl.numpy(), g[f"jlin.{link}"], what=f"jlin {link}")
        assert_close(ja.numpy(), g[f"jang.{link}"], what=f"jang {link}")
    for grav in (0, 1):
        for damp in (0, 1):
            tau = O.inverse_dynamics(robot, q, qd, qdd, bool(grav), bool(damp))
            # the reference's own inverse-dynamics tolerance is atol=1e-5 (tests/test_kinematics_dynamics.py:373-377)
            assert_close(tau.numpy(), g[f"tau.g{grav}d{damp}"], rtol=1e-5, atol=1e-5, what=f"tau g{grav} d{damp}")


def test_oracle_raw_quaternion_sign_matches_reference(robot_stem):
    """Away from branch boundaries even the raw (non-canonicalised) sign convention must agree."""
    g = load_golden(robot_stem)
    robot = O.load_robot(urdf_path(robot_stem), torch.float32)
    q = torch.tensor(g["q"])
    for link in g["fk_links"].tolist():
        _, quat = O.forward_kinematics(robot, q, link)
        ref = g[f"quat.{link}"]
        same = np.abs(quat.numpy() - ref).max(axis=1) < 1e-5
        assert same.mean() >= 0.85, f"{link}: raw quaternion sign differs on {np.sum(~same)} of {len(same)} rows"


def _grad_robot(stem, dtype):
    robot = O.load_robot(urdf_path(stem), dtype)
    for name in ("trans", "rpy", "mass", "com", "inertia", "damping"):
        getattr(robot, name).requires_grad_(True)
    return robot


def test_oracle_gradients_match_reference_autograd(robot_stem):
    """fp64 oracle + torch autograd vs the gradients of the reference's own fp32 autograd graph."""
    g = load_golden(robot_stem)
    dt = torch.float64
    q, qd, qdd = (t.requires_grad_(True) for t in _inputs(g, dt))
    G = {k: torch.tensor(g[k], dtype=dt) for k in ("G_pos", "G_jl", "G_ja", "G_tau")}
    param_of = {"trans": "trans", "rot_angles": "rpy", "mass": "mass", "com": "com", "inertia_mat": "inertia",
                "joint_damping": "damping"}

    def check(prefix, robot, loss, inputs):
        grads = torch.autograd.grad(loss, inputs + [robot.trans, robot.rpy, robot.mass, robot.com, robot.inertia,
                                                    robot.damping], allow_unused=True)
        n_in = len(inputs)
        by_name = dict(zip(["trans", "rpy", "mass", "com", "inertia", "damping"], grads[n_in:]))
        scale = max(float(np.abs(g[k]).max()) for k in g.files if k.startswith(prefix))
        tol = dict(rtol=2e-4, atol=2e-5 * max(scale, 1.0))     # fp32 reference autograd noise
        for t, key in zip(grads[:n_in], ("q", "qd", "qdd")):
            assert_close(t.numpy(), g[f"{prefix}.{key}"], what=f"{prefix}.{key}", **tol)
        checked = 0
        for key in g.files:
            if not key.startswith(prefix + "."):
                continue
            rest = key[len(prefix) + 1:]
            if "." not in rest:
                continue
            pname, idx = rest.rsplit(".", 1)
            if pname not in param_of:
                continue
            mine = by_name[param_of[pname]]
            mine = torch.zeros_like(getattr(robot, param_of[pname])) if mine is None else mine
            assert_close(mine[int(idx)].reshape(g[key].shape).numpy(), g[key], what=key, **tol)
            checked += 1
        assert checked > 0

    for link in g["fk_links"].tolist()[:2]:
        robot = _grad_robot(robot_stem, dt)
        pos, _ = O.forward_kinematics(robot, q, link)
        jl, ja = O.jacobian(robot, q, link)
        check(f"fkgrad.{link}", robot, (G["G_pos"] * pos).sum() + (G["G_jl"] * jl).sum() + (G["G_ja"] * ja).sum(), [q])
    robot = _grad_robot(robot_stem, dt)
    tau = O.inverse_dynamics(robot, q, qd, qdd, True, True)
    check("idgrad", robot, (G["G_tau"] * tau).sum(), [q, qd, qdd])


def test_known_answers():
    """Hand-checkable values quoted in SURVEY.md section 8(c) (Kuka, q_k = 0.1 k)."""
    robot = O.load_robot(urdf_path("iiwa7"), torch.float32)
    k = torch.arange(1, 8, dtype=torch.float32)
    q, qd, qdd = (0.1 * k)[None], (0.05 * k)[None], (-0.02 * k)[None]
    pos, quat = O.forward_kinematics(robot, torch.zeros(1, 7), "iiwa_link_ee")
    assert_close(pos.numpy(), [[0.0, 0.0, 1.266]], what="zero pose")
    assert_close(quat.numpy(), [[0.0, 0.0, 0.0, 1.0]], what="zero quat")
    pos, quat = O.forward_kinematics(robot, q, "iiwa_link_ee")
    assert_close(pos.numpy(), [[0.03738306, -0.00471164, 1.2391480]], what="pos")
    assert_close(quat.numpy(), [[-0.04092941, 0.19003928, 0.69464809, 0.69258499]], what="quat")
    tau = O.inverse_dynamics(robot, q, qd, qdd, True, True)
    assert_close(tau.numpy(), [[0.01269711, -8.3910265, -0.68258399, -4.2039914, 0.42146286, -1.0341269, 0.17390044]],
                 atol=1e-5, what="tau")
    tau = O.inverse_dynamics(robot, q, qd, qdd, False, False)
    assert_close(tau.numpy(), [[-0.01230314, -0.09288787, -0.00786535, 0.03695315, -0.00435376, -0.00580134,
                                -0.00109955]], atol=1e-5, what="tau no gravity")


# ------------------------------------------------------------------------------------------------
# the analytic adjoint recursions the backward kernels implement (oracle/adjoint_proto.py)
# ------------------------------------------------------------------------------------------------
from oracle import adjoint_proto as A  # noqa: E402

_PARAMS = ("trans", "rpy", "mass", "com", "inertia", "damping")


def _param_grads_via_table(robot, table_grad):
    table = O.link_table(robot)
    return torch.autograd.grad(table, [getattr(robot, p) for p in _PARAMS], grad_outputs=table_grad, allow_unused=True)


@pytest.mark.parametrize("stem", ["2link_robot", "iiwa7", "allegro_hand_description_left", "trifinger_edu",
                                  "jaco_clean", "fetch_arm_no_gripper", "iiwa7_allegro"])
def test_adjoint_recursions_match_autograd(stem):
    g = load_golden(stem)
    dt = torch.float64
    q, qd, qdd = (t[:5].clone().requires_grad_(True) for t in _inputs(g, dt))
    gen = torch.Generator().manual_seed(5)
    robot = _grad_robot(stem, dt)
    codes = O.axis_codes(robot)
    n = robot.n_dofs

    # inverse dynamics, all four flag combinations
    for grav, damp in ((True, True), (False, True), (True, False), (False, False)):
        G = torch.randn(5, n, generator=gen, dtype=dt)
        tau = O.inverse_dynamics(robot, q, qd, qdd, grav, damp)
        want = torch.autograd.grad((G * tau).sum(), [q, qd, qdd] + [getattr(robot, p) for p in _PARAMS],
                                   allow_unused=True)
        table = O.link_table(robot).detach()
        dq, dqd, dqdd, tg = A.inverse_dynamics_backward(table, robot.parent, codes, robot.dof, q.detach(), qd.detach(),
                                                        qdd.detach(), G, grav, damp)
        for got, w, name in zip((dq, dqd, dqdd), want[:3], ("q", "qd", "qdd")):
            assert_close(got.numpy(), w.numpy(), rtol=1e-9, atol=1e-9, what=f"{stem} id d{name}")
        for got, w, name in zip(_param_grads_via_table(robot, tg), want[3:], _PARAMS):
            w = torch.zeros_like(getattr(robot, name)) if w is None else w
            got = torch.zeros_like(w) if got is None else got
            assert_close(got[1:].numpy(), w[1:].numpy(), rtol=1e-9, atol=1e-9, what=f"{stem} id d{name}")

    # FK + Jacobian (+ quaternion) of every golden link
    for link in g["fk_links"].tolist():
        e = robot.index(link)
        Gp, Gq = torch.randn(5, 3, generator=gen, dtype=dt), torch.randn(5, 4, generator=gen, dtype=dt)
        Gl, Ga = torch.randn(5, 3, n, generator=gen, dtype=dt), torch.randn(5, 3, n, generator=gen, dtype=dt)
        pos, quat = O.forward_kinematics(robot, q, link)
        jl, ja = O.jacobian(robot, q, link)
        loss = (Gp * pos).sum() + (Gq * quat).sum() + (Gl * jl).sum() + (Ga * ja).sum()
        want = torch.autograd.grad(loss, [q] + [getattr(robot, p) for p in _PARAMS], allow_unused=True)
        table = O.link_table(robot).detach()
        dq, tg = A.fk_jacobian_backward(table, robot.parent, codes, robot.dof, e, q.detach(), Gp, Gq, Gl, Ga)
        assert_close(dq.numpy(), want[0].numpy(), rtol=1e-9, atol=1e-9, what=f"{stem} fk dq {link}")
        for got, w, name in zip(_param_grads_via_table(robot, tg), want[1:], _PARAMS):
            if name in ("trans", "rpy"):
                w = torch.zeros_like(getattr(robot, name)) if w is None else w
                # fixed-joint origins: the oracle (like the reference) treats them as constants baked in at
                # construction, the table carries their gradient -- compare movable links only
                mov = [i for i in range(len(robot.names)) if robot.dof[i] >= 0]
                assert_close(got[mov].numpy(), w[mov].numpy(), rtol=1e-9, atol=1e-9, what=f"{stem} fk d{name} {link}")


@pytest.mark.parametrize("stem", ["iiwa7", "panda_no_gripper"])
def test_two_sweep_chain_adjoint_matches_the_four_pass_recursion(stem):
    """oracle/adjoint_proto.py: inverse_dynamics_backward_chain (what csrc/backward_rnea.cu's chain kernel evaluates) against
    inverse_dynamics_backward (checked against autograd above), on robots whose rows are canonical as they stand (all
    movable axes +z) and on random chains with fixed links in the middle."""
    dt = torch.float64
    robot = _grad_robot(stem, dt)
    codes = O.axis_codes(robot)
    assert set(codes) <= {0, 3} and all(robot.parent[i] == i - 1 for i in range(1, len(robot.parent)))
    g = load_golden(stem)
    q, qd, qdd = (t[:7] for t in _inputs(g, dt))
    gen = torch.Generator().manual_seed(9)
    table = O.link_table(robot).detach()
    cases = [(table, list(robot.dof), q, qd, qdd)]
    for axis in ([0, 3, 3, 0, 3, 3, 3, 0, 3], [0, 3, 3], [0, 0, 3, 3, 0]):
        N = len(axis)
        dof, k = [-1] * N, 0
        for i in range(N):
            if axis[i] != 0:
                dof[i] = k
                k += 1
        t = torch.zeros(N, 28, dtype=dt)
        for i in range(1, N):
            t[i, 0:9] = torch.linalg.qr(torch.randn(3, 3, generator=gen, dtype=dt))[0].reshape(9)
            t[i, 9:24] = torch.randn(15, generator=gen, dtype=dt)
            t[i, 24:26] = torch.rand(2, generator=gen, dtype=dt) + 0.5
        cases.append((t, dof) + tuple(torch.randn(7, k, generator=gen, dtype=dt) for _ in range(3)))
    for t, dof, a, b, c in cases:
        N, n = t.shape[0], a.shape[1]
        axis = [3 if d >= 0 else 0 for d in dof]
        for grav, damp in ((True, True), (False, True), (True, False)):
            G = torch.randn(7, n, generator=gen, dtype=dt)
            want = A.inverse_dynamics_backward(t, [-1] + list(range(N - 1)), axis, dof, a, b, c, G, grav, damp)
            got = A.inverse_dynamics_backward_chain(t, dof, a, b, c, G, grav, damp)
            for x, y, name in zip(got, want, ("q", "qd", "qdd", "table")):
                assert_close(x.numpy(), y.numpy(), rtol=1e-9, atol=1e-9 * max(1.0, float(y.abs().max())), what=f"{stem} chain d{name}")


def test_two_sweep_tree_adjoint_matches_the_four_pass_recursion():
    """oracle/adjoint_proto.py: inverse_dynamics_backward_two_sweep_tree (the form DESIGN.md section 10 proposes for the
    tree kernel) against the four-pass recursion, on random canonical trees with branch points and fixed links."""
    dt = torch.float64
    gen = torch.Generator().manual_seed(21)
    trees = [[-1, 0, 1, 2, 1, 4, 5, 0, 7], [-1, 0, 1, 2, 3, 4, 5, 6, 7], [-1, 0, 0, 0, 1, 1, 2, 5, 5, 8], [-1, 0, 1, 1, 3, 3, 0, 6]]
    for parent in trees:
        N = len(parent)
        for with_fixed in (False, True):
            axis = [0] + [3] * (N - 1)
            if with_fixed:
                axis[2] = axis[N - 1] = 0
            dof, k = [-1] * N, 0
            for i in range(N):
                if axis[i] != 0:
                    dof[i] = k
                    k += 1
            t = torch.zeros(N, 28, dtype=dt)
            for i in range(1, N):
                t[i, 0:9] = torch.linalg.qr(torch.randn(3, 3, generator=gen, dtype=dt))[0].reshape(9)
                t[i, 9:24] = torch.randn(15, generator=gen, dtype=dt)
                t[i, 24:26] = torch.rand(2, generator=gen, dtype=dt) + 0.5
            a, b, c, G = (torch.randn(5, k, generator=gen, dtype=dt) for _ in range(4))
            for grav, damp in ((True, True), (False, False)):
                want = A.inverse_dynamics_backward(t, parent, axis, dof, a, b, c, G, grav, damp)
                got = A.inverse_dynamics_backward_two_sweep_tree(t, parent, dof, a, b, c, G, grav, damp)
                for x, y, name in zip(got, want, ("q", "qd", "qdd", "table")):
                    assert_close(x.numpy(), y.numpy(), rtol=1e-9, atol=1e-9 * max(1.0, float(y.abs().max())), what=f"tree {parent} d{name}")
