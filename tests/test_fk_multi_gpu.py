"""GPU: the multi-end-effector tree-walk kernel (csrc/fk_tree.cu) against
  * the single-link kernel, link by link (same arithmetic in the same order -> BIT-identical),
  * the reference's own outputs (tests/golden/large_allegro_left.npz: four fingertips, 512 rows; the 9-row goldens of
    every shipped URDF), and
  * its autograd Function against the single-link Functions (gradients w.r.t. q and learnable link parameters)."""
import pytest
import torch

from conftest import assert_close, canon_quat, load_golden, urdf_path
import differentiable_robot_model_b200 as drm
from differentiable_robot_model_b200 import engine
from differentiable_robot_model_b200.rigid_body_params import UnconstrainedTensor
from oracle import drm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CASES = {
    "allegro_hand_description_left": ["link_3.0_tip", "link_7.0_tip", "link_11.0_tip", "link_15.0_tip"],
    "trifinger_edu": ["finger_tip_link_0", "finger_tip_link_120", "finger_tip_link_240"],
    "iiwa7_allegro": ["link_15.0_tip", "link_3.0_tip", "palm_link", "iiwa_link_4", "link_7.0"],
    "iiwa7": ["iiwa_link_ee", "iiwa_link_4", "iiwa_link_0", "iiwa_link_7"],          # nested paths + the root itself
    "panda": ["panda_leftfinger", "panda_rightfinger", "panda_virtual_ee_link"],
    "2link_robot": ["endEffector", "arm2"],
}


def model(stem):
    return drm.DifferentiableRobotModel(urdf_path(stem), stem, device=DEV)


@pytest.mark.parametrize("stem", sorted(CASES))
@pytest.mark.parametrize("batch", [1, 31, 32, 1000, 4099, 40000])
def test_multi_equals_single_link_kernel_bit_for_bit(stem, batch):
    m = model(stem)
    robot = O.load_robot(urdf_path(stem), torch.float32)
    q = O.sample_inputs(robot, batch, seed=batch)[0].to(DEV)
    links = CASES[stem]
    got = m.compute_fk_and_jacobian_multi(q, links)
    jac_only = m.compute_endeffector_jacobians(q, links[:1])
    for name in links:
        want = m.compute_fk_and_jacobian(q, name)
        for a, b, what in zip(got[name], want, ("pos", "quat", "jlin", "jang")):
            assert torch.equal(a, b), f"{stem} {name} {what} differs at batch {batch}"
    for a, b in zip(jac_only[links[0]], m.compute_endeffector_jacobian(q, links[0])):
        assert torch.equal(a, b)


def test_multi_matches_reference_golden_fingertips():
    g = load_golden("large_allegro_left")
    m = model("allegro_hand_description_left")
    q = torch.tensor(g["q"], device=DEV)
    links = g["links"].tolist()
    m._link_table()
    base = engine.launch_count()
    out = m.compute_fk_and_jacobian_multi(q, links)
    assert engine.launch_count() - base == 1                                    # ONE launch for the four fingertips
    for name in links:
        pos, quat, jl, ja = (t.cpu().numpy() for t in out[name])
        assert_close(pos, g[f"pos.{name}"], what=f"pos {name}")
        assert_close(canon_quat(quat), canon_quat(g[f"quat.{name}"]), atol=2e-6, what=f"quat {name}")
        assert_close(jl, g[f"jlin.{name}"], what=f"jlin {name}")
        assert_close(ja, g[f"jang.{name}"], what=f"jang {name}")


def test_multi_small_goldens_every_urdf(robot_stem):
    g = load_golden(robot_stem)
    m = model(robot_stem)
    links = g["fk_links"].tolist()[:8]
    out = m.compute_fk_and_jacobian_multi(torch.tensor(g["q"], device=DEV), links)
    for name in links:
        pos, quat, jl, ja = (t.cpu().numpy() for t in out[name])
        assert_close(pos, g[f"pos.{name}"], what=f"{robot_stem} pos {name}")
        assert_close(canon_quat(quat), canon_quat(g[f"quat.{name}"]), atol=2e-6, what=f"{robot_stem} quat {name}")
        assert_close(jl, g[f"jlin.{name}"], what=f"{robot_stem} jlin {name}")
        assert_close(ja, g[f"jang.{name}"], what=f"{robot_stem} jang {name}")


def test_multi_gradients_equal_sum_of_single_link_gradients():
    stem, links = "allegro_hand_description_left", CASES["allegro_hand_description_left"]
    robot = O.load_robot(urdf_path(stem), torch.float32)
    q0 = O.sample_inputs(robot, 513, seed=9)[0].to(DEV)

    def run(multi):
        m = model(stem)
        body = m._bodies[m._name_to_idx_map["link_2.0"]]
        m.make_link_param_learnable("link_2.0", "trans", UnconstrainedTensor(dim1=1, dim2=3, init_tensor=body.trans().detach().clone()))
        q = q0.clone().requires_grad_(True)
        gen = torch.Generator().manual_seed(3)
        loss = 0.0
        outs = m.compute_fk_and_jacobian_multi(q, links) if multi else {n: m.compute_fk_and_jacobian(q, n) for n in links}
        for name in links:
            for t in outs[name]:
                loss = loss + (t * torch.randn(t.shape, generator=gen).to(DEV)).sum()
        loss.backward()
        return q.grad.clone(), [p.grad.clone() for p in m.parameters()]

    gq_m, gp_m = run(True)
    gq_s, gp_s = run(False)
    assert_close(gq_m.cpu().numpy(), gq_s.cpu().numpy(), rtol=1e-5, atol=1e-5, what="q grad")
    for a, b in zip(gp_m, gp_s):
        assert_close(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-4, atol=1e-4, what="param grad")


def test_multi_argument_errors():
    m = model("iiwa7")
    q = torch.zeros(4, 7, device=DEV)
    with pytest.raises(KeyError):
        m.compute_fk_and_jacobian_multi(q, ["no_such_link"])
    with pytest.raises(AssertionError):
        m.compute_fk_and_jacobian_multi(q, ["iiwa_link_ee", "iiwa_link_ee"])
    with pytest.raises(RuntimeError):
        m.compute_fk_and_jacobian_multi(q, [f"iiwa_link_{i}" for i in range(8)] + ["iiwa_link_ee"])     # 9 > 8 links


@pytest.mark.parametrize("stem", ["allegro_hand_description_left", "iiwa7_allegro", "trifinger_edu"])
def test_multi_output_subsets_and_unaligned_inputs(stem):
    """Pose-only / Jacobian-only launches (other template instantiations of the tree kernel) and a q whose base pointer is
    not 16-byte aligned (cooperative-copy staging instead of TMA bulk copies) give the same bits."""
    m = model(stem)
    robot = O.load_robot(urdf_path(stem), torch.float32)
    links = [m._name_to_idx_map[n] for n in CASES[stem]]
    table, topo = m._link_table(), m._topology
    for batch in (257, 2048):
        q = O.sample_inputs(robot, batch + 1, seed=3)[0].to(DEV)
        full = engine.fk_jacobian_multi_raw(topo, links, table, q[1:].contiguous())
        pos, quat, none1, none2 = engine.fk_jacobian_multi_raw(topo, links, table, q[1:].contiguous(), want_jac=False)
        assert none1 is None and none2 is None and torch.equal(pos, full[0]) and torch.equal(quat, full[1])
        p2, q2, jl, ja = engine.fk_jacobian_multi_raw(topo, links, table, q[1:].contiguous(), want_pos=False, want_quat=False)
        assert p2 is None and q2 is None and torch.equal(jl, full[2]) and torch.equal(ja, full[3])
        # unaligned: the same rows one float into a fresh allocation
        flat = torch.empty(batch * q.shape[1] + 1, device=DEV)
        flat[1:] = q[1:].reshape(-1)
        q_un = flat[1:].view(batch, q.shape[1])
        assert q_un.data_ptr() % 16 != 0 and q_un.is_contiguous()
        un = engine.fk_jacobian_multi_raw(topo, links, table, q_un)
        for a, b in zip(un, full):
            assert torch.equal(a, b)
