"""CPU: the composite-rigid-body identity behind "rnea_fold" (csrc/drm_common.cuh: stage_folded_table).

A link behind a fixed joint, at pose (R, p) in the frame A of its nearest movable ancestor, with spatial inertia
(Io, mc, m) in its own frame -- Io an ARBITRARY 3x3, like the reference's un-symmetrised inertia_mat -- loads A exactly like a
body with
    Io' = R Io R^T - S(p) S(c) - S(c) S(p) - m S(p) S(p),   mc' = c + m p,   m' = m,      c = R mc
attached to A directly: for every motion (w, v) of A the wrench "multiply in the link's frame, transform back"
(spatial_vector_algebra.py:321-338 then :281-291, what the reference does for every fixed link on every call) equals
"multiply by the folded inertia in A".  Chains of fixed links compose the same way."""
import numpy as np


def skew(a):
    return np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]], dtype=np.float64)


def rot(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    return q * np.sign(np.linalg.det(q))


def multiply(Io, mc, m, w, v):
    """(lin, ang) = I (w, v) with the reference's block structure (spatial_vector_algebra.py:321-338)."""
    return m * v - np.cross(mc, w), Io @ w + np.cross(mc, v)


def fold(Io, mc, m, R, p):
    c = R @ mc
    Io2 = R @ Io @ R.T - skew(p) @ skew(c) - skew(c) @ skew(p) - m * skew(p) @ skew(p)
    return Io2, c + m * p, m


def test_folded_inertia_gives_the_same_wrench_for_any_motion():
    rng = np.random.default_rng(0)
    for _ in range(50):
        R, p = rot(rng), rng.normal(size=3)
        Io, mc, m = rng.normal(size=(3, 3)), rng.normal(size=3), abs(rng.normal()) + 0.1      # Io not symmetric
        w, v = rng.normal(size=3), rng.normal(size=3)
        # through the link's own frame: motion A -> link, multiply, wrench link -> A
        w_l, v_l = R.T @ w, R.T @ (v + np.cross(w, p))
        lin_l, ang_l = multiply(Io, mc, m, w_l, v_l)
        lin_a, ang_a = R @ lin_l, R @ ang_l + np.cross(p, R @ lin_l)
        lin_f, ang_f = multiply(*fold(Io, mc, m, R, p), w, v)
        np.testing.assert_allclose(lin_f, lin_a, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(ang_f, ang_a, rtol=1e-12, atol=1e-12)


def test_folding_composes_along_a_chain_of_fixed_links():
    rng = np.random.default_rng(1)
    for _ in range(20):
        R1, p1, R2, p2 = rot(rng), rng.normal(size=3), rot(rng), rng.normal(size=3)
        Io, mc, m = rng.normal(size=(3, 3)), rng.normal(size=3), 1.3
        step = fold(*fold(Io, mc, m, R2, p2), R1, p1)                  # link -> fixed parent -> movable ancestor
        once = fold(Io, mc, m, R1 @ R2, p1 + R1 @ p2)                   # the composed pose used by the kernel
        for a, b in zip(step, once):
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)
