"""TEST INFRASTRUCTURE -- reverse-mode recursions of the FK/Jacobian and RNEA kernels, in batched torch.

The reference has no hand-written backward: it differentiates its per-link op graph with autograd.
The CUDA engine implements analytic adjoint kernels (csrc/backward.cu) instead; this file states the
exact recursions those kernels evaluate (re-derived), in the same order and
with the same intermediate quantities, so that

  * tests/test_oracle.py can check the MATHS against torch.autograd of oracle/drm_oracle.py (fp64), and
  * the kernels can be checked against both.

Conventions (link i, parent P, dof column k):  M = F Q(theta), theta = sign * q_k, E = M^T, r = trans,
s = signed axis,  wJ = s qd_k.  Table row layout as in include/drm_b200.h:
[F(9) r(3) Io(9) mc(3) m d pad pad].  All functions return (input grads..., table_grad [N,28]).
"""
import torch


def _skew_cross(a, b):
    return torch.cross(a, b, dim=-1)


def _elem(a, c, s):
    """Q(theta) about coordinate a and dQ/dtheta, batched [B,3,3]."""
    B = c.shape[0]
    Q = torch.zeros(B, 3, 3, dtype=c.dtype)
    dQ = torch.zeros(B, 3, 3, dtype=c.dtype)
    u, v = (a + 1) % 3, (a + 2) % 3
    Q[:, a, a] = 1
    Q[:, u, u] = c; Q[:, v, v] = c; Q[:, v, u] = s; Q[:, u, v] = -s
    dQ[:, u, u] = -s; dQ[:, v, v] = -s; dQ[:, v, u] = c; dQ[:, u, v] = -c
    return Q, dQ


def _joint(table, i, axis_code, q_col):
    """(M, F, Q, dQ, sign) for link i; fixed joints: Q = I."""
    F = table[i, 0:9].reshape(1, 3, 3)
    if axis_code == 0:
        B = q_col.shape[0]
        eye = torch.eye(3, dtype=table.dtype).expand(B, 3, 3)
        return F.expand(B, 3, 3), F, eye, None, 0.0
    sign = 1.0 if axis_code > 0 else -1.0
    a = abs(axis_code) - 1
    th = sign * q_col
    Q, dQ = _elem(a, torch.cos(th), torch.sin(th))
    return F @ Q, F, Q, dQ, sign


def _axis_vec(axis_code, dtype):
    s = torch.zeros(3, dtype=dtype)
    if axis_code != 0:
        s[abs(axis_code) - 1] = 1.0 if axis_code > 0 else -1.0
    return s


def fk_jacobian_backward(table, parent, axis, dof, ee, q, g_pos, g_quat, g_jl, g_ja):
    """Adjoint of (pos, quat, J_lin, J_ang) of link `ee`.  Any upstream gradient may be None.
    Returns (q_grad [B,n], table_grad [N,28])."""
    B, n = q.shape
    N = table.shape[0]
    dt = q.dtype
    path = []
    l = ee
    while l > 0:
        path.append(l)
        l = parent[l]
    path.reverse()
    # ---- forward recompute along the path, keeping R_i, p_i --------------------------------------
    R = [torch.eye(3, dtype=dt).expand(B, 3, 3)]
    p = [torch.zeros(B, 3, dtype=dt)]
    joints = []
    for i in path:
        M, F, Q, dQ, sign = _joint(table, i, axis[i], q[:, dof[i]] if axis[i] != 0 else q[:, 0])
        r = table[i, 9:12]
        p.append((R[-1] @ r) + p[-1])
        R.append(R[-1] @ M)
        joints.append((M, F, Q, dQ, sign, r))
    p_ee, R_ee = p[-1], R[-1]
    zero3 = torch.zeros(B, 3, dtype=dt)
    gl = (lambda k: g_jl[:, :, k]) if g_jl is not None else (lambda k: zero3)
    ga = (lambda k: g_ja[:, :, k]) if g_ja is not None else (lambda k: zero3)
    # ---- seeds -----------------------------------------------------------------------------------
    pbar = g_pos.clone() if g_pos is not None else zero3.clone()
    for idx, i in enumerate(path):
        if axis[i] != 0:
            z = R[idx + 1] @ _axis_vec(axis[i], dt)
            pbar = pbar + _skew_cross(gl(dof[i]), z)            # p_ee adjoint: J_lin = z x (p_ee - p_i)
    Rbar = torch.zeros(B, 3, 3, dtype=dt)
    if g_quat is not None:
        Rbar = Rbar + quaternion_backward(R_ee, g_quat)
    q_grad = torch.zeros(B, n, dtype=dt)
    table_grad = torch.zeros(N, 28, dtype=dt)
    # ---- reverse sweep ee -> root ----------------------------------------------------------------
    for idx in range(len(path) - 1, -1, -1):
        i = path[idx]
        Ri, pi, RP = R[idx + 1], p[idx + 1], R[idx]
        M, F, Q, dQ, sign, r = joints[idx]
        if axis[i] != 0:
            s = _axis_vec(axis[i], dt)
            z = Ri @ s
            d = p_ee - pi
            k = dof[i]
            zbar = _skew_cross(d, gl(k)) + ga(k)
            Rbar = Rbar + zbar[:, :, None] * s[None, None, :]
            pbar = pbar - _skew_cross(gl(k), z)
        Mbar = RP.transpose(1, 2) @ Rbar
        rbar = (RP.transpose(1, 2) @ pbar[:, :, None])[:, :, 0]
        Rbar = Rbar @ M.transpose(1, 2) + pbar[:, :, None] * r[None, None, :]
        table_grad[i, 9:12] += rbar.sum(0)
        table_grad[i, 0:9] += (Mbar @ Q.transpose(1, 2)).sum(0).reshape(9)
        if axis[i] != 0:
            q_grad[:, dof[i]] = sign * ((F.transpose(1, 2) @ Mbar) * dQ).sum((1, 2))
    return q_grad, table_grad


def quaternion_backward(R, g):
    """dL/dR for the xyzw quaternion of spatial_vector_algebra.py:116-135 given dL/dquat = g.
    Mathematically exact (the reference's autograd detaches the 0.5/sqrt(t) factor -- quirk 5)."""
    B = R.shape[0]
    out = torch.zeros_like(R)
    d0, d1, d2 = R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]
    tr = d0 + d1 + d2
    for b in range(B):
        Rb = R[b].detach().clone().requires_grad_(True)
        if tr[b] + 1 > 1:
            t = Rb[0, 0] + Rb[1, 1] + Rb[2, 2] + 1
            u = torch.stack([Rb[2, 1] - Rb[1, 2], Rb[0, 2] - Rb[2, 0], Rb[1, 0] - Rb[0, 1], t])
        elif d2[b] > max(d0[b], d1[b]):
            t = Rb[2, 2] - (Rb[0, 0] + Rb[1, 1]) + 1
            u = torch.stack([Rb[2, 0] + Rb[0, 2], Rb[1, 2] + Rb[2, 1], t, Rb[1, 0] - Rb[0, 1]])
        elif d1[b] > d0[b]:
            t = Rb[1, 1] - (Rb[2, 2] + Rb[0, 0]) + 1
            u = torch.stack([Rb[0, 1] + Rb[1, 0], t, Rb[1, 2] + Rb[2, 1], Rb[0, 2] - Rb[2, 0]])
        else:
            t = Rb[0, 0] - (Rb[1, 1] + Rb[2, 2]) + 1
            u = torch.stack([t, Rb[0, 1] + Rb[1, 0], Rb[2, 0] + Rb[0, 2], Rb[2, 1] - Rb[1, 2]])
        quat = u * 0.5 / torch.sqrt(t)
        (gr,) = torch.autograd.grad((quat * g[b]).sum(), Rb)
        out[b] = gr
    return out


def inverse_dynamics_backward(table, parent, axis, dof, q, qd, qdd, g_tau, gravity=True, damping=True):
    """Adjoint of RNEA.  Returns (q_grad, qd_grad, qdd_grad, table_grad)."""
    B, n = q.shape
    N = table.shape[0]
    dt = q.dtype
    z3 = torch.zeros(B, 3, dtype=dt)
    cr = _skew_cross

    # ---- forward recompute: motion state and accumulated wrenches ---------------------------------
    w, v, al, a = [z3] * N, [z3] * N, [z3] * N, [z3] * N
    a[0] = torch.tensor([0.0, 0.0, 9.81 if gravity else 0.0], dtype=dt).expand(B, 3)
    J = [None] * N
    f, nn = [z3] * N, [z3] * N
    Hl, Ha = [None] * N, [None] * N
    for i in range(1, N):
        P = parent[i]
        M, F, Q, dQ, sign = _joint(table, i, axis[i], q[:, dof[i]] if axis[i] != 0 else q[:, 0])
        r = table[i, 9:12]
        s = _axis_vec(axis[i], dt)
        qd_k = qd[:, dof[i]] if axis[i] != 0 else torch.zeros(B, dtype=dt)
        qdd_k = qdd[:, dof[i]] if axis[i] != 0 else torch.zeros(B, dtype=dt)
        wJ = qd_k[:, None] * s
        E = M.transpose(1, 2)
        w[i] = (E @ w[P][:, :, None])[:, :, 0] + wJ
        v[i] = (E @ (v[P] + cr(w[P], r.expand(B, 3)))[:, :, None])[:, :, 0]
        al[i] = (E @ al[P][:, :, None])[:, :, 0] + qdd_k[:, None] * s + cr(w[i], wJ)
        a[i] = (E @ (a[P] + cr(al[P], r.expand(B, 3)))[:, :, None])[:, :, 0] + cr(v[i], wJ)
        J[i] = (M, F, Q, dQ, sign, r, s, wJ)
        Io, mc, m = table[i, 12:21].reshape(3, 3), table[i, 21:24], table[i, 24]
        Hl[i] = m * v[i] - cr(mc.expand(B, 3), w[i])
        Ha[i] = w[i] @ Io.T + cr(mc.expand(B, 3), v[i])
        Bl = m * a[i] - cr(mc.expand(B, 3), al[i])
        Ba = al[i] @ Io.T + cr(mc.expand(B, 3), a[i])
        f[i] = Bl + cr(w[i], Hl[i])
        nn[i] = Ba + cr(w[i], Ha[i]) + cr(v[i], Hl[i])
    for i in range(N - 1, 0, -1):
        P = parent[i]
        if P > 0:
            M, r = J[i][0], J[i][5]
            Mf = (M @ f[i][:, :, None])[:, :, 0]
            f[P] = f[P] + Mf
            nn[P] = nn[P] + cr(r.expand(B, 3), Mf) + (M @ nn[i][:, :, None])[:, :, 0]

    q_grad, qd_grad, qdd_grad = (torch.zeros(B, n, dtype=dt) for _ in range(3))
    tg = torch.zeros(N, 28, dtype=dt)
    Mbar = [torch.zeros(B, 3, 3, dtype=dt) for _ in range(N)]
    rbar = [torch.zeros(B, 3, dtype=dt) for _ in range(N)]

    # ---- pass 1, root -> leaves: wrench adjoints lam = n-bar, mu = f-bar ---------------------------
    lam, mu = [z3] * N, [z3] * N
    for i in range(1, N):
        P = parent[i]
        M, F, Q, dQ, sign, r, s, wJ = J[i]
        E = M.transpose(1, 2)
        u = mu[P] + cr(lam[P], r.expand(B, 3))
        g_k = g_tau[:, dof[i]] if axis[i] != 0 else torch.zeros(B, dtype=dt)
        lam[i] = (E @ lam[P][:, :, None])[:, :, 0] + g_k[:, None] * s
        mu[i] = (E @ u[:, :, None])[:, :, 0]
        Mbar[i] = Mbar[i] + lam[P][:, :, None] * nn[i][:, None, :] + u[:, :, None] * f[i][:, None, :]
        Mf = (M @ f[i][:, :, None])[:, :, 0]
        rbar[i] = rbar[i] + cr(Mf, lam[P])
        if axis[i] != 0 and damping:
            qd_grad[:, dof[i]] += table[i, 25] * g_k
            tg[i, 25] += (g_k * qd[:, dof[i]]).sum()

    # ---- pass 2, leaves -> root: motion adjoints ---------------------------------------------------
    wb = [torch.zeros(B, 3, dtype=dt) for _ in range(N)]
    vb = [torch.zeros(B, 3, dtype=dt) for _ in range(N)]
    alb = [torch.zeros(B, 3, dtype=dt) for _ in range(N)]
    ab = [torch.zeros(B, 3, dtype=dt) for _ in range(N)]
    for i in range(N - 1, 0, -1):
        P = parent[i]
        M, F, Q, dQ, sign, r, s, wJ = J[i]
        Io, mc, m = table[i, 12:21].reshape(3, 3), table[i, 21:24].expand(B, 3), table[i, 24]
        L, U = lam[i], mu[i]
        # body part
        Hlb = cr(U, w[i]) + cr(L, v[i])
        Hab = cr(L, w[i])
        alb[i] = alb[i] + cr(mc, U) + L @ Io
        ab[i] = ab[i] + m * U + cr(L, mc)
        wb[i] = wb[i] + cr(Hl[i], U) + cr(Ha[i], L) + cr(mc, Hlb) + Hab @ Io
        vb[i] = vb[i] + cr(Hl[i], L) + m * Hlb + cr(Hab, mc)
        tg[i, 24] += ((U * a[i]).sum(1) + (Hlb * v[i]).sum(1)).sum()
        tg[i, 21:24] += (cr(U, al[i]) + cr(a[i], L) + cr(Hlb, w[i]) + cr(v[i], Hab)).sum(0)
        tg[i, 12:21] += (L[:, :, None] * al[i][:, None, :] + Hab[:, :, None] * w[i][:, None, :]).sum(0).reshape(9)
        # kinematic part, in the order a, alpha, v, omega
        rB = r.expand(B, 3)
        wJb = cr(ab[i], v[i])
        vb[i] = vb[i] + cr(wJ, ab[i])
        ua = (M @ ab[i][:, :, None])[:, :, 0]
        ab[P] = ab[P] + ua
        alb[P] = alb[P] + cr(rB, ua)
        rbar[i] = rbar[i] + cr(ua, al[P])
        Mbar[i] = Mbar[i] + (a[P] + cr(al[P], rB))[:, :, None] * ab[i][:, None, :]

        wb[i] = wb[i] + cr(wJ, alb[i])
        wJb = wJb + cr(alb[i], w[i])
        alb[P] = alb[P] + (M @ alb[i][:, :, None])[:, :, 0]
        Mbar[i] = Mbar[i] + al[P][:, :, None] * alb[i][:, None, :]
        if axis[i] != 0:
            qdd_grad[:, dof[i]] = (alb[i] * s).sum(1)

        uv = (M @ vb[i][:, :, None])[:, :, 0]
        vb[P] = vb[P] + uv
        wb[P] = wb[P] + cr(rB, uv)
        rbar[i] = rbar[i] + cr(uv, w[P])
        Mbar[i] = Mbar[i] + (v[P] + cr(w[P], rB))[:, :, None] * vb[i][:, None, :]

        wb[P] = wb[P] + (M @ wb[i][:, :, None])[:, :, 0]
        Mbar[i] = Mbar[i] + w[P][:, :, None] * wb[i][:, None, :]
        wJb = wJb + wb[i]
        if axis[i] != 0:
            qd_grad[:, dof[i]] += (wJb * s).sum(1)
            q_grad[:, dof[i]] = sign * ((F.transpose(1, 2) @ Mbar[i]) * dQ).sum((1, 2))
        tg[i, 0:9] += (Mbar[i] @ Q.transpose(1, 2)).sum(0).reshape(9)
        tg[i, 9:12] += rbar[i].sum(0)
    return q_grad, qd_grad, qdd_grad, tg
