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


def inverse_dynamics_backward_chain(table, dof, q, qd, qdd, g_tau, gravity=True, damping=True):
    """The same adjoint, restated the way rnea_backward_chain_kernel (csrc/backward_rnea.cu) evaluates it: TWO sweeps over a serial chain whose
    rows are canonical (every movable axis is +z; parent of link i is link i - 1), nothing per link kept except (cos, sin).

      sweep 1, root -> leaves: motion state and wrench adjoints lam, mu -- only to arrive at the last link's values;
      sweep 2, leaves -> root: the state AND the wrench adjoints of link i - 1 re-derived from link i's (both recursions
        are invertible), the body wrench recomputed from the state and accumulated on the way (f, n), the motion
        adjoints, and every gradient.
    M = F Rz(theta) is never formed: x -> Rz^T (F^T x) and x -> F (Rz x); with y = M^T x and adjoint ybar the joint-angle
    gradient of that product is (ybar x y).z and its F-gradient x (Rz ybar)^T.
    Returns (q_grad, qd_grad, qdd_grad, table_grad) like inverse_dynamics_backward (axis codes: 3 movable, 0 fixed)."""
    B, n = q.shape
    N = table.shape[0]
    dt = q.dtype
    cr = _skew_cross
    zero = torch.zeros(B, 3, dtype=dt)
    a_root = torch.tensor([0.0, 0.0, 9.81 if gravity else 0.0], dtype=dt).expand(B, 3)

    def rotz(x, c, s):        # Rz x
        return torch.stack([c * x[:, 0] - s * x[:, 1], c * x[:, 1] + s * x[:, 0], x[:, 2]], 1)

    def rotzT(x, c, s):       # Rz^T x
        return torch.stack([c * x[:, 0] + s * x[:, 1], c * x[:, 1] - s * x[:, 0], x[:, 2]], 1)

    def cz(a, b):             # (a x b).z
        return a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]

    def col(x, k, i):
        return x[:, k] if k >= 0 else torch.zeros(B, dtype=dt)

    ez = torch.tensor([0.0, 0.0, 1.0], dtype=dt)
    # ---- sweep 1 ----------------------------------------------------------------------------------
    w, v, al, a, lam, mu = zero, zero, zero, a_root, zero, zero
    keep = [None] * N
    for i in range(1, N):
        F, r = table[i, 0:9].reshape(3, 3), table[i, 9:12].expand(B, 3)
        k = dof[i]
        cs = torch.cos(q[:, k]) if k >= 0 else torch.ones(B, dtype=dt)
        sn = torch.sin(q[:, k]) if k >= 0 else torch.zeros(B, dtype=dt)
        qd_k, qdd_k, g_k = col(qd, k, i), col(qdd, k, i), col(g_tau, k, i)
        wn = rotzT(w @ F, cs, sn)                         # F^T x == x @ F for row vectors
        vn = rotzT((cr(w, r) + v) @ F, cs, sn)
        aln = rotzT(al @ F, cs, sn)
        an = rotzT((cr(al, r) + a) @ F, cs, sn)
        wn = wn + qd_k[:, None] * ez
        aln = aln + cr(wn, qd_k[:, None] * ez) + qdd_k[:, None] * ez
        an = an + cr(vn, qd_k[:, None] * ez)
        u = cr(lam, r) + mu
        lam = rotzT(lam @ F, cs, sn) + g_k[:, None] * ez
        mu = rotzT(u @ F, cs, sn)
        w, v, al, a = wn, vn, aln, an
        keep[i] = (cs, sn)

    # ---- sweep 2 ----------------------------------------------------------------------------------
    q_grad, qd_grad, qdd_grad = (torch.zeros(B, n, dtype=dt) for _ in range(3))
    tg = torch.zeros(N, 28, dtype=dt)
    c_wb, c_vb, c_alb, c_ab, carry_f, carry_n = zero, zero, zero, zero, zero, zero
    for i in range(N - 1, 0, -1):
        F, r = table[i, 0:9].reshape(3, 3), table[i, 9:12].expand(B, 3)
        Io, mc, m, d = table[i, 12:21].reshape(3, 3), table[i, 21:24].expand(B, 3), table[i, 24], table[i, 25]
        L, U = lam, mu
        cs, sn = keep[i]
        k = dof[i]
        qd_k, qdd_k, g_k = col(qd, k, i), col(qdd, k, i), col(g_tau, k, i)
        wJ = qd_k[:, None] * ez
        # (a, b) the parent's state through the inverted recursion
        tw = w - wJ
        tal = al - cr(w, wJ) - qdd_k[:, None] * ez
        apre = a - cr(v, wJ)
        if i > 1:
            wp, alp = rotz(tw, cs, sn) @ F.T, rotz(tal, cs, sn) @ F.T          # F x == x @ F^T
            Xv, Xa = rotz(v, cs, sn) @ F.T, rotz(apre, cs, sn) @ F.T          # = wp x r + vp, alp x r + ap
            vp, ap = Xv - cr(wp, r), Xa - cr(alp, r)
            LP = rotz(L - g_k[:, None] * ez, cs, sn) @ F.T                    # lam_i = M^T lam_p + g_k e_z
            u = rotz(U, cs, sn) @ F.T                                         # mu_i = M^T (lam_p x r + mu_p)
            UP = u - cr(LP, r)
        else:
            wp, alp, Xv, Xa, vp, ap = zero, zero, zero, a_root, zero, a_root
            LP, u, UP = zero, zero, zero
        # (c) body wrench from the state, plus what the child handed up
        Hl = m * v - cr(mc, w)
        Ha = w @ Io.T + cr(mc, v)
        Bl = m * a - cr(mc, al)
        Ba = al @ Io.T + cr(mc, a)
        f = Bl + cr(w, Hl) + carry_f
        nn = Ba + cr(w, Ha) + cr(v, Hl) + carry_n
        # (d) to the parent
        Rf, Rn = rotz(f, cs, sn), rotz(nn, cs, sn)
        fp = Rf @ F.T
        carry_f, carry_n = fp, cr(r, fp) + Rn @ F.T
        # (e) wrench-adjoint part
        th = cz(nn, L) + cz(f, U)
        Fbar = LP[:, :, None] * Rn[:, None, :] + u[:, :, None] * Rf[:, None, :]
        rbar = cr(fp, LP)
        qdv = torch.zeros(B, dtype=dt)
        if k >= 0 and damping:
            qdv = d * g_k
            tg[i, 25] += (g_k * qd_k).sum()
        # (f) body part of the motion adjoints
        wb, vb, alb, ab = c_wb, c_vb, c_alb, c_ab
        Hlb = cr(U, w) + cr(L, v)
        Hab = cr(L, w)
        alb = alb + cr(mc, U) + L @ Io
        ab = ab + m * U + cr(L, mc)
        wb = wb + cr(Hl, U) + cr(Ha, L) + cr(mc, Hlb) + Hab @ Io
        vb = vb + cr(Hl, L) + m * Hlb + cr(Hab, mc)
        tg[i, 24] += ((U * a).sum(1) + (Hlb * v).sum(1)).sum()
        tg[i, 21:24] += (cr(U, al) + cr(a, L) + cr(Hlb, w) + cr(v, Hab)).sum(0)
        tg[i, 12:21] += (L[:, :, None] * al[:, None, :] + Hab[:, :, None] * w[:, None, :]).sum(0).reshape(9)
        # (g) kinematic part
        vb = vb + cr(wJ, ab)
        wb = wb + cr(wJ, alb)
        wJb = cz(ab, v) + cz(alb, w) + wb[:, 2]
        th = th + cz(ab, apre) + cz(alb, tal) + cz(vb, v) + cz(wb, w)
        Rwb, Ralb, Rvb, Rab = rotz(wb, cs, sn), rotz(alb, cs, sn), rotz(vb, cs, sn), rotz(ab, cs, sn)
        uw, ual, uv, ua = Rwb @ F.T, Ralb @ F.T, Rvb @ F.T, Rab @ F.T
        c_vb, c_ab = uv, ua
        c_wb, c_alb = uw + cr(r, uv), ual + cr(r, ua)
        Fbar = Fbar + Xv[:, :, None] * Rvb[:, None, :] + Xa[:, :, None] * Rab[:, None, :] \
            + wp[:, :, None] * Rwb[:, None, :] + alp[:, :, None] * Ralb[:, None, :]
        rbar = rbar + cr(ua, alp) + cr(uv, wp)
        if k >= 0:
            q_grad[:, k] = th
            qd_grad[:, k] = wJb + qdv
            qdd_grad[:, k] = alb[:, 2]
        tg[i, 0:9] += Fbar.sum(0).reshape(9)
        tg[i, 9:12] += rbar.sum(0)
        w, v, al, a, lam, mu = wp, vp, alp, ap, LP, UP
    return q_grad, qd_grad, qdd_grad, tg


def inverse_dynamics_backward_two_sweep_tree(table, parent, dof, q, qd, qdd, g_tau, gravity=True, damping=True):
    """The two-sweep form for TREES (canonical rows, links in document order, parents before children) -- the executable
    statement of DESIGN.md section 10, item 2; no kernel evaluates it yet.  Differences from the chain:
      * sweep 1 reads the parent's (state, lam, mu) from wherever the forward RNEA kernel reads the parent's state
        (registers when the parent is the link before, a branch-point slot otherwise) and STORES them for every chain end
        (link i whose successor i + 1 is not its child): sweep 2 cannot re-derive those from a child;
      * sweep 2 walks the links backwards; link i's (state, lam, mu) come from its child i + 1 through the inverted
        recursions or from the chain-end store; wrenches and motion adjoints for a parent that is not the link before are
        accumulated in that parent's slot."""
    B, n = q.shape
    N = table.shape[0]
    dt = q.dtype
    cr = _skew_cross
    zero = torch.zeros(B, 3, dtype=dt)
    a_root = torch.tensor([0.0, 0.0, 9.81 if gravity else 0.0], dtype=dt).expand(B, 3)
    ez = torch.tensor([0.0, 0.0, 1.0], dtype=dt)

    def rotz(x, c, s):
        return torch.stack([c * x[:, 0] - s * x[:, 1], c * x[:, 1] + s * x[:, 0], x[:, 2]], 1)

    def rotzT(x, c, s):
        return torch.stack([c * x[:, 0] + s * x[:, 1], c * x[:, 1] - s * x[:, 0], x[:, 2]], 1)

    def cz(a, b):
        return a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]

    def col(x, k):
        return x[:, k] if k >= 0 else torch.zeros(B, dtype=dt)

    is_tip = [i == N - 1 or parent[i + 1] != i for i in range(N)]
    # ---- sweep 1 ----------------------------------------------------------------------------------
    st = {0: (zero, zero, zero, a_root, zero, zero)}       # what a kernel holds in registers / branch slots
    trig, tips = [None] * N, {}
    for i in range(1, N):
        w, v, al, a, lam, mu = st[parent[i]]
        F, r = table[i, 0:9].reshape(3, 3), table[i, 9:12].expand(B, 3)
        k = dof[i]
        cs = torch.cos(q[:, k]) if k >= 0 else torch.ones(B, dtype=dt)
        sn = torch.sin(q[:, k]) if k >= 0 else torch.zeros(B, dtype=dt)
        qd_k, qdd_k, g_k = col(qd, k), col(qdd, k), col(g_tau, k)
        wn = rotzT(w @ F, cs, sn) + qd_k[:, None] * ez
        vn = rotzT((cr(w, r) + v) @ F, cs, sn)
        aln = rotzT(al @ F, cs, sn) + cr(wn, qd_k[:, None] * ez) + qdd_k[:, None] * ez
        an = rotzT((cr(al, r) + a) @ F, cs, sn) + cr(vn, qd_k[:, None] * ez)
        lamn = rotzT(lam @ F, cs, sn) + g_k[:, None] * ez
        mun = rotzT((cr(lam, r) + mu) @ F, cs, sn)
        st[i] = (wn, vn, aln, an, lamn, mun)
        trig[i] = (cs, sn)
        if is_tip[i]:
            tips[i] = st[i]
    del st                                                   # sweep 2 may only use `tips` and what children hand down

    # ---- sweep 2 ----------------------------------------------------------------------------------
    q_grad, qd_grad, qdd_grad = (torch.zeros(B, n, dtype=dt) for _ in range(3))
    tg = torch.zeros(N, 28, dtype=dt)
    acc = [[zero] * 6 for _ in range(N)]                     # per link: wb, vb, alb, ab, f, n handed down by its children
    handed = None                                            # (state, lam, mu) of link i from its child i + 1
    for i in range(N - 1, 0, -1):
        P = parent[i]
        F, r = table[i, 0:9].reshape(3, 3), table[i, 9:12].expand(B, 3)
        Io, mc, m, d = table[i, 12:21].reshape(3, 3), table[i, 21:24].expand(B, 3), table[i, 24], table[i, 25]
        w, v, al, a, L, U = tips[i] if is_tip[i] else handed
        cs, sn = trig[i]
        k = dof[i]
        qd_k, qdd_k, g_k = col(qd, k), col(qdd, k), col(g_tau, k)
        wJ = qd_k[:, None] * ez
        tw = w - wJ
        tal = al - cr(w, wJ) - qdd_k[:, None] * ez
        apre = a - cr(v, wJ)
        if P > 0:
            wp, alp = rotz(tw, cs, sn) @ F.T, rotz(tal, cs, sn) @ F.T
            Xv, Xa = rotz(v, cs, sn) @ F.T, rotz(apre, cs, sn) @ F.T
            vp, ap = Xv - cr(wp, r), Xa - cr(alp, r)
            LP = rotz(L - g_k[:, None] * ez, cs, sn) @ F.T
            u = rotz(U, cs, sn) @ F.T
            UP = u - cr(LP, r)
        else:
            wp, alp, Xv, Xa, vp, ap, LP, u, UP = zero, zero, zero, a_root, zero, a_root, zero, zero, zero
        handed = (wp, vp, alp, ap, LP, UP)                  # used by iteration i - 1 iff parent(i) == i - 1
        c_wb, c_vb, c_alb, c_ab, carry_f, carry_n = acc[i]
        Hl = m * v - cr(mc, w)
        Ha = w @ Io.T + cr(mc, v)
        f = m * a - cr(mc, al) + cr(w, Hl) + carry_f
        nn = al @ Io.T + cr(mc, a) + cr(w, Ha) + cr(v, Hl) + carry_n
        Rf, Rn = rotz(f, cs, sn), rotz(nn, cs, sn)
        fp = Rf @ F.T
        npar = cr(r, fp) + Rn @ F.T
        th = cz(nn, L) + cz(f, U)
        Fbar = LP[:, :, None] * Rn[:, None, :] + u[:, :, None] * Rf[:, None, :]
        rbar = cr(fp, LP)
        qdv = torch.zeros(B, dtype=dt)
        if k >= 0 and damping:
            qdv = d * g_k
            tg[i, 25] += (g_k * qd_k).sum()
        Hlb = cr(U, w) + cr(L, v)
        Hab = cr(L, w)
        alb = c_alb + cr(mc, U) + L @ Io
        ab = c_ab + m * U + cr(L, mc)
        wb = c_wb + cr(Hl, U) + cr(Ha, L) + cr(mc, Hlb) + Hab @ Io
        vb = c_vb + cr(Hl, L) + m * Hlb + cr(Hab, mc)
        tg[i, 24] += ((U * a).sum(1) + (Hlb * v).sum(1)).sum()
        tg[i, 21:24] += (cr(U, al) + cr(a, L) + cr(Hlb, w) + cr(v, Hab)).sum(0)
        tg[i, 12:21] += (L[:, :, None] * al[:, None, :] + Hab[:, :, None] * w[:, None, :]).sum(0).reshape(9)
        vb = vb + cr(wJ, ab)
        wb = wb + cr(wJ, alb)
        wJb = cz(ab, v) + cz(alb, w) + wb[:, 2]
        th = th + cz(ab, apre) + cz(alb, tal) + cz(vb, v) + cz(wb, w)
        Rwb, Ralb, Rvb, Rab = rotz(wb, cs, sn), rotz(alb, cs, sn), rotz(vb, cs, sn), rotz(ab, cs, sn)
        uw, ual, uv, ua = Rwb @ F.T, Ralb @ F.T, Rvb @ F.T, Rab @ F.T
        Fbar = Fbar + Xv[:, :, None] * Rvb[:, None, :] + Xa[:, :, None] * Rab[:, None, :] \
            + wp[:, :, None] * Rwb[:, None, :] + alp[:, :, None] * Ralb[:, None, :]
        rbar = rbar + cr(ua, alp) + cr(uv, wp)
        if P > 0:                                            # hand down to the parent (registers or its slot)
            pw, pv, pal, pa, pf, pn = acc[P]
            acc[P] = [pw + uw + cr(r, uv), pv + uv, pal + ual + cr(r, ua), pa + ua, pf + fp, pn + npar]
        if k >= 0:
            q_grad[:, k] = th
            qd_grad[:, k] = wJb + qdv
            qdd_grad[:, k] = alb[:, 2]
        tg[i, 0:9] += Fbar.sum(0).reshape(9)
        tg[i, 9:12] += rbar.sum(0)
    return q_grad, qd_grad, qdd_grad, tg


# ------------------------------------------------------------------------------------------------
# articulated-body forward dynamics (robot_model.py:488-624) and its adjoint
# ------------------------------------------------------------------------------------------------
def _skew_m(r):
    """[..., 3] -> [..., 3, 3] with skew(r) b = r x b."""
    z = torch.zeros_like(r[..., 0])
    return torch.stack([torch.stack([z, -r[..., 2], r[..., 1]], -1),
                        torch.stack([r[..., 2], z, -r[..., 0]], -1),
                        torch.stack([-r[..., 1], r[..., 0], z], -1)], -2)


def _unskew(Sb):
    """Adjoint of _skew_m: [..., 3, 3] -> [..., 3]."""
    return torch.stack([Sb[..., 2, 1] - Sb[..., 1, 2], Sb[..., 0, 2] - Sb[..., 2, 0], Sb[..., 1, 0] - Sb[..., 0, 1]], -1)


def _mv(M, x):
    return (M @ x[:, :, None])[:, :, 0]


def _outer(x, y):
    return x[:, :, None] * y[:, None, :]


ABA_EPS = 1e-37


def forward_dynamics_with_backward(table, parent, axis, dof, q, qd, f, g_qdd, gravity=True, damping=False):
    """The reference's articulated-body algorithm on the link table (blocks [[A, B], [C, D]] of the 6x6 articulated
    inertia acting on [ang; lin]) followed by its hand-derived adjoint, exactly the recursions csrc/aba.cu and
    csrc/backward_aba.cu evaluate.  Returns (qdd, q_grad, qd_grad, f_grad, table_grad)."""
    B, n = q.shape
    N = table.shape[0]
    dt = q.dtype
    z3 = torch.zeros(B, 3, dtype=dt)
    cr = _skew_cross
    T = lambda X: X.transpose(1, 2)  # noqa: E731

    # ================================ forward ======================================================
    J = [None] * N
    w, v = [z3] * N, [z3] * N
    ca, cl = [z3] * N, [z3] * N
    hl, ha = [None] * N, [None] * N
    A, Bm, C, D = [None] * N, [None] * N, [None] * N, [None] * N
    p_ang, p_lin = [None] * N, [None] * N
    for i in range(1, N):
        P = parent[i]
        M, F, Q, dQ, sign = _joint(table, i, axis[i], q[:, dof[i]] if axis[i] != 0 else q[:, 0])
        r = table[i, 9:12].expand(B, 3)
        s = _axis_vec(axis[i], dt)
        qd_k = qd[:, dof[i]] if axis[i] != 0 else torch.zeros(B, dtype=dt)
        wJ = qd_k[:, None] * s
        E = T(M)
        w[i] = _mv(E, w[P]) + wJ
        v[i] = _mv(E, v[P] + cr(w[P], r))
        ca[i], cl[i] = cr(w[i], wJ), cr(v[i], wJ)
        Io, mc, m = table[i, 12:21].reshape(3, 3), table[i, 21:24].expand(B, 3), table[i, 24]
        hl[i] = m * v[i] - cr(mc, w[i])
        ha[i] = w[i] @ Io.T + cr(mc, v[i])
        p_ang[i] = cr(w[i], ha[i]) + cr(v[i], hl[i])
        p_lin[i] = cr(w[i], hl[i])
        Smc = _skew_m(table[i, 21:24])
        A[i] = Io.expand(B, 3, 3).clone()
        Bm[i] = Smc.expand(B, 3, 3).clone()
        C[i] = Smc.T.expand(B, 3, 3).clone()
        D[i] = (m * torch.eye(3, dtype=dt)).expand(B, 3, 3).clone()
        J[i] = (M, F, Q, dQ, sign, r, s, wJ)
    Ua, Ul, d, u, inv = [z3] * N, [z3] * N, [None] * N, [None] * N, [None] * N
    Ap, Bp, Cp, Dp = [None] * N, [None] * N, [None] * N, [None] * N      # IA' blocks
    pa_ang, pa_lin = [None] * N, [None] * N
    for i in range(N - 1, 0, -1):
        P = parent[i]
        M, F, Q, dQ, sign, r, s, wJ = J[i]
        mov = axis[i] != 0
        if mov:
            Ua[i], Ul[i] = A[i] @ s, C[i] @ s
            d[i] = Ua[i] @ s
            fk = f[:, dof[i]] - (table[i, 25] * qd[:, dof[i]] if damping else 0.0)
            u[i] = fk - p_ang[i] @ s
        if P > 0:
            if mov:
                inv[i] = 1.0 / (d[i] + ABA_EPS)
                k = inv[i][:, None, None]
                Ap[i] = A[i] - _outer(Ua[i], Ua[i]) * k
                Bp[i] = Bm[i] - _outer(Ua[i], Ul[i]) * k
                Cp[i] = C[i] - _outer(Ul[i], Ua[i]) * k
                Dp[i] = D[i] - _outer(Ul[i], Ul[i]) * k
                ud = (u[i] * inv[i])[:, None]
                pa_ang[i] = p_ang[i] + _mv(Ap[i], ca[i]) + _mv(Bp[i], cl[i]) + Ua[i] * ud
                pa_lin[i] = p_lin[i] + _mv(Cp[i], ca[i]) + _mv(Dp[i], cl[i]) + Ul[i] * ud
            else:
                Ap[i], Bp[i], Cp[i], Dp[i] = A[i], Bm[i], C[i], D[i]
                pa_ang[i], pa_lin[i] = p_ang[i], p_lin[i]
            S = _skew_m(r)
            Ah, Bh, Ch, Dh = (M @ X @ T(M) for X in (Ap[i], Bp[i], Cp[i], Dp[i]))
            D[P] = D[P] + Dh
            Bm[P] = Bm[P] + Bh + S @ Dh
            C[P] = C[P] + Ch - Dh @ S
            A[P] = A[P] + Ah + S @ Ch - Bh @ S - S @ Dh @ S
            Ql = _mv(M, pa_lin[i])
            p_lin[P] = p_lin[P] + Ql
            p_ang[P] = p_ang[P] + cr(r, Ql) + _mv(M, pa_ang[i])
    al, a = [z3] * N, [z3] * N
    alq, aq = [z3] * N, [z3] * N          # a' (before the joint acceleration is added)
    a[0] = torch.tensor([0.0, 0.0, 9.81 if gravity else 0.0], dtype=dt).expand(B, 3)
    qdd = torch.zeros(B, n, dtype=dt)
    for i in range(1, N):
        P = parent[i]
        M, F, Q, dQ, sign, r, s, wJ = J[i]
        E = T(M)
        alq[i] = _mv(E, al[P]) + ca[i]
        aq[i] = _mv(E, a[P] + cr(al[P], r)) + cl[i]
        al[i], a[i] = alq[i], aq[i]
        if axis[i] != 0:
            qdd_i = (1.0 / d[i]) * (u[i] - ((Ua[i] * alq[i]).sum(1) + (Ul[i] * aq[i]).sum(1)))
            qdd[:, dof[i]] = qdd_i
            al[i] = alq[i] + qdd_i[:, None] * s

    # ================================ adjoint ======================================================
    q_grad, qd_grad, f_grad = (torch.zeros(B, n, dtype=dt) for _ in range(3))
    tg = torch.zeros(N, 28, dtype=dt)
    Mbar = [torch.zeros(B, 3, 3, dtype=dt) for _ in range(N)]
    rbar = [torch.zeros(B, 3, dtype=dt) for _ in range(N)]
    Z = lambda: torch.zeros(B, 3, dtype=dt)  # noqa: E731
    alb, ab = [Z() for _ in range(N)], [Z() for _ in range(N)]
    cab, clb = [Z() for _ in range(N)], [Z() for _ in range(N)]
    Uab, Ulb = [Z() for _ in range(N)], [Z() for _ in range(N)]
    db = [torch.zeros(B, dtype=dt) for _ in range(N)]
    ub = [torch.zeros(B, dtype=dt) for _ in range(N)]

    # ---- reverse of pass 3 (leaves -> root) -----------------------------------------------------------
    for i in range(N - 1, 0, -1):
        P = parent[i]
        M, F, Q, dQ, sign, r, s, wJ = J[i]
        alq_b, aq_b = alb[i], ab[i]
        if axis[i] != 0:
            qb = g_qdd[:, dof[i]] + alb[i] @ s
            k = (qb / d[i])[:, None]
            ub[i] = ub[i] + qb / d[i]
            Uab[i] = Uab[i] - k * alq[i]
            Ulb[i] = Ulb[i] - k * aq[i]
            db[i] = db[i] - qb * qdd[:, dof[i]] / d[i]
            alq_b = alb[i] - k * Ua[i]
            aq_b = ab[i] - k * Ul[i]
        cab[i] = cab[i] + alq_b
        clb[i] = clb[i] + aq_b
        ua = _mv(M, aq_b)
        ab[P] = ab[P] + ua
        alb[P] = alb[P] + cr(r, ua) + _mv(M, alq_b)
        rbar[i] = rbar[i] + cr(ua, al[P])
        Mbar[i] = Mbar[i] + _outer(a[P] + cr(al[P], r), aq_b) + _outer(al[P], alq_b)

    # ---- reverse of pass 2 (root -> leaves) -----------------------------------------------------------
    ZM = lambda: torch.zeros(B, 3, 3, dtype=dt)  # noqa: E731
    Ab, Bb, Cb, Db = [ZM() for _ in range(N)], [ZM() for _ in range(N)], [ZM() for _ in range(N)], [ZM() for _ in range(N)]
    pb_ang, pb_lin = [Z() for _ in range(N)], [Z() for _ in range(N)]
    for i in range(1, N):
        P = parent[i]
        M, F, Q, dQ, sign, r, s, wJ = J[i]
        mov = axis[i] != 0
        if P > 0:
            E = T(M)
            # force transform
            Qa_b, Ql_b = pb_ang[P], pb_lin[P]
            t = Ql_b + cr(Qa_b, r)
            pal_b, paa_b = _mv(E, t), _mv(E, Qa_b)
            Ql = _mv(M, pa_lin[i])
            rbar[i] = rbar[i] + cr(Ql, Qa_b)
            Mbar[i] = Mbar[i] + _outer(t, pa_lin[i]) + _outer(Qa_b, pa_ang[i])
            # Y = T^T (rotated) T
            S = _skew_m(r)
            Ah, Bh, Ch, Dh = (M @ X @ T(M) for X in (Ap[i], Bp[i], Cp[i], Dp[i]))
            YA, YB, YC, YD = Ab[P], Bb[P], Cb[P], Db[P]
            Ah_b = YA
            Bh_b = YB + YA @ S
            Ch_b = YC - S @ YA
            Dh_b = YD - S @ YB + YC @ S - S @ YA @ S
            Sb = YB @ T(Dh) - T(Dh) @ YC + YA @ T(Ch) - T(Bh) @ YA + YA @ S @ T(Dh) + T(Dh) @ S @ YA
            rbar[i] = rbar[i] + _unskew(Sb)
            # rotation blocks
            Ap_b, Bp_b, Cp_b, Dp_b = (E @ X @ M for X in (Ah_b, Bh_b, Ch_b, Dh_b))
            for Xb, X in ((Ah_b, Ap[i]), (Bh_b, Bp[i]), (Ch_b, Cp[i]), (Dh_b, Dp[i])):
                Mbar[i] = Mbar[i] + Xb @ M @ T(X) + T(Xb) @ M @ X
            pb_ang[i] = pb_ang[i] + paa_b
            pb_lin[i] = pb_lin[i] + pal_b
            if mov:
                k = inv[i]
                Ap_b = Ap_b + _outer(paa_b, ca[i]); Bp_b = Bp_b + _outer(paa_b, cl[i])
                Cp_b = Cp_b + _outer(pal_b, ca[i]); Dp_b = Dp_b + _outer(pal_b, cl[i])
                cab[i] = cab[i] + _mv(T(Ap[i]), paa_b) + _mv(T(Cp[i]), pal_b)
                clb[i] = clb[i] + _mv(T(Bp[i]), paa_b) + _mv(T(Dp[i]), pal_b)
                sig = (Ua[i] * paa_b).sum(1) + (Ul[i] * pal_b).sum(1)
                Uab[i] = Uab[i] + paa_b * (u[i] * k)[:, None]
                Ulb[i] = Ulb[i] + pal_b * (u[i] * k)[:, None]
                ub[i] = ub[i] + sig * k
                inv_b = sig * u[i]
                Uab[i] = Uab[i] - k[:, None] * (_mv(Ap_b, Ua[i]) + _mv(T(Ap_b), Ua[i]) + _mv(Bp_b, Ul[i]) + _mv(T(Cp_b), Ul[i]))
                Ulb[i] = Ulb[i] - k[:, None] * (_mv(T(Bp_b), Ua[i]) + _mv(Cp_b, Ua[i]) + _mv(Dp_b, Ul[i]) + _mv(T(Dp_b), Ul[i]))
                inv_b = inv_b - ((Ua[i] * _mv(Ap_b, Ua[i])).sum(1) + (Ua[i] * _mv(Bp_b, Ul[i])).sum(1)
                                 + (Ul[i] * _mv(Cp_b, Ua[i])).sum(1) + (Ul[i] * _mv(Dp_b, Ul[i])).sum(1))
                db[i] = db[i] - inv_b * k * k
            Ab[i] = Ab[i] + Ap_b; Bb[i] = Bb[i] + Bp_b; Cb[i] = Cb[i] + Cp_b; Db[i] = Db[i] + Dp_b
        if mov:
            f_grad[:, dof[i]] = ub[i]
            if damping:
                qd_grad[:, dof[i]] += -table[i, 25] * ub[i]
                tg[i, 25] += -(ub[i] * qd[:, dof[i]]).sum()
            pb_ang[i] = pb_ang[i] - ub[i][:, None] * s
            Uab[i] = Uab[i] + db[i][:, None] * s
            Ab[i] = Ab[i] + _outer(Uab[i], s.expand(B, 3))
            Cb[i] = Cb[i] + _outer(Ulb[i], s.expand(B, 3))
        # rigid-body inertia IA0 = [[Io, mc^], [mc^T, m 1]]
        tg[i, 12:21] += Ab[i].sum(0).reshape(9)
        tg[i, 21:24] += _unskew(Bb[i] + T(Cb[i])).sum(0)
        tg[i, 24] += (Db[i][:, 0, 0] + Db[i][:, 1, 1] + Db[i][:, 2, 2]).sum()

    # ---- reverse of pass 1 (leaves -> root) -----------------------------------------------------------
    wb, vb = [Z() for _ in range(N)], [Z() for _ in range(N)]
    for i in range(N - 1, 0, -1):
        P = parent[i]
        M, F, Q, dQ, sign, r, s, wJ = J[i]
        Io, mc, m = table[i, 12:21].reshape(3, 3), table[i, 21:24].expand(B, 3), table[i, 24]
        pi_, rho = pb_ang[i], pb_lin[i]
        wb[i] = wb[i] + cr(ha[i], pi_) + cr(hl[i], rho)
        vb[i] = vb[i] + cr(hl[i], pi_)
        hab = cr(pi_, w[i])
        hlb = cr(pi_, v[i]) + cr(rho, w[i])
        vb[i] = vb[i] + m * hlb + cr(hab, mc)
        wb[i] = wb[i] + cr(mc, hlb) + hab @ Io
        tg[i, 24] += (hlb * v[i]).sum()
        tg[i, 21:24] += (cr(hlb, w[i]) + cr(v[i], hab)).sum(0)
        tg[i, 12:21] += _outer(hab, w[i]).sum(0).reshape(9)
        wb[i] = wb[i] + cr(wJ, cab[i])
        vb[i] = vb[i] + cr(wJ, clb[i])
        wJb = cr(cab[i], w[i]) + cr(clb[i], v[i]) + wb[i]
        uv = _mv(M, vb[i])
        vb[P] = vb[P] + uv
        wb[P] = wb[P] + cr(r, uv) + _mv(M, wb[i])
        rbar[i] = rbar[i] + cr(uv, w[P])
        Mbar[i] = Mbar[i] + _outer(v[P] + cr(w[P], r), vb[i]) + _outer(w[P], wb[i])
        if axis[i] != 0:
            qd_grad[:, dof[i]] += wJb @ s
            q_grad[:, dof[i]] = sign * ((T(F) @ Mbar[i]) * dQ).sum((1, 2))
        tg[i, 0:9] += (Mbar[i] @ T(Q)).sum(0).reshape(9)
        tg[i, 9:12] += rbar[i].sum(0)
    return qdd, q_grad, qd_grad, f_grad, tg
