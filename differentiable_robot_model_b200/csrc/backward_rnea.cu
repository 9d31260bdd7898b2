// backward_rnea.cu -- analytic reverse-mode kernels for RNEA inverse dynamics (sm_100a).
//
// See backward.cu for the common design (recompute instead of save, deterministic table-gradient reduction,
// canonical joint frames) and oracle/adjoint_proto.py for the recursions in executable form.
//
// Two kernels:
//   rnea_backward_kernel            the full adjoint: gradients w.r.t. q, qd, qdd and every table column;
//   rnea_backward_inertial_kernel   a single root->leaves sweep for the case that only the inertial columns
//                                   (I_o, mc, m) and the damping are wanted (DRMB200_INERTIAL_GRADS_ONLY).
//
// Shared-memory diet of the full adjoint.  Shared memory is what limits its occupancy.  The leaves->root sweep
// needs the motion state (w, v, al, a) of every link and of its parent, but the forward recursion is invertible
// (E = M^T is orthogonal):
//     w_p  = M (w_i - (0,0,qd))                    v_p = M v_i - w_p x r
//     al_p = M (al_i - (0,0,qdd) - w_i x (0,0,qd))  a_p = M (a_i - v_i x (0,0,qd)) - al_p x r
// so the sweep carries the state DOWN the tree in registers and only the "tips" (links whose successor in
// document order is not their child: the end of every chain) keep their 12 floats in shared memory.  Per link
// that leaves 8 floats (accumulated wrench -> wrench adjoints, cos, sin) instead of 20: 3x the resident warps
// for a 7-DoF arm.  The re-derived states differ from the forward ones by rounding only (~1e-7 relative).
#include "backward_common.cuh"

namespace drm {

struct RneaBwdArgs {
    const float* __restrict__ table;
    const float* __restrict__ q;
    const float* __restrict__ qd;
    const float* __restrict__ qdd;
    const float* __restrict__ g_tau;
    float* __restrict__ q_grad;
    float* __restrict__ qd_grad;
    float* __restrict__ qdd_grad;
    float* __restrict__ partials;
    int64_t batch;
    uint32_t flags;
    int32_t vec_ok;
};

// per-link per-thread state, slot-major: f n -> mu lambda (6) | cos sin (2)
constexpr int LSTATE = 8;

struct RneaBwdSmem {
    int q, qd, qdd, g, qg, qdg, qddg, table, link, slots, tips, scratch, acc, total_floats;
    __host__ __device__ RneaBwdSmem(int tile, int n, int n_links, int n_slots, int n_tips) {
        int o = 0;
        q = o; o += tile * n;
        qd = o; o += tile * n;
        qdd = o; o += tile * n;
        g = o; o += tile * n;
        qg = o; o += tile * n;
        qdg = o; o += tile * n;
        qddg = o; o += tile * n;
        table = o; o += n_links * DRMB200_TABLE_STRIDE;
        link = o; o += n_links * LSTATE * tile;
        slots = o; o += n_slots * 12 * tile;      // forward: branch-point motion states; backward: adjoint accumulators
        tips = o; o += n_tips * 12 * tile;        // motion state of every chain end
        scratch = o; o += block_accumulate_floats(25, tile);
        acc = o; o += n_links * DRMB200_TABLE_STRIDE;
        total_floats = o;
    }
};

template <bool NEED_TABLE, int T>
__global__ void __launch_bounds__(T)
rnea_backward_kernel(const __grid_constant__ TreeProgram prog, const RneaBwdArgs args) {
    extern __shared__ __align__(128) float smem[];
    const int n = prog.n_dofs, N = prog.n_links;
    const RneaBwdSmem L(T, n, N, prog.n_slots, prog.n_tips);
    float* s_q = smem + L.q;
    float* s_qd = smem + L.qd;
    float* s_qdd = smem + L.qdd;
    float* s_g = smem + L.g;
    float* s_qg = smem + L.qg;
    float* s_qdg = smem + L.qdg;
    float* s_qddg = smem + L.qddg;
    float* s_tab = smem + L.table;
    float* s_link = smem + L.link;
    float* s_slot = smem + L.slots;
    float* s_tips = smem + L.tips;
    float* s_scr = smem + L.scratch;
    float* s_acc = smem + L.acc;
    const int tid = threadIdx.x;
    const bool vec_ok = args.vec_ok;
    const float grav = (args.flags & DRMB200_GRAVITY) ? GRAVITY_B : 0.f;
    const bool damp = (args.flags & DRMB200_DAMPING) != 0;

    for (int i = tid; i < N * DRMB200_TABLE_STRIDE; i += T) {
        const int l = i / DRMB200_TABLE_STRIDE, e = i - l * DRMB200_TABLE_STRIDE;
        const int p = prog.parent[l];
        int src;
        const float sg = canon_map(e, p >= 0 ? (int)prog.axis[p] : 0, prog.axis[l], src);
        s_tab[i] = sg * __ldg(args.table + l * DRMB200_TABLE_STRIDE + src);
        if (NEED_TABLE) s_acc[i] = 0.f;
    }

    const int64_t n_tiles = (args.batch + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t start = tile * T;
        const int valid = (int)min((int64_t)T, args.batch - start);
        __syncthreads();
        coop_copy(s_q, args.q + start * n, valid * n, vec_ok);
        coop_copy(s_qd, args.qd + start * n, valid * n, vec_ok);
        coop_copy(s_qdd, args.qdd + start * n, valid * n, vec_ok);
        coop_copy(s_g, args.g_tau + start * n, valid * n, vec_ok);
        __syncthreads();

        const bool active = tid < valid;
        const float* qrow = s_q + tid * n;
        const float* qdrow = s_qd + tid * n;
        const float* qddrow = s_qdd + tid * n;
        const float* grow = s_g + tid * n;
        float* qg = s_qg + tid * n;
        float* qdg = s_qdg + tid * n;
        float* qddg = s_qddg + tid * n;
        float* lk = s_link + tid;
        const V3 zero = v3(0.f, 0.f, 0.f);
        const V3 a_root = v3(0.f, 0.f, grav);

        // ---- forward recompute, pass A: motion state (registers / branch slots) + body wrench ------
        {
            V3 w = zero, v = zero, al = zero, a = zero;
            for (int i = 1; i < N; ++i) {
                const LinkRow C = load_row(s_tab + i * DRMB200_TABLE_STRIDE);
                const int src = prog.psrc[i];
                V3 wp, vp, alp, ap;
                if (src == 0) { wp = w; vp = v; alp = al; ap = a; }
                else if (src < 0) { wp = vp = alp = zero; ap = a_root; }
                else {
                    const float* sl = s_slot + (src - 1) * 12 * T + tid;
                    wp = ldv(sl, T); vp = ldv(sl + 3 * T, T); alp = ldv(sl + 6 * T, T); ap = ldv(sl + 9 * T, T);
                }
                M3 M = C.F;
                const int c = prog.dof[i];
                float cs = 1.f, sn = 0.f, qd_k = 0.f, qdd_k = 0.f;
                if (c >= 0) {
                    qd_k = qdrow[c]; qdd_k = qddrow[c];
                    sincos_pi2(qrow[c], sn, cs);
                    rotate_z(M, cs, sn);
                }
                w = mulT(M, wp); w.z += qd_k;
                v = mulT(M, cross_add(wp, C.r, vp));
                al = mulT(M, alp) + cross_z(w, qd_k); al.z += qdd_k;
                a = mulT(M, cross_add(alp, C.r, ap)) + cross_z(v, qd_k);
                const V3 hl_a = C.m * a - cross(C.mc, al);
                const V3 ha_a = mul_add(C.Io, al, cross(C.mc, a));
                const V3 hl_v = C.m * v - cross(C.mc, w);
                const V3 ha_v = mul_add(C.Io, w, cross(C.mc, v));
                const V3 f = cross_add(w, hl_v, hl_a);
                const V3 nn = cross_add(w, ha_v, cross_add(v, hl_v, ha_a));
                float* s = lk + i * LSTATE * T;
                stv(s, T, f); stv(s + 3 * T, T, nn);
                s[6 * T] = cs; s[7 * T] = sn;
                const int sv = prog.save[i];
                if (sv >= 0) {
                    float* sl = s_slot + sv * 12 * T + tid;
                    stv(sl, T, w); stv(sl + 3 * T, T, v); stv(sl + 6 * T, T, al); stv(sl + 9 * T, T, a);
                }
                const int tp = prog.tip[i];
                if (tp >= 0) {
                    float* st = s_tips + tp * 12 * T + tid;
                    stv(st, T, w); stv(st + 3 * T, T, v); stv(st + 6 * T, T, al); stv(st + 9 * T, T, a);
                }
            }
        }
        // ---- forward recompute, pass B: accumulate wrenches leaves -> root -----------------------
        for (int i = N - 1; i >= 1; --i) {
            const int P = prog.parent[i];
            if (P <= 0) continue;
            const float* s = lk + i * LSTATE * T;
            M3 F; V3 r;
            load_Fr(s_tab + i * DRMB200_TABLE_STRIDE, F, r);
            const float cs = s[6 * T], sn = s[7 * T];
            const V3 fp = mul(F, rotz(ldv(s, T), cs, sn));
            const V3 np = cross_add(r, fp, mul(F, rotz(ldv(s + 3 * T, T), cs, sn)));
            float* sp = lk + P * LSTATE * T;
            stv(sp, T, ldv(sp, T) + fp);
            stv(sp + 3 * T, T, ldv(sp + 3 * T, T) + np);
        }

        // ---- adjoint pass 1, root -> leaves: lambda = n-bar, mu = f-bar ----------------------------
        for (int i = 1; i < N; ++i) {
            float* s = lk + i * LSTATE * T;
            const float* row = s_tab + i * DRMB200_TABLE_STRIDE;
            M3 M; V3 r;
            load_Fr(row, M, r);
            const int P = prog.parent[i];
            V3 lamP = zero, muP = zero;
            if (P > 0) { const float* sp = lk + P * LSTATE * T; muP = ldv(sp, T); lamP = ldv(sp + 3 * T, T); }
            const int c = prog.dof[i];
            const float cs = s[6 * T], sn = s[7 * T];
            float gk = 0.f;
            if (c >= 0) { rotate_z(M, cs, sn); gk = grow[c]; }
            const V3 f = ldv(s, T), nn = ldv(s + 3 * T, T);                 // accumulated wrenches
            const V3 u = cross_add(lamP, r, muP);
            V3 lam = mulT(M, lamP); lam.z += gk;                            // tau_k = n_i . e_z
            const V3 mu = mulT(M, u);
            stv(s, T, mu);                                                   // f slot -> mu
            stv(s + 3 * T, T, lam);                                          // n slot -> lambda
            M3 Mbar = zero3();
            add_outer(Mbar, lamP, nn);
            add_outer(Mbar, u, f);
            const V3 rbar = cross(mul(M, f), lamP);
            float dbar = 0.f;
            if (c >= 0) {
                qg[c] = theta_grad_z(Mbar, M);
                float qdv = 0.f;
                if (damp) { qdv = row[25] * gk; dbar = gk * qdrow[c]; }
                qdg[c] = qdv;
            }
            if (NEED_TABLE) {
                M3 Fbar = Mbar;
                if (c >= 0) rotate_z(Fbar, cs, -sn);
                float vals[13];
                m3_to_array(Fbar, vals);
                vals[9] = rbar.x; vals[10] = rbar.y; vals[11] = rbar.z; vals[12] = dbar;
                block_accumulate<13, T>(s_scr, s_acc + i * DRMB200_TABLE_STRIDE, vals, active,
                                        [](int j) { return j < 12 ? j : 25; });
            }
        }

        // ---- adjoint pass 2, leaves -> root: motion adjoints -------------------------------------
        V3 c_wb = zero, c_vb = zero, c_alb = zero, c_ab = zero;     // adjoints carried into link i from child i+1
        V3 s_w = zero, s_v = zero, s_al = zero, s_a = zero;         // motion state of link i re-derived from child i+1
        for (int i = N - 1; i >= 1; --i) {
            const float* s = lk + i * LSTATE * T;
            const LinkRow C = load_row(s_tab + i * DRMB200_TABLE_STRIDE);
            const int P = prog.parent[i];
            const int c = prog.dof[i];
            const float cs = s[6 * T], sn = s[7 * T];
            M3 M = C.F;
            float qd_k = 0.f, qdd_k = 0.f;
            if (c >= 0) { rotate_z(M, cs, sn); qd_k = qdrow[c]; qdd_k = qddrow[c]; }
            // this link's motion state: stored for chain ends, otherwise handed down by child i+1
            V3 w, v, al, a;
            const int tp = prog.tip[i];
            if (tp >= 0) {
                const float* st = s_tips + tp * 12 * T + tid;
                w = ldv(st, T); v = ldv(st + 3 * T, T); al = ldv(st + 6 * T, T); a = ldv(st + 9 * T, T);
            } else { w = s_w; v = s_v; al = s_al; a = s_a; }
            // the parent's state through the inverted recursion (root: constants)
            V3 wp = zero, vp = zero, alp = zero, ap = a_root;
            if (P > 0) {
                V3 t = w; t.z -= qd_k;
                wp = mul(M, t);
                vp = mul(M, v) - cross(wp, C.r);
                t = al - cross_z(w, qd_k); t.z -= qdd_k;
                alp = mul(M, t);
                ap = mul(M, a - cross_z(v, qd_k)) - cross(alp, C.r);
            }
            s_w = wp; s_v = vp; s_al = alp; s_a = ap;               // used by iteration i-1 iff parent(i) == i-1
            const V3 mu = ldv(s, T), lam = ldv(s + 3 * T, T);
            // incoming adjoints: child i+1 through registers, far children through the branch slot
            V3 wb = zero, vb = zero, alb = zero, ab = zero;
            if (i + 1 < N && prog.psrc[i + 1] == 0) { wb = c_wb; vb = c_vb; alb = c_alb; ab = c_ab; }
            const int sv = prog.save[i];
            if (sv >= 0) {
                const float* sl = s_slot + sv * 12 * T + tid;
                wb = wb + ldv(sl, T); vb = vb + ldv(sl + 3 * T, T); alb = alb + ldv(sl + 6 * T, T); ab = ab + ldv(sl + 9 * T, T);
            }
            // body part (momentum H, wrench adjoints lam / mu)
            const V3 Hl = C.m * v - cross(C.mc, w);
            const V3 Ha = mul_add(C.Io, w, cross(C.mc, v));
            const V3 Hlb = cross_add(mu, w, cross(lam, v));
            const V3 Hab = cross(lam, w);
            alb = alb + cross_add(C.mc, mu, mulT(C.Io, lam));
            ab = ab + cross_add(lam, C.mc, C.m * mu);
            wb = wb + cross_add(Hl, mu, cross_add(Ha, lam, cross_add(C.mc, Hlb, mulT(C.Io, Hab))));
            vb = vb + cross_add(Hl, lam, cross_add(Hab, C.mc, C.m * Hlb));
            float vals[25];
            if (NEED_TABLE) {
                M3 Iob = zero3();
                add_outer(Iob, lam, al);
                add_outer(Iob, Hab, w);
                m3_to_array(Iob, vals + 12);
                const V3 mcb = cross_add(mu, al, cross_add(a, lam, cross_add(Hlb, w, cross(v, Hab))));
                vals[21] = mcb.x; vals[22] = mcb.y; vals[23] = mcb.z;
                vals[24] = dot(mu, a) + dot(Hlb, v);
            }
            // kinematic part, in the order a, alpha, v, omega; wJ = (0, 0, qd_k)
            M3 Mbar = zero3();
            float wJb = ab.x * v.y - ab.y * v.x;            // (ab x v).z -- only the z component of wJ-bar matters
            vb = vb + z_cross(qd_k, ab);
            const V3 ua = mul(M, ab);
            const V3 abP = ua;
            V3 albP = cross(C.r, ua);
            V3 rbar = cross(ua, alp);
            add_outer(Mbar, cross_add(alp, C.r, ap), ab);

            wb = wb + z_cross(qd_k, alb);
            wJb += alb.x * w.y - alb.y * w.x;               // (alb x w).z
            albP = albP + mul(M, alb);
            add_outer(Mbar, alp, alb);

            const V3 uv = mul(M, vb);
            const V3 vbP = uv;
            V3 wbP = cross(C.r, uv);
            rbar = rbar + cross(uv, wp);
            add_outer(Mbar, cross_add(wp, C.r, vp), vb);

            wbP = wbP + mul(M, wb);
            add_outer(Mbar, wp, wb);
            wJb += wb.z;

            if (c >= 0) {
                qddg[c] = alb.z;
                qdg[c] += wJb;
                qg[c] += theta_grad_z(Mbar, M);
            }
            // route the parent contributions
            if (P == i - 1) { c_wb = wbP; c_vb = vbP; c_alb = albP; c_ab = abP; }
            else if (P > 0) {
                float* sl = s_slot + (int)prog.save[P] * 12 * T + tid;
                if (prog.accw[i] == 2) {
                    stv(sl, T, wbP); stv(sl + 3 * T, T, vbP); stv(sl + 6 * T, T, albP); stv(sl + 9 * T, T, abP);
                } else {
                    stv(sl, T, ldv(sl, T) + wbP); stv(sl + 3 * T, T, ldv(sl + 3 * T, T) + vbP);
                    stv(sl + 6 * T, T, ldv(sl + 6 * T, T) + albP); stv(sl + 9 * T, T, ldv(sl + 9 * T, T) + abP);
                }
            }
            if (NEED_TABLE) {
                M3 Fbar = Mbar;
                if (c >= 0) rotate_z(Fbar, cs, -sn);
                m3_to_array(Fbar, vals);
                vals[9] = rbar.x; vals[10] = rbar.y; vals[11] = rbar.z;
                block_accumulate<25, T>(s_scr, s_acc + i * DRMB200_TABLE_STRIDE, vals, active, [](int j) { return j; });
            }
        }
        __syncthreads();
        if (args.q_grad != nullptr) coop_copy(args.q_grad + start * n, s_qg, valid * n, vec_ok);
        if (args.qd_grad != nullptr) coop_copy(args.qd_grad + start * n, s_qdg, valid * n, vec_ok);
        if (args.qdd_grad != nullptr) coop_copy(args.qdd_grad + start * n, s_qddg, valid * n, vec_ok);
    }
    if (NEED_TABLE) {
        __syncthreads();
        float* out = args.partials + (size_t)blockIdx.x * N * DRMB200_TABLE_STRIDE;
        for (int i = tid; i < N * DRMB200_TABLE_STRIDE; i += T) {        // canonical -> natural (bijection per row)
            const int l = i / DRMB200_TABLE_STRIDE, e = i - l * DRMB200_TABLE_STRIDE;
            const int p = prog.parent[l];
            int src;
            const float sg = canon_map(e, p >= 0 ? (int)prog.axis[p] : 0, prog.axis[l], src);
            out[l * DRMB200_TABLE_STRIDE + src] = sg * s_acc[i];
        }
    }
}

// =============================================================================================
// RNEA backward, inertial parameters only (DRMB200_INERTIAL_GRADS_ONLY)
// =============================================================================================
// tau is LINEAR in (m, mc, I_o) and the damping, and the wrench adjoints (lambda, mu) obey a root->leaves
// recursion just like the motion state, so when only those table columns are wanted (the classic "learn the link
// inertias" setting, BASELINE config 5: nothing kinematic is learnable and no input gradients are requested) the
// whole backward collapses into ONE root->leaves sweep with no per-link storage:
//   lam_i = E lam_p + (0,0,g_k)        mu_i = E (mu_p + lam_p x r)
//   Io-bar = lam al^T + (lam x w) w^T   mc-bar = mu x al + a x lam + Hl-bar x w + v x Ha-bar
//   m-bar  = mu . a + Hl-bar . v        d-bar  = g_k qd_k           (Hl-bar = mu x w + lam x v, Ha-bar = lam x w)
// About 230 instructions per link instead of ~1350 for the full adjoint, and shared memory only for the I/O tiles.
struct RneaInertialSmem {
    int q, qd, qdd, g, table, slots, scratch, acc, total_floats;
    __host__ __device__ RneaInertialSmem(int tile, int n, int n_links, int n_slots) {
        int o = 0;
        q = o; o += tile * n;
        qd = o; o += tile * n;
        qdd = o; o += tile * n;
        g = o; o += tile * n;
        table = o; o += n_links * DRMB200_TABLE_STRIDE;
        slots = o; o += n_slots * 18 * tile;
        scratch = o; o += block_accumulate_floats(14, tile);
        acc = o; o += n_links * DRMB200_TABLE_STRIDE;
        total_floats = o;
    }
};

template <int T>
__global__ void __launch_bounds__(T)
rnea_backward_inertial_kernel(const __grid_constant__ TreeProgram prog, const RneaBwdArgs args) {
    extern __shared__ __align__(128) float smem[];
    const int n = prog.n_dofs, N = prog.n_links;
    const RneaInertialSmem L(T, n, N, prog.n_slots);
    float* s_q = smem + L.q;
    float* s_qd = smem + L.qd;
    float* s_qdd = smem + L.qdd;
    float* s_g = smem + L.g;
    float* s_tab = smem + L.table;
    float* s_slot = smem + L.slots;
    float* s_scr = smem + L.scratch;
    float* s_acc = smem + L.acc;
    const int tid = threadIdx.x;
    const bool vec_ok = args.vec_ok;
    const float grav = (args.flags & DRMB200_GRAVITY) ? GRAVITY_B : 0.f;
    const bool damp = (args.flags & DRMB200_DAMPING) != 0;

    for (int i = tid; i < N * DRMB200_TABLE_STRIDE; i += T) {
        const int l = i / DRMB200_TABLE_STRIDE, e = i - l * DRMB200_TABLE_STRIDE;
        const int p = prog.parent[l];
        int src;
        const float sg = canon_map(e, p >= 0 ? (int)prog.axis[p] : 0, prog.axis[l], src);
        s_tab[i] = sg * __ldg(args.table + l * DRMB200_TABLE_STRIDE + src);
        s_acc[i] = 0.f;
    }
    const int64_t n_tiles = (args.batch + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t start = tile * T;
        const int valid = (int)min((int64_t)T, args.batch - start);
        __syncthreads();
        coop_copy(s_q, args.q + start * n, valid * n, vec_ok);
        coop_copy(s_qd, args.qd + start * n, valid * n, vec_ok);
        coop_copy(s_qdd, args.qdd + start * n, valid * n, vec_ok);
        coop_copy(s_g, args.g_tau + start * n, valid * n, vec_ok);
        __syncthreads();
        const bool active = tid < valid;
        const float* qrow = s_q + tid * n;
        const float* qdrow = s_qd + tid * n;
        const float* qddrow = s_qdd + tid * n;
        const float* grow = s_g + tid * n;
        const V3 zero = v3(0.f, 0.f, 0.f);
        V3 w = zero, v = zero, al = zero, a = zero, lam = zero, mu = zero;      // state of the previous link
        for (int i = 1; i < N; ++i) {
            M3 M; V3 r;
            load_Fr(s_tab + i * DRMB200_TABLE_STRIDE, M, r);
            const int src = prog.psrc[i];
            V3 wp, vp, alp, ap, lamP, muP;
            if (src == 0) { wp = w; vp = v; alp = al; ap = a; lamP = lam; muP = mu; }
            else if (src < 0) { wp = vp = alp = lamP = muP = zero; ap = v3(0.f, 0.f, grav); }
            else {
                const float* sl = s_slot + (src - 1) * 18 * T + tid;
                wp = ldv(sl, T); vp = ldv(sl + 3 * T, T); alp = ldv(sl + 6 * T, T); ap = ldv(sl + 9 * T, T);
                lamP = ldv(sl + 12 * T, T); muP = ldv(sl + 15 * T, T);
            }
            const int c = prog.dof[i];
            float qd_k = 0.f, qdd_k = 0.f, gk = 0.f;
            if (c >= 0) {
                float sn, cs;
                sincos_pi2(qrow[c], sn, cs);
                rotate_z(M, cs, sn);
                qd_k = qdrow[c]; qdd_k = qddrow[c]; gk = grow[c];
            }
            w = mulT(M, wp); w.z += qd_k;
            v = mulT(M, cross_add(wp, r, vp));
            al = mulT(M, alp) + cross_z(w, qd_k); al.z += qdd_k;
            a = mulT(M, cross_add(alp, r, ap)) + cross_z(v, qd_k);
            lam = mulT(M, lamP); lam.z += gk;
            mu = mulT(M, cross_add(lamP, r, muP));
            const V3 Hlb = cross_add(mu, w, cross(lam, v));
            const V3 Hab = cross(lam, w);
            float vals[14];
            M3 Iob = zero3();
            add_outer(Iob, lam, al);
            add_outer(Iob, Hab, w);
            m3_to_array(Iob, vals);
            const V3 mcb = cross_add(mu, al, cross_add(a, lam, cross_add(Hlb, w, cross(v, Hab))));
            vals[9] = mcb.x; vals[10] = mcb.y; vals[11] = mcb.z;
            vals[12] = dot(mu, a) + dot(Hlb, v);
            vals[13] = (damp && c >= 0) ? gk * qd_k : 0.f;
            block_accumulate<14, T>(s_scr, s_acc + i * DRMB200_TABLE_STRIDE, vals, active,
                                    [](int j) { return j < 13 ? 12 + j : 25; });
            const int sv = prog.save[i];
            if (sv >= 0) {
                float* sl = s_slot + sv * 18 * T + tid;
                stv(sl, T, w); stv(sl + 3 * T, T, v); stv(sl + 6 * T, T, al); stv(sl + 9 * T, T, a);
                stv(sl + 12 * T, T, lam); stv(sl + 15 * T, T, mu);
            }
        }
    }
    __syncthreads();
    float* out = args.partials + (size_t)blockIdx.x * N * DRMB200_TABLE_STRIDE;
    for (int i = tid; i < N * DRMB200_TABLE_STRIDE; i += T) {            // canonical -> natural (bijection per row)
        const int l = i / DRMB200_TABLE_STRIDE, e = i - l * DRMB200_TABLE_STRIDE;
        const int p = prog.parent[l];
        int src;
        const float sg = canon_map(e, p >= 0 ? (int)prog.axis[p] : 0, prog.axis[l], src);
        out[l * DRMB200_TABLE_STRIDE + src] = sg * s_acc[i];
    }
}

int inverse_dynamics_backward_device(const drmb200_topology_t* topo, const float* table, const float* q,
                                     const float* qd, const float* qdd, int64_t batch, uint32_t flags,
                                     const float* g_tau, float* q_grad, float* qd_grad, float* qdd_grad,
                                     float* table_grad, void* workspace, cudaStream_t stream) {
    TreeProgram prog;
    int rc = build_tree_program(topo, &prog);
    if (rc != DRMB200_OK) return rc;
    if (batch < 0) { set_error("batch=%lld < 0", (long long)batch); return DRMB200_EINVAL; }
    if (batch == 0 || prog.n_dofs == 0) return DRMB200_OK;
    if (q_grad == nullptr && qd_grad == nullptr && qdd_grad == nullptr && table_grad == nullptr) return DRMB200_OK;
    if (table == nullptr || q == nullptr || qd == nullptr || qdd == nullptr || g_tau == nullptr) { set_error("null pointer argument"); return DRMB200_EINVAL; }
    if (table_grad != nullptr && workspace == nullptr) { set_error("table_grad requested without workspace"); return DRMB200_EINVAL; }

    RneaBwdArgs args;
    args.table = table; args.q = q; args.qd = qd; args.qdd = qdd; args.g_tau = g_tau;
    args.q_grad = q_grad; args.qd_grad = qd_grad; args.qdd_grad = qdd_grad;
    args.partials = static_cast<float*>(workspace); args.batch = batch; args.flags = flags;
    auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    args.vec_ok = (al16(q) && al16(qd) && al16(qdd) && al16(g_tau) && al16(q_grad) && al16(qd_grad) && al16(qdd_grad)) ? 1 : 0;

    if ((flags & DRMB200_INERTIAL_GRADS_ONLY) && table_grad != nullptr) {
        if (q_grad != nullptr || qd_grad != nullptr || qdd_grad != nullptr) {
            set_error("DRMB200_INERTIAL_GRADS_ONLY cannot be combined with input gradients");
            return DRMB200_EINVAL;
        }
        constexpr int TI = 128;
        const size_t sb = (size_t)RneaInertialSmem(TI, prog.n_dofs, prog.n_links, prog.n_slots).total_floats * sizeof(float);
        if (sb > 227 * 1024) { set_error("rnea inertial backward needs %zu B of shared memory (> 227 KB)", sb); return DRMB200_ELIMIT; }
        int g = 0;
        rc = persistent_grid(rnea_backward_inertial_kernel<TI>, TI, sb, (batch + TI - 1) / TI, &g, "rnea inertial backward");
        if (rc != DRMB200_OK) return rc;
        rnea_backward_inertial_kernel<TI><<<g, TI, sb, stream>>>(prog, args);
        cudaError_t ei = cudaGetLastError();
        if (ei != cudaSuccess) { set_error("rnea inertial backward launch: %s", cudaGetErrorString(ei)); return DRMB200_ECUDA; }
        count_launch();
        return launch_reduce(args.partials, g, topo, table_grad, stream);
    }

    // shared memory (8 floats per link + 12 per chain end, per configuration) is the occupancy limiter: pick the
    // tile that keeps the most warps resident per SM, larger tile on ties
    int tile = 32, best_warps = 0;
    for (int t = 128; t >= 32; t >>= 1) {
        const size_t b = (size_t)RneaBwdSmem(t, prog.n_dofs, prog.n_links, prog.n_slots, prog.n_tips).total_floats * sizeof(float) + 1024;
        const int warps = b > 227 * 1024 ? 0 : (int)((227 * 1024) / b) * (t / 32);
        if (warps > best_warps) { best_warps = warps; tile = t; }
    }
    const size_t smem_bytes = (size_t)RneaBwdSmem(tile, prog.n_dofs, prog.n_links, prog.n_slots, prog.n_tips).total_floats * sizeof(float);
    if (smem_bytes > 227 * 1024) { set_error("rnea backward needs %zu B of shared memory per CTA (> 227 KB): model too large", smem_bytes); return DRMB200_ELIMIT; }
    const int64_t tiles = (batch + tile - 1) / tile;
    int grid = 0;
    const bool need_table = table_grad != nullptr;
#define DRM_LAUNCH_IDB(NT, TT)                                                                              \
    do {                                                                                                    \
        rc = persistent_grid(rnea_backward_kernel<NT, TT>, TT, smem_bytes, tiles, &grid, "rnea backward");  \
        if (rc != DRMB200_OK) return rc;                                                                    \
        rnea_backward_kernel<NT, TT><<<grid, TT, smem_bytes, stream>>>(prog, args);                         \
    } while (0)
    if (need_table) { if (tile == 128) DRM_LAUNCH_IDB(true, 128); else if (tile == 64) DRM_LAUNCH_IDB(true, 64); else DRM_LAUNCH_IDB(true, 32); }
    else            { if (tile == 128) DRM_LAUNCH_IDB(false, 128); else if (tile == 64) DRM_LAUNCH_IDB(false, 64); else DRM_LAUNCH_IDB(false, 32); }
#undef DRM_LAUNCH_IDB
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("rnea backward launch: %s", cudaGetErrorString(e)); return DRMB200_ECUDA; }
    count_launch();
    return need_table ? launch_reduce(args.partials, grid, topo, table_grad, stream) : DRMB200_OK;
}

}  // namespace drm
