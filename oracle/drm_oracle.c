/* TEST INFRASTRUCTURE -- plain C restatement of the reference's FK / Jacobian / RNEA / articulated-body algorithms.
 *
 * A second, independent CPU oracle (the first is oracle/drm_oracle.py): scalar code, one joint
 * configuration at a time, pthreads over the batch, compiled twice (float and double) from this file
 * by oracle/Makefile.  It follows the reference's own structure -- per-link joint transform from
 * rpy + signed-axis elementary rotation, chain walk with parent lookup, body-frame spatial velocity /
 * acceleration propagation, per-link inertia product with the parallel-axis term, leaves-to-root
 * wrench propagation -- NOT the canonical-frame formulation of the CUDA kernels, so agreement between
 * the two is meaningful.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * link or call it; it is pinned against tests/golden/ (outputs of the reference itself) by
 * tests/test_oracle.py.
 *
 * Reference lines followed (relative to /root/reference/differentiable_robot_model/):
 *   joint_rot            rigid_body.py:138-156, spatial_vector_algebra.py:14-53
 *   walk (poses, vel)    robot_model.py:140-195, spatial_vector_algebra.py:92-106, 226-236
 *   quaternion           spatial_vector_algebra.py:108-136
 *   jacobian             robot_model.py:627-667
 *   rnea                 robot_model.py:251-375, spatial_vector_algebra.py:204-224, 281-291, 321-338
 *   forward dynamics     robot_model.py:488-624 (6x6 articulated inertias), spatial_vector_algebra.py:138-154, 340-372
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#ifndef REAL
#define REAL float
#endif
#ifndef SUFFIX
#define SUFFIX f32
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

#define MAXL 64

/* minimal parallel-for over the batch (libgomp is not linkable in this image): contiguous row ranges
 * on min(n_threads, online cores) pthreads; n_threads <= 0 means all online cores. */
typedef void (*FN(row_fn))(void* ctx, long b);
typedef struct { FN(row_fn) fn; void* ctx; long lo, hi; } FN(range_job);
static void* FN(range_main)(void* arg) {
    FN(range_job)* j = (FN(range_job)*)arg;
    for (long b = j->lo; b < j->hi; ++b) j->fn(j->ctx, b);
    return NULL;
}
static void FN(parallel_rows)(FN(row_fn) fn, void* ctx, long batch, int n_threads) {
    long cores = sysconf(_SC_NPROCESSORS_ONLN);
    if (n_threads <= 0 || n_threads > cores) n_threads = (int)cores;
    if (n_threads > batch) n_threads = batch > 0 ? (int)batch : 1;
    if (n_threads <= 1) { for (long b = 0; b < batch; ++b) fn(ctx, b); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    FN(range_job)* jobs = (FN(range_job)*)malloc(sizeof(FN(range_job)) * n_threads);
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].lo = batch * t / n_threads; jobs[t].hi = batch * (t + 1) / n_threads;
        pthread_create(&th[t], NULL, FN(range_main), &jobs[t]);
    }
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

typedef struct {
    int n_links, n_dofs;
    const int* parent;      /* [N], -1 root */
    const int* dof;         /* [N], -1 fixed */
    const REAL* axis;       /* [N,3] */
    const REAL* trans;      /* [N,3] */
    const REAL* rpy;        /* [N,3] */
    const REAL* mass;       /* [N] */
    const REAL* com;        /* [N,3] */
    const REAL* inertia;    /* [N,9] row-major, at the COM */
    const REAL* damping;    /* [N] */
} FN(drm_oracle_robot);

static void mat_mul(const REAL* a, const REAL* b, REAL* c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
static void mat_vec(const REAL* a, const REAL* v, REAL* r) {
    for (int i = 0; i < 3; ++i) r[i] = a[3 * i] * v[0] + a[3 * i + 1] * v[1] + a[3 * i + 2] * v[2];
}
static void matT_vec(const REAL* a, const REAL* v, REAL* r) {
    for (int i = 0; i < 3; ++i) r[i] = a[i] * v[0] + a[3 + i] * v[1] + a[6 + i] * v[2];
}
static void cross3(const REAL* a, const REAL* b, REAL* r) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
static void elem_rot(int k, REAL ang, REAL* R) {     /* x_rot / y_rot / z_rot */
    const REAL c = (REAL)cos((double)ang), s = (REAL)sin((double)ang);
    memset(R, 0, 9 * sizeof(REAL));
    if (k == 0) { R[0] = 1; R[4] = c; R[5] = -s; R[7] = s; R[8] = c; }
    else if (k == 1) { R[0] = c; R[2] = s; R[4] = 1; R[6] = -s; R[8] = c; }
    else { R[0] = c; R[1] = -s; R[3] = s; R[4] = c; R[8] = 1; }
}
/* joint pose rotation of link i: Rz(yaw) Ry(pitch) Rx(roll) R_axis(sign * q) */
static void joint_rot(const FN(drm_oracle_robot)* rb, int i, const REAL* q, REAL* Rj) {
    REAL Rx[9], Ry[9], Rz[9], T[9], F[9], Q[9];
    elem_rot(0, rb->rpy[3 * i], Rx);
    elem_rot(1, rb->rpy[3 * i + 1], Ry);
    elem_rot(2, rb->rpy[3 * i + 2], Rz);
    mat_mul(Rz, Ry, T);
    mat_mul(T, Rx, F);
    if (rb->dof[i] < 0) { memcpy(Rj, F, sizeof(F)); return; }
    const REAL* ax = rb->axis + 3 * i;
    int k = 2;
    if (fabs((double)ax[0]) == 1.0) k = 0; else if (fabs((double)ax[1]) == 1.0) k = 1;
    const REAL sg = (REAL)((ax[k] > 0) - (ax[k] < 0));
    elem_rot(k, sg * q[rb->dof[i]], Q);
    mat_mul(F, Q, Rj);
}

static void quat_from_rot(const REAL* R, REAL* q) {
    const REAL tr = R[0] + R[4] + R[8];
    REAL t;
    if (tr + 1 > 1) {
        t = tr + 1; q[3] = t; q[2] = R[3] - R[1]; q[1] = R[2] - R[6]; q[0] = R[7] - R[5];
    } else {
        int i = 0, j = 1, k = 2;
        if (R[4] > R[0]) { i = 1; j = 2; k = 0; }
        if (R[8] > R[4 * i]) { i = 2; j = 0; k = 1; }
        t = R[4 * i] - (R[4 * j] + R[4 * k]) + 1;
        q[i] = t; q[j] = R[3 * i + j] + R[3 * j + i]; q[k] = R[3 * k + i] + R[3 * i + k];
        q[3] = R[3 * k + j] - R[3 * j + k];
    }
    const REAL sc = (REAL)(0.5 / sqrt((double)t));
    for (int c = 0; c < 4; ++c) q[c] *= sc;
}

typedef struct { const FN(drm_oracle_robot)* rb; int ee; const REAL* q; REAL *pos, *quat, *jlin, *jang; } FN(fk_ctx);
static void FN(fk_row)(void* vctx, long b) {
    const FN(fk_ctx)* x = (const FN(fk_ctx)*)vctx;
    const FN(drm_oracle_robot)* rb = x->rb;
    const int ee = x->ee;
    const REAL* q = x->q;
    REAL *pos = x->pos, *quat = x->quat, *jlin = x->jlin, *jang = x->jang;
    const int N = rb->n_links, n = rb->n_dofs;
    {
        REAL R[MAXL][9], p[MAXL][3], Rj[9], t[3];
        memset(R[0], 0, sizeof(R[0])); R[0][0] = R[0][4] = R[0][8] = 1;
        p[0][0] = p[0][1] = p[0][2] = 0;
        const REAL* qb = q + b * n;
        for (int i = 1; i < N; ++i) {            /* the reference walks every link, not just the path */
            const int par = rb->parent[i];
            joint_rot(rb, i, qb, Rj);
            mat_mul(R[par], Rj, R[i]);
            mat_vec(R[par], rb->trans + 3 * i, t);
            for (int c = 0; c < 3; ++c) p[i][c] = t[c] + p[par][c];
        }
        if (pos) for (int c = 0; c < 3; ++c) pos[3 * b + c] = p[ee][c];
        if (quat) quat_from_rot(R[ee], quat + 4 * b);
        if (jlin && jang) {
            REAL* jl = jlin + b * 3 * n;
            REAL* ja = jang + b * 3 * n;
            memset(jl, 0, 3 * n * sizeof(REAL));
            memset(ja, 0, 3 * n * sizeof(REAL));
            for (int i = ee; i > 0; i = rb->parent[i]) {
                if (rb->dof[i] < 0) continue;
                REAL z[3], d[3], l[3];
                mat_vec(R[i], rb->axis + 3 * i, z);
                for (int c = 0; c < 3; ++c) d[c] = p[ee][c] - p[i][c];
                cross3(z, d, l);
                for (int c = 0; c < 3; ++c) { jl[c * n + rb->dof[i]] = l[c]; ja[c * n + rb->dof[i]] = z[c]; }
            }
        }
    }
}
void FN(drm_oracle_fk_jacobian)(const FN(drm_oracle_robot)* rb, int ee, const REAL* q, long batch, REAL* pos,
                                REAL* quat, REAL* jlin, REAL* jang, int n_threads) {
    FN(fk_ctx) ctx = {rb, ee, q, pos, quat, jlin, jang};
    FN(parallel_rows)(FN(fk_row), &ctx, batch, n_threads);
}

/* f = I v for the spatial inertia of link i (parallel-axis term recomputed, like the reference) */
static void inertia_times(const FN(drm_oracle_robot)* rb, int i, const REAL* ang, const REAL* lin, REAL* f_lin,
                          REAL* f_ang) {
    const REAL m = rb->mass[i];
    const REAL* c = rb->com + 3 * i;
    const REAL S[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
    REAL St[9], SSt[9], Io[9], mc[3], t[3], u[3];
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) St[3 * r + k] = S[3 * k + r];
    mat_mul(S, St, SSt);
    for (int e = 0; e < 9; ++e) Io[e] = rb->inertia[9 * i + e] + m * SSt[e];
    for (int e = 0; e < 3; ++e) mc[e] = m * c[e];
    cross3(mc, ang, t);
    for (int e = 0; e < 3; ++e) f_lin[e] = m * lin[e] - t[e];
    mat_vec(Io, ang, u);
    cross3(mc, lin, t);
    for (int e = 0; e < 3; ++e) f_ang[e] = u[e] + t[e];
}

typedef struct { const FN(drm_oracle_robot)* rb; const REAL *q, *qd, *qdd; int gravity, damping; REAL* tau; } FN(id_ctx);
static void FN(id_row)(void* vctx, long b) {
    const FN(id_ctx)* x = (const FN(id_ctx)*)vctx;
    const FN(drm_oracle_robot)* rb = x->rb;
    const REAL *q = x->q, *qd = x->qd, *qdd = x->qdd;
    const int gravity = x->gravity, damping = x->damping;
    REAL* tau = x->tau;
    const int N = rb->n_links, n = rb->n_dofs;
    {
        REAL Rj[MAXL][9], w[MAXL][3], v[MAXL][3], al[MAXL][3], a[MAXL][3], fl[MAXL][3], fa[MAXL][3];
        const REAL* qb = q + b * n;
        const REAL* qdb = qd + b * n;
        const REAL* qddb = qdd + b * n;
        memset(w[0], 0, sizeof(w[0])); memset(v[0], 0, sizeof(v[0])); memset(al[0], 0, sizeof(al[0]));
        a[0][0] = a[0][1] = 0; a[0][2] = gravity ? (REAL)9.81 : 0;
        for (int i = 1; i < N; ++i) {
            const int par = rb->parent[i];
            const REAL* r = rb->trans + 3 * i;
            REAL jw[3] = {0, 0, 0}, ja[3] = {0, 0, 0}, t[3], u[3];
            joint_rot(rb, i, qb, Rj[i]);
            if (rb->dof[i] >= 0)
                for (int c = 0; c < 3; ++c) { jw[c] = qdb[rb->dof[i]] * rb->axis[3 * i + c]; ja[c] = qddb[rb->dof[i]] * rb->axis[3 * i + c]; }
            /* parent motion transformed by the inverse joint pose (R^T, -R^T r) */
            matT_vec(Rj[i], w[par], w[i]);
            cross3(w[par], r, t);                                  /* v_p - r x w_p */
            for (int c = 0; c < 3; ++c) u[c] = v[par][c] + t[c];
            matT_vec(Rj[i], u, v[i]);
            for (int c = 0; c < 3; ++c) w[i][c] += jw[c];
            matT_vec(Rj[i], al[par], al[i]);
            cross3(al[par], r, t);
            for (int c = 0; c < 3; ++c) u[c] = a[par][c] + t[c];
            matT_vec(Rj[i], u, a[i]);
            cross3(w[i], jw, t);
            for (int c = 0; c < 3; ++c) al[i][c] += ja[c] + t[c];
            cross3(v[i], jw, t);
            for (int c = 0; c < 3; ++c) a[i][c] += t[c];
        }
        for (int i = 0; i < N; ++i) { memset(fl[i], 0, sizeof(fl[i])); memset(fa[i], 0, sizeof(fa[i])); }
        for (int i = N - 1; i >= 1; --i) {
            REAL il[3], ia[3], vl[3], va[3], t[3], u[3];
            inertia_times(rb, i, al[i], a[i], il, ia);
            inertia_times(rb, i, w[i], v[i], vl, va);
            cross3(w[i], vl, t);
            for (int c = 0; c < 3; ++c) fl[i][c] += il[c] + t[c];
            cross3(w[i], va, t);
            cross3(v[i], vl, u);
            for (int c = 0; c < 3; ++c) fa[i][c] += ia[c] + t[c] + u[c];
            const int par = rb->parent[i];
            REAL nl[3], na[3];
            mat_vec(Rj[i], fl[i], nl);
            mat_vec(Rj[i], fa[i], na);
            cross3(rb->trans + 3 * i, nl, t);
            for (int c = 0; c < 3; ++c) { fl[par][c] += nl[c]; fa[par][c] += t[c] + na[c]; }
        }
        for (int i = 1; i < N; ++i) {
            if (rb->dof[i] < 0) continue;
            const REAL* ax = rb->axis + 3 * i;
            REAL t = fa[i][0] * ax[0] + fa[i][1] * ax[1] + fa[i][2] * ax[2];
            if (damping) t += rb->damping[i] * qdb[rb->dof[i]];
            tau[b * n + rb->dof[i]] = t;
        }
    }
}
void FN(drm_oracle_inverse_dynamics)(const FN(drm_oracle_robot)* rb, const REAL* q, const REAL* qd, const REAL* qdd,
                                     long batch, int gravity, int damping, REAL* tau, int n_threads) {
    FN(id_ctx) ctx = {rb, q, qd, qdd, gravity, damping, tau};
    FN(parallel_rows)(FN(id_row), &ctx, batch, n_threads);
}

/* ---- articulated-body forward dynamics (robot_model.py:488-624), 6x6 matrices / 6-vectors in [ang; lin] order ---- */
static void mat6_vec(const REAL* A, const REAL* x, REAL* y) {
    for (int i = 0; i < 6; ++i) { REAL s = 0; for (int j = 0; j < 6; ++j) s += A[6 * i + j] * x[j]; y[i] = s; }
}
/* X = CoordinateTransform.to_matrix() of the joint pose (sva:138-154): [[R^T, 0], [-R^T t^, R^T]] */
static void motion_matrix(const REAL* Rj, const REAL* t, REAL* X) {
    const REAL S[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
    REAL Rt[9], RtS[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Rt[3 * r + c] = Rj[3 * c + r];
    mat_mul(Rt, S, RtS);
    memset(X, 0, 36 * sizeof(REAL));
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
        X[6 * r + c] = Rt[3 * r + c];
        X[6 * (r + 3) + c] = -RtS[3 * r + c];
        X[6 * (r + 3) + c + 3] = Rt[3 * r + c];
    }
}
/* get_spatial_mat (sva:340-372): [[Io, mc^], [(mc^)^T, m 1]], Io = I_c + m S(c) S(c)^T, NOT symmetrised */
static void spatial_inertia(const FN(drm_oracle_robot)* rb, int i, REAL* I6) {
    const REAL m = rb->mass[i];
    const REAL* c = rb->com + 3 * i;
    const REAL S[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
    REAL St[9], SSt[9];
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) St[3 * r + k] = S[3 * k + r];
    mat_mul(S, St, SSt);
    memset(I6, 0, 36 * sizeof(REAL));
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) {
        I6[6 * r + k] = rb->inertia[9 * i + 3 * r + k] + m * SSt[3 * r + k];
        I6[6 * r + k + 3] = m * S[3 * r + k];
        I6[6 * (r + 3) + k] = m * S[3 * k + r];
    }
    I6[21] = I6[28] = I6[35] = m;
}

typedef struct { const FN(drm_oracle_robot)* rb; const REAL *q, *qd, *f; int gravity, damping; REAL* qdd; } FN(fd_ctx);
static void FN(fd_row)(void* vctx, long b) {
    const FN(fd_ctx)* x = (const FN(fd_ctx)*)vctx;
    const FN(drm_oracle_robot)* rb = x->rb;
    const int N = rb->n_links, n = rb->n_dofs;
    const REAL* qb = x->q + b * n;
    const REAL* qdb = x->qd + b * n;
    const REAL* fb = x->f + b * n;
    REAL* out = x->qdd + b * n;
    static __thread REAL IA[MAXL][36], X[MAXL][36];
    REAL Rj[MAXL][9], w[MAXL][3], v[MAXL][3], c[MAXL][6], pA[MAXL][6], U[MAXL][6], d[MAXL], u[MAXL], S[MAXL][6], acc[MAXL][6];
    memset(w[0], 0, sizeof(w[0])); memset(v[0], 0, sizeof(v[0]));
    for (int i = 1; i < N; ++i) {                                   /* kinematic state + bias terms (:524-545) */
        const int par = rb->parent[i];
        const REAL* r = rb->trans + 3 * i;
        REAL jw[3] = {0, 0, 0}, t[3], s[3], hl[3], ha[3];
        joint_rot(rb, i, qb, Rj[i]);
        motion_matrix(Rj[i], r, X[i]);
        for (int k = 0; k < 3; ++k) S[i][k] = rb->axis[3 * i + k], S[i][k + 3] = 0;       /* zero for fixed links (:550-553) */
        if (rb->dof[i] >= 0) for (int k = 0; k < 3; ++k) jw[k] = qdb[rb->dof[i]] * rb->axis[3 * i + k];
        matT_vec(Rj[i], w[par], w[i]);
        cross3(w[par], r, t);
        for (int k = 0; k < 3; ++k) s[k] = v[par][k] + t[k];
        matT_vec(Rj[i], s, v[i]);
        for (int k = 0; k < 3; ++k) w[i][k] += jw[k];
        cross3(w[i], jw, c[i]);                                     /* cross_motion_vec, joint_vel.lin = 0 (:541) */
        cross3(v[i], jw, c[i] + 3);
        inertia_times(rb, i, w[i], v[i], hl, ha);
        cross3(w[i], ha, t); cross3(v[i], hl, s);                   /* cross_force_vec (:543) */
        for (int k = 0; k < 3; ++k) pA[i][k] = t[k] + s[k];
        cross3(w[i], hl, pA[i] + 3);
        spatial_inertia(rb, i, IA[i]);
    }
    for (int i = N - 1; i >= 1; --i) {                              /* articulated inertias (:547-596) */
        const int par = rb->parent[i];
        mat6_vec(IA[i], S[i], U[i]);
        d[i] = 0; u[i] = 0;
        for (int k = 0; k < 6; ++k) { d[i] += S[i][k] * U[i][k]; u[i] -= pA[i][k] * S[i][k]; }
        if (rb->dof[i] >= 0) {
            REAL fk = fb[rb->dof[i]];
            if (x->damping) fk -= rb->damping[i] * qdb[rb->dof[i]];
            u[i] += fk;
        }
        if (par > 0) {
            REAL IAp[36], T1[36], tmp[6], pa[6];
            const REAL dd = d[i] + (REAL)1e-37;
            for (int r = 0; r < 6; ++r) for (int k = 0; k < 6; ++k) IAp[6 * r + k] = IA[i][6 * r + k] - U[i][r] * (U[i][k] / dd);
            mat6_vec(IAp, c[i], tmp);
            for (int k = 0; k < 6; ++k) pa[k] = pA[i][k] + tmp[k] + U[i][k] * (u[i] / dd);
            for (int r = 0; r < 6; ++r) for (int k = 0; k < 6; ++k) {        /* X^T IA' */
                REAL s = 0; for (int m = 0; m < 6; ++m) s += X[i][6 * m + r] * IAp[6 * m + k]; T1[6 * r + k] = s;
            }
            for (int r = 0; r < 6; ++r) for (int k = 0; k < 6; ++k) {        /* (X^T IA') X */
                REAL s = 0; for (int m = 0; m < 6; ++m) s += T1[6 * r + m] * X[i][6 * m + k]; IA[par][6 * r + k] += s;
            }
            REAL nl[3], na[3], t[3];                                /* SpatialForceVec.transform (sva:281-291) */
            mat_vec(Rj[i], pa + 3, nl);
            mat_vec(Rj[i], pa, na);
            cross3(rb->trans + 3 * i, nl, t);
            for (int k = 0; k < 3; ++k) { pA[par][k] += t[k] + na[k]; pA[par][k + 3] += nl[k]; }
        }
    }
    memset(acc[0], 0, sizeof(acc[0]));
    acc[0][5] = x->gravity ? (REAL)9.81 : 0;
    for (int i = 1; i < N; ++i) {                                   /* accelerations (:604-622) */
        mat6_vec(X[i], acc[rb->parent[i]], acc[i]);
        for (int k = 0; k < 6; ++k) acc[i][k] += c[i][k];
        if (rb->dof[i] >= 0) {
            REAL s = 0;
            for (int k = 0; k < 6; ++k) s += U[i][k] * acc[i][k];
            const REAL qdd = ((REAL)1 / d[i]) * (u[i] - s);
            out[rb->dof[i]] = qdd;
            for (int k = 0; k < 6; ++k) acc[i][k] += S[i][k] * qdd;
        }
    }
}
void FN(drm_oracle_forward_dynamics)(const FN(drm_oracle_robot)* rb, const REAL* q, const REAL* qd, const REAL* f,
                                     long batch, int gravity, int damping, REAL* qdd, int n_threads) {
    FN(fd_ctx) ctx = {rb, q, qd, f, gravity, damping, qdd};
    FN(parallel_rows)(FN(fd_row), &ctx, batch, n_threads);
}
