// backward.cu -- analytic reverse-mode kernels for FK/Jacobian and RNEA (sm_100a).
//
// The reference differentiates its per-link op graph with torch autograd (about 7 k graph nodes for one
// Kuka RNEA).  Here each backward is ONE launch that re-walks the chain / tree per configuration and
// evaluates the adjoint recursions of SURVEY.md Appendix B; the exact recursions (same order, same
// intermediate quantities) are stated in executable form in oracle/adjoint_proto.py and verified there
// against autograd of the fp64 oracle.  Like the forward kernels they work in the canonical joint
// frames of drm_common.cuh (every joint a +z rotation), and map the table gradient back through the
// inverse signed permutation when it is written out.
//
// Outputs:
//   * gradients w.r.t. the per-configuration inputs (q, qd, qdd): one row per configuration, staged in
//     shared memory and streamed out exactly like the forward outputs;
//   * the gradient of the link table [n_links, 28]: a SUM OVER THE BATCH.  Per link, every thread of a
//     CTA drops its partial values into a padded shared-memory scratch matrix, the warps reduce its rows
//     with shuffles into a per-CTA accumulator (deterministic order), CTAs are persistent over tiles, and
//     each CTA writes one [n_links, 28] partial to a workspace that a second tiny kernel sums in fixed
//     order -- no atomics, bitwise reproducible.
//
// Nothing per-link is saved by the forward kernels (that would add >= 15 floats per link per
// configuration of HBM traffic); everything is recomputed here from q, qd, qdd.
#include "drm_common.cuh"

namespace drm {

// configurations per CTA per tile == threads per CTA: a template parameter T in {128, 64, 32}, the largest
// whose shared-memory footprint fits (big trees such as the 21-link Allegro hand need the smaller tiles)
constexpr int BWD_MAX_GRID = 148 * 8;         // upper bound of persistent CTAs (workspace sizing)
constexpr size_t BWD_SMEM_BUDGET = 200 * 1024;
constexpr float GRAVITY_B = 9.81f;

// ---------------------------------------------------------------------------------------------
// block-level sum of NV per-thread values into the CTA accumulator row `acc_row` (entries map(j))
// ---------------------------------------------------------------------------------------------
template <int NV, int T, typename Map>
__device__ __forceinline__ void block_accumulate(float* scratch, float* acc_row, const float (&vals)[NV], bool active,
                                                 Map map) {
    constexpr int SCR_LD = T + 1;             // padded leading dimension of the reduction scratch
    const int tid = threadIdx.x;
#pragma unroll
    for (int j = 0; j < NV; ++j) scratch[j * SCR_LD + tid] = active ? vals[j] : 0.f;
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < NV; j += T / 32) {
        float x = 0.f;
#pragma unroll
        for (int c = 0; c < T / 32; ++c) x += scratch[j * SCR_LD + lane + 32 * c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) acc_row[map(j)] += x;
    }
    __syncthreads();
}

__device__ __forceinline__ void tile_load_or_zero(float* dst, const float* src, int nfloats, int total, bool vec_ok) {
    if (src != nullptr) {
        coop_copy(dst, src, nfloats, vec_ok);
    } else {
        for (int i = threadIdx.x; i < total; i += blockDim.x) dst[i] = 0.f;
    }
}

// dL/dR of the xyzw quaternion (branch structure of spatial_vector_algebra.py:116-135), exact derivative
// (the reference's autograd treats the 0.5/sqrt(t) factor as a constant -- SURVEY.md quirk 5).
__device__ __forceinline__ M3 quat_backward(const M3& R, float4 g) {
    M3 b = zero3();
    const float tr = (R.a00 + R.a11) + R.a22;
    const float t4 = tr + 1.0f;
    if (t4 > 1.0f) {
        const float y = 0.5f * rsqrt_nr(t4);
        const float u0 = R.a21 - R.a12, u1 = R.a02 - R.a20, u2 = R.a10 - R.a01;
        const float tb = g.w * y - (y / (2.f * t4)) * (g.x * u0 + g.y * u1 + g.z * u2 + g.w * t4);
        b.a00 = b.a11 = b.a22 = tb;
        b.a21 = g.x * y; b.a12 = -g.x * y; b.a02 = g.y * y; b.a20 = -g.y * y; b.a10 = g.z * y; b.a01 = -g.z * y;
    } else if (R.a22 > fmaxf(R.a00, R.a11)) {
        const float t = R.a22 - (R.a00 + R.a11) + 1.0f;
        const float y = 0.5f * rsqrt_nr(t);
        const float u0 = R.a20 + R.a02, u1 = R.a12 + R.a21, u3 = R.a10 - R.a01;
        const float tb = g.z * y - (y / (2.f * t)) * (g.x * u0 + g.y * u1 + g.z * t + g.w * u3);
        b.a22 = tb; b.a00 = -tb; b.a11 = -tb;
        b.a20 = b.a02 = g.x * y; b.a12 = b.a21 = g.y * y; b.a10 = g.w * y; b.a01 = -g.w * y;
    } else if (R.a11 > R.a00) {
        const float t = R.a11 - (R.a22 + R.a00) + 1.0f;
        const float y = 0.5f * rsqrt_nr(t);
        const float u0 = R.a01 + R.a10, u2 = R.a12 + R.a21, u3 = R.a02 - R.a20;
        const float tb = g.y * y - (y / (2.f * t)) * (g.x * u0 + g.y * t + g.z * u2 + g.w * u3);
        b.a11 = tb; b.a22 = -tb; b.a00 = -tb;
        b.a01 = b.a10 = g.x * y; b.a12 = b.a21 = g.z * y; b.a02 = g.w * y; b.a20 = -g.w * y;
    } else {
        const float t = R.a00 - (R.a11 + R.a22) + 1.0f;
        const float y = 0.5f * rsqrt_nr(t);
        const float u1 = R.a01 + R.a10, u2 = R.a20 + R.a02, u3 = R.a21 - R.a12;
        const float tb = g.x * y - (y / (2.f * t)) * (g.x * t + g.y * u1 + g.z * u2 + g.w * u3);
        b.a00 = tb; b.a11 = -tb; b.a22 = -tb;
        b.a01 = b.a10 = g.y * y; b.a20 = b.a02 = g.z * y; b.a21 = g.w * y; b.a12 = -g.w * y;
    }
    return b;
}

// =============================================================================================
// FK + Jacobian backward
// =============================================================================================
struct FkBwdArgs {
    const float* __restrict__ table;
    const float* __restrict__ q;
    const float* __restrict__ g_pos;
    const float* __restrict__ g_quat;
    const float* __restrict__ g_jl;
    const float* __restrict__ g_ja;
    float* __restrict__ q_grad;
    float* __restrict__ partials;        // [gridDim.x, n_links * 28] or null
    int64_t batch;
    int32_t n_links;
    int32_t vec_ok;
};

struct FkBwdSmem {
    int q, qg, gpos, gquat, gjl, gja, table, link, scratch, acc, total_floats;
    __host__ __device__ FkBwdSmem(int tile, int n, int len) {
        int o = 0;
        gquat = o; o += tile * 4;
        q = o; o += tile * n;
        qg = o; o += tile * n;
        gpos = o; o += tile * 3;
        gjl = o; o += tile * 3 * n;
        gja = o; o += tile * 3 * n;
        table = o; o += len * 12;
        link = o; o += len * 14 * tile;                  // per path link: R~ (9), p (3), cos, sin -- slot-major
        scratch = o; o += 12 * (tile + 1);
        acc = o; o += len * 12;                          // canonical (F~, r~) gradient per path link
        total_floats = o;
    }
};

template <bool NEED_TABLE, int T>
__global__ void __launch_bounds__(T)
fk_jacobian_backward_kernel(const __grid_constant__ PathProgram prog, const FkBwdArgs args) {
    extern __shared__ __align__(128) float smem[];
    const int n = prog.n_dofs, len = prog.len;
    const FkBwdSmem L(T, n, len);
    float* s_q = smem + L.q;
    float* s_qg = smem + L.qg;
    float* s_gpos = smem + L.gpos;
    float* s_gquat = smem + L.gquat;
    float* s_gjl = smem + L.gjl;
    float* s_gja = smem + L.gja;
    float* s_tab = smem + L.table;
    float* s_link = smem + L.link;
    float* s_scr = smem + L.scratch;
    float* s_acc = smem + L.acc;
    const int tid = threadIdx.x;
    const bool vec_ok = args.vec_ok;

    for (int i = tid; i < len * 12; i += T) {
        const int k = i / 12, e = i - k * 12;
        int src;
        const float sg = canon_map(e, prog.paxis[k], prog.axis[k], src);
        s_tab[i] = sg * __ldg(args.table + (int)prog.link[k] * DRMB200_TABLE_STRIDE + src);
        if (NEED_TABLE) s_acc[i] = 0.f;
    }

    const int64_t n_tiles = (args.batch + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t start = tile * T;
        const int valid = (int)min((int64_t)T, args.batch - start);
        __syncthreads();                                    // previous tile fully consumed
        coop_copy(s_q, args.q + start * n, valid * n, vec_ok);
        tile_load_or_zero(s_gpos, args.g_pos ? args.g_pos + start * 3 : nullptr, valid * 3, T * 3, vec_ok);
        tile_load_or_zero(s_gquat, args.g_quat ? args.g_quat + start * 4 : nullptr, valid * 4, T * 4, vec_ok);
        tile_load_or_zero(s_gjl, args.g_jl ? args.g_jl + start * 3 * n : nullptr, valid * 3 * n, T * 3 * n, vec_ok);
        tile_load_or_zero(s_gja, args.g_ja ? args.g_ja + start * 3 * n : nullptr, valid * 3 * n, T * 3 * n, vec_ok);
        for (int i = tid; i < T * n; i += T) s_qg[i] = 0.f;
        __syncthreads();

        const bool active = tid < valid;
        const float* qrow = s_q + tid * n;
        const float* gl = s_gjl + tid * 3 * n;
        const float* ga = s_gja + tid * 3 * n;
        float* lk = s_link + tid;

        // ---- forward recompute along the path; keep R~_k, p_k, (cos, sin) in smem ----------------
        M3 R = identity3();
        V3 p = v3(0.f, 0.f, 0.f);
        V3 pbar = v3(s_gpos[tid * 3], s_gpos[tid * 3 + 1], s_gpos[tid * 3 + 2]);
        for (int k = 0; k < len; ++k) {
            M3 F; V3 r;
            load_Fr(s_tab + k * 12, F, r);
            p = mul_add(R, r, p);
            R = mul(R, F);
            const int c = prog.dof[k];
            float cs = 1.f, sn = 0.f;
            if (c >= 0) {
                sincos_pi2(qrow[c], sn, cs);
                const V3 g = v3(gl[c], gl[n + c], gl[2 * n + c]);
                pbar = cross_add(g, col2(R), pbar);         // adjoint of p_ee: J_lin = z x (p_ee - p_i)
                rotate_z(R, cs, sn);
            }
            float* s = lk + k * 14 * T;
            stm(s, T, R);
            stv(s + 9 * T, T, p);
            s[12 * T] = cs; s[13 * T] = sn;
        }
        const V3 p_ee = p;
        const float4 gq = reinterpret_cast<const float4*>(s_gquat)[tid];
        M3 Rbar = permute_cols_adjoint(quat_backward(unpermute_cols(R, prog.ee_axis), gq), prog.ee_axis);

        // ---- reverse sweep ee -> root ------------------------------------------------------------
        M3 Rk = R;                                          // R~_k of the link being processed
        for (int k = len - 1; k >= 0; --k) {
            const float* s = lk + k * 14 * T;
            const V3 pk = ldv(s + 9 * T, T);
            const float cs = s[12 * T], sn = s[13 * T];
            const M3 RP = (k > 0) ? ldm(lk + (k - 1) * 14 * T, T) : identity3();
            M3 F; V3 r;
            load_Fr(s_tab + k * 12, F, r);
            const int c = prog.dof[k];
            M3 M = F;
            if (c >= 0) {
                const V3 z = col2(Rk);
                const V3 g = v3(gl[c], gl[n + c], gl[2 * n + c]);
                const V3 h = v3(ga[c], ga[n + c], ga[2 * n + c]);
                const V3 zbar = cross_add(p_ee - pk, g, h);   // d x G_l + G_a
                Rbar.a02 += zbar.x; Rbar.a12 += zbar.y; Rbar.a22 += zbar.z;    // z = R~_k e_z
                pbar = pbar - cross(g, z);                    // adjoint of p_k
                rotate_z(M, cs, sn);
            }
            const M3 Mbar = mulTN(RP, Rbar);
            const V3 rbar = mulT(RP, pbar);
            M3 Rbar_P = mulNT(Rbar, M);
            add_outer(Rbar_P, pbar, r);
            if (c >= 0) s_qg[tid * n + c] = theta_grad_z(Mbar, M);
            if (NEED_TABLE) {
                M3 Fbar = Mbar;
                if (c >= 0) rotate_z(Fbar, cs, -sn);          // Mbar Rz^T
                float vals[12];
                m3_to_array(Fbar, vals);
                vals[9] = rbar.x; vals[10] = rbar.y; vals[11] = rbar.z;
                block_accumulate<12, T>(s_scr, s_acc + k * 12, vals, active, [](int j) { return j; });
            }
            Rbar = Rbar_P;
            Rk = RP;
        }
        if (args.q_grad != nullptr) {
            __syncthreads();
            coop_copy(args.q_grad + start * n, s_qg, valid * n, vec_ok);
        }
    }
    if (NEED_TABLE) {
        __syncthreads();
        float* out = args.partials + (size_t)blockIdx.x * args.n_links * DRMB200_TABLE_STRIDE;
        for (int i = tid; i < args.n_links * DRMB200_TABLE_STRIDE; i += T) out[i] = 0.f;
        __syncthreads();
        for (int i = tid; i < len * 12; i += T) {            // canonical -> natural entries (bijection per row)
            const int k = i / 12, e = i - k * 12;
            int src;
            const float sg = canon_map(e, prog.paxis[k], prog.axis[k], src);
            out[(int)prog.link[k] * DRMB200_TABLE_STRIDE + src] = sg * s_acc[i];
        }
    }
}

// sums the per-CTA partial tables in fixed order and ADDS them to table_grad
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int n_blocks, int n_entries,
                                       float* __restrict__ table_grad) {
    // one warp per table entry: lanes stride over the CTA partials (fixed assignment -> fixed summation order),
    // then a shuffle tree.  (One thread per entry looping over ~1200 partials took 35 us, longer than the
    // single-sweep backward kernel it follows.)
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (e >= n_entries) return;
    float s = 0.f;
    for (int b = lane; b < n_blocks; b += 32) s += partials[(size_t)b * n_entries + e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) table_grad[e] += s;
}

int64_t table_grad_workspace_bytes(const drmb200_topology_t* topo, int64_t batch) {
    if (topo == nullptr || topo->n_links < 1 || topo->n_links > DRMB200_MAX_LINKS) return 0;
    int64_t tiles = (batch + 31) / 32;            // smallest tile the launchers may pick
    if (tiles < 1) tiles = 1;
    const int64_t grid = tiles < BWD_MAX_GRID ? tiles : BWD_MAX_GRID;
    return grid * topo->n_links * DRMB200_TABLE_STRIDE * (int64_t)sizeof(float);
}

template <typename Kern>
static int persistent_grid(Kern kern, int block, size_t smem_bytes, int64_t tiles, int* grid_out, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", what, smem_bytes, cudaGetErrorString(e)); return DRMB200_ECUDA; }
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem_bytes);
    if (e != cudaSuccess || per_sm < 1) { set_error("%s: kernel does not fit on an SM (%zu B smem)", what, smem_bytes); return DRMB200_ECUDA; }
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > BWD_MAX_GRID) grid = BWD_MAX_GRID;
    if (grid > tiles) grid = tiles;
    *grid_out = (int)grid;
    return DRMB200_OK;
}

static int launch_reduce(const float* partials, int grid, const drmb200_topology_t* topo, float* table_grad,
                         cudaStream_t stream) {
    const int entries = topo->n_links * DRMB200_TABLE_STRIDE;
    reduce_partials_kernel<<<(entries * 32 + 255) / 256, 256, 0, stream>>>(partials, grid, entries, table_grad);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("reduce launch: %s", cudaGetErrorString(e)); return DRMB200_ECUDA; }
    count_launch();
    return DRMB200_OK;
}

int fk_jacobian_backward_device(const drmb200_topology_t* topo, int32_t ee_link, const float* table, const float* q,
                                int64_t batch, const float* g_pos, const float* g_quat, const float* g_jl,
                                const float* g_ja, float* q_grad, float* table_grad, void* workspace,
                                cudaStream_t stream) {
    PathProgram prog;
    int rc = build_path_program(topo, ee_link, &prog);
    if (rc != DRMB200_OK) return rc;
    if (batch < 0) { set_error("batch=%lld < 0", (long long)batch); return DRMB200_EINVAL; }
    if (batch == 0 || (q_grad == nullptr && table_grad == nullptr)) return DRMB200_OK;
    if (table == nullptr || q == nullptr) { set_error("table / q is null"); return DRMB200_EINVAL; }
    if (table_grad != nullptr && workspace == nullptr) { set_error("table_grad requested without workspace"); return DRMB200_EINVAL; }

    FkBwdArgs args;
    args.table = table; args.q = q; args.g_pos = g_pos; args.g_quat = g_quat; args.g_jl = g_jl; args.g_ja = g_ja;
    args.q_grad = q_grad; args.partials = static_cast<float*>(workspace); args.batch = batch;
    args.n_links = topo->n_links;
    auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    args.vec_ok = (al16(q) && al16(g_pos) && al16(g_quat) && al16(g_jl) && al16(g_ja) && al16(q_grad)) ? 1 : 0;

    int tile = 32, best_warps = 0;           // the tile that keeps the most warps resident per SM
    for (int t = 128; t >= 32; t >>= 1) {
        const size_t b = (size_t)FkBwdSmem(t, prog.n_dofs, prog.len).total_floats * sizeof(float) + 1024;
        const int warps = b > 227 * 1024 ? 0 : (int)((227 * 1024) / b) * (t / 32);
        if (warps > best_warps) { best_warps = warps; tile = t; }
    }
    const size_t smem_bytes = (size_t)FkBwdSmem(tile, prog.n_dofs, prog.len).total_floats * sizeof(float);
    if (smem_bytes > 227 * 1024) { set_error("fk backward needs %zu B of shared memory per CTA (> 227 KB)", smem_bytes); return DRMB200_ELIMIT; }
    const int64_t tiles = (batch + tile - 1) / tile;
    int grid = 0;
    const bool need_table = table_grad != nullptr;
#define DRM_LAUNCH_FKB(NT, TT)                                                                                  \
    do {                                                                                                        \
        rc = persistent_grid(fk_jacobian_backward_kernel<NT, TT>, TT, smem_bytes, tiles, &grid, "fk backward"); \
        if (rc != DRMB200_OK) return rc;                                                                        \
        fk_jacobian_backward_kernel<NT, TT><<<grid, TT, smem_bytes, stream>>>(prog, args);                      \
    } while (0)
    if (need_table) { if (tile == 128) DRM_LAUNCH_FKB(true, 128); else if (tile == 64) DRM_LAUNCH_FKB(true, 64); else DRM_LAUNCH_FKB(true, 32); }
    else            { if (tile == 128) DRM_LAUNCH_FKB(false, 128); else if (tile == 64) DRM_LAUNCH_FKB(false, 64); else DRM_LAUNCH_FKB(false, 32); }
#undef DRM_LAUNCH_FKB
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("fk backward launch: %s", cudaGetErrorString(e)); return DRMB200_ECUDA; }
    count_launch();
    return need_table ? launch_reduce(args.partials, grid, topo, table_grad, stream) : DRMB200_OK;
}

// =============================================================================================
// RNEA backward
// =============================================================================================
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

// per-link per-thread state, slot-major: w v al a (12) | f n -> mu lambda (6) | cos sin (2)
constexpr int LSTATE = 20;

struct RneaBwdSmem {
    int q, qd, qdd, g, qg, qdg, qddg, table, link, slots, scratch, acc, total_floats;
    __host__ __device__ RneaBwdSmem(int tile, int n, int n_links, int n_slots) {
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
        slots = o; o += n_slots * 12 * tile;
        scratch = o; o += 25 * (tile + 1);
        acc = o; o += n_links * DRMB200_TABLE_STRIDE;
        total_floats = o;
    }
};

template <bool NEED_TABLE, int T>
__global__ void __launch_bounds__(T)
rnea_backward_kernel(const __grid_constant__ TreeProgram prog, const RneaBwdArgs args) {
    extern __shared__ __align__(128) float smem[];
    const int n = prog.n_dofs, N = prog.n_links;
    const RneaBwdSmem L(T, n, N, prog.n_slots);
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

        // ---- forward recompute, pass A: motion state + body wrench --------------------------------
        for (int i = 1; i < N; ++i) {
            const LinkRow C = load_row(s_tab + i * DRMB200_TABLE_STRIDE);
            const int P = prog.parent[i];
            V3 wp = zero, vp = zero, alp = zero, ap = a_root;
            if (P > 0) {
                const float* sp = lk + P * LSTATE * T;
                wp = ldv(sp, T); vp = ldv(sp + 3 * T, T); alp = ldv(sp + 6 * T, T); ap = ldv(sp + 9 * T, T);
            }
            M3 M = C.F;
            const int c = prog.dof[i];
            float cs = 1.f, sn = 0.f, qd_k = 0.f, qdd_k = 0.f;
            if (c >= 0) {
                qd_k = qdrow[c]; qdd_k = qddrow[c];
                sincos_pi2(qrow[c], sn, cs);
                rotate_z(M, cs, sn);
            }
            V3 w = mulT(M, wp); w.z += qd_k;
            const V3 v = mulT(M, cross_add(wp, C.r, vp));
            V3 al = mulT(M, alp) + cross_z(w, qd_k); al.z += qdd_k;
            const V3 a = mulT(M, cross_add(alp, C.r, ap)) + cross_z(v, qd_k);
            const V3 hl_a = C.m * a - cross(C.mc, al);
            const V3 ha_a = mul_add(C.Io, al, cross(C.mc, a));
            const V3 hl_v = C.m * v - cross(C.mc, w);
            const V3 ha_v = mul_add(C.Io, w, cross(C.mc, v));
            const V3 f = cross_add(w, hl_v, hl_a);
            const V3 nn = cross_add(w, ha_v, cross_add(v, hl_v, ha_a));
            float* s = lk + i * LSTATE * T;
            stv(s, T, w); stv(s + 3 * T, T, v); stv(s + 6 * T, T, al); stv(s + 9 * T, T, a);
            stv(s + 12 * T, T, f); stv(s + 15 * T, T, nn);
            s[18 * T] = cs; s[19 * T] = sn;
        }
        // ---- forward recompute, pass B: accumulate wrenches leaves -> root -----------------------
        for (int i = N - 1; i >= 1; --i) {
            const int P = prog.parent[i];
            if (P <= 0) continue;
            const float* s = lk + i * LSTATE * T;
            M3 F; V3 r;
            load_Fr(s_tab + i * DRMB200_TABLE_STRIDE, F, r);
            const float cs = s[18 * T], sn = s[19 * T];
            const V3 fp = mul(F, rotz(ldv(s + 12 * T, T), cs, sn));
            const V3 np = cross_add(r, fp, mul(F, rotz(ldv(s + 15 * T, T), cs, sn)));
            float* sp = lk + P * LSTATE * T;
            stv(sp + 12 * T, T, ldv(sp + 12 * T, T) + fp);
            stv(sp + 15 * T, T, ldv(sp + 15 * T, T) + np);
        }

        // ---- adjoint pass 1, root -> leaves: lambda = n-bar, mu = f-bar ----------------------------
        for (int i = 1; i < N; ++i) {
            float* s = lk + i * LSTATE * T;
            const float* row = s_tab + i * DRMB200_TABLE_STRIDE;
            M3 M; V3 r;
            load_Fr(row, M, r);
            const int P = prog.parent[i];
            V3 lamP = zero, muP = zero;
            if (P > 0) { const float* sp = lk + P * LSTATE * T; muP = ldv(sp + 12 * T, T); lamP = ldv(sp + 15 * T, T); }
            const int c = prog.dof[i];
            const float cs = s[18 * T], sn = s[19 * T];
            float gk = 0.f;
            if (c >= 0) { rotate_z(M, cs, sn); gk = grow[c]; }
            const V3 f = ldv(s + 12 * T, T), nn = ldv(s + 15 * T, T);       // accumulated wrenches
            const V3 u = cross_add(lamP, r, muP);
            V3 lam = mulT(M, lamP); lam.z += gk;                            // tau_k = n_i . e_z
            const V3 mu = mulT(M, u);
            stv(s + 12 * T, T, mu);                                          // f slot -> mu
            stv(s + 15 * T, T, lam);                                         // n slot -> lambda
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
        V3 c_wb = zero, c_vb = zero, c_alb = zero, c_ab = zero;     // carry into link i from child i+1
        for (int i = N - 1; i >= 1; --i) {
            const float* s = lk + i * LSTATE * T;
            const LinkRow C = load_row(s_tab + i * DRMB200_TABLE_STRIDE);
            const int P = prog.parent[i];
            const int c = prog.dof[i];
            const float cs = s[18 * T], sn = s[19 * T];
            M3 M = C.F;
            float qd_k = 0.f;
            if (c >= 0) { rotate_z(M, cs, sn); qd_k = qdrow[c]; }
            const V3 w = ldv(s, T), v = ldv(s + 3 * T, T), al = ldv(s + 6 * T, T), a = ldv(s + 9 * T, T);
            const V3 mu = ldv(s + 12 * T, T), lam = ldv(s + 15 * T, T);
            V3 wp = zero, vp = zero, alp = zero, ap = a_root;
            if (P > 0) {
                const float* sp = lk + P * LSTATE * T;
                wp = ldv(sp, T); vp = ldv(sp + 3 * T, T); alp = ldv(sp + 6 * T, T); ap = ldv(sp + 9 * T, T);
            }
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
        scratch = o; o += 14 * (tile + 1);
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

    if ((flags & DRMB200_INERTIAL_GRADS_ONLY) && table_grad != nullptr) {
        if (q_grad != nullptr || qd_grad != nullptr || qdd_grad != nullptr) {
            set_error("DRMB200_INERTIAL_GRADS_ONLY cannot be combined with input gradients");
            return DRMB200_EINVAL;
        }
        RneaBwdArgs ia;
        ia.table = table; ia.q = q; ia.qd = qd; ia.qdd = qdd; ia.g_tau = g_tau;
        ia.q_grad = ia.qd_grad = ia.qdd_grad = nullptr;
        ia.partials = static_cast<float*>(workspace); ia.batch = batch; ia.flags = flags;
        auto al16i = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
        ia.vec_ok = (al16i(q) && al16i(qd) && al16i(qdd) && al16i(g_tau)) ? 1 : 0;
        constexpr int TI = 128;
        const size_t sb = (size_t)RneaInertialSmem(TI, prog.n_dofs, prog.n_links, prog.n_slots).total_floats * sizeof(float);
        if (sb > 227 * 1024) { set_error("rnea inertial backward needs %zu B of shared memory (> 227 KB)", sb); return DRMB200_ELIMIT; }
        int g = 0;
        rc = persistent_grid(rnea_backward_inertial_kernel<TI>, TI, sb, (batch + TI - 1) / TI, &g, "rnea inertial backward");
        if (rc != DRMB200_OK) return rc;
        rnea_backward_inertial_kernel<TI><<<g, TI, sb, stream>>>(prog, ia);
        cudaError_t ei = cudaGetLastError();
        if (ei != cudaSuccess) { set_error("rnea inertial backward launch: %s", cudaGetErrorString(ei)); return DRMB200_ECUDA; }
        count_launch();
        return launch_reduce(ia.partials, g, topo, table_grad, stream);
    }

    RneaBwdArgs args;
    args.table = table; args.q = q; args.qd = qd; args.qdd = qdd; args.g_tau = g_tau;
    args.q_grad = q_grad; args.qd_grad = qd_grad; args.qdd_grad = qdd_grad;
    args.partials = static_cast<float*>(workspace); args.batch = batch; args.flags = flags;
    auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    args.vec_ok = (al16(q) && al16(qd) && al16(qdd) && al16(g_tau) && al16(q_grad) && al16(qd_grad) && al16(qdd_grad)) ? 1 : 0;

    // shared memory (20 floats per link per configuration) is the occupancy limiter: pick the tile that keeps the
    // most warps resident per SM (Kuka: 128 -> 1 CTA = 4 warps, 64 -> 3 CTAs = 6 warps), larger tile on ties
    int tile = 32, best_warps = 0;
    for (int t = 128; t >= 32; t >>= 1) {
        const size_t b = (size_t)RneaBwdSmem(t, prog.n_dofs, prog.n_links, prog.n_slots).total_floats * sizeof(float) + 1024;
        const int warps = b > 227 * 1024 ? 0 : (int)((227 * 1024) / b) * (t / 32);
        if (warps > best_warps) { best_warps = warps; tile = t; }
    }
    const size_t smem_bytes = (size_t)RneaBwdSmem(tile, prog.n_dofs, prog.n_links, prog.n_slots).total_floats * sizeof(float);
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
