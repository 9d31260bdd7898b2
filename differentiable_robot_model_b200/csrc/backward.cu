// backward.cu -- analytic adjoint kernels (placeholder until implemented)
#include "drm_common.cuh"
namespace drm {
int64_t table_grad_workspace_bytes(const drmb200_topology_t*, int64_t) { return 0; }
int fk_jacobian_backward_device(const drmb200_topology_t*, int32_t, const float*, const float*, int64_t,
                                const float*, const float*, const float*, const float*, float*, float*, void*,
                                cudaStream_t) { set_error("fk_jacobian_backward: not implemented"); return DRMB200_EINVAL; }
int inverse_dynamics_backward_device(const drmb200_topology_t*, const float*, const float*, const float*,
                                     const float*, int64_t, uint32_t, const float*, float*, float*, float*,
                                     float*, void*, cudaStream_t) { set_error("inverse_dynamics_backward: not implemented"); return DRMB200_EINVAL; }
}
