// Host-side check (no GPU needed): the select-based row permutation used by the kernels' table staging
// (canonical_row, drm_common.cuh) must equal the element-wise definition canon_map() for every pair of axis codes.
#include <cstdio>
#include "drm_common.cuh"

int main() {
    int bad = 0;
    for (int cp = -3; cp <= 3; ++cp)
        for (int ci = -3; ci <= 3; ++ci) {
            float x[DRMB200_TABLE_STRIDE], y[DRMB200_TABLE_STRIDE];
            for (int e = 0; e < DRMB200_TABLE_STRIDE; ++e) x[e] = 1.0f + 0.37f * e + 0.011f * e * e;   // all distinct
            drm::canonical_row(x, cp, ci, y);
            for (int e = 0; e < DRMB200_TABLE_STRIDE; ++e) {
                int src;
                const float sg = drm::canon_map(e, cp, ci, src);
                if (y[e] != sg * x[src]) { ++bad; std::printf("mismatch cp=%d ci=%d e=%d\n", cp, ci, e); }
            }
        }
    std::printf("%s\n", bad ? "FAIL" : "canonical_row == canon_map for all 49 axis-code pairs");
    return bad ? 1 : 0;
}
