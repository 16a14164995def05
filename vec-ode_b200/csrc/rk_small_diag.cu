#include "rk_small_launch.cuh"
int32_t launch_small_diag(const SmallLaunch& L, int d) {
    switch (d) {
        case 1: return launch_family<RhsF<VO_RHS_DIAG_LINEAR, 1>>(L);
        case 2: return launch_family<RhsF<VO_RHS_DIAG_LINEAR, 2>>(L);
        case 3: return launch_family<RhsF<VO_RHS_DIAG_LINEAR, 3>>(L);
        case 4: return launch_family<RhsF<VO_RHS_DIAG_LINEAR, 4>>(L);
    }
    return VO_ERR_UNSUPPORTED;
}
