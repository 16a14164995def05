#include "rk_small_launch.cuh"
int32_t launch_small_lorenz(const SmallLaunch& L) { return launch_family<RhsF<VO_RHS_LORENZ63, 3>>(L); }
