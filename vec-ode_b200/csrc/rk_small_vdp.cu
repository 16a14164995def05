#include "rk_small_launch.cuh"
int32_t launch_small_vdp(const SmallLaunch& L) { return launch_family<RhsF<VO_RHS_VDP, 2>>(L); }
