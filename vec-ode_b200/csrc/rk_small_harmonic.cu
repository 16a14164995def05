#include "rk_small_launch.cuh"
int32_t launch_small_harmonic(const SmallLaunch& L) { return launch_family<RhsF<VO_RHS_HARMONIC2D, 2>>(L); }
