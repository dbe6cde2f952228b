// walk_masked.cu - the masked group walk kernel (sm_100a): GravityTree_t::EvaluatePotential / BindingEnergy
// (src/gravity_tree.cpp:79-175) for segments of at least `group_min` targets.  The algorithm is in walk_masked.cuh, the
// kernel in walk_masked_kernel.cuh; this file instantiates the 128-target variants (two slice pairs per warp) and holds the
// dispatcher.  Roofline: FP32 issue.  Tensor cores are deliberately not used.
#include "walk_masked_kernel.cuh"

namespace hbt
{

void launch_walk_masked_np1(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, int blocks); // walk_masked_np1.cu

void launch_walk_masked(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls)
{
  if (a.nwarps <= 0) return;
  const int blocks = walk_tuning().masked_blocks;
  if (a.targets_per_lane == kWalkGroup2)
    launch_walk_masked_np1(a, cfg, stream, blocks);
  else if (blocks <= 4) launch_masked_variant<4, 2>(a, cfg, stream);
  else if (blocks >= 6) launch_masked_variant<6, 2>(a, cfg, stream);
  else launch_masked_variant<5, 2>(a, cfg, stream);
  HBT_CHECK_LAUNCH();
  ls.launches++;
}

} // namespace hbt
