// walk_masked_np1.cu - the 64-target variants of the masked group walk (one slice pair per warp, two targets per lane):
// half the register state and 16-byte chain entries, hence more resident warps per SM than the 128-target kernel.
#include "walk_masked_kernel.cuh"

namespace hbt
{

void launch_walk_masked_np1(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, int blocks)
{
  if (blocks <= 6) launch_masked_variant<6, 1>(a, cfg, stream);
  else if (blocks == 7) launch_masked_variant<7, 1>(a, cfg, stream);
  else if (blocks == 8) launch_masked_variant<8, 1>(a, cfg, stream);
  else launch_masked_variant<9, 1>(a, cfg, stream);
}

} // namespace hbt
