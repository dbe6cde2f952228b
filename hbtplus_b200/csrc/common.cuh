// common.cuh - error handling, device arena and shared PODs of libhbtunbind (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "hbt_unbind.h"
#include "tree_core.cuh"

namespace hbt
{

struct CudaError
{
  int code; // HBTU_ERR_*
  std::string what;
};

#define HBT_CUDA(call)                                                                                              \
  do                                                                                                                \
  {                                                                                                                 \
    cudaError_t e_ = (call);                                                                                        \
    if (e_ != cudaSuccess)                                                                                          \
      throw hbt::CudaError{e_ == cudaErrorMemoryAllocation ? HBTU_ERR_NOMEM : HBTU_ERR_CUDA,                        \
                           std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +             \
                               std::to_string(__LINE__) + ")"};                                                     \
  } while (0)

#define HBT_CHECK_LAUNCH() HBT_CUDA(cudaGetLastError())

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Grow-only device pool with bump allocation; reset() at the start of every round.  All transient
// per-round arrays (keys, sorted copies, node arrays, CUB temp storage) live here, so a 1.8e8-particle
// batch does a handful of cudaMalloc calls in total instead of dozens per round.
class Arena
{
  char *base_ = nullptr;
  int64_t cap_ = 0, used_ = 0, high_ = 0;
  std::vector<char *> retired_; // older, smaller blocks kept alive until the round finishes

public:
  ~Arena() { release(); }
  void release()
  {
    if (base_) cudaFree(base_);
    for (char *p : retired_) cudaFree(p);
    retired_.clear();
    base_ = nullptr;
    cap_ = used_ = 0;
  }
  void reserve(int64_t bytes)
  {
    if (bytes <= cap_) return;
    // only legal when nothing of the current round lives in the pool
    if (used_ != 0) throw CudaError{HBTU_ERR_NOMEM, "arena grown while in use"};
    if (base_) cudaFree(base_);
    base_ = nullptr;
    cap_ = 0;
    HBT_CUDA(cudaMalloc(&base_, (size_t)bytes));
    cap_ = bytes;
  }
  void reset()
  {
    used_ = 0;
    for (char *p : retired_) cudaFree(p);
    retired_.clear();
  }
  template <class T>
  T *alloc(int64_t count)
  {
    int64_t bytes = align_up((count > 0 ? count : 1) * (int64_t)sizeof(T), 256);
    if (used_ + bytes > cap_)
    { // overflow of the estimate: chain a new block (keeps earlier pointers valid)
      int64_t ncap = cap_ * 3 / 2 + bytes;
      char *nb = nullptr;
      HBT_CUDA(cudaMalloc(&nb, (size_t)ncap));
      if (base_) retired_.push_back(base_);
      base_ = nb;
      cap_ = ncap;
      used_ = 0;
    }
    T *p = reinterpret_cast<T *>(base_ + used_);
    used_ += bytes;
    if (used_ > high_) high_ = used_;
    return p;
  }
  int64_t capacity() const { return cap_; }
  int64_t high_water() const { return high_; }
};

// order-preserving encodings used by the bbox atomics
__device__ __forceinline__ void atomic_min_float(uint32_t *addr, float v) { atomicMin(addr, float_to_ordered(v)); }
__device__ __forceinline__ void atomic_max_float(uint32_t *addr, float v) { atomicMax(addr, float_to_ordered(v)); }

// NEAREST() of src/config_parser.h:142 in fp32 (HBTReal)
__device__ __forceinline__ float nearest_f(float x, float box, float half)
{
  return x > half ? x - box : (x < -half ? x + box : x);
}
__device__ __forceinline__ double nearest_d(double x, double box, double half)
{
  return x > half ? x - box : (x < -half ? x + box : x);
}

// Constants of one context in the storage types of the reference's V32 build (HBTReal=float).
struct DevConfig
{
  float box_size, box_half, softening, theta2, resolution, resolution_half, G, bound_mass_precision, relax_factor;
  float scale_factor, hz;
  int periodic, min_num_part, snapshot_index;
  int no_stripping;   // HBTU_FLAG_NO_STRIPPING of the staged batch (-DNO_STRIPPING builds of the reference)
  int thermal_energy; // HBTU_FLAG_THERMAL_ENERGY: vel.w is Particle_t::InternalEnergy and is added to E in full evaluations
  int64_t max_sample;
};

} // namespace hbt
