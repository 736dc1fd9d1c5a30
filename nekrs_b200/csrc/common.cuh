// common.cuh -- shared types and helpers for the nrsb200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/nrsb200.h"

namespace nrsb {

// nrssys.hpp: dfloat=double, pfloat=float, dlong=int, hlong=long long
using dfloat = double;
using pfloat = float;
using dlong = int32_t;
using hlong = int64_t;

constexpr int kBlockSize = 256;   // BLOCKSIZE (nrssys.hpp)
constexpr int kMaxNq = 14;        // N <= 13 (determineMGLevels.cpp level tables go to 15; FDM needs Nq+2)
constexpr int kNumSMs = 148;      // B200

void set_last_error(const std::string& s);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define NRSB_CUDA(call)                                                      \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) return ::nrsb::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define NRSB_CHECK_LAUNCH() NRSB_CUDA(cudaGetLastError())

#define NRSB_REQUIRE(cond, msg)                                              \
  do {                                                                       \
    if (!(cond)) {                                                           \
      ::nrsb::set_last_error(std::string(msg) + " [" #cond "]");             \
      return NRSB_ERR_INVALID;                                               \
    }                                                                        \
  } while (0)

// D-matrix passed BY VALUE: kernel parameters live in the constant bank, and every
// index below is a compile-time constant after unrolling, so each D entry is an
// immediate constant-bank operand of the FMA (no register, no shared-memory load).
template <typename T, int Nq>
struct DMat {
  T v[Nq * Nq];  // row-major D[i][m] = l_m'(r_i)
};

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may become resident while the
// previous kernel of the stream is still running (once every CTA of that kernel has executed pdl_trigger(),
// or exited).  Everything before pdl_wait() must not touch data the previous kernel writes; pdl_wait()
// returns when the previous kernel has completed and its stores are visible.  Launch latency, block
// scheduling and the index-table loads of the dependent kernel are thereby hidden behind the producer.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled(bool producer);  // see capi.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool on, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, Args&&... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// launch_pdl: the persistent axhelm kernel (its prologue and geometric-factor prefetch overlap the tail of the
// previous kernel); launch_pdl_consumer: gather-scatter style kernels behind axhelm (off by default, see capi.cu)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args)
{
  return launch_pdl_if(pdl_enabled(true), kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_consumer(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                       cudaStream_t stream, Args&&... args)
{
  return launch_pdl_if(pdl_enabled(false), kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

template <typename T>
__device__ __forceinline__ T ldg_stream(const T* p)
{
  return __ldcs(p);  // ld.global.cs: streaming, evict-first (geometric factors are read once)
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace nrsb
