// common.cuh -- shared helpers for the sm_100a kernels of libdbb200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dbb200.h"

#ifndef DBB_NUM_SMS
#define DBB_NUM_SMS 148   // B200: 2 dies x 74 SMs; grids are sized in multiples of this
#endif

namespace dbb {

extern thread_local char g_last_error[512];
extern uint64_t g_launch_count;

int set_cuda_error(cudaError_t e, const char* where);
// per-kernel CUDA-event timing (off by default; bench.py turns it on for the roofline figure)
int prof_begin(const char* name, cudaStream_t s);   // returns slot or -1 when profiling is off
void prof_end(int slot, cudaStream_t s);
}  // namespace dbb
#include <string>
namespace dbb {
const char* prof_label(const std::string& s);
bool prof_enabled();   // interned label (stable pointer)
int set_error(int code, const char* msg);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// call after every kernel launch
#define DBB_CHECK_LAUNCH(where)                                             \
  do {                                                                      \
    ++::dbb::g_launch_count;                                                \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) return ::dbb::set_cuda_error(e__, where);       \
  } while (0)

// launch + account + (optionally) time one kernel:  DBB_LAUNCH("name", stream, kernel<<<g, b, smem, stream>>>(args...));
#define DBB_LAUNCH(name, stream, ...)                                       \
  do {                                                                      \
    const int prof__ = ::dbb::prof_begin(name, stream);                     \
    __VA_ARGS__;                                                            \
    if (prof__ >= 0) ::dbb::prof_end(prof__, stream);                       \
    DBB_CHECK_LAUNCH(name);                                                 \
  } while (0)

#define DBB_CUDA(call)                                                      \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess) return ::dbb::set_cuda_error(e__, #call);       \
  } while (0)

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit load: read-only path, do not allocate in L1 (data is touched once)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace dbb
