// warp_prims.cuh — the handful of warp-level primitives the per-graph device code is written against.
//
// The graph / transfrag / flow stages (asm_device.cuh) give one WARP to one splice graph: control flow is
// warp-uniform (every lane holds the same scalars), the inner loops that dominate — bases of a region, transfrags of a
// node, words of a bit pattern — are strided over the lanes, and order-dependent float sums are taken in index order
// from ballots.  On the device the primitives below are the hardware ones; compiled for the host (tools/asm_hostcheck.cu,
// a development aid that steps through the same code on a CPU-only box) a "warp" has ONE lane and they degenerate to
// the identity, so the very same source runs sequentially.  libstgpu only ever calls this code from kernels.
#pragma once
#include <stdint.h>

#define STG_HD __host__ __device__ __forceinline__
#define STG_HDN __host__ __device__

#ifdef __CUDA_ARCH__
#define WP_N 32
#define WP_FULL 0xffffffffu
__device__ __forceinline__ int wp_lane() { return (int)(threadIdx.x & 31u); }
__device__ __forceinline__ uint32_t wp_ballot(bool p) { return __ballot_sync(WP_FULL, p); }
template <class T> __device__ __forceinline__ T wp_shfl(T v, int src) { return __shfl_sync(WP_FULL, v, src); }
__device__ __forceinline__ void wp_sync() { __syncwarp(); }
__device__ __forceinline__ bool wp_any(bool p) { return __any_sync(WP_FULL, p); }
__device__ __forceinline__ bool wp_all(bool p) { return __all_sync(WP_FULL, p); }
__device__ __forceinline__ int wp_ffs(uint32_t m) { return __ffs((int)m); }
__device__ __forceinline__ int wp_popc(uint32_t m) { return __popc(m); }
__device__ __forceinline__ int wp_popcll(unsigned long long m) { return __popcll(m); }
__device__ __forceinline__ float wp_fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float wp_fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float wp_fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double wp_dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double wp_dmul(double a, double b) { return __dmul_rn(a, b); }
#else
#define WP_N 1
#define WP_FULL 1u
inline int wp_lane() { return 0; }
inline uint32_t wp_ballot(bool p) { return p ? 1u : 0u; }
template <class T> inline T wp_shfl(T v, int) { return v; }
inline void wp_sync() {}
inline bool wp_any(bool p) { return p; }
inline bool wp_all(bool p) { return p; }
inline int wp_ffs(uint32_t m) { return __builtin_ffs((int)m); }
inline int wp_popc(uint32_t m) { return __builtin_popcount(m); }
inline int wp_popcll(unsigned long long m) { return __builtin_popcountll(m); }
inline float wp_fadd(float a, float b) { return a + b; }
inline float wp_fsub(float a, float b) { return a - b; }
inline float wp_fmul(float a, float b) { return a * b; }
inline double wp_dadd(double a, double b) { return a + b; }
inline double wp_dmul(double a, double b) { return a * b; }
#endif
