// ordered_sum.cuh — order-preserving keyed float sums for one CTA (used by the commit stage of row a11 and by the very
// deep bundles of rows a8 + a9).
//
// The reference accumulates float sums (CTransfrag::abundance, CGraphnode::cov, CGroup::cov_sum / multi) with "+=" inside
// a sequential loop over the reads, so the result of every sum depends on the ORDER of its terms.  What is NOT
// order-dependent is which terms belong to which sum.  So the terms are first brought into key order by a STABLE counting
// sort (every warp counts the keys of its contiguous slice of the records, a prefix over the slices gives every
// (slice, key) its first slot, the warps scatter their slices in record order) and then one warp per key adds its run
// from first to last — the chain of dependent float adds of one key is the only serial part left.
#pragma once
#include <stdint.h>

// acc[k] = (...((acc[k] + v_i1) + v_i2) + ...) over the records i1 < i2 < ... of key k, i.e. the reference's "+=" loop
// over the records in order, for all keys at once.  One CTA; scratch: cnt [W * K] (zeroed), kbase [K + 1], sorted [n].
template <class KeyFn, class ValFn>
__device__ void cta_ordered_keyed_sum(uint32_t n, KeyFn key, ValFn val, uint32_t K, float* acc, uint32_t* cnt, uint32_t* kbase, float* sorted) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t W = blockDim.x >> 5;
  uint32_t S = (n + W - 1) / W;
  S = (S + 31u) & ~31u;                                  // slices start on a multiple of 32: whole warps step together
  const uint32_t lo = min(n, w * S), hi = min(n, lo + S);
  uint32_t* row = cnt + (size_t)w * K;
  // 1. records per key in this warp's slice
  for (uint32_t i0 = lo; i0 < hi; i0 += 32) {
    const uint32_t i = i0 + lane;
    const bool have = i < hi;
    const uint32_t k = have ? key(i) : 0xffffffffu - lane;
    const uint32_t peers = __match_any_sync(0xffffffffu, k);
    if (have && (__ffs(peers) - 1) == lane) row[k] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // 2. first slot of every (key, slice): prefix over the slices, then over the keys
  for (uint32_t k = threadIdx.x; k < K; k += blockDim.x) {
    uint32_t run = 0;
    for (uint32_t x = 0; x < W; x++) { const uint32_t c = cnt[(size_t)x * K + k]; cnt[(size_t)x * K + k] = run; run += c; }
    kbase[k + 1] = run;
  }
  __syncthreads();
  if (w == 0) {   // exclusive scan of the totals (K is small next to n)
    uint32_t carry = 0;
    for (uint32_t k0 = 0; k0 < K; k0 += 32) {
      const uint32_t k = k0 + lane;
      uint32_t x = k < K ? kbase[k + 1] : 0u;
      uint32_t inc = x;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
      if (k < K) kbase[k + 1] = carry + inc;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) kbase[0] = 0;
  }
  __syncthreads();
  // 3. stable scatter of the values
  for (uint32_t i0 = lo; i0 < hi; i0 += 32) {
    const uint32_t i = i0 + lane;
    const bool have = i < hi;
    const uint32_t k = have ? key(i) : 0xffffffffu - lane;
    const uint32_t peers = __match_any_sync(0xffffffffu, k);
    uint32_t at = 0;
    if (have) at = row[k];
    __syncwarp();
    if (have) {
      sorted[kbase[k] + at + __popc(peers & ((1u << lane) - 1u))] = val(i);
      if ((__ffs(peers) - 1) == lane) row[k] = at + __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  // 4. one warp per key: its run added in order (every lane carries the same sum; values come 32 at a time)
  for (uint32_t k = w; k < K; k += W) {
    const uint32_t a = kbase[k], z = kbase[k + 1];
    if (a == z) continue;
    float s = acc[k];
    float xnext = a + lane < z ? sorted[a + lane] : 0.f;
    for (uint32_t j0 = a; j0 < z; j0 += 32) {
      const float x = xnext;
      xnext = j0 + 32 + lane < z ? sorted[j0 + 32 + lane] : 0.f;   // the next 32 values are on their way while these are added
      const int m = (int)min(32u, z - j0);
      if (m == 32) {
#pragma unroll
        for (int l = 0; l < 32; l++) s = __fadd_rn(s, __shfl_sync(0xffffffffu, x, l));
      } else {
        for (int l = 0; l < m; l++) s = __fadd_rn(s, __shfl_sync(0xffffffffu, x, l));
      }
    }
    if (lane == 0) acc[k] = s;
  }
}


// The same for two sums per key that share the records (CGroup::cov_sum and CGroup::multi).  Keys >= K_sum are sorted
// like the others but not added (padding entries).  val(i) returns a float2; sorted: [n] float2.
template <class KeyFn, class ValFn>
__device__ void cta_ordered_keyed_sum2(uint32_t n, KeyFn key, ValFn val, uint32_t K, uint32_t K_sum, float* acc0, float* acc1, uint32_t* cnt,
                                       uint32_t* kbase, float2* sorted) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t W = blockDim.x >> 5;
  uint32_t S = (n + W - 1) / W;
  S = (S + 31u) & ~31u;
  const uint32_t lo = min(n, w * S), hi = min(n, lo + S);
  uint32_t* row = cnt + (size_t)w * K;
  for (size_t x = threadIdx.x; x < (size_t)W * K; x += blockDim.x) cnt[x] = 0;
  __syncthreads();
  for (uint32_t i0 = lo; i0 < hi; i0 += 32) {
    const uint32_t i = i0 + lane;
    const bool have = i < hi;
    const uint32_t k = have ? key(i) : 0xffffffffu - lane;
    const uint32_t peers = __match_any_sync(0xffffffffu, k);
    if (have && (__ffs(peers) - 1) == lane) row[k] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < K; k += blockDim.x) {
    uint32_t run = 0;
    for (uint32_t x = 0; x < W; x++) { const uint32_t c = cnt[(size_t)x * K + k]; cnt[(size_t)x * K + k] = run; run += c; }
    kbase[k + 1] = run;
  }
  __syncthreads();
  if (w == 0) {
    uint32_t carry = 0;
    for (uint32_t k0 = 0; k0 < K; k0 += 32) {
      const uint32_t k = k0 + lane;
      uint32_t inc = k < K ? kbase[k + 1] : 0u;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
      if (k < K) kbase[k + 1] = carry + inc;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) kbase[0] = 0;
  }
  __syncthreads();
  for (uint32_t i0 = lo; i0 < hi; i0 += 32) {
    const uint32_t i = i0 + lane;
    const bool have = i < hi;
    const uint32_t k = have ? key(i) : 0xffffffffu - lane;
    const uint32_t peers = __match_any_sync(0xffffffffu, k);
    uint32_t at = 0;
    if (have) at = row[k];
    __syncwarp();
    if (have) {
      sorted[kbase[k] + at + __popc(peers & ((1u << lane) - 1u))] = val(i);
      if ((__ffs(peers) - 1) == lane) row[k] = at + __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  for (uint32_t k = w; k < K_sum; k += W) {
    const uint32_t a = kbase[k], z = kbase[k + 1];
    if (a == z) continue;
    float s0 = acc0[k], s1 = acc1[k];
    float2 xnext = a + lane < z ? sorted[a + lane] : make_float2(0.f, 0.f);
    for (uint32_t j0 = a; j0 < z; j0 += 32) {
      const float2 x = xnext;
      xnext = j0 + 32 + lane < z ? sorted[j0 + 32 + lane] : make_float2(0.f, 0.f);
      const int m = (int)min(32u, z - j0);
      for (int l = 0; l < m; l++) {
        s0 = __fadd_rn(s0, __shfl_sync(0xffffffffu, x.x, l));
        s1 = __fadd_rn(s1, __shfl_sync(0xffffffffu, x.y, l));
      }
    }
    if (lane == 0) { acc0[k] = s0; acc1[k] = s1; }
  }
  __syncthreads();
}
