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
#include "exact_chain.cuh"
#define OS_BATCH 8
#define OS_E 8        // consecutive values a lane of a chain warp holds: a block of the chain is 32 * OS_E values

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
  for (uint32_t i0 = lo; i0 < hi; i0 += 32 * OS_BATCH) {   // OS_BATCH loads in flight per lane: the pass is bound by load latency
    uint32_t kk[OS_BATCH];
#pragma unroll
    for (int u = 0; u < OS_BATCH; u++) { const uint32_t i = i0 + 32u * u + lane; kk[u] = i < hi ? key(i) : 0xffffffffu; }
#pragma unroll
    for (int u = 0; u < OS_BATCH; u++) {
      if (i0 + 32u * u >= hi) break;
      const bool have = kk[u] != 0xffffffffu;
      const uint32_t k = have ? kk[u] : 0xffffffffu - lane;
      const uint32_t peers = __match_any_sync(0xffffffffu, k);
      if (have && (__ffs(peers) - 1) == lane) row[k] += __popc(peers);
      __syncwarp();
    }
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
  for (uint32_t i0 = lo; i0 < hi; i0 += 32 * OS_BATCH) {
    uint32_t kk[OS_BATCH]; float vv[OS_BATCH];
#pragma unroll
    for (int u = 0; u < OS_BATCH; u++) {
      const uint32_t i = i0 + 32u * u + lane;
      kk[u] = i < hi ? key(i) : 0xffffffffu; vv[u] = i < hi ? val(i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < OS_BATCH; u++) {
      if (i0 + 32u * u >= hi) break;
      const bool have = kk[u] != 0xffffffffu;
      const uint32_t k = have ? kk[u] : 0xffffffffu - lane;
      const uint32_t peers = __match_any_sync(0xffffffffu, k);
      uint32_t at = 0;
      if (have) at = row[k];
      __syncwarp();
      if (have) {
        sorted[kbase[k] + at + __popc(peers & ((1u << lane) - 1u))] = vv[u];
        if ((__ffs(peers) - 1) == lane) row[k] = at + __popc(peers);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // 4. one warp per key: its run added in order (every lane carries the same sum), 32 * OS_E values at a time through
  //    warp_chain_add (exact_chain.cuh): inside a binade of the running sum the chain of dependent adds is an integer sum
  //    the warp takes with one scan; a block that leaves the binade is added the plain way.  The next block is loaded
  //    while this one is added.
  for (uint32_t k = w; k < K; k += W) {
    const uint32_t a = kbase[k], z = kbase[k + 1];
    if (a == z) continue;
    if (z - a >= EC_CTA_MIN) continue;    // below: the whole CTA
    const EcFloatList ld{sorted + a, (int)(z - a)};
    const float s = warp_chain_list(acc[k], ld.n, ec_misalign8(ld.list), ld, lane);
    if (lane == 0) acc[k] = s;
  }
  __shared__ EcCtaState cst;
  for (uint32_t k = 0; k < K; k++) {      // the few long runs, one after the other, every warp preparing blocks (CTA-uniform loop)
    const uint32_t a = kbase[k], z = kbase[k + 1];
    if (z - a < EC_CTA_MIN) continue;
    const EcFloatList ld{sorted + a, (int)(z - a)};
    const float s = cta_chain_list(acc[k], ld.n, ec_misalign8(ld.list), ld, cst);
    if (threadIdx.x == 0) acc[k] = s;
  }
}


// The same for two sums per key that share the records (CGroup::cov_sum and CGroup::multi), for keys that are SPARSE in a
// large range (group slots of a bundle: thousands were created, a handful are visited by the records at hand).  The
// distinct keys are first collected in a shared-memory hash table and numbered densely; the counting sort runs on the
// dense numbers, the sums go to acc0 / acc1 [key].  key(i) < 0 marks padding.  Returns false (nothing done) when the
// records visit more than OS_DENSE_MAX distinct keys: the caller falls back to its serial loop.
#define OS_TAB 1024            // hash slots
#define OS_DENSE_MAX 448       // distinct keys the shared-memory counts can take
struct OrderedSumSmem {
  uint32_t tkey[OS_TAB]; int32_t tidx[OS_TAB];
  uint32_t dkey[OS_DENSE_MAX];
  uint32_t cnt[8 * OS_DENSE_MAX]; uint32_t kbase[OS_DENSE_MAX + 1];
  uint32_t ndist; int overflow;
};

// component C of a float2 list
template <int C> struct EcFloat2List {
  const float2* list; int n;
  __device__ __forceinline__ void load(int rel, float (&x)[8]) const {
    float2 t[8];
    ec_load8(list, n, rel, t, make_float2(0.f, 0.f));
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = C ? t[e].y : t[e].x;
  }
  __device__ __forceinline__ void prefetch(int rel) const { if (rel < n) ec_prefetch_l2(list + rel); }
};

template <class KeyFn, class ValFn>
__device__ bool cta_ordered_keyed_sum2(uint32_t n, KeyFn key, ValFn val, float* acc0, float* acc1, OrderedSumSmem& M, float2* sorted) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t W = blockDim.x >> 5;      // <= 8
  // 0. the distinct keys
  for (uint32_t x = threadIdx.x; x < OS_TAB; x += blockDim.x) M.tkey[x] = 0xffffffffu;
  if (threadIdx.x == 0) { M.ndist = 0; M.overflow = 0; }
  __syncthreads();
  for (uint32_t i0 = threadIdx.x; i0 < n; i0 += blockDim.x * OS_BATCH) {
    int32_t kk[OS_BATCH];
#pragma unroll
    for (int u = 0; u < OS_BATCH; u++) { const uint32_t i = i0 + u * blockDim.x; kk[u] = i < n ? key(i) : -1; }
#pragma unroll
    for (int u = 0; u < OS_BATCH; u++) {
      if (kk[u] < 0) continue;
      uint32_t h = ((uint32_t)kk[u] * 2654435761u) >> 22;
      for (int probe = 0; probe < OS_TAB; probe++) {
        const uint32_t old = M.tkey[h];
        if (old == (uint32_t)kk[u]) break;
        if (old == 0xffffffffu) { const uint32_t was = atomicCAS(&M.tkey[h], 0xffffffffu, (uint32_t)kk[u]); if (was == 0xffffffffu || was == (uint32_t)kk[u]) break; }
        h = (h + 1) & (OS_TAB - 1);
        if (probe == OS_TAB - 1) M.overflow = 1;
      }
    }
  }
  __syncthreads();
  for (uint32_t x = threadIdx.x; x < OS_TAB; x += blockDim.x) {
    if (M.tkey[x] != 0xffffffffu) {
      const uint32_t id = atomicAdd(&M.ndist, 1u);
      M.tidx[x] = (int32_t)id;
      if (id < OS_DENSE_MAX) M.dkey[id] = M.tkey[x];
    }
  }
  __syncthreads();
  const uint32_t K = M.ndist;
  if (M.overflow || K > OS_DENSE_MAX) { __syncthreads(); return false; }
  if (K == 0) return true;
  auto dense = [&](int32_t k) -> uint32_t {
    uint32_t h = ((uint32_t)k * 2654435761u) >> 22;
    while (M.tkey[h] != (uint32_t)k) h = (h + 1) & (OS_TAB - 1);
    return (uint32_t)M.tidx[h];
  };
  uint32_t S = (n + W - 1) / W;
  S = (S + 31u) & ~31u;
  const uint32_t lo = min(n, w * S), hi = min(n, lo + S);
  uint32_t* row = M.cnt + (size_t)w * K;
  for (uint32_t x = threadIdx.x; x < W * K; x += blockDim.x) M.cnt[x] = 0;
  __syncthreads();
  for (uint32_t i0 = lo; i0 < hi; i0 += 32 * OS_BATCH) {
    int32_t kk[OS_BATCH];
#pragma unroll
    for (int u = 0; u < OS_BATCH; u++) { const uint32_t i = i0 + 32u * u + lane; kk[u] = i < hi ? key(i) : -1; }
#pragma unroll
    for (int u = 0; u < OS_BATCH; u++) {
      if (i0 + 32u * u >= hi) break;
      const bool have = kk[u] >= 0;
      const uint32_t k = have ? dense(kk[u]) : 0xffffffffu - lane;
      const uint32_t peers = __match_any_sync(0xffffffffu, k);
      if (have && (__ffs(peers) - 1) == lane) row[k] += __popc(peers);
      __syncwarp();
    }
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < K; k += blockDim.x) {
    uint32_t run = 0;
    for (uint32_t x = 0; x < W; x++) { const uint32_t c = M.cnt[(size_t)x * K + k]; M.cnt[(size_t)x * K + k] = run; run += c; }
    M.kbase[k + 1] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t run = 0; for (uint32_t k = 0; k < K; k++) { const uint32_t c = M.kbase[k + 1]; M.kbase[k] = run; run += c; } M.kbase[K] = run; }
  __syncthreads();
  for (uint32_t i0 = lo; i0 < hi; i0 += 32 * OS_BATCH) {
    int32_t kk[OS_BATCH]; float2 vv[OS_BATCH];
#pragma unroll
    for (int u = 0; u < OS_BATCH; u++) {
      const uint32_t i = i0 + 32u * u + lane;
      kk[u] = i < hi ? key(i) : -1; vv[u] = i < hi ? val(i) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < OS_BATCH; u++) {
      if (i0 + 32u * u >= hi) break;
      const bool have = kk[u] >= 0;
      const uint32_t k = have ? dense(kk[u]) : 0xffffffffu - lane;
      const uint32_t peers = __match_any_sync(0xffffffffu, k);
      uint32_t at = 0;
      if (have) at = row[k];
      __syncwarp();
      if (have) {
        sorted[M.kbase[k] + at + __popc(peers & ((1u << lane) - 1u))] = vv[u];
        if ((__ffs(peers) - 1) == lane) row[k] = at + __popc(peers);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (uint32_t k = w; k < K; k += W) {
    const uint32_t a = M.kbase[k], z = M.kbase[k + 1];
    if (a == z) continue;
    const uint32_t real = M.dkey[k];
    if (z - a >= EC_CTA_MIN) continue;    // below: the whole CTA
    const float2* list = sorted + a;
    const int n = (int)(z - a), mis = ec_misalign8(list);
    const float s0 = warp_chain_list(acc0[real], n, mis, EcFloat2List<0>{list, n}, lane);
    const float s1 = warp_chain_list(acc1[real], n, mis, EcFloat2List<1>{list, n}, lane);
    if (lane == 0) { acc0[real] = s0; acc1[real] = s1; }
  }
  __shared__ EcCtaState cst;
  for (uint32_t k = 0; k < K; k++) {      // the few long runs: every warp prepares blocks (CTA-uniform loop)
    const uint32_t a = M.kbase[k], z = M.kbase[k + 1];
    if (z - a < EC_CTA_MIN) continue;
    const uint32_t real = M.dkey[k];
    const float2* list = sorted + a;
    const int n = (int)(z - a), mis = ec_misalign8(list);
    const float s0 = cta_chain_list(acc0[real], n, mis, EcFloat2List<0>{list, n}, cst);
    const float s1 = cta_chain_list(acc1[real], n, mis, EcFloat2List<1>{list, n}, cst);
    if (threadIdx.x == 0) { acc0[real] = s0; acc1[real] = s1; }
  }
  __syncthreads();
  return true;
}
