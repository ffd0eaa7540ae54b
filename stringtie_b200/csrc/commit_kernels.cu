// commit_kernels.cu — the order-dependent half of SURVEY.md §8 row a11 (update_abundance, rlink.cpp:4680-4867, and the
// node coverage of get_read_pattern, :4395): the keys and contributions frag_write produced, in read order, become
//   * transfrags numbered in order of first appearance, each with its abundance = float sum of its records IN READ ORDER;
//   * node coverages = float sums of the contributions IN READ ORDER.
// The reference does this with one sequential loop per bundle over a trie.  A bundle of 1.7 M reads (the deepest of the
// C2 workload: 1.8 % of all reads in ONE locus) makes any one-warp-per-bundle loop the whole stage (340 ms of 415 ms
// measured, profiles/r3_*), so nothing here is per-bundle sequential except the float chains themselves:
//
//   tf0_seed / tf_insert     one THREAD per record, whole batch: keys enter the bundle's open-addressing table with
//                            atomicCAS on the 64-bit key hash; atomicMin keeps the first record of every key
//   tf_first -> scan -> tf_ids   first occurrences flagged, exclusive scan = rank of a key among the bundle's new keys in
//                            order of first appearance = its transfrag id (the reference's transfrag[s][g].Add order);
//                            every record learns its transfrag; keys are verified against the first record's (a hash
//                            collision raises an error instead of merging two transfrags)
//   commit_plan              bundles with many records are handed to the CTA path below (scratch from a pool)
//   *_sum_shallow            one warp per bundle, 32 records per step, equal keys added in lane order
//   *_sum_deep               one CTA per deep bundle: STABLE counting sort of the values by key (per-warp slices, key
//                            counts per slice, prefix over the slices), then one warp per key adds its run in order —
//                            the only serial part left is the chain of dependent float adds of one key
#include "asm_device.cuh"
#include "frag_device.cuh"
#include "ordered_sum.cuh"

#define ASM_ERR_HASH 0x1000u      // two different transfrag keys of one bundle share a 64-bit hash
#define DEEP_RECORDS 8192u        // bundles with at least this many records take the CTA path
#define DEEP_THREADS 512
#define DEEP_WARPS (DEEP_THREADS / 32)

namespace {

unsigned blocks(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

__device__ __forceinline__ uint32_t tf_slot0(const DevBatchView& v, const AsmView& A, uint32_t b) {
  return A.tf0_off[A.gr_first[2 * b]] + A.r_ntf[v.b_read_off[b]];
}

__device__ __forceinline__ uint32_t table_insert(const AsmView& A, uint32_t b, unsigned long long h) {
  const uint32_t t0 = A.tab_off[b], cap = A.tab_off[b + 1] - t0;
  uint32_t slot = (uint32_t)(h >> 17) & (cap - 1);
  for (;;) {
    const unsigned long long old = atomicCAS(A.slot_hash + t0 + slot, 0ull, h);
    if (old == 0ull || old == h) return slot;
    slot = (slot + 1) & (cap - 1);
  }
}

__device__ __forceinline__ bool same_key(const AsmView& A, uint32_t a, uint32_t b) {
  if (A.rec_graph[a] != A.rec_graph[b] || A.rec_klen[a] != A.rec_klen[b]) return false;
  const uint32_t* ka = A.key + A.rec_koff[a];
  const uint32_t* kb = A.key + A.rec_koff[b];
  for (uint32_t i = 0; i < A.rec_klen[a]; i++) if (ka[i] != kb[i]) return false;
  return true;
}

// the transfrags create_graph left: each is a transfrag of its own (ids 0 .. nt0-1 of the bundle, in graph order); the
// ones that do not start at the source can be met again by reads — the LAST one with a given key, as in the reference's
// trie (construct_treepat, rlink.cpp:4486: tree->tr is overwritten)
__global__ void __launch_bounds__(256) tf0_seed_kernel(DevBatchView v, AsmView A) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n_tf0) return;
  const int gi = A.rec_graph[r];
  const uint32_t b = (uint32_t)A.g_bundle[gi];
  const uint32_t id = (uint32_t)r - A.tf0_off[A.gr_first[2 * b]];
  const uint32_t s0 = tf_slot0(v, A, b);
  A.tf_rep[s0 + id] = (int32_t)r; A.tf_graph[s0 + id] = gi; A.tf_abund[s0 + id] = A.rec_ab[r]; A.rec_tf[r] = (int32_t)id;
  if ((A.key[A.rec_koff[r]] & ~FR_FLAG) != 0u) {
    const uint32_t slot = table_insert(A, b, A.rec_hash[r]);
    atomicMax(A.slot_tf0 + A.tab_off[b] + slot, (int32_t)id);
  }
}

__global__ void __launch_bounds__(256) tf_insert_kernel(DevBatchView v, AsmView A) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n_rec) return;
  const uint32_t r = (uint32_t)(A.n_tf0 + i);
  const uint32_t b = (uint32_t)A.g_bundle[A.rec_graph[r]];
  const uint32_t slot = table_insert(A, b, A.rec_hash[r]);
  atomicMin(A.slot_minrec + A.tab_off[b] + slot, r);
  A.rec_slot[i] = slot;
}

__global__ void __launch_bounds__(256) tf_first_kernel(AsmView A) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > A.n_rec) return;
  uint32_t first = 0;
  if (i < A.n_rec) {
    const uint32_t r = (uint32_t)(A.n_tf0 + i);
    const uint32_t t = A.tab_off[A.g_bundle[A.rec_graph[r]]] + A.rec_slot[i];
    first = (A.slot_tf0[t] < 0 && A.slot_minrec[t] == r) ? 1u : 0u;
  }
  A.rec_first[i] = first;
}

// rec_first: exclusive prefix.  The id of a new key = (transfrags create_graph left) + its rank among the bundle's new keys.
__global__ void __launch_bounds__(256) tf_ids_kernel(DevBatchView v, AsmView A) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n_rec) return;
  const uint32_t r = (uint32_t)(A.n_tf0 + i);
  const int gi = A.rec_graph[r];
  const uint32_t b = (uint32_t)A.g_bundle[gi];
  const uint32_t t = A.tab_off[b] + A.rec_slot[i];
  const uint32_t g0 = A.gr_first[2 * b], g1 = A.gr_first[2 * b + 2];
  const uint32_t nt0 = A.tf0_off[g1] - A.tf0_off[g0];
  const uint32_t s0 = tf_slot0(v, A, b);
  uint32_t id, rep;
  const int32_t seed = A.slot_tf0[t];
  if (seed >= 0) { id = (uint32_t)seed; rep = A.tf0_off[g0] + id; }
  else {
    rep = A.slot_minrec[t];
    id = nt0 + (A.rec_first[rep - (uint32_t)A.n_tf0] - A.rec_first[A.r_ntf[v.b_read_off[b]]]);
    if (rep == r) { A.tf_rep[s0 + id] = (int32_t)r; A.tf_graph[s0 + id] = gi; A.tf_abund[s0 + id] = 0.f; }
  }
  A.rec_tf[r] = (int32_t)id;
  if (rep != r && !same_key(A, rep, r)) atomicOr(v.err_flag, ASM_ERR_HASH);
}

// per bundle: number of transfrags; deep bundles get scratch from the pool and go on the CTA lists
__global__ void __launch_bounds__(256) commit_plan_kernel(DevBatchView v, AsmView A) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= v.n_bundles) return;
  const uint32_t g0 = A.gr_first[2 * b], g1 = A.gr_first[2 * b + 2];
  const uint32_t nt0 = A.tf0_off[g1] - A.tf0_off[g0];
  const uint32_t q0 = A.r_ntf[v.b_read_off[b]], q1 = A.r_ntf[v.b_read_off[b + 1]];
  const uint32_t ntf = nt0 + (A.rec_first[q1] - A.rec_first[q0]);
  A.b_ntf[b] = (int32_t)ntf;
  uint8_t deep = 0;
  if (q1 - q0 >= DEEP_RECORDS && ntf) {
    const unsigned long long need = (unsigned long long)DEEP_WARPS * ntf + ntf + 1;
    const unsigned long long at = atomicAdd(A.pool_ctr, need);
    if (at + need <= A.pool_cap) {
      const uint32_t d = atomicAdd(A.deep_cnt, 1u);
      if (d < A.deep_cap) { A.deep_tf_bundle[d] = (uint32_t)b; A.deep_tf_off[d] = at; deep |= 1; }
    }
  }
  const uint32_t c0 = A.r_ncov[v.b_read_off[b]], c1 = A.r_ncov[v.b_read_off[b + 1]];
  const uint32_t K = A.x_node[g1] - A.x_node[g0];
  if (c1 - c0 >= DEEP_RECORDS && K) {
    const unsigned long long need = (unsigned long long)DEEP_WARPS * K + K + 1;
    const unsigned long long at = atomicAdd(A.pool_ctr, need);
    if (at + need <= A.pool_cap) {
      const uint32_t d = atomicAdd(A.deep_cnt + 1, 1u);
      if (d < A.deep_cap) { A.deep_cov_bundle[d] = (uint32_t)b; A.deep_cov_off[d] = at; deep |= 2; }
    }
  }
  A.b_deep[b] = deep;
}

__global__ void __launch_bounds__(DEEP_THREADS) tf_sum_deep_kernel(DevBatchView v, AsmView A) {
  if (blockIdx.x >= min(A.deep_cnt[0], A.deep_cap)) return;
  const uint32_t b = A.deep_tf_bundle[blockIdx.x];
  const uint32_t K = (uint32_t)A.b_ntf[b];
  uint32_t* cnt = A.pool + A.deep_tf_off[blockIdx.x];
  uint32_t* kbase = cnt + (size_t)DEEP_WARPS * K;
  const uint32_t q0 = A.r_ntf[v.b_read_off[b]], q1 = A.r_ntf[v.b_read_off[b + 1]];
  const uint32_t r0 = (uint32_t)A.n_tf0 + q0;
  for (size_t x = threadIdx.x; x < (size_t)DEEP_WARPS * K; x += blockDim.x) cnt[x] = 0;
  __syncthreads();
  const int32_t* rec_tf = A.rec_tf + r0;
  const float* rec_ab = A.rec_ab + r0;
  cta_ordered_keyed_sum(q1 - q0, [&](uint32_t i) { return (uint32_t)rec_tf[i]; }, [&](uint32_t i) { return rec_ab[i]; }, K,
                        A.tf_abund + tf_slot0(v, A, b), cnt, kbase, A.ab_sorted + q0);
}

__global__ void __launch_bounds__(DEEP_THREADS) nodecov_deep_kernel(DevBatchView v, AsmView A) {
  if (blockIdx.x >= min(A.deep_cnt[1], A.deep_cap)) return;
  const uint32_t b = A.deep_cov_bundle[blockIdx.x];
  const uint32_t g0 = A.gr_first[2 * b], g1 = A.gr_first[2 * b + 2];
  const uint32_t n_lo = A.x_node[g0], K = A.x_node[g1] - n_lo;
  uint32_t* cnt = A.pool + A.deep_cov_off[blockIdx.x];
  uint32_t* kbase = cnt + (size_t)DEEP_WARPS * K;
  const uint32_t c0 = A.r_ncov[v.b_read_off[b]], c1 = A.r_ncov[v.b_read_off[b + 1]];
  for (size_t x = threadIdx.x; x < (size_t)DEEP_WARPS * K; x += blockDim.x) cnt[x] = 0;
  __syncthreads();
  const uint32_t* cov_node = A.cov_node + c0;
  const float* cov_val = A.cov_val + c0;
  cta_ordered_keyed_sum(c1 - c0, [&](uint32_t i) { return cov_node[i] - n_lo; }, [&](uint32_t i) { return cov_val[i]; }, K,
                        A.n_cov + n_lo, cnt, kbase, A.cov_sorted + c0);
}

// shallow bundles: one warp walks the bundle's records 32 at a time; the first lane of every group of equal keys adds the
// group's values in lane order
__global__ void __launch_bounds__(128) tf_sum_shallow_kernel(DevBatchView v, AsmView A) {
  __shared__ float stage[4][32];
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= v.n_bundles || (A.b_deep[b] & 1)) return;
  const int lane = threadIdx.x & 31;
  float* st = stage[(threadIdx.x >> 5) & 3];
  const uint32_t q0 = A.r_ntf[v.b_read_off[b]], q1 = A.r_ntf[v.b_read_off[b + 1]];
  float* acc = A.tf_abund + tf_slot0(v, A, (uint32_t)b);
  for (uint32_t q = q0; q < q1; q += 32) {
    const uint32_t r = (uint32_t)A.n_tf0 + q + lane;
    const bool have = q + lane < q1;
    const int id = have ? A.rec_tf[r] : -1 - lane;
    st[lane] = have ? A.rec_ab[r] : 0.f;
    __syncwarp();
    const uint32_t peers = __match_any_sync(0xffffffffu, id);
    if (have && (__ffs(peers) - 1) == lane) {
      float a = acc[id];
      uint32_t mm = peers;
      while (mm) { const int l = __ffs(mm) - 1; mm &= mm - 1; a = __fadd_rn(a, st[l]); }
      acc[id] = a;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) nodecov_shallow_kernel(DevBatchView v, AsmView A) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= v.n_bundles || (A.b_deep[b] & 2)) return;
  const uint32_t c0 = A.r_ncov[v.b_read_off[b]], c1 = A.r_ncov[v.b_read_off[b + 1]];
  if (c1 > c0) nodecov_commit_warp(A.cov_node, A.cov_val, c0, c1, A.n_cov);
}

}  // namespace

cudaError_t launch_asm_tf_commit(const DevBatchView& v, const AsmView& A, cudaStream_t st) {
  if (A.n_tf0) tf0_seed_kernel<<<blocks(A.n_tf0, 256), 256, 0, st>>>(v, A);
  if (A.n_rec) tf_insert_kernel<<<blocks(A.n_rec, 256), 256, 0, st>>>(v, A);
  tf_first_kernel<<<blocks(A.n_rec + 1, 256), 256, 0, st>>>(A);
  return cudaGetLastError();
}
cudaError_t launch_asm_tf_ids(const DevBatchView& v, const AsmView& A, cudaStream_t st) {   // after the scan of rec_first
  if (A.n_rec) tf_ids_kernel<<<blocks(A.n_rec, 256), 256, 0, st>>>(v, A);
  if (v.n_bundles) commit_plan_kernel<<<blocks(v.n_bundles, 256), 256, 0, st>>>(v, A);
  return cudaGetLastError();
}
cudaError_t launch_asm_tf_sums(const DevBatchView& v, const AsmView& A, cudaStream_t st) {
  if (!v.n_bundles) return cudaSuccess;
  tf_sum_deep_kernel<<<A.deep_cap, DEEP_THREADS, 0, st>>>(v, A);
  tf_sum_shallow_kernel<<<blocks(v.n_bundles * 32, 128), 128, 0, st>>>(v, A);
  return cudaGetLastError();
}
cudaError_t launch_asm_nodecov_commit(const DevBatchView& v, const AsmView& A, cudaStream_t st) {
  if (!v.n_bundles) return cudaSuccess;
  nodecov_deep_kernel<<<A.deep_cap, DEEP_THREADS, 0, st>>>(v, A);
  nodecov_shallow_kernel<<<blocks(v.n_bundles * 32, 128), 128, 0, st>>>(v, A);
  return cudaGetLastError();
}
int64_t asm_deep_records() { return DEEP_RECORDS; }
