// exact_chain.cuh — a chain of dependent float additions  s = fl(fl(fl(s + x0) + x1) + ...)  without the chain.
//
// The reference accumulates its float sums with "+=" in read order (cov_edge_add rlink.cpp:133-142, CGraphnode::cov and
// CTransfrag::abundance in update_abundance :4680, CGroup::cov_sum / multi in merge_read_to_group :1281); the bit pattern
// of the result depends on the order, so the terms of one sum cannot be reassociated.  What CAN be done: while the running
// sum stays inside one binade [2^k, 2^(k+1)) every float in reach lies on the grid q·Z, q = 2^(k-23), and IEEE
// round-to-nearest of  s + x  onto that grid is  s + q·rint(x / q)  whenever x / q is not exactly halfway between two
// integers (s / q is an integer, and rounding to nearest commutes with integer shifts).  So inside a binade the chain is
// an INTEGER sum of per-term increments rint(x_i / q), which is associative: a warp takes 32·E terms at a time, every
// lane turns its E consecutive terms into increments (shifts on the bit patterns, no floating point), one warp scan of
// the lane totals gives every lane its prefix, and the block is accepted if every partial sum stays strictly inside the
// binade; otherwise the block is added the plain way.  For a monotone sum a binade is left ~25 times in the life of the
// chain, whatever its length.
//
// ec_chain_incr and the acceptance bounds are plain integer code shared with the CPU check (tests/exact_chain_check.cc:
// the same functions against the plain loop on random and adversarial chains).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define EC_HD __host__ __device__ __forceinline__
#else
#define EC_HD static inline
#endif

// A tie (x / q exactly halfway) rounds to the EVEN neighbour, which looks at the parity of the sum in front of it.  That
// is one bit of state.  A lane runs its E terms as if the sum in front of them were even; had it been odd, everything is
// the same up to the lane's first tie, which then goes the other way (+-1), and from there on both runs sit on even sums
// and move in step: the odd-start run is the even-start run + delta, delta in {-1, 0, +1}.  The warp scan composes the
// lanes' "parity in -> increment" maps (d0, d1 = d0 + delta).

// partial sums, in units of q, must stay in [EC_LO, EC_HI]: then the exact sum is > 2^23 + 0.5 and < 2^24 - 0.5 grid
// steps, i.e. inside the binade whose grid is q on both sides of it (the odd-start run is checked with its extremes
// widened by one step instead of tracked exactly)
#define EC_LO 0x800002
#define EC_HI 0xfffffe
#define EC_MIN_SHIFT 3      // |x| < |s| / 4: an increment is below 2^21, 256 of them cannot overflow 32 bits

// rint(x / q) for the term with bit pattern xb, q = the grid step of a normal sum with biased exponent ex_s; the sign is
// relative to the sum's (the chain is run on |s|).  A tie returns floor(x / q) and sets `tie`: the increment is then
// floor + ((sum + floor) & 1).  Sets `bad` when the term has to go the plain way: NaN / Inf, denormal, or too large next
// to the sum.  Branch-free: x / q = mant / 2^sh with sh clamped to 26 (anything smaller than half a step adds 0).
EC_HD int32_t ec_chain_incr(uint32_t xb, int ex_s, uint32_t neg_s, bool& tie, bool& bad) {
  const uint32_t ex_x = (xb >> 23) & 0xffu;
  const int sh0 = ex_s - (int)ex_x;     // x = mant · 2^(ex_x - 150), q = 2^(ex_s - 150)
  const int sh = sh0 > 26 ? 26 : (sh0 < 1 ? 1 : sh0);
  const uint32_t mant = ex_x ? ((xb & 0x7fffffu) | 0x800000u) : 0u;   // +-0 adds nothing
  bad = bad || ex_x == 255u || (ex_x != 0u && sh0 < EC_MIN_SHIFT) || (ex_x == 0u && (xb << 1) != 0u);
  const uint32_t half = 1u << (sh - 1);
  tie = (mant & ((half << 1) - 1u)) == half;
  const int32_t r = (int32_t)((mant + (tie ? 0u : half)) >> sh);      // nearest, or the floor of a tie
  const bool neg = (xb >> 31) != neg_s;
  return neg ? -r - (tie ? 1 : 0) : r;
}

// One lane's E terms on an even sum in front of them: run = total increment, mn / mx = extremes of the partial sums after
// each term (relative to the sum in front), delta = what an odd sum in front adds to everything after the first tie.
struct EcLane { int32_t run, mn, mx, delta; bool seen; };
EC_HD void ec_lane_step(EcLane& L, bool first, int32_t r, bool tie) {
  const int32_t cur = first ? 0 : L.run;
  const int32_t odd = (cur + r) & 1;
  if (first) { L.delta = 0; L.seen = false; }
  if (tie && !L.seen) { L.delta = odd ? -1 : 1; L.seen = true; }
  const int32_t nx = cur + r + (tie ? odd : 0);
  L.run = nx;
  L.mn = first ? nx : (nx < L.mn ? nx : L.mn);
  L.mx = first ? nx : (nx > L.mx ? nx : L.mx);
}
// the map of lanes [.. F ..][.. G ..] from the maps of F (earlier) and G (later): d[b] = increment when the sum in front has parity b
EC_HD void ec_compose(int32_t f0, int32_t f1, int32_t g0, int32_t g1, int32_t& h0, int32_t& h1) {
  h0 = f0 + ((f0 & 1) ? g1 : g0);
  h1 = f1 + (((1 + f1) & 1) ? g1 : g0);
}

#if defined(__CUDACC__)
// A lane's 8 consecutive elements list[rel .. rel + 8) of a list of n elements (T of 4 or 8 bytes), as 128-bit loads.
// Chain warps cut a list into blocks of 256 that start on a 32- / 64-byte boundary IN MEMORY, so `rel` may be negative for
// the first lane of the first block (elements in front of the list read as `pad`; the address is aligned and lies in
// the same allocation) and the last group may be ragged (loaded element by element).
template <class T>
__device__ __forceinline__ int ec_misalign8(const T* list) { return (int)(((uintptr_t)list / sizeof(T)) & 7u); }
template <class T>
__device__ __forceinline__ void ec_load8(const T* __restrict__ list, int n, int rel, T (&out)[8], T pad) {
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "ec_load8");
  if (rel >= n || rel <= -8) {
#pragma unroll
    for (int e = 0; e < 8; e++) out[e] = pad;
    return;
  }
  if (rel + 8 <= n) {
    constexpr int NV = (int)sizeof(T) / 2;      // 128-bit vectors per 8 elements
    union { uint4 v[NV]; T t[8]; } u;
    const uint4* p = reinterpret_cast<const uint4*>(list + rel);
#pragma unroll
    for (int k = 0; k < NV; k++) u.v[k] = p[k];
#pragma unroll
    for (int e = 0; e < 8; e++) out[e] = rel + e < 0 ? pad : u.t[e];
  } else {
#pragma unroll
    for (int e = 0; e < 8; e++) out[e] = (rel + e >= 0 && rel + e < n) ? list[rel + e] : pad;
  }
}
#endif

#if defined(__CUDACC__)
// What a block of 32·E terms (lane-major: lane 0 holds the first E, lane 1 the next E, ...) does to a running sum of a
// given sign and binade (`key` = its sign and exponent bits), prepared WITHOUT the sum: per lane the two runs, their
// extremes, and the composed map of the lanes in front.  ec_post then needs a handful of dependent instructions per
// block, so the preparation of the next block overlaps the completion of this one (warp_chain_list).
#ifdef EC_STATS
__device__ unsigned long long ec_stats[4];   // blocks in integer form, blocks added plainly, of those: prepared for the right binade, lanes with a bad term
#endif
struct EcPre {
  int32_t run, mn, mx, delta;     // this lane's terms on an even sum in front of them (EcLane)
  int32_t ex0, ex1;               // increment of the lanes in front of this one, by parity of the sum in front of the block
  bool bad;                       // this lane holds a term the integer form cannot take
  bool allzero;                   // warp-uniform: the block holds nothing but +-0
};

template <int E>
__device__ __forceinline__ EcPre ec_pre(const float (&x)[E], uint32_t key, int lane) {
  EcPre P;
  const int ex_s = (int)((key >> 23) & 0xffu);
  const uint32_t neg_s = key >> 31;
  uint32_t nz = 0u;
  P.bad = !(ex_s >= 1 && ex_s <= 254);
  EcLane L;
#pragma unroll
  for (int e = 0; e < E; e++) {
    const uint32_t xb = __float_as_uint(x[e]);
    nz |= xb << 1;
    bool tie;
    const int32_t r = ec_chain_incr(xb, ex_s, neg_s, tie, P.bad);
    ec_lane_step(L, e == 0, r, tie);
  }
  P.allzero = !__any_sync(0xffffffffu, nz != 0u);
  int32_t i0 = L.run, i1 = L.run + L.delta;    // inclusive map of the lanes up to this one
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t f0 = __shfl_up_sync(0xffffffffu, i0, o), f1 = __shfl_up_sync(0xffffffffu, i1, o);
    if (lane >= o) { int32_t h0, h1; ec_compose(f0, f1, i0, i1, h0, h1); i0 = h0; i1 = h1; }
  }
  P.ex0 = __shfl_up_sync(0xffffffffu, i0, 1); P.ex1 = __shfl_up_sync(0xffffffffu, i1, 1);
  if (lane == 0) { P.ex0 = 0; P.ex1 = 0; }
  P.run = L.run; P.mn = L.mn; P.mx = L.mx; P.delta = L.delta;
  return P;
}

// s after the block: the integer form if the block was prepared for the binade s is in and every partial sum stays inside
// it, else the plain chain of adds over the first `lanes` lanes.
template <int E>
__device__ __forceinline__ float ec_post(const EcPre& P, uint32_t key, float s, const float (&x)[E], int lane, int lanes) {
  const uint32_t sb = __float_as_uint(s);
  if (P.allzero && sb != 0x80000000u) return s;          // only -0 + +0 changes a sum
  if ((sb & 0xff800000u) == key) {
    const int32_t A = (int32_t)((sb & 0x7fffffu) | 0x800000u);
    const int32_t at = A + ((A & 1) ? P.ex1 : P.ex0);    // the sum in front of this lane's terms
    const int32_t run = P.run + ((at & 1) ? P.delta : 0);
    const bool bad = P.bad || at + P.mn < EC_LO || at + P.mx > EC_HI;
    if (!__any_sync(0xffffffffu, bad)) {
      const int32_t fin = __shfl_sync(0xffffffffu, at + run, 31);
      return __uint_as_float(key | ((uint32_t)fin & 0x7fffffu));
    }
  }
#pragma unroll 1
  for (int l = 0; l < lanes; l++) {
#pragma unroll
    for (int e = 0; e < E; e++) s = __fadd_rn(s, __shfl_sync(0xffffffffu, x[e], l));
  }
  return s;
}

// One block on its own (no overlap).  Terms past the end of a list are passed as +0.0 and only the first `lanes` lanes hold
// terms.  Returns the new sum in every lane.
template <int E>
__device__ __forceinline__ float warp_chain_add(float s, const float (&x)[E], int lane, int lanes = 32) {
  const uint32_t key = __float_as_uint(s) & 0xff800000u;
  const EcPre P = ec_pre<E>(x, key, lane);
  return ec_post<E>(P, key, s, x, lane, lanes);
}

__device__ __forceinline__ void ec_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// A whole list of n terms added to s in order by one warp.  ld.load(rel, x) gives the lane its 8 terms rel .. rel + 7
// (+0.0 outside [0, n)); the first block starts at rel = -mis so that the loader can use aligned 128-bit loads;
// ld.prefetch(rel) may pull a far-ahead block towards the L2.  Three blocks are in registers: one being completed, one
// being prepared (for the binade of the sum in front of the first: re-prepared in the rare case the first one leaves
// it), one in flight from memory.
template <class Loader>
__device__ __forceinline__ float warp_chain_list(float s, int n, int mis, const Loader& ld, int lane) {
  constexpr int E = 8, B = 32 * E;
  float xa[E], xb[E], xc[E];
  int b0 = -mis;
  ld.load(b0 + lane * E, xa);
  ld.load(b0 + B + lane * E, xb);
  uint32_t key = __float_as_uint(s) & 0xff800000u;
  EcPre pa = ec_pre<E>(xa, key, lane);
  for (; b0 < n; b0 += B) {
    ld.load(b0 + 2 * B + lane * E, xc);
    ld.prefetch(b0 + 24 * B + lane * E);
    EcPre pb = ec_pre<E>(xb, key, lane);
    s = ec_post<E>(pa, key, s, xa, lane, min(32, (n - b0 + E - 1) / E));
    const uint32_t key2 = __float_as_uint(s) & 0xff800000u;
    if (key2 != key) { key = key2; pb = ec_pre<E>(xb, key, lane); }
    pa = pb;
#pragma unroll
    for (int e = 0; e < E; e++) { xa[e] = xb[e]; xb[e] = xc[e]; }
  }
  return s;
}

// The same by a whole CTA, for the few lists long enough to matter on their own (one position of a deep locus, one node of
// a deep graph: 10^6 terms): the preparation of a block (ec_pre: ~85 % of the work) does not need the sum, so warp w
// prepares blocks w, w + W, ... while the sum travels from block to block through shared memory (ec_post: a vote and a
// shuffle).  Called by every thread of the CTA; returns the new sum to all.
struct EcCtaState { float s; int turn; };
template <class Loader>
__device__ __forceinline__ float cta_chain_list(float s_in, int n, int mis, const Loader& ld, volatile EcCtaState& st) {
  constexpr int E = 8, B = 32 * E;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int nb = (n + mis + B - 1) / B;
  __syncthreads();
  if (threadIdx.x == 0) { st.s = s_in; st.turn = 0; }
  __syncthreads();
  float xa[E], xn[E];
  int t = w;
  if (t < nb) ld.load(t * B - mis + lane * E, xa);
  for (; t < nb; t += W) {
    ld.load((t + W) * B - mis + lane * E, xn);           // past the end: zeros, no memory access
    ld.prefetch((t + 4 * W) * B - mis + lane * E);
    uint32_t key = __float_as_uint(st.s) & 0xff800000u;  // the binade the chain is in right now: a guess for block t
    EcPre P = ec_pre<E>(xa, key, lane);
    while (st.turn != t) { }
    float s = st.s;
    const uint32_t key2 = __float_as_uint(s) & 0xff800000u;
    if (key2 != key) { key = key2; P = ec_pre<E>(xa, key, lane); }
    s = ec_post<E>(P, key, s, xa, lane, min(32, (n + mis - t * B + E - 1) / E));
    if (lane == 0) { st.s = s; __threadfence_block(); st.turn = t + 1; }
#pragma unroll
    for (int e = 0; e < E; e++) xa[e] = xn[e];
  }
  __syncthreads();
  return st.s;
}
#define EC_CTA_MIN 4096     // terms from which a list is worth a CTA

// plain float lists
struct EcFloatList {
  const float* list; int n;
  __device__ __forceinline__ void load(int rel, float (&x)[8]) const { ec_load8(list, n, rel, x, 0.f); }
  __device__ __forceinline__ void prefetch(int rel) const { if (rel < n) ec_prefetch_l2(list + rel); }
};
#endif
