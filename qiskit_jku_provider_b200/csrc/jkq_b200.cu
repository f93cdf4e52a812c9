// libjkq_b200.so -- kernels and the C ABI (include/jkq.h) of the B200-native QMDD engine.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <array>
#include <algorithm>
#include <chrono>
#include <thread>
#include <map>
#include <string>
#include <vector>

#include "../../include/jkq.h"
#include "jkq_apply.cuh"
#include "jkq_sample.cuh"

// =============================================================================== kernels
extern __shared__ __align__(16) unsigned char jkq_dyn_smem[];  // MODE_CTA: CtaState + ThinState; MODE_GRID k_run_gates: ThinState
struct CtaSmem {
  CtaState s;
  ThinState t;
};

// 227 KB of shared memory per CTA on sm_100; the kernels add up to ~200 bytes of static shared memory on top
static_assert(sizeof(CtaSmem) + 512 <= 227 * 1024, "CtaState + ThinState no longer fit one SM's shared memory");

template <int MODE> __device__ __forceinline__ Ctx make_ctx(JEng *E) {
  Ctx c;
  c.E = E;
  if constexpr (MODE == MODE_CTA) {
    c.sh = &reinterpret_cast<CtaSmem *>(jkq_dyn_smem)->s;
    c.T = &reinterpret_cast<CtaSmem *>(jkq_dyn_smem)->t;
  } else {
    c.sh = nullptr;
    c.T = reinterpret_cast<ThinState *>(jkq_dyn_smem);  // the cursor runs in warp 0 of block 0 (k_run_gates only)
  }
  c.items = MODE == MODE_CTA ? c.sh->items : E->items;
  c.phase = 0; c.ring = 0; c.gprof = 0;
  c.ltid = threadIdx.x;
  c.lnth = blockDim.x;
  if (MODE == MODE_CTA) { c.tid = threadIdx.x; c.nth = blockDim.x; c.block0 = 1; }
  else { c.tid = blockIdx.x * blockDim.x + threadIdx.x; c.nth = gridDim.x * blockDim.x; c.block0 = blockIdx.x == 0; }
  return c;
}

template <int MODE> __global__ void __launch_bounds__(256) k_init(JEng *E) {
  Ctx c = make_ctx<MODE>(E);
  init_state<MODE>(c);
}

// Persistent circuit kernel: all gates of a segment in one launch, no host round trip per gate.
// The grid shape is bound by memory-level parallelism: every thread follows ONE dependent chain of random accesses, so
// throughput = resident threads / latency.  Two CTAs per SM (128 registers per thread, the rest spills to L1-cached local
// memory) instead of one doubles the accesses in flight (JKQ_GRID_MINB: A/B builds).
#ifndef JKQ_GRID_MINB
#define JKQ_GRID_MINB 1
#endif
template <int MODE, bool SUMS> __global__ void __launch_bounds__(256, (MODE == MODE_GRID ? JKQ_GRID_MINB : 1)) k_run_gates(JEng *E, const JGate *gates, int n_gates) {
  Ctx c = make_ctx<MODE>(E);
  run_gates<MODE, SUMS>(c, gates, n_gates);
}

template <int MODE> __global__ void __launch_bounds__(256) k_build_sampling(JEng *E, JSNode *sn, double *prob) {
  Ctx c = make_ctx<MODE>(E);
  build_sampling<MODE>(c, sn, prob);
}

__global__ void k_amplitudes(JEng *E, uint64_t first, uint64_t count, double2 *out) {
  Ctx c;
  c.E = E; c.sh = nullptr; c.T = nullptr; c.items = E->items; c.gprof = 0; c.phase = 0; c.ring = 0; c.block0 = 0; c.ltid = threadIdx.x; c.lnth = blockDim.x;
  c.tid = blockIdx.x * blockDim.x + threadIdx.x; c.nth = gridDim.x * blockDim.x;
  for (uint64_t i = c.tid; i < count; i += c.nth) {
    double re, im;
    amp_walk<MODE_GRID>(c, first + i, &re, &im);
    out[i] = make_double2(re, im);
  }
}

// the same in a caller-chosen bit order: bit b of the output index is qubit bp.q[b]; from the dense vector of a snapshot
// when there is one, else by walking the diagram E names
struct BitPerm { uint8_t q[JKQ_MAXQ]; };
__global__ void k_amplitudes_ordered(JEng *E, const double2 *dense, BitPerm bp, int n, uint64_t first, uint64_t count, double2 *out) {
  Ctx c;
  c.E = E; c.sh = nullptr; c.T = nullptr; c.items = nullptr; c.gprof = 0; c.phase = 0; c.ring = 0; c.block0 = 0; c.ltid = threadIdx.x; c.lnth = blockDim.x;
  c.tid = blockIdx.x * blockDim.x + threadIdx.x; c.nth = gridDim.x * blockDim.x;
  for (uint64_t i = c.tid; i < count; i += c.nth) {
    const uint64_t idx = first + i;
    uint64_t e = 0;
    for (int b = 0; b < n; b++) e |= ((idx >> b) & 1ull) << bp.q[b];
    if (dense) { out[i] = dense[e]; continue; }
    double re, im;
    amp_walk<MODE_GRID>(c, e, &re, &im);
    out[i] = make_double2(re, im);
  }
}

// Simulator::AssignProbsOne (Simulator.cpp:291-339): probability that the qubit on `level` reads 0.  The arena
// is level-ordered and the sampling records hold every node's branch probability, so the path mass is pushed
// down from the root one level at a time; at `level` the masses times q0 add up to P(0).  One CTA.
__device__ double marginal_block(const JEng *E, const JSNode *sn, double *mass, int level) {
  __shared__ double s_sum;
  const int n = (int)E->n;
  const uint32_t total = E->node_count;
  for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) mass[i] = 0.0;
  if (threadIdx.x == 0) s_sum = 0.0;
  __syncthreads();
  if (threadIdx.x == 0) mass[E->root] = 1.0;
  __syncthreads();
  for (int L = n - 1; L > level; L--) {
    const uint32_t off = E->lvl_off[L], cnt = E->lvl_cnt[L];
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
      const JSNode s = sn[off + i];
      const double m = mass[off + i];
      if (m == 0.0) continue;
      if (s.q0 > 0.0) atomicAdd(&mass[s.c0], m * s.q0);
      if (s.q0 < 1.0) atomicAdd(&mass[s.c1], m * (1.0 - s.q0));
    }
    __syncthreads();
  }
  double acc = 0.0;
  const uint32_t off = E->lvl_off[level], cnt = E->lvl_cnt[level];
  for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) acc += mass[off + i] * sn[off + i].q0;
  atomicAdd(&s_sum, acc);
  __syncthreads();
  const double r = s_sum;
  __syncthreads();
  return r;
}
__global__ void __launch_bounds__(256) k_marginal(const JEng *E, const JSNode *sn, double *mass, int level, double *out) {
  const double p0 = marginal_block(E, sn, mass, level);
  if (threadIdx.x == 0) *out = p0;
}

// Simulator::GetProbability for every assignment of a set of qubits (the "probabilities" of a snapshot,
// QASMsimulator.cpp:990-1011).  bit_of_level[L] = index bit that fixes the qubit on level L, -1 = level not listed.
struct ProbParams {
  int8_t bit_of_level[JKQ_MAXQ];
};
// all qubits listed: one thread follows one path of conditional branch probabilities
__global__ void __launch_bounds__(256) k_probs_full(const JEng *E, const JSNode *sn, ProbParams pp, uint64_t count, double *out) {
  const int n = (int)E->n;
  const double norm = E->norm;
  const uint32_t root = E->root;
  for (uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; a < count; a += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t node = root;
    double p = norm;
    for (int L = n - 1; L >= 0 && p != 0.0; L--) {
      const JSNode s = sn[node];
      const int b = (int)((a >> pp.bit_of_level[L]) & 1ull);
      p *= b ? 1.0 - s.q0 : s.q0;
      node = b ? s.c1 : s.c0;
    }
    out[a] = p;
  }
}
// a subset: one CTA per assignment pushes the path mass down level by level (k_marginal with fixed branches on the
// listed levels) and stops at the lowest listed level -- below it the conditional probabilities sum to one
__global__ void __launch_bounds__(256) k_probs_subset(const JEng *E, const JSNode *sn, ProbParams pp, int low, uint64_t count,
                                                      double *scratch, double *out) {
  __shared__ double s_sum;
  const int n = (int)E->n;
  const uint32_t total = E->node_count;
  double *mass = scratch + (size_t)blockIdx.x * total;
  for (uint64_t a = blockIdx.x; a < count; a += gridDim.x) {
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) mass[i] = 0.0;
    if (threadIdx.x == 0) s_sum = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) mass[E->root] = E->norm;
    __syncthreads();
    for (int L = n - 1; L > low; L--) {
      const uint32_t off = E->lvl_off[L], cnt = E->lvl_cnt[L];
      const int bl = pp.bit_of_level[L];
      const int b = bl < 0 ? -1 : (int)((a >> bl) & 1ull);
      for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const JSNode s = sn[off + i];
        const double m = mass[off + i];
        if (m == 0.0) continue;
        if (b != 1 && s.q0 > 0.0) atomicAdd(&mass[s.c0], m * s.q0);
        if (b != 0 && s.q0 < 1.0) atomicAdd(&mass[s.c1], m * (1.0 - s.q0));
      }
      __syncthreads();
    }
    const uint32_t off = E->lvl_off[low], cnt = E->lvl_cnt[low];
    const int b = (int)((a >> pp.bit_of_level[low]) & 1ull);
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) acc += mass[off + i] * (b ? 1.0 - sn[off + i].q0 : sn[off + i].q0);
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_sum, acc);
    __syncthreads();
    if (threadIdx.x == 0) out[a] = s_sum;
    __syncthreads();
  }
}

struct SampleParams {
  uint8_t bitpos[JKQ_MAXQ];
};

#define JH_EMPTY 0xFFFFFFFFFFFFFFFFull
#define JH_SMEM 2048

// Count histogram in HBM: open addressing over FUSED 16-byte slots (key, count) -- the probe that finds (or claims)
// the key and the add that counts the shot touch the same 32-byte sector, where separate key / count arrays cost two
// to three sectors per shot.  The all-ones outcome of a 64-bit register collides with the empty
// marker: it is counted in the extra slot hs[slots] behind the table.
struct __align__(16) JHSlot {
  unsigned long long key, cnt;
};
// `fresh`: a counter in shared memory that takes one for every slot this call claims (the number of distinct outcomes
// then costs no pass over a 34 GB table: the kernels add their block's count to a device word at the end)
__device__ __forceinline__ void hist_global_add(JHSlot *hs, uint32_t hmask, uint64_t key, unsigned long long cnt, int *overflow,
                                                unsigned int *fresh) {
  if (key == JH_EMPTY) { atomicAdd(&hs[(size_t)hmask + 1].cnt, cnt); return; }
  uint32_t h = (uint32_t)jmix(key) & hmask;
  for (uint32_t probes = 0; probes <= hmask; probes++) {
    unsigned long long old = hs[h].key;
    if (old != key) {
      if (old != JH_EMPTY) { h = (h + 1) & hmask; continue; }
      old = atomicCAS(&hs[h].key, JH_EMPTY, (unsigned long long)key);
      if (old != JH_EMPTY && old != key) { h = (h + 1) & hmask; continue; }
      if (old == JH_EMPTY) atomicAdd(fresh, 1u);
    }
    atomicAdd(&hs[h].cnt, cnt);
    return;
  }
  *overflow = 1;
}

// PACKED format for registers of at most 32 classical bits (QFT-32: 2^32 outcomes) and fewer than 2^32 shots: one
// 64-bit word per slot, key << 32 | count.  A new outcome costs ONE atomic (the CAS that claims the slot carries the
// count), a repeated one a load and one add -- the fused format needs two atomics per shot, and the sampling kernel
// runs at the atomic throughput of the L2 (ncu r2: 19.6 G atomics/s against 20.8 G CAS/s measured).  The word of
// key 0xFFFFFFFF never equals the empty marker because its count stays below 2^32 - 1.
template <bool PK> __device__ __forceinline__ void hist_add(void *tab, uint32_t hmask, uint64_t key, unsigned long long cnt, int *overflow,
                                                           unsigned int *fresh) {
  if (!PK) { hist_global_add((JHSlot *)tab, hmask, key, cnt, overflow, fresh); return; }
  unsigned long long *t = (unsigned long long *)tab;
  uint32_t h = (uint32_t)jmix(key) & hmask;
  for (uint32_t probes = 0; probes <= hmask; probes++) {
    unsigned long long w = t[h];
    if (w == JH_EMPTY) {
      const unsigned long long old = atomicCAS(&t[h], JH_EMPTY, (key << 32) | cnt);
      if (old == JH_EMPTY) { atomicAdd(fresh, 1u); return; }
      w = old;
    }
    if ((w >> 32) == key) { atomicAdd(&t[h], cnt); return; }
    h = (h + 1) & hmask;
  }
  *overflow = 1;
}
// slot i of either format: false = empty
template <bool PK> __device__ __forceinline__ bool hist_get(const void *tab, uint64_t i, unsigned long long *key, unsigned long long *cnt) {
  if (PK) {
    const unsigned long long w = ((const unsigned long long *)tab)[i];
    if (w == JH_EMPTY) return false;
    *key = w >> 32; *cnt = w & 0xFFFFFFFFull;
    return true;
  }
  const JHSlot e = ((const JHSlot *)tab)[i];
  if (e.key == JH_EMPTY) return false;
  *key = e.key; *cnt = e.cnt;
  return true;
}
// the all-ones outcome of a 64-bit register (fused format only: extra slot behind the table)
template <bool PK> __device__ __forceinline__ unsigned long long hist_all_ones(const void *tab, uint64_t slots) {
  return PK ? 0ull : ((const JHSlot *)tab)[slots].cnt;
}

// MeasureAll x shots (Simulator.cpp:148-224, QASMsimulator.cpp:487-500): one thread per shot, n
// dependent 16-byte loads through L1/L2, equal outcomes merged inside the warp (__match_any_sync),
// then in a shared-memory hash, then in the HBM hash histogram.
// nodes of the sampling diagram with a real choice (0 < q0 < 1): decides whether k_sample draws its random numbers lazily
__global__ void k_sn_choice(const JSNode *__restrict__ sn, uint32_t n_nodes, unsigned int *out) {
  unsigned int mine = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x + 1; i < n_nodes; i += gridDim.x * blockDim.x) {  // (record 0 is the terminal)
    const double q0 = sn[i].q0;
    mine += (q0 > 0.0 && q0 < 1.0) ? 1u : 0u;
  }
  mine = __reduce_add_sync(0xFFFFFFFFu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, mine);
}

template <bool PK, bool LAZY>
__global__ void __launch_bounds__(256) k_sample(const JSNode *__restrict__ sn, uint32_t root, int n, SampleParams sp, uint64_t seed,
                                               uint64_t shot0, uint64_t nshots, void *hs, uint32_t hmask, int *overflow,
                                               unsigned long long *n_distinct) {
  __shared__ unsigned long long skeys[JH_SMEM];
  __shared__ unsigned int scnts[JH_SMEM];
  __shared__ unsigned int s_fresh;
  if (threadIdx.x == 0) s_fresh = 0;
  for (int i = threadIdx.x; i < JH_SMEM; i += blockDim.x) { skeys[i] = JH_EMPTY; scnts[i] = 0; }
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t rounded = (nshots + 31) & ~31ull;
  const uint32_t lane = threadIdx.x & 31;
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < rounded; s += stride) {
    const bool active = s < nshots;
    uint64_t key = active ? sample_walk<LAZY>(sn, root, n, sp.bitpos, seed, shot0 + s) : JH_EMPTY;
    uint32_t peers = __match_any_sync(0xFFFFFFFFu, key) & __ballot_sync(0xFFFFFFFFu, active);
    if (!active || lane != (uint32_t)(__ffs(peers) - 1)) continue;
    uint32_t cnt = __popc(peers);
    if (key == JH_EMPTY) { hist_add<PK>(hs, hmask, key, cnt, overflow, &s_fresh); continue; }
    uint32_t h = (uint32_t)jmix(key) & (JH_SMEM - 1);
    bool done = false;
    for (int p = 0; p < 8 && !done; p++) {
      unsigned long long old = skeys[h];
      if (old != key) {
        if (old == JH_EMPTY) old = atomicCAS(&skeys[h], JH_EMPTY, (unsigned long long)key);
        if (old != JH_EMPTY && old != key) { h = (h + 1) & (JH_SMEM - 1); continue; }
      }
      atomicAdd(&scnts[h], cnt);
      done = true;
    }
    if (!done) hist_add<PK>(hs, hmask, key, cnt, overflow, &s_fresh);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < JH_SMEM; i += blockDim.x)
    if (skeys[i] != JH_EMPTY && scnts[i]) hist_add<PK>(hs, hmask, skeys[i], scnts[i], overflow, &s_fresh);
  __syncthreads();
  if (threadIdx.x == 0 && s_fresh) atomicAdd(n_distinct, (unsigned long long)s_fresh);
}

template <bool PK> __global__ void k_hist_fill(void *hs, uint64_t slots, unsigned long long *n_distinct) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_distinct = 0;
  if (PK) {
    unsigned long long *t = (unsigned long long *)hs;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= slots; i += (uint64_t)gridDim.x * blockDim.x) t[i] = JH_EMPTY;
    return;
  }
  JHSlot e;
  e.key = JH_EMPTY; e.cnt = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= slots; i += (uint64_t)gridDim.x * blockDim.x) ((JHSlot *)hs)[i] = e;
}

template <bool PK>
__global__ void k_hist_compact(const void *hs, uint64_t slots, unsigned long long *ok, unsigned long long *oc, unsigned long long *n_out,
                               uint64_t cap) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long key, cnt;
    if (hist_get<PK>(hs, i, &key, &cnt)) {
      unsigned long long p = atomicAdd(n_out, 1ull);
      if (p < cap) { ok[p] = key; oc[p] = cnt; }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && hist_all_ones<PK>(hs, slots)) {
    unsigned long long p = atomicAdd(n_out, 1ull);
    if (p < cap) { ok[p] = JH_EMPTY; oc[p] = hist_all_ones<PK>(hs, slots); }
  }
}

// ---- sparse histogram exchange between GPUs ---------------------------------------------------
// A histogram with up to 10^9 distinct outcomes does not fit a dense all-reduce: every outcome key has an OWNER rank
// (hash of the key, independent of the slot hash), each rank sorts its (key, count) pairs by owner into one send
// buffer (count pass, prefix on the host, scatter pass), the pairs travel in one all-to-all, and the owner adds what
// it received into its histogram with the same insert the sampler uses.
__device__ __forceinline__ uint32_t hist_owner(uint64_t key, uint32_t world) { return (uint32_t)((jmix(key ^ 0x5851F42D4C957F2Dull) >> 32) % world); }

template <bool PK>
__global__ void __launch_bounds__(256) k_hist_owner_count(const void *hs, uint64_t slots, uint32_t world, unsigned long long *per_owner) {
  __shared__ unsigned int s_cnt[64];
  for (uint32_t i = threadIdx.x; i < 64; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (uint64_t)gridDim.x * blockDim.x)
  {
    unsigned long long key, cnt;
    if (hist_get<PK>(hs, i, &key, &cnt)) atomicAdd(&s_cnt[hist_owner(key, world)], 1u);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && hist_all_ones<PK>(hs, slots)) atomicAdd(&s_cnt[hist_owner(JH_EMPTY, world)], 1u);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < world; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(per_owner + i, (unsigned long long)s_cnt[i]);
}

template <bool PK>
__global__ void __launch_bounds__(256) k_hist_owner_scatter(const void *hs, uint64_t slots, uint32_t world, unsigned long long *cursor,
                                                            unsigned long long *out_keys, unsigned long long *out_counts) {
  // one atomic per (warp, owner) instead of one per pair: the `world` cursors are the only contended words
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, rounded = (slots + 31) & ~31ull;
  const uint32_t lane = threadIdx.x & 31;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += stride) {
    unsigned long long key = JH_EMPTY, ecnt = 0;
    if (i < slots && !hist_get<PK>(hs, i, &key, &ecnt)) key = JH_EMPTY;
    const bool live = key != JH_EMPTY;
    const uint32_t owner = live ? hist_owner(key, world) : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, owner);
    if (!live) continue;
    const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(cursor + owner, (unsigned long long)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const unsigned long long p = base + __popc(peers & ((1u << lane) - 1u));
    out_keys[p] = key; out_counts[p] = ecnt;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && hist_all_ones<PK>(hs, slots)) {
    const unsigned long long p = atomicAdd(cursor + hist_owner(JH_EMPTY, world), 1ull);
    out_keys[p] = JH_EMPTY; out_counts[p] = hist_all_ones<PK>(hs, slots);
  }
}

template <bool PK>
__global__ void __launch_bounds__(256) k_hist_merge(void *hs, uint32_t hmask, const unsigned long long *in_keys, const unsigned long long *in_counts,
                                                    uint64_t n, int *overflow, unsigned long long *n_distinct) {
  __shared__ unsigned int s_fresh;
  if (threadIdx.x == 0) s_fresh = 0;
  __syncthreads();
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    if (in_counts[i]) hist_add<PK>(hs, hmask, in_keys[i], in_counts[i], overflow, &s_fresh);
  __syncthreads();
  if (threadIdx.x == 0 && s_fresh) atomicAdd(n_distinct, (unsigned long long)s_fresh);
}

// Roofline denominators measured on the device in use: uniform random 32-byte sector reads
// (ld.global.cg, one sector per load) and random 64-bit atomicCAS.
__global__ void k_random_sector(const uint4 *buf, uint64_t n_sectors2, int per_thread, uint64_t seed, unsigned long long *sink) {
  uint64_t x = jmix(seed + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
  uint32_t acc = 0;
  for (int i = 0; i < per_thread; i++) {
    x = x * 6364136223846793005ull + 1442695040888963407ull;
    uint64_t idx = (x >> 20) % n_sectors2;
    uint4 v = __ldcg(buf + idx * 2);  // 16 bytes out of a distinct 32-byte sector
    acc += v.x ^ v.w;
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}
__global__ void k_random_cas(unsigned long long *buf, uint64_t n_words, int per_thread, uint64_t seed) {
  uint64_t x = jmix(seed + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
  for (int i = 0; i < per_thread; i++) {
    x = x * 6364136223846793005ull + 1442695040888963407ull;
    uint64_t idx = (x >> 20) % n_words;
    atomicCAS(buf + idx, 0ull, x | 1ull);
  }
}

// =============================================================================== host side
struct jkq_handle {
  jkq_config cfg;
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  std::string err;
  JEng host;       // host mirror of the control block
  JEng *dev = nullptr;
  void *slab = nullptr;  // the one device allocation behind the arenas, tables, control block and sampling records
  JSNode *snodes = nullptr;
  double *prob = nullptr;
  uint32_t *lvl_log = nullptr;
  JGate *gates_dev = nullptr;
  uint64_t *mat_u = nullptr;
  uint32_t *mat_epoch = nullptr;
  uint32_t n_mats = 0, mats_cap = 0;
  size_t gates_cap = 0;
  JGate *gates_pin = nullptr;       // pinned staging copy of `gates` (truly asynchronous H2D)
  std::vector<jkq_op> ops;
  std::vector<JGate> gates;         // GATE ops only
  std::vector<int> gate_of_op;      // ops index -> number of gates before it
  uint64_t measure_seed = 1;        // Philox key of the mid-circuit measurement outcomes
  uint64_t mid_bits = 0, mid_mask = 0;  // outcomes of MEASURE_ONE ops by classical bit
  double *scalar_dev = nullptr;
  uint32_t n = 0;
  bool loaded = false, applied = false, sampling_ready = false, imported = false, level_log = false;
  uint64_t h2d_bytes = 0, d2h_bytes = 0;  // bytes copied by this handle since it was created (jcopy)
  uint32_t sums_ask = 8;            // gates that would have used thin_probe_sums before the kernel that has it is launched (option sums_ask)
  bool sums_now = false;            // ... and it is the build with thin_probe_sums (it asked for it: want_sums)
  bool wide_now = false;            // the single-CTA kernel of this circuit is launched with cfg.cta_threads (it asked for them)
  SampleParams sp;
  uint32_t dd_root = 0, dd_nodes = 0;
  std::map<uint32_t, double2 *> snapshots;  // slot -> device buffer of 2^n amplitudes
  // Above snapshot_dense_max qubits a snapshot keeps the DIAGRAM instead of 2^n amplitudes: a copy of the node arena and
  // the reals as they are when the statement is reached (16 B per node, 8 B per real) plus a control block that names
  // them; amplitudes are produced range by range when they are read (jkq_get_snapshot_amplitudes)
  struct DDSnap { JNode *nodes = nullptr; double *vals = nullptr; JEng *eng = nullptr; uint32_t n_nodes = 0, n_vals = 0; };
  std::map<uint32_t, DDSnap> dd_snapshots;
  uint32_t snapshot_dense_max = 26;
  std::map<uint32_t, std::pair<double *, uint32_t>> prob_slots;  // slot -> (device buffer of 2^k probabilities, k)
  bool physical_reset = false;      // RESET = measure-then-X instead of the reference's projection
  uint64_t placeholder_gates = 0;   // projector / X slots of MEASURE_ONE and RESET applied by the last jkq_apply
  bool debug_stop_at_handover = false;  // testing aid: jkq_apply returns where the single-CTA path would hand over
  // histogram
  void *hs = nullptr;               // hslots + 1 slots: fused 16-byte (key, count), or packed 8-byte key << 32 | count
  bool hist_packed = false;         // packed: every classical bit below 32 and fewer than 2^32 shots in the histogram
  uint64_t hist_shots = 0;          // shots counted into the histogram so far
  bool force_fused = false;         // option "histogram_fused": never use the packed format (A/B runs, tests)
  unsigned long long *ok = nullptr, *oc = nullptr, *n_out = nullptr;
  unsigned long long *n_distinct = nullptr;  // slots claimed since the histogram was cleared (k_sample / k_hist_merge)
  int *overflow = nullptr;
  unsigned int *choice_dev = nullptr;  // k_sn_choice's counter
  uint64_t sn_epoch = 1, sn_choice_epoch = 0;  // the sampling records changed / were last looked at by k_sn_choice
  bool sn_lazy = false;
  uint64_t hslots = 0, out_cap = 0;
  bool hist_valid = false;
  jkq_stats stats;
  int grid_blocks = 0;
};

// every host <-> device copy of a handle goes through here: the bytes are counted (jkq_stats.h2d_bytes / d2h_bytes,
// what bench.py reports as the measured copies of a call)
static cudaError_t jcopy(jkq_handle *h, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) {
  if (kind == cudaMemcpyHostToDevice) h->h2d_bytes += bytes;
  else if (kind == cudaMemcpyDeviceToHost) h->d2h_bytes += bytes;
  return cudaMemcpyAsync(dst, src, bytes, kind, h->stream);
}

static int fail(jkq_handle *h, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}
#define CU(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) return fail(h, JKQ_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static uint32_t pow2_at_least(uint64_t v, uint32_t lo) {
  uint32_t p = lo;
  while (p < v && p < 0x80000000u) p <<= 1;
  return p;
}
static const char *status_text(int s) {
  switch (s) {
    case JKQ_ERR_NODE_FULL: return "node arena full (raise jkq_config.node_capacity)";
    case JKQ_ERR_UTAB_FULL: return "unique table full";
    case JKQ_ERR_VAL_FULL: return "complex-value table full (raise jkq_config.value_capacity)";
    case JKQ_ERR_ITEM_FULL: return "frontier pool full (raise jkq_config.item_capacity)";
    case JKQ_ERR_DTAB_FULL: return "frontier de-dup table full";
    case JKQ_ERR_PEND_FULL: return "pending-value buffer full";
    case JKQ_ERR_BARRIER_TIMEOUT: return "grid barrier timeout";
    case JKQ_ERR_ZERO_STATE: return "numerical instabilities led to a 0-vector";
    case JKQ_ERR_RETRY_LIMIT: return "a level kept producing new values (retry limit of the value table reached)";
    case JKQ_ERR_VAL_RANGE: return "a weight of absurd magnitude (>= 1e8) or NaN: numerical breakdown";
    default: return "internal error";
  }
}

extern "C" uint32_t jkq_abi_version(void) { return JKQ_ABI_VERSION; }
extern "C" const char *jkq_last_error(const jkq_handle *h) { return h ? h->err.c_str() : "null handle"; }

static void free_hist(jkq_handle *h) {
  cudaFree(h->hs); cudaFree(h->ok); cudaFree(h->oc);
  h->hs = nullptr;
  h->hist_packed = false;
  h->ok = h->oc = nullptr;
  h->hslots = 0; h->out_cap = 0; h->hist_valid = false;
}
static void free_snapshots(jkq_handle *h) {
  for (auto &kv : h->snapshots) cudaFree(kv.second);
  h->snapshots.clear();
  for (auto &kv : h->prob_slots) cudaFree(kv.second.first);
  h->prob_slots.clear();
  for (auto &kv : h->dd_snapshots) { cudaFree(kv.second.nodes); cudaFree(kv.second.vals); cudaFree(kv.second.eng); }
  h->dd_snapshots.clear();
}

static void multi_free(jkq_handle *src);
extern "C" void jkq_destroy(jkq_handle *h) {
  if (h) multi_free(h);
  if (!h) return;
  cudaSetDevice(h->device);
  JEng &E = h->host;
  (void)E;
  cudaFree(h->slab);  // arenas, tables, control block, sampling records: one allocation (jkq_create)
  cudaFree(h->lvl_log); cudaFree(h->gates_dev);
  cudaFree(h->mat_u); cudaFree(h->mat_epoch);
  if (h->gates_pin) cudaFreeHost(h->gates_pin);
  free_hist(h);
  free_snapshots(h);
  if (h->ev0) { cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->ev2); cudaEventDestroy(h->ev3); }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" int jkq_create(jkq_handle **out, const jkq_config *cfg_in) {
  if (!out) return JKQ_E_ARGUMENT;
  *out = nullptr;
  jkq_handle *h = new jkq_handle();
  jkq_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  if (cfg_in) cfg = *cfg_in;
  if (cfg.tolerance <= 0) cfg.tolerance = 1e-10;
  cfg.node_capacity = pow2_at_least(cfg.node_capacity ? cfg.node_capacity : (1u << 22), 1u << 10);
  cfg.value_capacity = pow2_at_least(cfg.value_capacity ? cfg.value_capacity : (1u << 20), 1u << 10);
  cfg.item_capacity = pow2_at_least(cfg.item_capacity ? cfg.item_capacity : cfg.node_capacity, 1u << 10);
  if (!cfg.cta_threads) cfg.cta_threads = 256;
  if (cfg.cta_threads > 256) cfg.cta_threads = 256;
  cfg.cta_threads = (cfg.cta_threads + 31) & ~31u;
  if (!cfg.cta_item_limit || cfg.cta_item_limit > JKQ_SITEMS / 2) cfg.cta_item_limit = JKQ_SITEMS / 2;
  if (!cfg.gc_min) cfg.gc_min = 1u << 15;
  if (!cfg.grid_blocks_per_sm) cfg.grid_blocks_per_sm = JKQ_GRID_MINB;
  h->cfg = cfg;
  h->device = cfg.device;
  memset(&h->stats, 0, sizeof(h->stats));
  memset(&h->host, 0, sizeof(h->host));
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0 || cfg.device >= ndev) {
    // no CPU fallback: the product path is CUDA only
    fprintf(stderr, "jkq_create: no usable CUDA device (%s)\n", e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range");
    delete h;
    return JKQ_E_CUDA;
  }
  *out = h;  // from here on the caller can read jkq_last_error and must jkq_destroy
  CU(cudaSetDevice(h->device));
  CU(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device));
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&h->ev0)); CU(cudaEventCreate(&h->ev1)); CU(cudaEventCreate(&h->ev2)); CU(cudaEventCreate(&h->ev3));
  JEng &E = h->host;
  E.tol = cfg.tolerance;
  E.inv_tol = 1.0 / cfg.tolerance;
  E.node_cap = cfg.node_capacity; E.utab_mask = cfg.node_capacity * 2 - 1;
  E.val_cap = cfg.value_capacity; E.vtab_mask = cfg.value_capacity * 4 - 1;
  E.pend_cap = cfg.value_capacity;
  E.item_cap = cfg.item_capacity; E.dtab_mask = cfg.item_capacity * 2 - 1;
  E.cta_item_limit = cfg.cta_item_limit; E.gc_min = cfg.gc_min;
  // One slab per handle, carved into the arenas and tables (a pool of 148 small engines would otherwise spend
  // most of its creation time in ~25 driver allocations each).
  {
    struct Part { void **ptr; size_t bytes; };
    const Part parts[] = {
        {(void **)&E.nodes, (size_t)E.node_cap * sizeof(JNode)}, {(void **)&E.nodes_alt, (size_t)E.node_cap * sizeof(JNode)},
        {(void **)&E.utab, ((size_t)E.utab_mask + 1) * 8}, {(void **)&E.vals, (size_t)E.val_cap * 8}, {(void **)&E.vals_alt, (size_t)E.val_cap * 8},
        {(void **)&E.vtab, ((size_t)E.vtab_mask + 1) * sizeof(JVSlot)}, {(void **)&E.vbm, ((((size_t)E.vtab_mask + 1) >> 8) + 1) * 4},
        {(void **)&E.pend, (size_t)E.pend_cap * 4},
        {(void **)&E.items, (size_t)E.item_cap * sizeof(JItem)}, {(void **)&E.dtab, ((size_t)E.dtab_mask + 1) * 8},
        {(void **)&E.mdir, (size_t)E.node_cap * 8}, {(void **)&E.thin_spill, sizeof(ThinState)}, {(void **)&E.fwd, (size_t)E.node_cap * 4},
        {(void **)&E.vfwd, (size_t)E.val_cap * 4}, {(void **)&h->dev, sizeof(JEng)}, {(void **)&h->snodes, (size_t)E.node_cap * sizeof(JSNode)},
        {(void **)&h->prob, (size_t)E.node_cap * 8}, {(void **)&h->n_out, 8}, {(void **)&h->n_distinct, 8}, {(void **)&h->overflow, 4}, {(void **)&h->choice_dev, 4}, {(void **)&h->scalar_dev, 8}};
    size_t total = 0;
    for (const Part &p : parts) total += (p.bytes + 255) & ~(size_t)255;
    CU(cudaMalloc(&h->slab, total));
    size_t off = 0;
    for (const Part &p : parts) { *p.ptr = (char *)h->slab + off; off += (p.bytes + 255) & ~(size_t)255; }
  }
  // generation-tagged de-dup tables: zeroed once, never cleared again (the generation counts across circuits)
  CU(cudaMemsetAsync(E.utab, 0, ((size_t)E.utab_mask + 1) * 8, h->stream));  // (generation 0 = never written)
  E.ut_gen = 1; E.vt_cur = E.vtab_mask; E.vt_dirty_all = 1;
  CU(cudaMemsetAsync(E.dtab, 0, ((size_t)E.dtab_mask + 1) * 8, h->stream));
  CU(cudaMemsetAsync(E.mdir, 0, (size_t)E.node_cap * 8, h->stream));
  CU(cudaFuncSetAttribute(k_run_gates<MODE_CTA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaSmem)));
  CU(cudaFuncSetAttribute(k_run_gates<MODE_CTA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaSmem)));
  CU(cudaFuncSetAttribute(k_build_sampling<MODE_CTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaSmem)));
  // cooperative grid size: every CTA must be resident
  int per_sm = 0;
  CU(cudaFuncSetAttribute(k_run_gates<MODE_GRID, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ThinState)));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_run_gates<MODE_GRID, true>, 256, sizeof(ThinState)));
  if (per_sm < 1) return fail(h, JKQ_E_CUDA, "grid kernel does not fit on an SM");
  if ((uint32_t)per_sm > cfg.grid_blocks_per_sm) per_sm = (int)cfg.grid_blocks_per_sm;
  h->grid_blocks = per_sm * h->sm_count;
  return JKQ_E_OK;
}

extern "C" int jkq_set_option(jkq_handle *h, const char *name, int64_t value) {
  if (!h || !name) return JKQ_E_ARGUMENT;
  if (!strcmp(name, "level_log")) { h->level_log = value != 0; return JKQ_E_OK; }
  if (!strcmp(name, "force_mode")) { h->cfg.force_mode = (int)value; return JKQ_E_OK; }
  if (!strcmp(name, "measure_seed")) { h->measure_seed = (uint64_t)value; return JKQ_E_OK; }
  if (!strcmp(name, "physical_reset")) { h->physical_reset = value != 0; return JKQ_E_OK; }
  if (!strcmp(name, "cta_node_limit")) { h->host.cta_node_limit = (uint32_t)value; return JKQ_E_OK; }
  if (!strcmp(name, "snapshot_dense_max")) { h->snapshot_dense_max = value < 0 ? 0u : (value > 30 ? 30u : (uint32_t)value); return JKQ_E_OK; }
  if (!strcmp(name, "solo_max")) { h->host.solo_max = (uint32_t)value; return JKQ_E_OK; }
  if (!strcmp(name, "cta_always_wide")) { h->cfg.reserved = value ? 1u : 0u; return JKQ_E_OK; }
  if (!strcmp(name, "cta_no_reserve")) { h->host.cta_no_reserve = (uint32_t)value; return JKQ_E_OK; }
  if (!strcmp(name, "histogram_fused")) { h->force_fused = value != 0; return JKQ_E_OK; }
  if (!strcmp(name, "thin_off")) { h->host.thin_off = (uint32_t)value; return JKQ_E_OK; }
  if (!strcmp(name, "sums_ask")) { h->sums_ask = (uint32_t)value; return JKQ_E_OK; }  // testing aid: JEng.sums_ask at the start of a run (0: never)
  if (!strcmp(name, "cta_val_limit")) { h->host.cta_val_limit = (uint32_t)value; return JKQ_E_OK; }
  if (!strcmp(name, "debug_stop_at_handover")) { h->debug_stop_at_handover = value != 0; return JKQ_E_OK; }
  if (!strcmp(name, "cta_item_limit")) {
    if (value < 1 || value > JKQ_SITEMS / 2) value = JKQ_SITEMS / 2;
    h->cfg.cta_item_limit = (uint32_t)value; h->host.cta_item_limit = (uint32_t)value;
    return JKQ_E_OK;
  }
  return fail(h, JKQ_E_ARGUMENT, "unknown option %s", name);
}

// ops -> gate descriptors (level roles, index of the matrix among the distinct matrices of the circuit) and the
// measurement map.  Returns 0 or JKQ_E_ARGUMENT with a message in *err.
// idle-level mask and parking level of the gates [first, gates.size()) of one circuit (JGate.idle / JGate.park)
static void finish_gates(std::vector<JGate> &gates, size_t first) {
  for (size_t i = first; i < gates.size(); i++) {
    JGate &g = gates[i];
    g.idle = 0;
    for (int L = 0; L < JKQ_MAXQ; L++)
      if (g.role[L] == JR_IDLE) g.idle |= 1ull << L;
    g.park = g.top;
    for (size_t j = i + 1; j < gates.size() && j <= i + JKQ_PARK_WINDOW; j++)
      if (gates[j].top > g.park) g.park = gates[j].top;
    // diagonal or anti-diagonal: local on any node (thin_gate).  (The projector / X slots of MEASURE_ONE and RESET
    // are filled in later with matrices of the same two shapes.)
    const bool diag = g.m[2] == 0.0 && g.m[3] == 0.0 && g.m[4] == 0.0 && g.m[5] == 0.0;
    const bool anti = g.m[0] == 0.0 && g.m[1] == 0.0 && g.m[6] == 0.0 && g.m[7] == 0.0;
    g.rsv = (diag || anti) ? 1 : 0;
  }
}

static int parse_ops(std::string *err, uint32_t n_qubits, const jkq_op *ops, size_t n_ops, std::vector<JGate> &gates, std::vector<int> *gate_of_op,
                     SampleParams &sp, uint32_t *n_mats, bool *has_snapshot) {
  char buf[128];
  const size_t first_gate = gates.size();
  if (n_qubits < 1 || n_qubits > JKQ_MAXQ) { snprintf(buf, sizeof(buf), "n_qubits must be 1..%d", JKQ_MAXQ); *err = buf; return JKQ_E_ARGUMENT; }
  bool any_measure = false;
  if (has_snapshot) *has_snapshot = false;
  std::vector<std::array<double, 8>> mats;
  for (int L = 0; L < JKQ_MAXQ; L++) sp.bitpos[L] = 255;
  const uint64_t qmask = n_qubits == 64 ? ~0ull : ((1ull << n_qubits) - 1);
  for (size_t i = 0; i < n_ops; i++) {
    const jkq_op &o = ops[i];
    if (gate_of_op) gate_of_op->push_back((int)gates.size());
    if (o.kind == JKQ_OP_GATE) {
      if (o.target >= n_qubits || ((o.ctrl_pos | o.ctrl_neg) & ~qmask) || ((o.ctrl_pos | o.ctrl_neg) >> o.target & 1) ||
          (o.ctrl_pos & o.ctrl_neg)) {
        snprintf(buf, sizeof(buf), "op %zu: bad target/control qubits", i); *err = buf;
        return JKQ_E_ARGUMENT;
      }
      JGate g;
      memset(&g, 0, sizeof(g));
      memcpy(g.m, o.m, sizeof(g.m));
      g.target = (int)(n_qubits - 1 - o.target);  // qubit q -> level n-1-q, QASMsimulator.cpp:300
      g.low = g.top = g.target;
      g.role[g.target] = JR_TARGET;
      for (uint32_t q = 0; q < n_qubits; q++) {
        int L = (int)(n_qubits - 1 - q);
        if (o.ctrl_pos >> q & 1) g.role[L] = JR_CPOS;
        else if (o.ctrl_neg >> q & 1) g.role[L] = JR_CNEG;
        else continue;
        if (L < g.low) g.low = L;
        if (L > g.top) g.top = L;
      }
      // index among the distinct matrices of this circuit (bit-exact compare of the 8 doubles)
      uint32_t mid = 0;
      for (; mid < mats.size(); mid++)
        if (!memcmp(mats[mid].data(), o.m, sizeof(o.m))) break;
      if (mid == mats.size()) { std::array<double, 8> a; memcpy(a.data(), o.m, sizeof(o.m)); mats.push_back(a); }
      g.pad = (int32_t)mid;
      gates.push_back(g);
    } else if (o.kind == JKQ_OP_MEASURE) {
      if (o.target >= n_qubits || o.arg >= 64) { snprintf(buf, sizeof(buf), "op %zu: bad measure", i); *err = buf; return JKQ_E_ARGUMENT; }
      sp.bitpos[n_qubits - 1 - o.target] = (uint8_t)o.arg;
      any_measure = true;
    } else if (o.kind == JKQ_OP_SNAPSHOT) {
      if (has_snapshot) *has_snapshot = true;
    } else if (o.kind == JKQ_OP_MEASURE_ONE || o.kind == JKQ_OP_RESET) {
      // MeasureOne / ResetQubit (Simulator.cpp:226-289, 341-413): one (two) single-qubit gate slot(s) whose
      // matrix the host fills in once the outcome is drawn -- the projector scaled by 1/sqrt(p), and for a
      // reset the X that takes |1> back to |0>
      if (o.target >= n_qubits || (o.kind == JKQ_OP_MEASURE_ONE && o.arg >= 64)) { snprintf(buf, sizeof(buf), "op %zu: bad measure/reset", i); *err = buf; return JKQ_E_ARGUMENT; }
      if (has_snapshot) *has_snapshot = true;  // (not served by the batch path either)
      for (int rep = 0; rep < (o.kind == JKQ_OP_RESET ? 2 : 1); rep++) {
        JGate g;
        memset(&g, 0, sizeof(g));
        g.m[0] = 1.0; g.m[6] = 1.0;
        g.target = (int)(n_qubits - 1 - o.target);
        g.low = g.top = g.target;
        g.role[g.target] = JR_TARGET;
        std::array<double, 8> a;  // a matrix slot of its own: the entries change from run to run
        for (int k = 0; k < 8; k++) a[k] = -1.0 - (double)mats.size() - 0.001 * k;
        mats.push_back(a);
        g.pad = (int32_t)(mats.size() - 1);
        gates.push_back(g);
      }
    } else if (o.kind == JKQ_OP_IF) {
      const size_t cnt = o.arg ? o.arg : 1;
      bool ok = i + cnt < n_ops;
      for (size_t j = 1; ok && j <= cnt; j++) {
        const uint32_t nk = ops[i + j].kind;
        ok = nk == JKQ_OP_GATE || nk == JKQ_OP_MEASURE_ONE || nk == JKQ_OP_RESET || nk == JKQ_OP_BARRIER;
      }
      if (!ok) {
        snprintf(buf, sizeof(buf), "op %zu: IF must be followed by the gate, measure-one or reset ops it guards", i); *err = buf;
        return JKQ_E_ARGUMENT;
      }
      if (has_snapshot) *has_snapshot = true;  // (needs the host between segments: not served by the batch path)
    } else if (o.kind == JKQ_OP_PROBABILITIES) {
      if (o.ctrl_pos == 0 || (o.ctrl_pos & ~qmask) || __builtin_popcountll(o.ctrl_pos) > 26) {
        snprintf(buf, sizeof(buf), "op %zu: probabilities need 1..26 qubits of the circuit", i); *err = buf;
        return JKQ_E_ARGUMENT;
      }
      if (has_snapshot) *has_snapshot = true;
    } else if (o.kind != JKQ_OP_BARRIER) {
      snprintf(buf, sizeof(buf), "op %zu: unknown kind %u", i, o.kind); *err = buf;
      return JKQ_E_ARGUMENT;
    }
  }
  if (!any_measure)  // engine-level default: every qubit, key bit q = qubit q
    for (uint32_t q = 0; q < n_qubits && q < 64; q++) sp.bitpos[n_qubits - 1 - q] = (uint8_t)q;
  if (n_mats) *n_mats = (uint32_t)mats.size();
  finish_gates(gates, first_gate);
  return JKQ_E_OK;
}

extern "C" int jkq_load_circuit(jkq_handle *h, uint32_t n_qubits, const jkq_op *ops, size_t n_ops) {
  if (!h) return JKQ_E_ARGUMENT;
  if (n_ops && !ops) return JKQ_E_ARGUMENT;
  CU(cudaSetDevice(h->device));
  h->gates.clear();
  h->gate_of_op.clear();
  uint32_t n_mats = 0;
  {
    std::string err;
    int rc = parse_ops(&err, n_qubits, ops, n_ops, h->gates, &h->gate_of_op, h->sp, &n_mats, nullptr);
    if (rc) return fail(h, rc, "%s", err.c_str());
  }
  h->n = n_qubits;
  h->ops.assign(ops, ops + n_ops);
  // device copies of the gate list: buffers are kept and only grow (cudaFree synchronises the whole device,
  // which would serialise the engines of a batch)
  h->n_mats = n_mats;
  if (h->n_mats > h->mats_cap) {
    cudaFree(h->mat_u); cudaFree(h->mat_epoch);
    h->mat_u = nullptr; h->mat_epoch = nullptr;
    h->mats_cap = pow2_at_least(h->n_mats, 64);
    CU(cudaMalloc(&h->mat_u, (size_t)h->mats_cap * 4 * 8));
    CU(cudaMalloc(&h->mat_epoch, (size_t)h->mats_cap * 4));
    // never cleared again: an entry is valid only if it carries the CURRENT value epoch, which init bumps per circuit
    CU(cudaMemsetAsync(h->mat_epoch, 0, (size_t)h->mats_cap * 4, h->stream));
  }
  if (h->gates.size() > h->gates_cap) {
    cudaFree(h->gates_dev);
    if (h->gates_pin) cudaFreeHost(h->gates_pin);
    h->gates_dev = nullptr; h->gates_pin = nullptr;
    h->gates_cap = pow2_at_least(h->gates.size(), 256);
    CU(cudaMalloc(&h->gates_dev, h->gates_cap * sizeof(JGate)));
    CU(cudaMallocHost(&h->gates_pin, h->gates_cap * sizeof(JGate)));
  }
  if (!h->gates.empty()) {
    // (the previous circuit's copy has completed: every path synchronises the stream before a new load)
    memcpy(h->gates_pin, h->gates.data(), h->gates.size() * sizeof(JGate));
    CU(jcopy(h, h->gates_dev, h->gates_pin, h->gates.size() * sizeof(JGate), cudaMemcpyHostToDevice));
  }
  h->loaded = true; h->applied = false; h->sampling_ready = false; h->imported = false; h->hist_valid = false;
  return JKQ_E_OK;
}

static int pull_state(jkq_handle *h) {
  CU(jcopy(h, &h->host, h->dev, sizeof(JEng), cudaMemcpyDeviceToHost));
  CU(cudaStreamSynchronize(h->stream));
  return JKQ_E_OK;
}

// the grid barrier's counter and flag ring must start clean at every cooperative launch
static int reset_barrier(jkq_handle *h) {
  CU(cudaMemsetAsync(&h->dev->bar_count, 0, sizeof(unsigned int), h->stream));
  CU(cudaMemsetAsync(&h->dev->bar_flags[0], 0, 3 * sizeof(unsigned int), h->stream));
  CU(cudaMemsetAsync(&h->dev->pend_n[0], 0, 3 * sizeof(uint32_t), h->stream));
  return JKQ_E_OK;
}

static int launch_gates(jkq_handle *h, bool grid, int upto) {
  if (grid) {
    int rb = reset_barrier(h);
    if (rb) return rb;
    JEng *E = h->dev;
    const JGate *g = h->gates_dev;
    void *args[] = {&E, &g, &upto};
    CU(cudaLaunchCooperativeKernel((void *)k_run_gates<MODE_GRID, true>, dim3(h->grid_blocks), dim3(256), args, sizeof(ThinState), h->stream));
    h->stats.grid_launches++;
  } else {
    // one warp while every level of every gate fits it (QFT, GHZ, Grover ...: the other warps would only wait at the
    // gate barrier); the wide shape once the kernel asked for it (want_wide)
    const unsigned threads = (h->wide_now || h->cfg.reserved == 1) ? h->cfg.cta_threads : 32u;
    if (h->sums_now) k_run_gates<MODE_CTA, true><<<1, threads, sizeof(CtaSmem), h->stream>>>(h->dev, h->gates_dev, upto);
    else k_run_gates<MODE_CTA, false><<<1, threads, sizeof(CtaSmem), h->stream>>>(h->dev, h->gates_dev, upto);
    CU(cudaGetLastError());
    h->stats.cta_launches++;
  }
  return JKQ_E_OK;
}

// run gates [host.next_gate, upto): single-CTA path until a frontier outgrows it, then the grid path
static int run_segment(jkq_handle *h, int upto) {
  if (h->host.next_gate >= upto && !h->host.cursor_saved) return JKQ_E_OK;  // nothing left to apply
  bool grid = h->cfg.force_mode == 2;
  for (;;) {
    int rc = launch_gates(h, grid, upto);
    if (rc) return rc;
    rc = pull_state(h);
    if (rc) return rc;
    if (h->host.status == JKQ_ERR_BARRIER_TIMEOUT) {
      std::string s;
      char tmp[64];
      for (int b = 0; b < h->grid_blocks && b < 1024; b++) {
        if (b == 0 || h->host.dbg_phase[b] != h->host.dbg_phase[b - 1]) {
          snprintf(tmp, sizeof(tmp), " b%d:%u", b, h->host.dbg_phase[b]);
          s += tmp;
        }
      }
      return fail(h, h->host.status, "gate %d: grid barrier timeout (block %u gave up; bar_count=%u; last phase per block:%s)",
                  h->host.next_gate, h->host.dbg_timeout_block, h->host.bar_count, s.c_str());
    }
    if (h->host.status) return fail(h, h->host.status, "gate %d: %s", h->host.next_gate, status_text(h->host.status));
    if (h->host.next_gate >= upto && !h->host.cursor_saved) return JKQ_E_OK;
    if (h->host.want_grid && h->host.want_sums && !h->host.cursor_saved && !grid && !h->sums_now) {
      h->sums_now = true;
      h->host.want_grid = 0; h->host.want_sums = 0;
      CU(jcopy(h, &h->dev->want_grid, &h->host.want_grid, sizeof(int32_t), cudaMemcpyHostToDevice));
      CU(jcopy(h, &h->dev->want_sums, &h->host.want_sums, sizeof(uint32_t), cudaMemcpyHostToDevice));
      continue;
    }
    if (h->host.want_grid && h->host.want_wide && !h->host.cursor_saved && !grid && !h->wide_now) {
      h->wide_now = true;
      h->host.want_grid = 0; h->host.want_wide = 0;
      CU(jcopy(h, &h->dev->want_grid, &h->host.want_grid, sizeof(int32_t), cudaMemcpyHostToDevice));
      CU(jcopy(h, &h->dev->want_wide, &h->host.want_wide, sizeof(uint32_t), cudaMemcpyHostToDevice));
      continue;
    }
    if ((h->host.want_grid || h->host.cursor_saved) && !grid) {
      if (h->debug_stop_at_handover) {
        fprintf(stderr, "debug_stop_at_handover: next_gate %d cursor_saved %u nodes %u vals %u\n", h->host.next_gate, h->host.cursor_saved, h->host.node_count,
                h->host.val_count);
        return JKQ_E_OK;
      }
      if (h->cfg.force_mode == 1) return fail(h, JKQ_E_ITEM_FULL, "gate %d: state or frontier exceeds the single-CTA path and force_mode=1", h->host.next_gate);
      grid = true;
      h->host.want_grid = 0;
      CU(jcopy(h, &h->dev->want_grid, &h->host.want_grid, sizeof(int32_t), cudaMemcpyHostToDevice));
      continue;
    }
    return fail(h, JKQ_E_INTERNAL, "gate kernel stopped at gate %d of %d without a reason", h->host.next_gate, upto);
  }
}

// Philox-4x32-10 on the host: the same stream the device sampler uses (jkq_sample.cuh), for the outcomes of
// mid-circuit measurements.
static double philox_uniform(uint64_t seed, uint64_t counter) {
  uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), 0x4D454153u /* "MEAS" */, 0};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return ((double)c[0] + 0.5) * (1.0 / 4294967296.0);
}

static int build_sampling_now(jkq_handle *h);

// P(qubit reads 0) in the state as it is now (AssignProbsOne, Simulator.cpp:291-339)
static int qubit_probability(jkq_handle *h, uint32_t qubit, double *p0) {
  int rc = build_sampling_now(h);
  if (rc) return rc;
  k_marginal<<<1, 256, 0, h->stream>>>(h->dev, h->snodes, h->prob, (int)(h->n - 1 - qubit), h->scalar_dev);
  CU(cudaGetLastError());
  CU(jcopy(h, p0, h->scalar_dev, 8, cudaMemcpyDeviceToHost));
  CU(cudaStreamSynchronize(h->stream));
  h->sampling_ready = false;  // (k_marginal used the probability array as scratch)
  return JKQ_E_OK;
}

// MeasureOne / ResetQubit at op i (Simulator.cpp:226-289, 291-339): apply what precedes, get P(0), then fill the
// matrices of the op's gate slot(s).
//   MEASURE_ONE            draw the outcome; slot 0 = projector on it / sqrt(p)
//   RESET (reference)      no draw: slot 0 = X when P(0) = 0, else identity; slot 1 = |0><0| / sqrt(P(0))
//   RESET (physical_reset) draw the outcome; slot 0 = projector / sqrt(p); slot 1 = X when the outcome was 1
static void set_diag(JGate &g, double d0, double d1) { memset(g.m, 0, sizeof(g.m)); g.m[0] = d0; g.m[6] = d1; }
static void set_x(JGate &g) { memset(g.m, 0, sizeof(g.m)); g.m[2] = 1.0; g.m[4] = 1.0; }
static int measure_collapse(jkq_handle *h, size_t i) {
  const jkq_op &o = h->ops[i];
  const int g0 = h->gate_of_op[i];
  int rc = run_segment(h, g0);
  if (rc) return rc;
  double p0 = 0;
  rc = qubit_probability(h, o.target, &p0);
  if (rc) return rc;
  if (p0 < 0) p0 = 0;
  if (p0 > 1) p0 = 1;
  const double norm = h->host.norm > 0 ? h->host.norm : 1.0;  // (k_marginal works on normalised records)
  int n_slots = 1;
  if (o.kind == JKQ_OP_RESET && !h->physical_reset) {
    JGate &first = h->gates[g0], &proj = h->gates[g0 + 1];
    double p = p0;
    if (p0 <= 0.0) { set_x(first); p = 1.0; }
    else set_diag(first, 1.0, 1.0);
    set_diag(proj, 1.0 / sqrt(p * norm), 0.0);
    n_slots = 2;
  } else {
    const double u = philox_uniform(h->measure_seed, (uint64_t)i);
    const int outcome = u < p0 ? 0 : 1;
    const double p = outcome ? 1.0 - p0 : p0;
    if (!(p > 0)) return fail(h, JKQ_E_ZERO_STATE, "measurement of qubit %u: outcome %d has probability 0", o.target, outcome);
    const double f = 1.0 / sqrt(p * norm);
    set_diag(h->gates[g0], outcome ? 0.0 : f, outcome ? f : 0.0);
    if (o.kind == JKQ_OP_RESET) {
      if (outcome) set_x(h->gates[g0 + 1]);
      else set_diag(h->gates[g0 + 1], 1.0, 1.0);
      n_slots = 2;
    } else {
      h->mid_mask |= 1ull << o.arg;
      if (outcome) h->mid_bits |= 1ull << o.arg;
      else h->mid_bits &= ~(1ull << o.arg);
    }
  }
  h->placeholder_gates += (uint64_t)n_slots;
  CU(jcopy(h, h->gates_dev + g0, &h->gates[g0], n_slots * sizeof(JGate), cudaMemcpyHostToDevice));
  CU(cudaStreamSynchronize(h->stream));  // (the host copies of the gates are pageable)
  return run_segment(h, g0 + n_slots);
}

// the 2^k probabilities of a PROBABILITIES op, into its slot
static int record_probabilities(jkq_handle *h, const jkq_op &o) {
  int rc = build_sampling_now(h);
  if (rc) return rc;
  ProbParams pp;
  int k = 0, low = -1;
  for (int L = 0; L < JKQ_MAXQ; L++) pp.bit_of_level[L] = -1;
  for (uint32_t q = 0; q < h->n; q++) {
    if (!(o.ctrl_pos >> q & 1)) continue;
    const int L = (int)(h->n - 1 - q);
    pp.bit_of_level[L] = (int8_t)k++;
    if (low < 0 || L < low) low = L;
  }
  const uint64_t count = 1ull << k;
  double *buf = nullptr;
  CU(cudaMalloc(&buf, count * sizeof(double)));
  auto it = h->prob_slots.find(o.arg);
  if (it != h->prob_slots.end()) { cudaFree(it->second.first); h->prob_slots.erase(it); }
  h->prob_slots[o.arg] = std::make_pair(buf, (uint32_t)k);
  if ((uint32_t)k == h->n) {
    int blocks = (int)((count + 255) / 256);
    if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
    k_probs_full<<<blocks, 256, 0, h->stream>>>(h->dev, h->snodes, pp, count, buf);
    CU(cudaGetLastError());
  } else {
    const uint64_t per = (uint64_t)h->host.node_count * sizeof(double);
    uint64_t blocks = count < (uint64_t)h->sm_count * 4 ? count : (uint64_t)h->sm_count * 4;
    while (blocks > 1 && blocks * per > (256ull << 20)) blocks /= 2;
    double *scratch = nullptr;
    CU(cudaMalloc(&scratch, blocks * per));
    k_probs_subset<<<(int)blocks, 256, 0, h->stream>>>(h->dev, h->snodes, pp, low, count, scratch, buf);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(scratch);
  }
  h->sampling_ready = false;  // (gates may follow: the records are rebuilt when shots are asked for)
  return JKQ_E_OK;
}

// SNAPSHOT on a register too wide for a dense vector: copy the arena (nodes and reals up to their counters, root edge
// in the control block) -- the state is a rooted DD here, run_segment has just returned
static int snapshot_dd(jkq_handle *h, uint32_t slot) {
  int rc = pull_state(h);
  if (rc) return rc;
  auto it = h->dd_snapshots.find(slot);
  if (it != h->dd_snapshots.end()) {
    cudaFree(it->second.nodes); cudaFree(it->second.vals); cudaFree(it->second.eng);
    h->dd_snapshots.erase(it);
  }
  jkq_handle::DDSnap d;
  d.n_nodes = h->host.node_count; d.n_vals = h->host.val_count;
  CU(cudaMalloc((void **)&d.nodes, (size_t)d.n_nodes * sizeof(JNode)));
  CU(cudaMalloc((void **)&d.vals, (size_t)d.n_vals * sizeof(double)));
  CU(cudaMalloc((void **)&d.eng, sizeof(JEng)));
  CU(cudaMemcpyAsync(d.nodes, h->host.nodes, (size_t)d.n_nodes * sizeof(JNode), cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaMemcpyAsync(d.vals, h->host.vals, (size_t)d.n_vals * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  JEng *clone = new JEng(h->host);
  clone->nodes = d.nodes; clone->vals = d.vals;
  cudaError_t e = cudaMemcpyAsync(d.eng, clone, sizeof(JEng), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  delete clone;
  h->dd_snapshots[slot] = d;
  if (e != cudaSuccess) return fail(h, JKQ_E_CUDA, "snapshot: %s", cudaGetErrorString(e));
  return JKQ_E_OK;
}

extern "C" int jkq_apply(jkq_handle *h) {
  if (!h) return JKQ_E_ARGUMENT;
  if (!h->loaded) return fail(h, JKQ_E_STATE, "jkq_apply before jkq_load_circuit");
  CU(cudaSetDevice(h->device));
  free_snapshots(h);
  h->applied = false; h->sampling_ready = false; h->imported = false; h->hist_valid = false;
  JEng &E = h->host;
  E.n = h->n;
  if (h->lvl_log) { cudaFree(h->lvl_log); h->lvl_log = nullptr; }
  if (h->level_log && !h->gates.empty()) {
    CU(cudaMalloc(&h->lvl_log, h->gates.size() * h->n * 4));
    CU(cudaMemsetAsync(h->lvl_log, 0, h->gates.size() * h->n * 4, h->stream));
  }
  E.lvl_log = h->lvl_log;
  E.mat_u = h->mat_u; E.mat_epoch = h->mat_epoch;
  E.status = 0; E.next_gate = 0; E.want_grid = 0; E.cursor_saved = 0; E.want_wide = 0; E.want_sums = 0; E.sums_ask = h->sums_ask;
  E.cta_wide_ok = (h->cfg.cta_threads > 32 && h->cfg.reserved != 1) ? 1u : 0u;
  h->wide_now = false;
  h->sums_now = false;
  E.bar_count = 0; E.bar_gen = 0; E.bar_flags[0] = E.bar_flags[1] = E.bar_flags[2] = 0;
  E.pend_n[0] = E.pend_n[1] = E.pend_n[2] = 0;
  CU(cudaEventRecord(h->ev0, h->stream));
  CU(jcopy(h, h->dev, &E, sizeof(JEng), cudaMemcpyHostToDevice));
  k_init<MODE_GRID><<<1, 32, 0, h->stream>>>(h->dev);
  CU(cudaGetLastError());
  const uint64_t cta0 = h->stats.cta_launches, grid0 = h->stats.grid_launches;
  h->mid_bits = 0; h->mid_mask = 0; h->placeholder_gates = 0;
  size_t i = 0;
  auto boundary = [](uint32_t k) {
    return k == JKQ_OP_SNAPSHOT || k == JKQ_OP_MEASURE_ONE || k == JKQ_OP_RESET || k == JKQ_OP_IF || k == JKQ_OP_PROBABILITIES;
  };
  while (i < h->ops.size()) {
    const uint32_t kind = h->ops[i].kind;
    if (kind == JKQ_OP_GATE) {
      size_t j = i;
      while (j < h->ops.size() && !boundary(h->ops[j].kind)) j++;
      int upto = j < h->ops.size() ? h->gate_of_op[j] : (int)h->gates.size();
      int rc = run_segment(h, upto);
      if (rc) return rc;
      i = j;
    } else if (kind == JKQ_OP_SNAPSHOT) {
      int rc0 = run_segment(h, h->gate_of_op[i]);
      if (rc0) return rc0;
      if (h->n > h->snapshot_dense_max) {  // keep the diagram, not the vector
        int rcs = snapshot_dd(h, h->ops[i].arg);
        if (rcs) return rcs;
        i++;
        continue;
      }
      double2 *buf = nullptr;
      uint64_t len = 1ull << h->n;
      CU(cudaMalloc(&buf, len * sizeof(double2)));
      auto it = h->snapshots.find(h->ops[i].arg);
      if (it != h->snapshots.end()) { cudaFree(it->second); h->snapshots.erase(it); }
      h->snapshots[h->ops[i].arg] = buf;
      int blocks = (int)((len + 255) / 256);
      if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
      k_amplitudes<<<blocks, 256, 0, h->stream>>>(h->dev, 0, len, buf);
      CU(cudaGetLastError());
      i++;
    } else if (kind == JKQ_OP_MEASURE_ONE || kind == JKQ_OP_RESET) {
      int rc = measure_collapse(h, i);
      if (rc) return rc;
      i++;
    } else if (kind == JKQ_OP_PROBABILITIES) {
      int rc = run_segment(h, h->gate_of_op[i]);
      if (!rc) rc = record_probabilities(h, h->ops[i]);
      if (rc) return rc;
      i++;
    } else if (kind == JKQ_OP_IF) {
      // every MEASURE_ONE before this op has been resolved: the condition is known on the host
      if ((h->mid_bits & h->ops[i].ctrl_pos) == h->ops[i].ctrl_neg) { i++; continue; }
      const size_t cnt = h->ops[i].arg ? h->ops[i].arg : 1;
      int rc = run_segment(h, h->gate_of_op[i + 1]);
      if (rc) return rc;
      // skip the guarded ops: their gate slots are stepped over
      h->host.next_gate = i + 1 + cnt < h->ops.size() ? h->gate_of_op[i + 1 + cnt] : (int)h->gates.size();
      CU(jcopy(h, &h->dev->next_gate, &h->host.next_gate, sizeof(h->host.next_gate), cudaMemcpyHostToDevice));
      CU(cudaStreamSynchronize(h->stream));
      i += 1 + cnt;
    } else {
      i++;
    }
  }
  CU(cudaEventRecord(h->ev1, h->stream));
  int rc = pull_state(h);
  if (rc) return rc;
  if (h->host.status) return fail(h, h->host.status, "%s", status_text(h->host.status));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->stats.apply_ms = ms;
  (void)cta0; (void)grid0;
  h->applied = true;
  return JKQ_E_OK;
}

static int build_sampling_now(jkq_handle *h);
static int ensure_sampling(jkq_handle *h) {
  if (h->sampling_ready) return JKQ_E_OK;
  if (!h->applied) return fail(h, JKQ_E_STATE, "no state: call jkq_apply or jkq_dd_import first");
  return build_sampling_now(h);
}

// level-ordered compaction + bottom-up probabilities of the CURRENT state (also between the segments of a circuit)
static int build_sampling_now(jkq_handle *h) {
  // level-ordered compaction + bottom-up probabilities; the CTA path first unless the DD is clearly too big for it
  bool grid = h->cfg.force_mode == 2 || (h->cfg.force_mode != 1 && h->host.node_count > 8 * JKQ_SNODES);
  JEng *E = h->dev;
  JSNode *sn = h->snodes;
  double *pr = h->prob;
  int rc = JKQ_E_OK;
  for (;;) {
    if (grid) {
      int rb = reset_barrier(h);
      if (rb) return rb;
      void *args[] = {&E, &sn, &pr};
      CU(cudaLaunchCooperativeKernel((void *)k_build_sampling<MODE_GRID>, dim3(h->grid_blocks), dim3(256), args, 0, h->stream));
    } else {
      k_build_sampling<MODE_CTA><<<1, h->cfg.cta_threads, sizeof(CtaSmem), h->stream>>>(E, sn, pr);
      CU(cudaGetLastError());
    }
    rc = pull_state(h);
    if (rc) return rc;
    if (!grid && !h->host.status && h->host.want_grid) {
      if (h->cfg.force_mode == 1) return fail(h, JKQ_E_NODE_FULL, "state does not fit the single-CTA path and force_mode=1");
      grid = true;
      h->host.want_grid = 0;
      CU(jcopy(h, &h->dev->want_grid, &h->host.want_grid, sizeof(int32_t), cudaMemcpyHostToDevice));
      continue;
    }
    break;
  }
  if (h->host.status) return fail(h, h->host.status, "build_sampling: %s", status_text(h->host.status));
  h->dd_root = h->host.root;
  h->dd_nodes = h->host.node_count;
  h->stats.norm = h->host.norm;
  h->sampling_ready = true;
  h->sn_epoch++;
  return JKQ_E_OK;
}

static int sample_impl(jkq_handle *h, uint64_t shots, uint64_t seed, uint64_t shot_offset, int accumulate, uint64_t expect_shots);
extern "C" int jkq_sample(jkq_handle *h, uint64_t shots, uint64_t seed, uint64_t shot_offset, int accumulate) {
  return sample_impl(h, shots, seed, shot_offset, accumulate, shots);
}
// expect_shots: what the histogram is sized for (>= shots: the shots other GPUs will merge into this one)
static int sample_impl(jkq_handle *h, uint64_t shots, uint64_t seed, uint64_t shot_offset, int accumulate, uint64_t expect_shots) {
  if (!h) return JKQ_E_ARGUMENT;
  CU(cudaSetDevice(h->device));
  CU(cudaEventRecord(h->ev2, h->stream));
  if (!h->imported) {
    int rc = ensure_sampling(h);
    if (rc) return rc;
  }
  if (fabs(h->stats.norm - 1.0) > 0.01 && !h->imported)  // the reference only warns, Simulator.cpp:157-163
    fprintf(stderr, "WARNING in measurement: numerical instability occurred during simulation: |alpha|^2 + |beta|^2 = %g, but should be 1!\n",
            h->stats.norm);
  int measured = 0;
  for (uint32_t L = 0; L < h->n; L++) measured += h->sp.bitpos[L] < 64;
  if (expect_shots < shots) expect_shots = shots;
  uint64_t need = expect_shots < (1ull << 40) ? expect_shots : (1ull << 40);
  if (measured < 40 && (1ull << measured) < need) need = 1ull << measured;
  uint64_t slots = pow2_at_least(need * 2, 1u << 12);
  bool narrow = true;  // every measured classical bit below 32?
  for (uint32_t L = 0; L < h->n; L++) narrow = narrow && (h->sp.bitpos[L] >= 64 || h->sp.bitpos[L] < 32);
  if (!accumulate || !h->hist_valid || slots > h->hslots) {
    if (accumulate && h->hist_valid) return fail(h, JKQ_E_STATE, "accumulating histogram would need to grow; sample the largest batch first");
    const bool packed = narrow && expect_shots < 0xFFFFFFFEull && !h->force_fused;
    if (slots > h->hslots || packed != h->hist_packed) {
      free_hist(h);
      CU(cudaMalloc(&h->hs, (slots + 1) * (packed ? 8 : sizeof(JHSlot))));
      h->hslots = slots;
      h->hist_packed = packed;
    }
    if (h->hist_packed) k_hist_fill<true><<<h->sm_count * 4, 256, 0, h->stream>>>(h->hs, h->hslots, h->n_distinct);
    else k_hist_fill<false><<<h->sm_count * 4, 256, 0, h->stream>>>(h->hs, h->hslots, h->n_distinct);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(h->overflow, 0, 4, h->stream));
    h->hist_valid = true;
    h->hist_shots = 0;
  }
  if (h->hist_packed && (!narrow || h->hist_shots + shots >= 0xFFFFFFFEull))
    return fail(h, JKQ_E_STATE, "accumulating into a packed histogram beyond 2^32 shots or 32 classical bits; sample the largest batch first");
  h->hist_shots += shots;
  // Random numbers on demand pay when most nodes leave no choice (GHZ: one node of 47; 19.7 -> 29 G shots/s) and cost
  // ~15 % when every node has one (QFT-32): counted once per sampling diagram.
  if (shots && h->sn_choice_epoch != h->sn_epoch) {
    CU(cudaMemsetAsync(h->choice_dev, 0, 4, h->stream));
    k_sn_choice<<<h->sm_count, 256, 0, h->stream>>>(h->snodes, h->dd_nodes, h->choice_dev);
    CU(cudaGetLastError());
    unsigned int n_choice = 0;
    CU(jcopy(h, &n_choice, h->choice_dev, 4, cudaMemcpyDeviceToHost));
    CU(cudaStreamSynchronize(h->stream));
    h->sn_lazy = (uint64_t)n_choice * 2 < (uint64_t)h->dd_nodes;
    h->sn_choice_epoch = h->sn_epoch;
  }
  const bool lazy = h->sn_lazy;
  if (shots) {
    uint64_t want = (shots + 255) / 256;
    int blocks = (int)(want < (uint64_t)h->sm_count * 8 ? want : (uint64_t)h->sm_count * 8);
    CU(cudaEventRecord(h->ev3, h->stream));
#define JKQ_LAUNCH_SAMPLE(PK, LZ) k_sample<PK, LZ><<<blocks, 256, 0, h->stream>>>(h->snodes, h->dd_root, (int)h->n, h->sp, seed, shot_offset, shots, \
                                                                                   h->hs, (uint32_t)(h->hslots - 1), h->overflow, h->n_distinct)
    if (h->hist_packed) { if (lazy) JKQ_LAUNCH_SAMPLE(true, true); else JKQ_LAUNCH_SAMPLE(true, false); }
    else { if (lazy) JKQ_LAUNCH_SAMPLE(false, true); else JKQ_LAUNCH_SAMPLE(false, false); }
#undef JKQ_LAUNCH_SAMPLE
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev1, h->stream));
    h->stats.sample_launches++;
  }
  int ovf = 0;
  CU(jcopy(h, &ovf, h->overflow, 4, cudaMemcpyDeviceToHost));
  CU(cudaStreamSynchronize(h->stream));
  if (ovf) return fail(h, JKQ_E_CAPACITY, "count histogram overflow");
  if (shots) {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev3, h->ev1));
    h->stats.sample_kernel_ms = ms;
    CU(cudaEventElapsedTime(&ms, h->ev2, h->ev1));
    h->stats.sample_ms = ms;
  }
  return JKQ_E_OK;
}

extern "C" int jkq_release_histogram(jkq_handle *h) {
  if (!h) return JKQ_E_ARGUMENT;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  free_hist(h);
  return JKQ_E_OK;
}

extern "C" int jkq_run(jkq_handle *h, uint64_t shots, uint64_t seed) {
  int rc = jkq_apply(h);
  if (rc) return rc;
  return jkq_sample(h, shots, seed, 0, 0);
}

extern "C" int jkq_get_counts(jkq_handle *h, uint64_t *keys, uint64_t *counts, size_t cap, size_t *n) {
  if (!h || !n) return JKQ_E_ARGUMENT;
  if (!h->hist_valid) return fail(h, JKQ_E_STATE, "no histogram: call jkq_sample first");
  CU(cudaSetDevice(h->device));
  if (cap == 0 && !keys) {
    // only the number of distinct outcomes: the kernels counted the slots they claimed
    unsigned long long fresh = 0, all_ones = 0;
    CU(jcopy(h, &fresh, h->n_distinct, 8, cudaMemcpyDeviceToHost));
    if (!h->hist_packed) CU(jcopy(h, &all_ones, &((JHSlot *)h->hs)[h->hslots].cnt, 8, cudaMemcpyDeviceToHost));
    CU(cudaStreamSynchronize(h->stream));
    *n = (size_t)(fresh + (all_ones ? 1 : 0));
    h->stats.distinct_outcomes = *n;
    return JKQ_E_OK;
  }
  uint64_t want = cap ? cap : 1;
  if (want > h->out_cap) {
    cudaFree(h->ok); cudaFree(h->oc);
    h->ok = h->oc = nullptr;
    CU(cudaMalloc(&h->ok, want * 8));
    CU(cudaMalloc(&h->oc, want * 8));
    h->out_cap = want;
  }
  CU(cudaMemsetAsync(h->n_out, 0, 8, h->stream));
  if (h->hist_packed) k_hist_compact<true><<<h->sm_count * 4, 256, 0, h->stream>>>(h->hs, h->hslots, h->ok, h->oc, h->n_out, cap);
  else k_hist_compact<false><<<h->sm_count * 4, 256, 0, h->stream>>>(h->hs, h->hslots, h->ok, h->oc, h->n_out, cap);
  CU(cudaGetLastError());
  unsigned long long cnt = 0;
  CU(jcopy(h, &cnt, h->n_out, 8, cudaMemcpyDeviceToHost));
  CU(cudaStreamSynchronize(h->stream));
  *n = (size_t)cnt;
  h->stats.distinct_outcomes = cnt;
  if (cnt > cap) return keys ? fail(h, JKQ_E_CAPACITY, "%llu distinct outcomes, buffer holds %zu", cnt, cap) : JKQ_E_OK;
  if (cnt && keys && counts) {
    CU(jcopy(h, keys, h->ok, cnt * 8, cudaMemcpyDeviceToHost));
    CU(jcopy(h, counts, h->oc, cnt * 8, cudaMemcpyDeviceToHost));
    CU(cudaStreamSynchronize(h->stream));
  }
  return JKQ_E_OK;
}

// sort this handle's histogram by owner rank into caller-provided DEVICE buffers (sparse exchange)
extern "C" int jkq_counts_partition(jkq_handle *h, uint32_t world, uint64_t *dev_keys, uint64_t *dev_counts, size_t cap, uint64_t *per_owner) {
  if (!h || !per_owner || world < 1 || world > 64) return JKQ_E_ARGUMENT;
  if (!h->hist_valid) return fail(h, JKQ_E_STATE, "no histogram: call jkq_sample first");
  CU(cudaSetDevice(h->device));
  unsigned long long *cnt = nullptr;
  CU(cudaMalloc(&cnt, 64 * 8));
  CU(cudaMemsetAsync(cnt, 0, 64 * 8, h->stream));
  if (h->hist_packed) k_hist_owner_count<true><<<h->sm_count * 4, 256, 0, h->stream>>>(h->hs, h->hslots, world, cnt);
  else k_hist_owner_count<false><<<h->sm_count * 4, 256, 0, h->stream>>>(h->hs, h->hslots, world, cnt);
  unsigned long long host_cnt[64];
  cudaError_t e = jcopy(h, host_cnt, cnt, 64 * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { cudaFree(cnt); return fail(h, JKQ_E_CUDA, "jkq_counts_partition: %s", cudaGetErrorString(e)); }
  uint64_t total = 0;
  unsigned long long offs[64];
  for (uint32_t r = 0; r < world; r++) { per_owner[r] = host_cnt[r]; offs[r] = total; total += host_cnt[r]; }
  h->stats.distinct_outcomes = total;
  if (!dev_keys || !dev_counts) { cudaFree(cnt); return JKQ_E_OK; }  // size query
  if (cap < total) { cudaFree(cnt); return fail(h, JKQ_E_CAPACITY, "%llu distinct outcomes, buffers hold %zu", (unsigned long long)total, cap); }
  e = jcopy(h, cnt, offs, world * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    if (h->hist_packed) k_hist_owner_scatter<true><<<h->sm_count * 4, 256, 0, h->stream>>>(h->hs, h->hslots, world, cnt, (unsigned long long *)dev_keys,
                                                                  (unsigned long long *)dev_counts);
    else k_hist_owner_scatter<false><<<h->sm_count * 4, 256, 0, h->stream>>>(h->hs, h->hslots, world, cnt, (unsigned long long *)dev_keys,
                                                                  (unsigned long long *)dev_counts);
    e = cudaStreamSynchronize(h->stream);
  }
  cudaFree(cnt);
  if (e != cudaSuccess) return fail(h, JKQ_E_CUDA, "jkq_counts_partition: %s", cudaGetErrorString(e));
  return JKQ_E_OK;
}

// add (key, count) pairs from DEVICE buffers into the histogram; reset = 1 starts an empty one sized for `expect` keys
extern "C" int jkq_counts_merge(jkq_handle *h, const uint64_t *dev_keys, const uint64_t *dev_counts, size_t n, size_t expect, int reset) {
  if (!h || (n && (!dev_keys || !dev_counts))) return JKQ_E_ARGUMENT;
  CU(cudaSetDevice(h->device));
  if (reset || !h->hist_valid) {
    const uint64_t slots = pow2_at_least((uint64_t)(expect > n ? expect : n) * 2, 1u << 12);
    if (slots > h->hslots || h->hist_packed) {  // (a merged histogram takes whatever keys arrive: fused format)
      free_hist(h);
      CU(cudaMalloc(&h->hs, (slots + 1) * sizeof(JHSlot)));
      h->hslots = slots;
      h->hist_packed = false;
    }
    k_hist_fill<false><<<h->sm_count * 4, 256, 0, h->stream>>>(h->hs, h->hslots, h->n_distinct);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(h->overflow, 0, 4, h->stream));
    h->hist_valid = true;
  }
  if (n) {
    uint64_t want = (n + 255) / 256;
    int blocks = (int)(want < (uint64_t)h->sm_count * 8 ? want : (uint64_t)h->sm_count * 8);
    if (h->hist_packed) k_hist_merge<true><<<blocks, 256, 0, h->stream>>>(h->hs, (uint32_t)(h->hslots - 1), (const unsigned long long *)dev_keys,
                                                (const unsigned long long *)dev_counts, n, h->overflow, h->n_distinct);
    else k_hist_merge<false><<<blocks, 256, 0, h->stream>>>(h->hs, (uint32_t)(h->hslots - 1), (const unsigned long long *)dev_keys,
                                                (const unsigned long long *)dev_counts, n, h->overflow, h->n_distinct);
    CU(cudaGetLastError());
  }
  int ovf = 0;
  CU(jcopy(h, &ovf, h->overflow, 4, cudaMemcpyDeviceToHost));
  CU(cudaStreamSynchronize(h->stream));
  if (ovf) return fail(h, JKQ_E_CAPACITY, "count histogram overflow while merging (raise `expect`)");
  return JKQ_E_OK;
}

extern "C" int jkq_get_amplitudes(jkq_handle *h, uint64_t first, uint64_t count, double *re_im) {
  if (!h || (count && !re_im)) return JKQ_E_ARGUMENT;
  if (!h->applied) return fail(h, JKQ_E_STATE, "no state: call jkq_apply first");
  if (h->n < 64 && first + count > (1ull << h->n)) return fail(h, JKQ_E_ARGUMENT, "amplitude range exceeds 2^n");
  if (!count) return JKQ_E_OK;
  CU(cudaSetDevice(h->device));
  double2 *buf = nullptr;
  CU(cudaMalloc(&buf, count * sizeof(double2)));
  int blocks = (int)((count + 255) / 256);
  if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
  k_amplitudes<<<blocks, 256, 0, h->stream>>>(h->dev, first, count, buf);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = jcopy(h, re_im, buf, count * sizeof(double2), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(buf);
  if (e != cudaSuccess) return fail(h, JKQ_E_CUDA, "amplitudes: %s", cudaGetErrorString(e));
  return JKQ_E_OK;
}

// amplitudes [first, first + count) of the state a SNAPSHOT op recorded in `slot` (0xFFFFFFFF: the current state):
// a slice of the dense vector, or -- for a snapshot kept as a diagram -- one walk per amplitude, so a register of any
// width is read out range by range (the streaming read-out of QASMsimulator.cpp:1012-1030's state vectors)
extern "C" int jkq_get_snapshot_amplitudes(jkq_handle *h, uint32_t slot, uint64_t first, uint64_t count, double *re_im) {
  if (!h || (count && !re_im)) return JKQ_E_ARGUMENT;
  if (!h->applied) return fail(h, JKQ_E_STATE, "no state: call jkq_apply first");
  if (slot == 0xFFFFFFFFu) return jkq_get_amplitudes(h, first, count, re_im);
  if (h->n < 64 && (first > (1ull << h->n) || count > (1ull << h->n) - first)) return fail(h, JKQ_E_ARGUMENT, "amplitude range exceeds 2^n");
  if (!count) return JKQ_E_OK;
  CU(cudaSetDevice(h->device));
  auto it = h->snapshots.find(slot);
  if (it != h->snapshots.end()) {
    CU(jcopy(h, re_im, it->second + first, count * sizeof(double2), cudaMemcpyDeviceToHost));
    CU(cudaStreamSynchronize(h->stream));
    return JKQ_E_OK;
  }
  auto dt = h->dd_snapshots.find(slot);
  if (dt == h->dd_snapshots.end()) return fail(h, JKQ_E_ARGUMENT, "no snapshot in slot %u", slot);
  double2 *buf = nullptr;
  CU(cudaMalloc(&buf, count * sizeof(double2)));
  int blocks = (int)((count + 255) / 256);
  if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
  k_amplitudes<<<blocks, 256, 0, h->stream>>>(dt->second.eng, first, count, buf);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = jcopy(h, re_im, buf, count * sizeof(double2), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(buf);
  if (e != cudaSuccess) return fail(h, JKQ_E_CUDA, "snapshot amplitudes: %s", cudaGetErrorString(e));
  return JKQ_E_OK;
}

// ... with the bits of the index in the caller's order: bit b of (first + i) is qubit qubit_of_bit[b] (n entries, a
// permutation of 0..n-1).  The reference prints a snapshot's state vector with the FIRST listed qubit as the most
// significant bit (QASMsimulator.cpp:1016-1028); the executable streams that order through this call.
extern "C" int jkq_get_snapshot_amplitudes_ordered(jkq_handle *h, uint32_t slot, const uint32_t *qubit_of_bit, uint64_t first, uint64_t count,
                                                   double *re_im) {
  if (!h || !qubit_of_bit || (count && !re_im)) return JKQ_E_ARGUMENT;
  if (!h->applied) return fail(h, JKQ_E_STATE, "no state: call jkq_apply first");
  if (h->n < 64 && (first > (1ull << h->n) || count > (1ull << h->n) - first)) return fail(h, JKQ_E_ARGUMENT, "amplitude range exceeds 2^n");
  BitPerm bp;
  uint64_t seen = 0;
  for (uint32_t b = 0; b < h->n; b++) {
    if (qubit_of_bit[b] >= h->n || (seen >> qubit_of_bit[b] & 1)) return fail(h, JKQ_E_ARGUMENT, "qubit_of_bit is not a permutation of the %u qubits", h->n);
    seen |= 1ull << qubit_of_bit[b];
    bp.q[b] = (uint8_t)qubit_of_bit[b];
  }
  if (!count) return JKQ_E_OK;
  JEng *eng = h->dev;
  const double2 *dense = nullptr;
  if (slot != 0xFFFFFFFFu) {
    auto it = h->snapshots.find(slot);
    auto dt = h->dd_snapshots.find(slot);
    if (it != h->snapshots.end()) dense = it->second;
    else if (dt != h->dd_snapshots.end()) eng = dt->second.eng;
    else return fail(h, JKQ_E_ARGUMENT, "no snapshot in slot %u", slot);
  }
  CU(cudaSetDevice(h->device));
  double2 *buf = nullptr;
  CU(cudaMalloc(&buf, count * sizeof(double2)));
  int blocks = (int)((count + 255) / 256);
  if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
  k_amplitudes_ordered<<<blocks, 256, 0, h->stream>>>(eng, dense, bp, (int)h->n, first, count, buf);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = jcopy(h, re_im, buf, count * sizeof(double2), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(buf);
  if (e != cudaSuccess) return fail(h, JKQ_E_CUDA, "snapshot amplitudes: %s", cudaGetErrorString(e));
  return JKQ_E_OK;
}

extern "C" int jkq_get_statevector(jkq_handle *h, uint32_t slot, double *re_im, size_t cap_amplitudes) {
  if (!h || !re_im) return JKQ_E_ARGUMENT;
  if (!h->applied) return fail(h, JKQ_E_STATE, "no state: call jkq_apply first");
  if (h->n >= 48) return fail(h, JKQ_E_CAPACITY, "2^%u amplitudes: read ranges with jkq_get_snapshot_amplitudes", h->n);
  uint64_t len = 1ull << h->n;
  if (cap_amplitudes < len) return fail(h, JKQ_E_CAPACITY, "need room for %llu amplitudes", (unsigned long long)len);
  if (slot == 0xFFFFFFFFu) return jkq_get_amplitudes(h, 0, len, re_im);
  if (h->dd_snapshots.count(slot)) return jkq_get_snapshot_amplitudes(h, slot, 0, len, re_im);
  auto it = h->snapshots.find(slot);
  if (it == h->snapshots.end()) return fail(h, JKQ_E_ARGUMENT, "no snapshot in slot %u", slot);
  CU(cudaSetDevice(h->device));
  CU(jcopy(h, re_im, it->second, len * sizeof(double2), cudaMemcpyDeviceToHost));
  CU(cudaStreamSynchronize(h->stream));
  return JKQ_E_OK;
}

// entries [first, first + count) of the probabilities in `slot`
extern "C" int jkq_get_probabilities_range(jkq_handle *h, uint32_t slot, uint64_t first, uint64_t count, double *out) {
  if (!h || (count && !out)) return JKQ_E_ARGUMENT;
  auto it = h->prob_slots.find(slot);
  if (it == h->prob_slots.end()) return fail(h, JKQ_E_ARGUMENT, "no probabilities in slot %u", slot);
  const uint64_t len = 1ull << it->second.second;
  if (first > len || count > len - first) return fail(h, JKQ_E_ARGUMENT, "range exceeds the %llu probabilities of slot %u", (unsigned long long)len, slot);
  if (!count) return JKQ_E_OK;
  CU(cudaSetDevice(h->device));
  CU(jcopy(h, out, it->second.first + first, count * sizeof(double), cudaMemcpyDeviceToHost));
  CU(cudaStreamSynchronize(h->stream));
  return JKQ_E_OK;
}

extern "C" int jkq_get_probabilities(jkq_handle *h, uint32_t slot, double *out, size_t cap, size_t *n) {
  if (!h) return JKQ_E_ARGUMENT;
  auto it = h->prob_slots.find(slot);
  if (it == h->prob_slots.end()) return fail(h, JKQ_E_ARGUMENT, "no probabilities in slot %u", slot);
  const size_t len = (size_t)1 << it->second.second;
  if (n) *n = len;
  if (!out) return JKQ_E_OK;
  if (cap < len) return fail(h, JKQ_E_CAPACITY, "need room for %zu probabilities", len);
  CU(cudaSetDevice(h->device));
  CU(jcopy(h, out, it->second.first, len * sizeof(double), cudaMemcpyDeviceToHost));
  CU(cudaStreamSynchronize(h->stream));
  return JKQ_E_OK;
}


extern "C" int jkq_get_level_counts(jkq_handle *h, uint32_t *per_level, size_t n) {
  if (!h || !per_level) return JKQ_E_ARGUMENT;
  if (n < h->n) return fail(h, JKQ_E_CAPACITY, "need %u entries", h->n);
  CU(cudaSetDevice(h->device));
  int rc = ensure_sampling(h);  // runs the collector: lvl_cnt = live nodes per level
  if (rc) return rc;
  for (uint32_t L = 0; L < h->n; L++) per_level[L] = h->host.lvl_cnt[L];
  return JKQ_E_OK;
}

extern "C" int jkq_get_level_log(jkq_handle *h, uint32_t *out, size_t cap, size_t *n_written) {
  if (!h || !n_written) return JKQ_E_ARGUMENT;
  if (!h->applied || !h->lvl_log) return fail(h, JKQ_E_STATE, "level log not recorded (set option level_log=1 before jkq_apply)");
  size_t total = h->gates.size() * h->n;
  if (!out) { *n_written = total; return JKQ_E_OK; }  // size query
  if (cap < total) return fail(h, JKQ_E_CAPACITY, "need %zu entries", total);
  CU(cudaSetDevice(h->device));
  CU(jcopy(h, out, h->lvl_log, total * 4, cudaMemcpyDeviceToHost));
  CU(cudaStreamSynchronize(h->stream));
  *n_written = total;
  return JKQ_E_OK;
}

extern "C" int jkq_get_stats(jkq_handle *h, jkq_stats *out) {
  if (!h || !out) return JKQ_E_ARGUMENT;
  const JEng &E = h->host;
#ifdef JKQ_PROFILE
  fprintf(stderr, "JKQ_PROFILE cycles/gate over %llu gates: total %llu | thin-descent+setup %llu  expand %llu  reduce %llu  barrier %llu  commit %llu  "
          "thin-ascent+root %llu | generic phases/gate %.2f  collector total %llu\n",
          E.prof[6], E.prof[6] ? E.prof[8] / E.prof[6] : 0, E.prof[6] ? E.prof[0] / E.prof[6] : 0, E.prof[6] ? E.prof[2] / E.prof[6] : 0,
          E.prof[6] ? E.prof[3] / E.prof[6] : 0, E.prof[6] ? E.prof[4] / E.prof[6] : 0, E.prof[6] ? E.prof[5] / E.prof[6] : 0,
          E.prof[6] ? E.prof[1] / E.prof[6] : 0, E.prof[6] ? (double)E.prof[7] / E.prof[6] : 0.0, E.prof[9]);
  if (E.prof[6])
    fprintf(stderr, "  per gate: cursor_prepare %llu  solo expand %llu  solo reduce %llu  cursor_finish %llu | solo levels/gate %.2f  wide gates %llu\n",
            E.prof[11] / E.prof[6], E.prof[12] / E.prof[6], E.prof[13] / E.prof[6], E.prof[14] / E.prof[6], (double)E.prof[15] / E.prof[6], E.prof[10]);
  if (E.prof_x[1] + E.prof_x[3])
    fprintf(stderr, "  grid item loops (cycles per thread and phase): first run without a miss %.0f (%llu), first run with a value miss %.0f (%llu), rerun %.0f (%llu)\n",
            E.prof_x[1] ? (double)E.prof_x[0] / E.prof_x[1] : 0.0, E.prof_x[1], E.prof_x[3] ? (double)E.prof_x[2] / E.prof_x[3] : 0.0, E.prof_x[3],
            E.prof_x[5] ? (double)E.prof_x[4] / E.prof_x[5] : 0.0, E.prof_x[5]);
  for (int k = 0; k < 8; k++)
    if (E.prof_w[8 + k])
      fprintf(stderr, "  grid level phases of up to %d items: %llu levels, %llu items, %llu barriers, %.1f Mcycles (%.0f cycles per level)\n", 4 << (2 * k),
              E.prof_w[8 + k], E.prof_w[24 + k], E.prof_w[16 + k], E.prof_w[k] * 1e-6, (double)E.prof_w[k] / E.prof_w[8 + k]);
#endif
#ifdef JKQ_PROFILE
  {
    unsigned long long p[64], z[64] = {0};
    cudaMemcpyFromSymbol(p, g_gprof, sizeof(p));
    cudaMemcpyToSymbol(g_gprof, z, sizeof(z));
    static const char *nm[32] = {"item quad load", "expand_item_g", "reduce_item_g", "barrier wait (item thread)", "barrier wait (tid 0)", "item loop (item thread)", "phase loop+barrier (tid 0)", "",
                                 "expand: header loads", "expand: batch 1 (products)", "expand: magnitudes", "expand: batch 2 (div/add)", "expand: de-dup x2", "expand: M above target", "", "",
                                 "reduce: header loads", "reduce: arithmetic", "reduce: unique table", "", "batch: operands+compute", "batch: probes", "batch: value insert", "", "", "", "", "", "", "", "", ""};
    for (int k = 0; k < 32; k++)
      if (p[32 + k]) fprintf(stderr, "  narrow grid phases: %-30s %10llu x %8.0f cycles\n", nm[k], p[32 + k], (double)p[k] / p[32 + k]);
  }
#endif
#ifdef JKQ_PROFILE2
  {
    unsigned long long p[32], z[32] = {0};
    cudaMemcpyFromSymbol(p, g_fine, sizeof(p));
    cudaMemcpyToSymbol(g_fine, z, sizeof(z));
    if (p[17] || p[19])
      fprintf(stderr, "  cursor_prepare: %llu up-steps, %llu cycles each | %llu levels down, %llu cycles each\n", p[17], p[17] ? p[16] / p[17] : 0, p[19],
              p[19] ? p[18] / p[19] : 0);
    if (p[1])
      fprintf(stderr, "JKQ_PROFILE2 lane 0: expand_item %llu calls, %llu cycles each | chain path %llu times: %llu cycles each, %.1f levels each | expand_child x2 %llu cycles/plain call, de-dup %llu | reduce_item %llu calls, %llu cycles each; chain rebuild %llu cycles per chain item\n",
              p[1], p[0] / p[1], p[10], p[10] ? p[2] / p[10] : 0, p[10] ? (double)p[3] / p[10] : 0.0, (p[1] - p[10]) ? p[4] / (p[1] - p[10]) : 0,
              (p[1] - p[10]) ? p[5] / (p[1] - p[10]) : 0, p[7], p[7] ? p[6] / p[7] : 0, p[10] ? p[8] / p[10] : 0);
    if (p[11])
      fprintf(stderr, "  plain expand_child pairs: %llu (avg %llu cycles) | M at target %llu (avg %llu) | F same-node %llu (avg %llu) | other %llu (avg %llu); de-dup per pair %llu\n",
              p[11], p[4] / p[11], p[13], p[13] ? p[12] / p[13] : 0, p[15], p[15] ? p[14] / p[15] : 0, p[11] - p[13] - p[15],
              (p[11] - p[13] - p[15]) ? (p[4] - p[12] - p[14]) / (p[11] - p[13] - p[15]) : 0, p[5] / p[11]);
  }
#endif
  // (ApplyGate calls: MeasureOne / ResetQubit multiply their projectors in directly)
  h->stats.gates = E.st[JST_GATES] - h->placeholder_gates; h->stats.items = E.st[JST_ITEMS]; h->stats.nodes_created = E.st[JST_NODES];
  h->stats.value_inserts = E.st[JST_VALS]; h->stats.gc_runs = E.st[JST_GC_RUNS]; h->stats.gc_nodes_copied = E.st[JST_GC_COPIED];
  h->stats.retry_rounds = E.st[JST_RETRY]; h->stats.max_frontier = E.st[JST_MAXF]; h->stats.thin_gates = E.st[JST_THIN];
  h->stats.h2d_bytes = h->h2d_bytes; h->stats.d2h_bytes = h->d2h_bytes;
  h->stats.live_nodes = E.node_count ? E.node_count - 1 : 0; h->stats.live_values = E.val_count;
  *out = h->stats;
  return JKQ_E_OK;
}

extern "C" int jkq_get_qubit_probability(jkq_handle *h, uint32_t qubit, double *p0) {
  if (!h || !p0) return JKQ_E_ARGUMENT;
  if (!h->applied) return fail(h, JKQ_E_STATE, "no state: call jkq_apply first");
  if (qubit >= h->n) return fail(h, JKQ_E_ARGUMENT, "qubit %u out of range", qubit);
  CU(cudaSetDevice(h->device));
  return qubit_probability(h, qubit, p0);
}

extern "C" int jkq_get_mid_measurements(jkq_handle *h, uint64_t *bits, uint64_t *mask) {
  if (!h || !bits || !mask) return JKQ_E_ARGUMENT;
  *bits = h->mid_bits; *mask = h->mid_mask;
  return JKQ_E_OK;
}

extern "C" int jkq_dd_export(jkq_handle *h, jkq_snode *host_nodes, size_t cap, uint32_t *n_nodes, uint32_t *root) {
  if (!h || !n_nodes || !root) return JKQ_E_ARGUMENT;
  CU(cudaSetDevice(h->device));
  if (!h->imported) {
    int rc = ensure_sampling(h);
    if (rc) return rc;
  }
  *n_nodes = h->dd_nodes;
  *root = h->dd_root;
  if (!host_nodes) return JKQ_E_OK;
  if (cap < h->dd_nodes) return fail(h, JKQ_E_CAPACITY, "need room for %u nodes", h->dd_nodes);
  CU(jcopy(h, host_nodes, h->snodes, (size_t)h->dd_nodes * sizeof(JSNode), cudaMemcpyDeviceToHost));
  CU(cudaStreamSynchronize(h->stream));
  return JKQ_E_OK;
}

extern "C" int jkq_dd_import(jkq_handle *h, uint32_t n_qubits, const jkq_snode *host_nodes, uint32_t n_nodes, uint32_t root,
                             const uint8_t *clbit_of_level) {
  if (!h || !host_nodes || !clbit_of_level) return JKQ_E_ARGUMENT;
  if (n_qubits < 1 || n_qubits > JKQ_MAXQ || n_nodes > h->host.node_cap || root >= n_nodes) return fail(h, JKQ_E_ARGUMENT, "bad DD");
  CU(cudaSetDevice(h->device));
  CU(jcopy(h, h->snodes, host_nodes, (size_t)n_nodes * sizeof(JSNode), cudaMemcpyHostToDevice));
  CU(cudaStreamSynchronize(h->stream));
  h->n = n_qubits; h->dd_root = root; h->dd_nodes = n_nodes;
  memcpy(h->sp.bitpos, clbit_of_level, n_qubits);
  for (uint32_t L = n_qubits; L < JKQ_MAXQ; L++) h->sp.bitpos[L] = 255;
  h->imported = true; h->applied = false; h->sampling_ready = false; h->hist_valid = false; h->sn_epoch++;
  h->stats.norm = 1.0;
  return JKQ_E_OK;
}

// the same two calls with DEVICE buffers of the caller (any device of this process): what an NCCL broadcast of the
// sampling DD uses on both ends -- no host staging
extern "C" int jkq_dd_export_device(jkq_handle *h, void *dev_nodes, size_t cap, uint32_t *n_nodes, uint32_t *root) {
  if (!h || !n_nodes || !root) return JKQ_E_ARGUMENT;
  CU(cudaSetDevice(h->device));
  if (!h->imported) {
    int rc = ensure_sampling(h);
    if (rc) return rc;
  }
  *n_nodes = h->dd_nodes;
  *root = h->dd_root;
  if (!dev_nodes) return JKQ_E_OK;
  if (cap < h->dd_nodes) return fail(h, JKQ_E_CAPACITY, "need room for %u nodes", h->dd_nodes);
  CU(cudaMemcpyAsync(dev_nodes, h->snodes, (size_t)h->dd_nodes * sizeof(JSNode), cudaMemcpyDefault, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return JKQ_E_OK;
}

extern "C" int jkq_dd_import_device(jkq_handle *h, uint32_t n_qubits, const void *dev_nodes, uint32_t n_nodes, uint32_t root,
                                    const uint8_t *clbit_of_level) {
  if (!h || !dev_nodes || !clbit_of_level) return JKQ_E_ARGUMENT;
  if (n_qubits < 1 || n_qubits > JKQ_MAXQ || n_nodes > h->host.node_cap || root >= n_nodes) return fail(h, JKQ_E_ARGUMENT, "bad DD");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(h->snodes, dev_nodes, (size_t)n_nodes * sizeof(JSNode), cudaMemcpyDefault, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->n = n_qubits; h->dd_root = root; h->dd_nodes = n_nodes;
  memcpy(h->sp.bitpos, clbit_of_level, n_qubits);
  for (uint32_t L = n_qubits; L < JKQ_MAXQ; L++) h->sp.bitpos[L] = 255;
  h->imported = true; h->applied = false; h->sampling_ready = false; h->hist_valid = false; h->sn_epoch++;
  h->stats.norm = 1.0;
  return JKQ_E_OK;
}

// ---- shot sampling over the GPUs of ONE process ------------------------------------------------
// The handle that holds the state applies nothing here: its level-ordered sampling DD (16 bytes per node) goes to one
// helper handle per other device by peer copy over NVLink (cudaMemcpyPeerAsync, no host staging), device i draws the
// shots [i S / G, (i + 1) S / G) of the one Philox stream on its own stream and host thread, and the helpers'
// histograms come back compacted to (key, count) pairs, peer-copied and merged on the device (k_hist_merge) into the
// histogram of `src` -- which jkq_get_counts then returns, identical to a one-GPU run of the same seed.  An outcome
// space too large for one GPU stays distributed (st->merged = 0, distinct_total = the sum).
struct MultiPeer { int device = -1; jkq_handle *h = nullptr; };
static std::map<jkq_handle *, std::vector<MultiPeer>> g_multi_peers;  // helper handles, kept for the life of `src`

static void multi_free(jkq_handle *src) {
  auto it = g_multi_peers.find(src);
  if (it == g_multi_peers.end()) return;
  for (MultiPeer &p : it->second) if (p.h) jkq_destroy(p.h);
  g_multi_peers.erase(it);
}

extern "C" int jkq_sample_multi_gpu(jkq_handle *h, const int32_t *devices, uint32_t n_devices, uint64_t shots, uint64_t seed,
                                    jkq_multi_stats *st) {
  if (!h || !devices || n_devices < 1 || n_devices > 64) return JKQ_E_ARGUMENT;
  jkq_handle *const src = h;
  const auto t_begin = std::chrono::steady_clock::now();
  CU(cudaSetDevice(src->device));
  if (!src->imported) {
    int rc = ensure_sampling(src);
    if (rc) return rc;
  }
  // helper handles: small pools, room for the DD
  std::vector<MultiPeer> &peers = g_multi_peers[src];
  std::vector<jkq_handle *> hs(n_devices, nullptr);
  for (uint32_t i = 0; i < n_devices; i++) {
    if (devices[i] == src->device && hs[i] == nullptr) {
      bool taken = false;
      for (uint32_t j = 0; j < i; j++) taken |= hs[j] == src;
      if (!taken) { hs[i] = src; continue; }
    }
    MultiPeer *found = nullptr;
    for (MultiPeer &p : peers) {
      bool used = false;
      for (uint32_t j = 0; j < i; j++) used |= hs[j] == p.h;
      if (p.device == devices[i] && !used && p.h->host.node_cap >= src->dd_nodes) { found = &p; break; }
    }
    if (!found) {
      jkq_config cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.device = devices[i];
      cfg.node_capacity = pow2_at_least((uint64_t)src->dd_nodes + 1, 1u << 12);
      cfg.value_capacity = 1u << 10; cfg.item_capacity = 1u << 10;
      jkq_handle *nh = nullptr;
      int rc = jkq_create(&nh, &cfg);
      if (rc) {
        std::string msg = nh ? nh->err : "out of memory";
        if (nh) jkq_destroy(nh);
        return fail(src, rc, "helper engine on device %d: %s", devices[i], msg.c_str());
      }
      MultiPeer mp; mp.device = devices[i]; mp.h = nh;
      peers.push_back(mp);
      found = &peers.back();
      int can = 0;  // direct NVLink path when the topology offers it (errors: already enabled / not supported -- copies still work)
      if (cudaDeviceCanAccessPeer(&can, devices[i], src->device) == cudaSuccess && can) { cudaSetDevice(devices[i]); cudaDeviceEnablePeerAccess(src->device, 0); }
      if (cudaDeviceCanAccessPeer(&can, src->device, devices[i]) == cudaSuccess && can) { cudaSetDevice(src->device); cudaDeviceEnablePeerAccess(devices[i], 0); }
      cudaGetLastError();
    }
    hs[i] = found->h;
  }
  // broadcast of the sampling DD: device to device
  const auto t_bc0 = std::chrono::steady_clock::now();
  CU(cudaSetDevice(src->device));
  CU(cudaStreamSynchronize(src->stream));
  for (uint32_t i = 0; i < n_devices; i++) {
    jkq_handle *ph = hs[i];
    if (ph == src) continue;
    cudaSetDevice(ph->device);
    cudaError_t e = cudaMemcpyPeerAsync(ph->snodes, ph->device, src->snodes, src->device, (size_t)src->dd_nodes * sizeof(JSNode), ph->stream);
    if (e != cudaSuccess) return fail(src, JKQ_E_CUDA, "peer copy to device %d: %s", ph->device, cudaGetErrorString(e));
    ph->n = src->n; ph->dd_root = src->dd_root; ph->dd_nodes = src->dd_nodes; ph->sp = src->sp;
    ph->imported = true; ph->applied = false; ph->sampling_ready = false; ph->hist_valid = false; ph->stats.norm = 1.0; ph->sn_epoch++;
  }
  for (uint32_t i = 0; i < n_devices; i++)
    if (hs[i] != src) { cudaSetDevice(hs[i]->device); cudaStreamSynchronize(hs[i]->stream); }
  const auto t_bc1 = std::chrono::steady_clock::now();
  // can everything be merged on src?  (2 x 16 bytes per possible outcome against the free memory there)
  int measured = 0;
  for (uint32_t L = 0; L < src->n; L++) measured += src->sp.bitpos[L] < 64;
  uint64_t outcomes = shots;
  if (measured < 40 && (1ull << measured) < outcomes) outcomes = 1ull << measured;
  size_t free_b = 0, total_b = 0;
  cudaSetDevice(src->device);
  cudaMemGetInfo(&free_b, &total_b);
  // (beyond 2^26 possible outcomes the merged histogram is a result no caller reads back as a dictionary, and merging
  // it on one device would serialise what the devices did side by side)
  const bool merge = n_devices == 1 || (outcomes <= (1ull << 26) && outcomes * 2 * sizeof(JHSlot) + outcomes * 32 < free_b / 2);
  // sample: one host thread per device (jkq_sample blocks on its stream)
  std::vector<int> rcs(n_devices, 0);
  std::vector<size_t> compacted(n_devices, 0);
  std::vector<std::thread> th;
  for (uint32_t i = 0; i < n_devices; i++) {
    const uint64_t lo = shots / n_devices * i + (i < shots % n_devices ? i : shots % n_devices);
    const uint64_t cnt = shots / n_devices + (i < shots % n_devices ? 1 : 0);
    jkq_handle *ph = hs[i];
    const uint64_t expect = (ph == src && merge) ? shots : cnt;
    // (a helper also compacts its histogram right away, in its own thread: the merge below only copies and adds)
    th.emplace_back([=, &rcs, &compacted]() {
      rcs[i] = sample_impl(ph, cnt, seed, lo, 0, expect);
      if (!rcs[i] && merge && ph != src) {
        size_t n = 0;
        rcs[i] = jkq_get_counts(ph, nullptr, nullptr, 0, &n);
        if (!rcs[i] && n) rcs[i] = jkq_get_counts(ph, nullptr, nullptr, n, &n);  // into ph->ok / ph->oc on its device (keys == NULL: no host copy)
        compacted[i] = n;
      }
    });
  }
  for (std::thread &t : th) t.join();
  for (uint32_t i = 0; i < n_devices; i++)
    if (rcs[i]) return hs[i] == src ? rcs[i] : fail(src, rcs[i], "device %d: %s", hs[i]->device, hs[i]->err.c_str());
  const auto t_s1 = std::chrono::steady_clock::now();
  double k_ms = 0;
  for (uint32_t i = 0; i < n_devices; i++) if (hs[i]->stats.sample_kernel_ms > k_ms) k_ms = hs[i]->stats.sample_kernel_ms;
  // merge on the device: compact on the helper, peer copy, k_hist_merge on src
  uint64_t distinct_total = 0;
  if (merge) {
    CU(cudaSetDevice(src->device));
    size_t most = 0;
    for (uint32_t i = 0; i < n_devices; i++) if (compacted[i] > most) most = compacted[i];
    if (most > src->out_cap) {
      cudaFree(src->ok); cudaFree(src->oc);
      src->ok = src->oc = nullptr;
      CU(cudaMalloc(&src->ok, most * 8));
      CU(cudaMalloc(&src->oc, most * 8));
      src->out_cap = most;
    }
    for (uint32_t i = 0; i < n_devices; i++) {
      jkq_handle *ph = hs[i];
      const size_t n = compacted[i];
      if (ph == src || n == 0) continue;
      CU(cudaMemcpyPeerAsync(src->ok, src->device, ph->ok, ph->device, n * 8, src->stream));
      CU(cudaMemcpyPeerAsync(src->oc, src->device, ph->oc, ph->device, n * 8, src->stream));
      int blocks = (int)((n + 255) / 256);
      if (blocks > src->sm_count * 8) blocks = src->sm_count * 8;
      if (src->hist_packed) k_hist_merge<true><<<blocks, 256, 0, src->stream>>>(src->hs, (uint32_t)(src->hslots - 1), src->ok, src->oc, n, src->overflow, src->n_distinct);
      else k_hist_merge<false><<<blocks, 256, 0, src->stream>>>(src->hs, (uint32_t)(src->hslots - 1), src->ok, src->oc, n, src->overflow, src->n_distinct);
      CU(cudaGetLastError());  // (the next peer's pairs reuse src->ok / src->oc: same stream, in order)
    }
    int ovf = 0;
    CU(jcopy(src, &ovf, src->overflow, 4, cudaMemcpyDeviceToHost));
    CU(cudaStreamSynchronize(src->stream));
    if (ovf) return fail(src, JKQ_E_CAPACITY, "count histogram overflow in the merge");
    size_t n = 0;
    int rc = jkq_get_counts(src, nullptr, nullptr, 0, &n);
    if (rc) return rc;
    distinct_total = n;
  } else {
    for (uint32_t i = 0; i < n_devices; i++) {
      size_t n = 0;
      int rc = jkq_get_counts(hs[i], nullptr, nullptr, 0, &n);
      if (rc) return fail(src, rc, "device %d: %s", hs[i]->device, hs[i]->err.c_str());
      distinct_total += n;  // (an outcome drawn on two devices counts twice: an upper bound)
    }
  }
  CU(cudaSetDevice(src->device));
  const auto t_end = std::chrono::steady_clock::now();
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  if (st) {
    st->n_devices = n_devices; st->merged = merge ? 1u : 0u; st->distinct_total = distinct_total;
    st->broadcast_ms = ms(t_bc0, t_bc1); st->sample_wall_ms = ms(t_bc1, t_s1); st->sample_kernel_ms_max = k_ms;
    st->merge_ms = ms(t_s1, t_end); st->wall_ms = ms(t_begin, t_end); st->dd_bytes = (uint64_t)src->dd_nodes * sizeof(JSNode);
  }
  return JKQ_E_OK;
}

extern "C" int jkq_get_measure_map(jkq_handle *h, uint8_t *clbit_of_level, size_t n) {
  if (!h || !clbit_of_level || n < h->n) return JKQ_E_ARGUMENT;
  memcpy(clbit_of_level, h->sp.bitpos, h->n);
  return JKQ_E_OK;
}

// =============================================================================== batches of circuits
// Independent circuits shard over SMs.  ONE persistent kernel, one CTA per engine of the pool
// (at most one per SM: a CTA owns ~216 KB of shared memory); every CTA pulls circuits from a queue and runs
// the whole path for each -- init, gates (single-CTA path), sampling records, shots, histogram compaction
// into the result block -- without any host involvement in between.  A circuit the single-CTA path cannot
// finish (frontier or DD too large) is reported in status[] and left to the caller's full engine.
struct JBatchEng {   // device-side view of one engine of the pool
  JEng *E;
  JSNode *sn;
  double *prob;
  JHSlot *hs;
};
struct JBatchCirc {
  uint32_t n, n_gates;
  uint64_t gate_off, seed;
  SampleParams sp;
};
struct JBatchOut {
  int32_t status;
  uint32_t n_distinct;
  uint64_t gates;
  uint64_t t0_ns, t1_ns;  // JKQ_PROFILE: %globaltimer when the instance started / ended (else 0)
};
__device__ __forceinline__ uint64_t jglobaltimer() {
#ifdef JKQ_PROFILE
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#else
  return 0;
#endif
}

__global__ void __launch_bounds__(256, 1) k_batch(const JBatchEng *engs, const JBatchCirc *circ, const JGate *gates, unsigned int n_circ,
                                                 const unsigned int *order, unsigned int *next, uint64_t shots, uint32_t hslots, uint32_t cap,
                                                 JBatchOut *out, unsigned long long *keys, unsigned long long *counts) {
  __shared__ unsigned int s_i, s_cnt;
  __shared__ int s_ovf;
  __shared__ unsigned int s_fresh_unused;
  const JBatchEng be = engs[blockIdx.x];
  JEng *E = be.E;
  for (;;) {
    if (threadIdx.x == 0) s_i = atomicAdd(next, 1u);
    __syncthreads();
    const unsigned int q = s_i;
    __syncthreads();
    if (q >= n_circ) break;
    const unsigned int i = order[q];  // longest gate lists first: a heavy circuit must not be the last one to start
    const uint64_t t_begin = jglobaltimer();
    const JBatchCirc cd = circ[i];
    if (threadIdx.x == 0) { E->n = cd.n; E->lvl_log = nullptr; E->mat_u = nullptr; E->mat_epoch = nullptr; E->cursor_saved = 0; }
    {
      Ctx c = make_ctx<MODE_CTA>(E);
      init_state<MODE_CTA>(c);
      __threadfence();
      __syncthreads();
      if (cd.n_gates) run_gates<MODE_CTA>(c, gates + cd.gate_off, (int)cd.n_gates);
    }
    __threadfence();
    __syncthreads();
    int st = *(volatile int *)&E->status;
    if (!st && (*(volatile int *)&E->next_gate < (int)cd.n_gates || *(volatile int *)&E->want_grid || *(volatile unsigned int *)&E->cursor_saved))
      st = JKQ_ERR_ITEM_FULL;  // outgrew the single-CTA path
    if (!st) {
      Ctx c = make_ctx<MODE_CTA>(E);
      build_sampling<MODE_CTA>(c, be.sn, be.prob);
      __threadfence();
      __syncthreads();
      st = *(volatile int *)&E->status;
      if (!st && *(volatile int *)&E->want_grid) st = JKQ_ERR_NODE_FULL;
    }
    unsigned int nd = 0;
    if (!st) {
      // MeasureAll x shots (Simulator.cpp:148-224): this CTA's threads walk the shots of the circuit
      for (uint32_t k = threadIdx.x; k <= hslots; k += blockDim.x) { be.hs[k].key = JH_EMPTY; be.hs[k].cnt = 0; }
      if (threadIdx.x == 0) { s_cnt = 0; s_ovf = 0; }
      __syncthreads();
      const uint64_t rounded = (shots + 31) & ~31ull;
      const uint32_t lane = threadIdx.x & 31;
      for (uint64_t s = threadIdx.x; s < rounded; s += blockDim.x) {
        const bool active = s < shots;
        const uint64_t key = active ? sample_walk(be.sn, 1u, (int)cd.n, cd.sp.bitpos, cd.seed, s) : JH_EMPTY;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key) & __ballot_sync(0xFFFFFFFFu, active);
        if (active && lane == (uint32_t)(__ffs(peers) - 1)) hist_global_add(be.hs, hslots - 1, key, __popc(peers), &s_ovf, &s_fresh_unused);
      }
      __syncthreads();
      for (uint32_t k = threadIdx.x; k < hslots; k += blockDim.x)
        if (be.hs[k].key != JH_EMPTY) {
          const unsigned int p = atomicAdd(&s_cnt, 1u);
          if (p < cap) { keys[(size_t)i * cap + p] = be.hs[k].key; counts[(size_t)i * cap + p] = be.hs[k].cnt; }
        }
      if (threadIdx.x == 0 && be.hs[hslots].cnt) {
        const unsigned int p = atomicAdd(&s_cnt, 1u);
        if (p < cap) { keys[(size_t)i * cap + p] = JH_EMPTY; counts[(size_t)i * cap + p] = be.hs[hslots].cnt; }
      }
      __syncthreads();
      nd = s_cnt;
      if (s_ovf || nd > cap) st = 103;  // JKQ_E_CAPACITY
    }
    if (threadIdx.x == 0) {
      JBatchOut o;
      o.status = st; o.n_distinct = nd; o.gates = (uint64_t)E->st[JST_GATES];
      o.t0_ns = t_begin; o.t1_ns = jglobaltimer();
      out[i] = o;
      E->status = 0; E->want_grid = 0; E->cursor_saved = 0;
    }
    __syncthreads();
  }
}

// ---- per-shot re-simulation as one kernel ------------------------------------------------------------------
// A circuit with a gate behind a measurement cannot reuse its final state: the reference parses and simulates it
// again for every shot (QASMsimulator.cpp:501-522).  Here `n_shots` instances of the SAME gate list go through the
// queue of k_batch's scheme, one CTA per SM; an instance handles MeasureOne / ResetQubit / `if` itself: it stops
// before the op's gate slot, builds the sampling records, takes the qubit's marginal (marginal_block), draws the
// outcome from the Philox stream (measure seed of the instance, index of the op) -- the same numbers the host path
// of jkq_apply draws -- writes the projector (and X) into two private gate slots and runs them as ordinary gates.
struct JShotSlot {   // one per gate slot of the circuit
  uint32_t kind;     // 0 gate | 1 measure projector | 2,3 reset (reference: X-or-identity, projector) | 4,5 reset (physical: projector, X-or-identity)
  uint32_t clbit;
  uint32_t op_index; // Philox counter of the draw
  uint32_t if_skip;  // > 0: an `if` guards the if_skip slots starting here
  uint64_t if_mask, if_value;
};

__device__ double philox_uniform_dev(uint64_t seed, uint64_t counter) {  // twin of philox_uniform (host)
  uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), 0x4D454153u, 0};
  philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return ((double)c[0] + 0.5) * (1.0 / 4294967296.0);
}

#ifdef JKQ_PROFILE
__device__ unsigned long long g_shot_prof[8];  // cycles: 0 plain gates 1 build_sampling 2 marginal 3 projector gates 4 final sampling 5 init; 6 shots
#define SPROF_T(v) long long v = clock64()
#define SPROF_ADD(slot, a, b) do { if (threadIdx.x == 0) atomicAdd(&g_shot_prof[slot], (unsigned long long)((b) - (a))); } while (0)
#else
#define SPROF_T(v) do {} while (0)
#define SPROF_ADD(slot, a, b) do {} while (0)
#endif
__global__ void __launch_bounds__(256, 1) k_batch_shots(const JBatchEng *engs, JBatchCirc cd, const JGate *gates, const JShotSlot *slots,
                                                       JGate *priv_all, unsigned int n_shots, unsigned int *next, uint64_t measure_seed0,
                                                       uint64_t shot_offset, JBatchOut *out, unsigned long long *keys,
                                                       unsigned long long *mid_bits) {
  __shared__ unsigned int s_i;
  __shared__ unsigned long long s_bits;
  __shared__ int s_h;
  const JBatchEng be = engs[blockIdx.x];
  JEng *E = be.E;
  JGate *priv = priv_all + 2 * blockIdx.x;
  const int n_gates = (int)cd.n_gates;
  for (;;) {
    if (threadIdx.x == 0) s_i = atomicAdd(next, 1u);
    __syncthreads();
    const unsigned int inst = s_i;
    __syncthreads();
    if (inst >= n_shots) break;
    SPROF_T(q0);
    if (threadIdx.x == 0) { E->n = cd.n; E->lvl_log = nullptr; E->mat_u = nullptr; E->mat_epoch = nullptr; E->cursor_saved = 0; s_bits = 0; }
    Ctx c = make_ctx<MODE_CTA>(E);
    init_state<MODE_CTA>(c);
    __threadfence();
    __syncthreads();
    SPROF_T(q1);
    SPROF_ADD(5, q0, q1);
    SPROF_ADD(6, 0, 1);
    // The whole shot runs on this CTA's shared memory: the first segment loads the state (mode_enter), the later
    // ones find it where the previous one left it (run_gates enter = false / leave = false), and the measurements
    // work on the shared-memory DD (cta_norms / cta_marginal).  Nothing goes back to the arena.
    ShotScratch *X = reinterpret_cast<ShotScratch *>(c.sh->items);
    int st = 0, g = 0;
    bool resident = false;
    uint64_t placeholders = 0;  // projector / X slots applied: not ApplyGate calls
    // after a segment: did the single-CTA path finish it?
    auto segment_status = [&](int upto) {
      __threadfence();
      __syncthreads();
      int r = *(volatile int *)&E->status;
      if (!r && (c.sh->next_gate < upto || *(volatile int *)&E->want_grid || *(volatile unsigned int *)&E->cursor_saved))
        r = JKQ_ERR_ITEM_FULL;  // outgrew the single-CTA path
      return r;
    };
    auto segment = [&](const JGate *gl, int upto) {
      c = make_ctx<MODE_CTA>(E);
      run_gates<MODE_CTA>(c, gl, upto, !resident, false);
      resident = true;
      return segment_status(upto);
    };
    while (!st && g < n_gates) {
      const JShotSlot sl = slots[g];
      if (sl.if_skip && (s_bits & sl.if_mask) != sl.if_value) {  // `if` is false: step over the guarded slots
        g += (int)sl.if_skip;
        __syncthreads();
        if (threadIdx.x == 0) { if (resident) c.sh->next_gate = g; else E->next_gate = g; }
        __threadfence();
        __syncthreads();
        continue;
      }
      if (sl.kind == 0) {
        // plain gates up to the next slot that needs a decision
        if (threadIdx.x == 0) {
          int h = g + 1;
          while (h < n_gates && slots[h].kind == 0 && slots[h].if_skip == 0) h++;
          s_h = h;
        }
        __syncthreads();
        const int h = s_h;
        __syncthreads();
        SPROF_T(a0);
        st = segment(gates, h);
        SPROF_T(a1);
        SPROF_ADD(0, a0, a1);
        g = h;
        continue;
      }
      // MeasureOne / ResetQubit: P(qubit reads 0) in the state as it is now
      SPROF_T(b0);
      if (!resident) {  // a measurement before any gate: only load the state
        c = make_ctx<MODE_CTA>(E);
        const uint32_t f = mode_enter<MODE_CTA>(c);
        resident = true;
        if (f) { st = (f & JF_ERROR) ? *(volatile int *)&E->status : JKQ_ERR_NODE_FULL; if (!st) st = JKQ_ERR_INTERNAL; break; }
      }
      if (!cta_norms(c, X)) { st = JKQ_ERR_NODE_FULL; break; }  // (more than 1024 live nodes: left to the full engine)
      SPROF_T(b1);
      SPROF_ADD(1, b0, b1);
      const int level = gates[g].target;
      double p0 = cta_marginal(c, X, level);
      SPROF_T(b2);
      SPROF_ADD(2, b1, b2);
      if (p0 < 0.0) p0 = 0.0;
      if (p0 > 1.0) p0 = 1.0;
      const double norm = X->norm > 0.0 ? X->norm : 1.0;
      if (threadIdx.x == 0) {
        JGate ga = gates[g], gb = gates[g + (sl.kind == 1 ? 0 : 1)];
        for (int k = 0; k < 8; k++) { ga.m[k] = 0.0; gb.m[k] = 0.0; }
        if (sl.kind == 2) {  // the reference's reset: projection on |0>, X first when P(0) = 0
          double p = p0;
          if (p0 <= 0.0) { ga.m[2] = 1.0; ga.m[4] = 1.0; p = 1.0; }
          else { ga.m[0] = 1.0; ga.m[6] = 1.0; }
          gb.m[0] = 1.0 / sqrt(p * norm);
        } else {
          const double u = philox_uniform_dev(measure_seed0 + inst, (uint64_t)sl.op_index);
          const int outcome = u < p0 ? 0 : 1;
          const double p = outcome ? 1.0 - p0 : p0;
          if (!(p > 0.0)) E->status = JKQ_ERR_ZERO_STATE;
          else ga.m[outcome ? 6 : 0] = 1.0 / sqrt(p * norm);
          if (sl.kind == 1) {
            if (outcome) s_bits |= 1ull << sl.clbit;
            else s_bits &= ~(1ull << sl.clbit);
          } else if (outcome) { gb.m[2] = 1.0; gb.m[4] = 1.0; }
          else { gb.m[0] = 1.0; gb.m[6] = 1.0; }
        }
        priv[0] = ga;
        priv[1] = gb;
      }
      __threadfence();
      __syncthreads();
      st = *(volatile int *)&E->status;
      if (st) break;
      const int n_slots = sl.kind == 1 ? 1 : 2;
      st = segment(priv - g, g + n_slots);  // gate index g is priv[0]
      SPROF_T(b3);
      SPROF_ADD(3, b2, b3);
      g += n_slots;
      placeholders += (uint64_t)n_slots;
    }
    unsigned long long key = 0;
    SPROF_T(f0);
    if (!st) {
      if (!resident) {
        c = make_ctx<MODE_CTA>(E);
        const uint32_t f = mode_enter<MODE_CTA>(c);
        resident = true;
        if (f) st = (f & JF_ERROR) ? JKQ_ERR_INTERNAL : JKQ_ERR_NODE_FULL;
      }
      if (!st && !cta_norms(c, X)) st = JKQ_ERR_NODE_FULL;
      // MeasureAll once (Simulator.cpp:148-224): shot (shot_offset + inst) of the sampling stream, the walk of
      // sample_walk on the shared-memory lists
      if (!st && threadIdx.x == 0) {
        const uint64_t shot = shot_offset + inst;
        const int n = (int)cd.n;
        uint32_t idx = 0, pc[4] = {0, 0, 0, 0};
        for (int step = 0; step < n; step++) {
          if ((step & 3) == 0) {
            pc[0] = (uint32_t)shot; pc[1] = (uint32_t)(shot >> 32); pc[2] = (uint32_t)(step >> 2); pc[3] = 0;
            philox4x32(pc, (uint32_t)cd.seed, (uint32_t)(cd.seed >> 32));
          }
          const uint32_t word = (step & 3) == 0 ? pc[0] : (step & 3) == 1 ? pc[1] : (step & 3) == 2 ? pc[2] : pc[3];
          const double u = ((double)word + 0.5) * (1.0 / 4294967296.0);
          const JNode nd = c.sh->nodes[X->list[idx]];
          const uint32_t bit = !(u < X->q0[idx]);
          const uint32_t bp = cd.sp.bitpos[n - 1 - step];
          if (bit && bp < 64) key |= 1ull << bp;
          const uint32_t ch = bit ? nd.c1 : nd_c0(nd);
          if (step + 1 < n) idx = X->pos[ch] - 1u;
        }
      }
    }
    SPROF_T(f1);
    SPROF_ADD(4, f0, f1);
    __syncthreads();
    if (threadIdx.x == 0) {
      JBatchOut o;
      o.t0_ns = 0; o.t1_ns = 0;
      o.status = st; o.n_distinct = st ? 0u : 1u;
      o.gates = (uint64_t)E->st[JST_GATES] + (resident ? (uint64_t)c.sh->st[JST_GATES] : 0ull) - placeholders;
      out[inst] = o;
      keys[inst] = key;
      if (mid_bits) mid_bits[inst] = s_bits;
      E->status = 0; E->want_grid = 0; E->cursor_saved = 0;
    }
    __syncthreads();
  }
}

struct jkq_batch {
  int device = 0, sm_count = 0;
  jkq_config cfg;
  std::vector<jkq_handle *> eng;
  std::string err;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  JBatchEng *engs_dev = nullptr;
  unsigned int *next_dev = nullptr;
  // grow-only device / pinned buffers of a run
  JGate *gates_dev = nullptr, *gates_pin = nullptr;
  size_t gates_cap = 0;
  JBatchCirc *circ_dev = nullptr, *circ_pin = nullptr;
  unsigned int *order_dev = nullptr, *order_pin = nullptr;  // queue order of k_batch (same capacity as circ_*)
  JBatchOut *out_dev = nullptr, *out_pin = nullptr;
  size_t circ_cap = 0;
  unsigned long long *res_dev = nullptr, *res_pin = nullptr;  // keys then counts
  size_t res_cap = 0;
  uint64_t hslots = 0;
  JGate *priv_dev = nullptr;          // two private gate slots per engine (k_batch_shots)
  JShotSlot *slots_dev = nullptr;
  size_t slots_cap = 0;
  bool engs_uploaded = false;
};

extern "C" const char *jkq_batch_last_error(const jkq_batch *b) { return b ? b->err.c_str() : "null batch"; }

extern "C" void jkq_batch_destroy(jkq_batch *b) {
  if (!b) return;
  cudaSetDevice(b->device);
  for (jkq_handle *h : b->eng) jkq_destroy(h);
  cudaFree(b->engs_dev); cudaFree(b->next_dev); cudaFree(b->gates_dev); cudaFree(b->circ_dev); cudaFree(b->out_dev); cudaFree(b->res_dev);
  cudaFree(b->priv_dev); cudaFree(b->slots_dev); cudaFree(b->order_dev);
  if (b->order_pin) cudaFreeHost(b->order_pin);
  if (b->gates_pin) cudaFreeHost(b->gates_pin);
  if (b->circ_pin) cudaFreeHost(b->circ_pin);
  if (b->out_pin) cudaFreeHost(b->out_pin);
  if (b->res_pin) cudaFreeHost(b->res_pin);
  if (b->ev0) { cudaEventDestroy(b->ev0); cudaEventDestroy(b->ev1); }
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

static int bfail(jkq_batch *b, int code, const char *what, cudaError_t e) {
  b->err = std::string(what) + ": " + cudaGetErrorString(e);
  return code;
}
#define BCU(call)                                                       \
  do {                                                                  \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess) return bfail(b, JKQ_E_CUDA, #call, e_);      \
  } while (0)

extern "C" int jkq_batch_create(jkq_batch **out, int device, int n_engines, const jkq_config *cfg_in) {
  if (!out || n_engines < 1) return JKQ_E_ARGUMENT;
  *out = nullptr;
  jkq_batch *b = new jkq_batch();
  *out = b;
  b->device = device;
  jkq_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  if (cfg_in) cfg = *cfg_in;
  cfg.device = device;
  if (!cfg.node_capacity) cfg.node_capacity = 1u << 15;
  if (!cfg.value_capacity) cfg.value_capacity = 1u << 13;
  if (!cfg.item_capacity) cfg.item_capacity = 1u << 12;  // (the HBM frontier pool only serves the collector's scratch list here)
  if (!cfg.gc_min) cfg.gc_min = 1u << 12;
  if (!cfg.cta_threads) cfg.cta_threads = 128;
  b->cfg = cfg;
  BCU(cudaSetDevice(device));
  BCU(cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, device));
  if (n_engines > b->sm_count) n_engines = b->sm_count;  // one CTA per SM: more engines would only queue
  for (int i = 0; i < n_engines; i++) {
    jkq_handle *h = nullptr;
    int rc = jkq_create(&h, &cfg);
    if (rc) {
      b->err = h ? h->err : "jkq_create failed";
      if (h) jkq_destroy(h);
      return rc;
    }
    b->eng.push_back(h);
  }
  BCU(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  BCU(cudaEventCreate(&b->ev0));
  BCU(cudaEventCreate(&b->ev1));
  BCU(cudaMalloc(&b->engs_dev, b->eng.size() * sizeof(JBatchEng)));
  BCU(cudaMalloc(&b->next_dev, 4));
  BCU(cudaFuncSetAttribute(k_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaSmem)));
  BCU(cudaFuncSetAttribute(k_batch_shots, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaSmem)));
  BCU(cudaMalloc(&b->priv_dev, b->eng.size() * 2 * sizeof(JGate)));
  // the control blocks of the engines: written once (the kernel initialises the per-circuit part itself)
  for (jkq_handle *h : b->eng) {
    JEng &E = h->host;
    E.lvl_log = nullptr; E.mat_u = nullptr; E.mat_epoch = nullptr;
    BCU(cudaMemcpy(h->dev, &E, sizeof(JEng), cudaMemcpyHostToDevice));
  }
  return JKQ_E_OK;
}

extern "C" int jkq_batch_run(jkq_batch *b, size_t n_circuits, const uint32_t *n_qubits, const jkq_op *ops, const uint64_t *op_offsets,
                             uint64_t shots, const uint64_t *seeds, size_t cap, uint64_t *keys, uint64_t *counts, uint32_t *n_distinct,
                             int32_t *status, uint64_t *gates_applied, double *apply_ms_sum) {
  if (!b || !n_qubits || !ops || !op_offsets || !seeds || !keys || !counts || !n_distinct || !status || cap == 0) return JKQ_E_ARGUMENT;
  BCU(cudaSetDevice(b->device));
  b->err.clear();
  if (gates_applied) *gates_applied = 0;
  if (apply_ms_sum) *apply_ms_sum = 0;
  if (n_circuits == 0) return JKQ_E_OK;
  // ---- host: parse every circuit into gate descriptors (concatenated) + a descriptor per circuit
  std::vector<JGate> gates;
  std::vector<JBatchCirc> cd(n_circuits);
  std::vector<int32_t> pre(n_circuits, 0);
  int max_measured = 0;
  for (size_t i = 0; i < n_circuits; i++) {
    JBatchCirc &c = cd[i];
    memset(&c, 0, sizeof(c));
    c.n = n_qubits[i]; c.gate_off = gates.size(); c.seed = seeds[i];
    std::string err;
    bool snap = false;
    const size_t before = gates.size();
    int rc = parse_ops(&err, n_qubits[i], ops + op_offsets[i], (size_t)(op_offsets[i + 1] - op_offsets[i]), gates, nullptr, c.sp, nullptr, &snap);
    if (!rc && snap) { rc = JKQ_E_ARGUMENT; err = "snapshots are not served by the batch path"; }
    if (rc) {  // left out of the launch
      gates.resize(before);
      c.n = 1; c.n_gates = 0;
      pre[i] = rc;
      if (b->err.empty()) b->err = err;
      continue;
    }
    c.n_gates = (uint32_t)(gates.size() - before);
    int measured = 0;
    for (uint32_t L = 0; L < c.n; L++) measured += c.sp.bitpos[L] < 64;
    if (measured > max_measured) max_measured = measured;
  }
  // ---- buffers (grow-only)
  uint64_t distinct = shots < (1ull << 30) ? shots : (1ull << 30);
  if (max_measured < 30 && (1ull << max_measured) < distinct) distinct = 1ull << max_measured;
  const uint64_t hslots = pow2_at_least(distinct * 2, 1u << 10);
  if (hslots > b->hslots) {
    for (jkq_handle *h : b->eng) {
      free_hist(h);
      CU(cudaMalloc(&h->hs, (hslots + 1) * sizeof(JHSlot)));
      h->hslots = hslots;
      h->hist_packed = false;
    }
    b->hslots = hslots;
    b->engs_uploaded = true;
    std::vector<JBatchEng> be(b->eng.size());
    for (size_t e = 0; e < b->eng.size(); e++) {
      jkq_handle *h = b->eng[e];
      be[e].E = h->dev; be[e].sn = h->snodes; be[e].prob = h->prob; be[e].hs = (JHSlot *)h->hs;
    }
    BCU(cudaMemcpy(b->engs_dev, be.data(), be.size() * sizeof(JBatchEng), cudaMemcpyHostToDevice));
  }
  if (gates.size() > b->gates_cap) {
    cudaFree(b->gates_dev);
    if (b->gates_pin) cudaFreeHost(b->gates_pin);
    b->gates_cap = pow2_at_least(gates.size(), 1024);
    BCU(cudaMalloc(&b->gates_dev, b->gates_cap * sizeof(JGate)));
    BCU(cudaMallocHost(&b->gates_pin, b->gates_cap * sizeof(JGate)));
  }
  if (n_circuits > b->circ_cap) {
    cudaFree(b->circ_dev); cudaFree(b->out_dev);
    if (b->circ_pin) cudaFreeHost(b->circ_pin);
    if (b->out_pin) cudaFreeHost(b->out_pin);
    b->circ_cap = pow2_at_least(n_circuits, 64);
    cudaFree(b->order_dev);
    if (b->order_pin) cudaFreeHost(b->order_pin);
    BCU(cudaMalloc(&b->order_dev, b->circ_cap * sizeof(unsigned int)));
    BCU(cudaMallocHost(&b->order_pin, b->circ_cap * sizeof(unsigned int)));
    BCU(cudaMalloc(&b->circ_dev, b->circ_cap * sizeof(JBatchCirc)));
    BCU(cudaMalloc(&b->out_dev, b->circ_cap * sizeof(JBatchOut)));
    BCU(cudaMallocHost(&b->circ_pin, b->circ_cap * sizeof(JBatchCirc)));
    BCU(cudaMallocHost(&b->out_pin, b->circ_cap * sizeof(JBatchOut)));
  }
  const size_t res_words = 2 * n_circuits * cap;
  if (res_words > b->res_cap) {
    cudaFree(b->res_dev);
    if (b->res_pin) cudaFreeHost(b->res_pin);
    b->res_cap = res_words;
    BCU(cudaMalloc(&b->res_dev, res_words * 8));
    BCU(cudaMallocHost(&b->res_pin, res_words * 8));
  }
  // ---- one launch
  memcpy(b->gates_pin, gates.data(), gates.size() * sizeof(JGate));
  memcpy(b->circ_pin, cd.data(), n_circuits * sizeof(JBatchCirc));
  {
    // queue order: longest gate list first.  The gate count is the only cost known up front; it keeps the circuits
    // most likely to be slow (large DDs come with long gate lists) away from the end of the queue, where one late
    // starter would leave 147 SMs idle.
    std::vector<unsigned int> ord(n_circuits);
    for (size_t i = 0; i < n_circuits; i++) ord[i] = (unsigned int)i;
    std::stable_sort(ord.begin(), ord.end(), [&](unsigned int x, unsigned int y) { return cd[x].n_gates > cd[y].n_gates; });
    memcpy(b->order_pin, ord.data(), n_circuits * sizeof(unsigned int));
  }
  BCU(cudaMemcpyAsync(b->order_dev, b->order_pin, n_circuits * sizeof(unsigned int), cudaMemcpyHostToDevice, b->stream));
  BCU(cudaMemcpyAsync(b->gates_dev, b->gates_pin, gates.size() * sizeof(JGate), cudaMemcpyHostToDevice, b->stream));
  BCU(cudaMemcpyAsync(b->circ_dev, b->circ_pin, n_circuits * sizeof(JBatchCirc), cudaMemcpyHostToDevice, b->stream));
  BCU(cudaMemsetAsync(b->next_dev, 0, 4, b->stream));
  const unsigned int blocks = (unsigned int)(n_circuits < b->eng.size() ? n_circuits : b->eng.size());
  BCU(cudaEventRecord(b->ev0, b->stream));
  k_batch<<<blocks, b->cfg.cta_threads, sizeof(CtaSmem), b->stream>>>(b->engs_dev, b->circ_dev, b->gates_dev, (unsigned int)n_circuits, b->order_dev,
                                                                     b->next_dev, shots, (uint32_t)hslots, (uint32_t)cap, b->out_dev, b->res_dev,
                                                                     b->res_dev + n_circuits * cap);
  BCU(cudaGetLastError());
  BCU(cudaEventRecord(b->ev1, b->stream));
  BCU(cudaMemcpyAsync(b->out_pin, b->out_dev, n_circuits * sizeof(JBatchOut), cudaMemcpyDeviceToHost, b->stream));
  BCU(cudaMemcpyAsync(b->res_pin, b->res_dev, res_words * 8, cudaMemcpyDeviceToHost, b->stream));
  BCU(cudaStreamSynchronize(b->stream));
  uint64_t total = 0;
  for (size_t i = 0; i < n_circuits; i++) {
    const JBatchOut &o = b->out_pin[i];
    status[i] = pre[i] ? pre[i] : o.status;
    n_distinct[i] = 0;
    if (status[i]) {
      if (b->err.empty()) b->err = "circuit " + std::to_string(i) + ": " + (o.status == 103 ? "more distinct outcomes than `cap`" : status_text(o.status));
      continue;
    }
    n_distinct[i] = o.n_distinct;
    memcpy(keys + i * cap, b->res_pin + i * cap, (size_t)o.n_distinct * 8);
    memcpy(counts + i * cap, b->res_pin + (n_circuits + i) * cap, (size_t)o.n_distinct * 8);
    total += o.gates;
  }
  if (gates_applied) *gates_applied = total;
  float ms = 0;
  if (apply_ms_sum && cudaEventElapsedTime(&ms, b->ev0, b->ev1) == cudaSuccess) *apply_ms_sum = ms;
#ifdef JKQ_PROFILE
  {
    uint64_t first = ~0ull, last = 0, sum = 0, longest = 0, late = 0;
    size_t longest_i = 0;
    for (size_t i = 0; i < n_circuits; i++) {
      const JBatchOut &o = b->out_pin[i];
      if (!o.t1_ns) continue;
      if (o.t0_ns < first) first = o.t0_ns;
      if (o.t1_ns > last) { last = o.t1_ns; late = i; }
      sum += o.t1_ns - o.t0_ns;
      if (o.t1_ns - o.t0_ns > longest) { longest = o.t1_ns - o.t0_ns; longest_i = i; }
    }
    if (last > first)
      fprintf(stderr, "k_batch: span %.2f ms, sum of circuit times %.2f ms (/%u CTAs = %.2f ms), longest circuit %zu: %.2f ms (status %d, %u gates), last to finish %zu: started at %.2f ms, took %.2f ms\n",
              (last - first) * 1e-6, sum * 1e-6, blocks, sum * 1e-6 / blocks, longest_i, longest * 1e-6, b->out_pin[longest_i].status, cd[longest_i].n_gates,
              (size_t)late, (b->out_pin[late].t0_ns - first) * 1e-6, (b->out_pin[late].t1_ns - b->out_pin[late].t0_ns) * 1e-6);
  }
#endif
  return JKQ_E_OK;
}

extern "C" int jkq_batch_run_shots(jkq_batch *b, uint32_t n_qubits, const jkq_op *ops, size_t n_ops, uint64_t shots, uint64_t seed,
                                   uint64_t shot_offset, uint64_t measure_seed0, int physical_reset, uint64_t *keys, uint64_t *mid_bits,
                                   int32_t *status, uint64_t *gates_applied, double *kernel_ms) {
  if (!b || (n_ops && !ops) || !keys || !status) return JKQ_E_ARGUMENT;
  BCU(cudaSetDevice(b->device));
  b->err.clear();
  if (gates_applied) *gates_applied = 0;
  if (kernel_ms) *kernel_ms = 0;
  if (shots == 0) return JKQ_E_OK;
  if (shots > 0xFFFFFFF0ull) { b->err = "too many shots for one launch"; return JKQ_E_ARGUMENT; }
  // ---- host: gate descriptors (placeholder slots for measure / reset) and what each slot is
  std::vector<JGate> gates;
  std::vector<int> gate_of_op;
  JBatchCirc cd;
  memset(&cd, 0, sizeof(cd));
  cd.n = n_qubits; cd.seed = seed;
  {
    std::string err;
    int rc = parse_ops(&err, n_qubits, ops, n_ops, gates, &gate_of_op, cd.sp, nullptr, nullptr);
    if (rc) { b->err = err; return rc; }
  }
  cd.n_gates = (uint32_t)gates.size();
  gate_of_op.push_back((int)gates.size());
  std::vector<JShotSlot> slots(gates.size() + 1);
  memset(slots.data(), 0, slots.size() * sizeof(JShotSlot));
  for (size_t i = 0; i < n_ops; i++) {
    const jkq_op &o = ops[i];
    const int g = gate_of_op[i];
    if (o.kind == JKQ_OP_MEASURE_ONE) {
      slots[g].kind = 1; slots[g].clbit = o.arg; slots[g].op_index = (uint32_t)i;
    } else if (o.kind == JKQ_OP_RESET) {
      slots[g].kind = physical_reset ? 4 : 2; slots[g + 1].kind = physical_reset ? 5 : 3;
      slots[g].op_index = (uint32_t)i;
    } else if (o.kind == JKQ_OP_IF) {
      const size_t cnt = o.arg ? o.arg : 1;
      const int first = gate_of_op[i + 1], last = gate_of_op[i + 1 + cnt];
      if (last > first) { slots[first].if_skip = (uint32_t)(last - first); slots[first].if_mask = o.ctrl_pos; slots[first].if_value = o.ctrl_neg; }
    }  // (SNAPSHOT / PROBABILITIES ops record nothing on this path)
  }
  // ---- buffers (grow-only)
  if (!b->engs_uploaded) {
    std::vector<JBatchEng> be(b->eng.size());
    for (size_t e = 0; e < b->eng.size(); e++) {
      jkq_handle *h = b->eng[e];
      be[e].E = h->dev; be[e].sn = h->snodes; be[e].prob = h->prob; be[e].hs = (JHSlot *)h->hs;
    }
    BCU(cudaMemcpy(b->engs_dev, be.data(), be.size() * sizeof(JBatchEng), cudaMemcpyHostToDevice));
    b->engs_uploaded = true;
  }
  if (gates.size() + 1 > b->gates_cap) {
    cudaFree(b->gates_dev);
    if (b->gates_pin) cudaFreeHost(b->gates_pin);
    b->gates_cap = pow2_at_least(gates.size() + 1, 1024);
    BCU(cudaMalloc(&b->gates_dev, b->gates_cap * sizeof(JGate)));
    BCU(cudaMallocHost(&b->gates_pin, b->gates_cap * sizeof(JGate)));
  }
  if (slots.size() > b->slots_cap) {
    cudaFree(b->slots_dev);
    b->slots_cap = pow2_at_least(slots.size(), 1024);
    BCU(cudaMalloc(&b->slots_dev, b->slots_cap * sizeof(JShotSlot)));
  }
  if (shots > b->circ_cap) {
    cudaFree(b->circ_dev); cudaFree(b->out_dev);
    if (b->circ_pin) cudaFreeHost(b->circ_pin);
    if (b->out_pin) cudaFreeHost(b->out_pin);
    b->circ_cap = pow2_at_least(shots, 64);
    cudaFree(b->order_dev);
    if (b->order_pin) cudaFreeHost(b->order_pin);
    BCU(cudaMalloc(&b->order_dev, b->circ_cap * sizeof(unsigned int)));
    BCU(cudaMallocHost(&b->order_pin, b->circ_cap * sizeof(unsigned int)));
    BCU(cudaMalloc(&b->circ_dev, b->circ_cap * sizeof(JBatchCirc)));
    BCU(cudaMalloc(&b->out_dev, b->circ_cap * sizeof(JBatchOut)));
    BCU(cudaMallocHost(&b->circ_pin, b->circ_cap * sizeof(JBatchCirc)));
    BCU(cudaMallocHost(&b->out_pin, b->circ_cap * sizeof(JBatchOut)));
  }
  const size_t res_words = 2 * (size_t)shots;
  if (res_words > b->res_cap) {
    cudaFree(b->res_dev);
    if (b->res_pin) cudaFreeHost(b->res_pin);
    b->res_cap = res_words;
    BCU(cudaMalloc(&b->res_dev, res_words * 8));
    BCU(cudaMallocHost(&b->res_pin, res_words * 8));
  }
  // ---- one launch
  if (!gates.empty()) {
    memcpy(b->gates_pin, gates.data(), gates.size() * sizeof(JGate));
    BCU(cudaMemcpyAsync(b->gates_dev, b->gates_pin, gates.size() * sizeof(JGate), cudaMemcpyHostToDevice, b->stream));
  }
  BCU(cudaMemcpyAsync(b->slots_dev, slots.data(), slots.size() * sizeof(JShotSlot), cudaMemcpyHostToDevice, b->stream));
  BCU(cudaMemsetAsync(b->next_dev, 0, 4, b->stream));
  const unsigned int blocks = (unsigned int)(shots < b->eng.size() ? shots : b->eng.size());
  BCU(cudaEventRecord(b->ev0, b->stream));
  k_batch_shots<<<blocks, b->cfg.cta_threads, sizeof(CtaSmem), b->stream>>>(b->engs_dev, cd, b->gates_dev, b->slots_dev, b->priv_dev, (unsigned int)shots,
                                                                           b->next_dev, measure_seed0, shot_offset, b->out_dev, b->res_dev,
                                                                           b->res_dev + shots);
  BCU(cudaGetLastError());
  BCU(cudaEventRecord(b->ev1, b->stream));
  BCU(cudaMemcpyAsync(b->out_pin, b->out_dev, shots * sizeof(JBatchOut), cudaMemcpyDeviceToHost, b->stream));
  BCU(cudaMemcpyAsync(b->res_pin, b->res_dev, res_words * 8, cudaMemcpyDeviceToHost, b->stream));
  BCU(cudaStreamSynchronize(b->stream));  // (also covers the pageable copy of `slots`)
  uint64_t total = 0;
  for (uint64_t i = 0; i < shots; i++) {
    const JBatchOut &o = b->out_pin[i];
    status[i] = o.status;
    keys[i] = b->res_pin[i];
    if (mid_bits) mid_bits[i] = b->res_pin[shots + i];
    if (o.status) {
      if (b->err.empty()) b->err = "shot " + std::to_string(i) + ": " + status_text(o.status);
      continue;
    }
    total += o.gates;
  }
  if (gates_applied) *gates_applied = total;
  float ms = 0;
  if (kernel_ms && cudaEventElapsedTime(&ms, b->ev0, b->ev1) == cudaSuccess) *kernel_ms = ms;
#ifdef JKQ_PROFILE
  {
    unsigned long long p[8], z[8] = {0};
    cudaMemcpyFromSymbol(p, g_shot_prof, sizeof(p));
    cudaMemcpyToSymbol(g_shot_prof, z, sizeof(z));
    if (p[6])
      fprintf(stderr, "k_batch_shots cycles/shot over %llu shots: init %llu  plain gates %llu  build_sampling %llu  marginal %llu  projector gates %llu  final sampling %llu\n",
              p[6], p[5] / p[6], p[0] / p[6], p[1] / p[6], p[2] / p[6], p[3] / p[6], p[4] / p[6]);
  }
#endif
  return JKQ_E_OK;
}

extern "C" int jkq_random_sector_bench(jkq_handle *h, uint64_t bytes, int iters, double *gsectors_per_s, double *gcas_per_s) {
  if (!h || !gsectors_per_s || !gcas_per_s) return JKQ_E_ARGUMENT;
  CU(cudaSetDevice(h->device));
  if (bytes < (1ull << 20)) bytes = 1ull << 20;
  if (iters < 1) iters = 1;
  uint4 *buf = nullptr;
  unsigned long long *sink = nullptr;
  CU(cudaMalloc(&buf, bytes));
  CU(cudaMalloc(&sink, 8));
  CU(cudaMemsetAsync(buf, 0, bytes, h->stream));
  CU(cudaMemsetAsync(sink, 0, 8, h->stream));
  const int blocks = h->sm_count * 8, threads = 256, per_thread = 256;
  const double ops = (double)blocks * threads * per_thread;
  float best_rd = 1e30f, best_cas = 1e30f;
  for (int it = 0; it < iters + 1; it++) {
    CU(cudaEventRecord(h->ev2, h->stream));
    k_random_sector<<<blocks, threads, 0, h->stream>>>(buf, bytes / 32, per_thread, 0x1234 + it, sink);
    CU(cudaEventRecord(h->ev3, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float ms;
    CU(cudaEventElapsedTime(&ms, h->ev2, h->ev3));
    if (it > 0 && ms < best_rd) best_rd = ms;
  }
  for (int it = 0; it < iters + 1; it++) {
    CU(cudaMemsetAsync(buf, 0, bytes, h->stream));
    CU(cudaEventRecord(h->ev2, h->stream));
    k_random_cas<<<blocks, threads, 0, h->stream>>>((unsigned long long *)buf, bytes / 8, per_thread, 0x9876 + it);
    CU(cudaEventRecord(h->ev3, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float ms;
    CU(cudaEventElapsedTime(&ms, h->ev2, h->ev3));
    if (it > 0 && ms < best_cas) best_cas = ms;
  }
  cudaFree(buf); cudaFree(sink);
  *gsectors_per_s = ops / (best_rd * 1e-3) * 1e-9;
  *gcas_per_s = ops / (best_cas * 1e-3) * 1e-9;
  return JKQ_E_OK;
}
