// Device side of the B200 QMDD engine: value table, unique table, frontier de-dup, the per-item
// expand / reduce steps of the level-synchronous gate application.  Hand-written CUDA for sm_100a.
//
// The same code runs in two launch shapes (template parameter MODE):
//   MODE_CTA   one CTA, __syncthreads between phases.  The whole working set of a SMALL decision
//              diagram lives in that SM's shared memory (~210 KB): node records, unique table, value
//              table, frontier pool, de-dup tables, the current gate.  A gate on a 32-node DD is a
//              dependent chain of ~2n level steps, so its cost is latency: an L2 round trip is ~0.3 us
//              on B200, a shared-memory access ~15 ns.  The HBM arena is written through (nodes, reals)
//              so the state is always complete in HBM; the HBM hash tables go stale and are rebuilt by
//              the collector when the grid path takes over.
//   MODE_GRID  cooperative grid, software grid barrier, every table in HBM, ld.global.cg everywhere --
//              the throughput path for blown-up DDs.
// With JKQ_EMU defined the header compiles as plain C++ and runs as ONE emulated thread (warp width 1).
// That build exists only under tests/emu (logic debugging in a GPU-less container); the product never
// loads it.
#ifndef JKQ_ENGINE_CUH
#define JKQ_ENGINE_CUH

#include "jkq_types.h"

#define MODE_CTA 0
#define MODE_GRID 1

// flags exchanged at a barrier
#define JF_RETRY 1u
#define JF_ERROR 2u
#define JF_GRID 4u

#ifdef JKQ_EMU
#include <math.h>
#include <string.h>
#define JD static inline
#define JDM inline
#define JNI static
#define JKQ_UNROLL
#define JKQ_WARP 1u
#define JKQ_DMUL(a, b) ((a) * (b))
#define JKQ_DADD(a, b) ((a) + (b))
#define JKQ_DSUB(a, b) ((a) - (b))
#define JKQ_DDIV(a, b) ((a) / (b))
template <class T> static inline T jatomic_add(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T jatomic_cas(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
static inline void jatomic_min(unsigned long long *p, unsigned long long v) { if (v < *p) *p = v; }
template <int MODE> static inline void jfence() {}
static inline unsigned long long jd2u(double x) { unsigned long long u; memcpy(&u, &x, 8); return u; }
static inline double ju2d(unsigned long long u) { double x; memcpy(&x, &u, 8); return x; }
static inline uint32_t jballot(bool p) { return p ? 1u : 0u; }
static inline uint32_t jshfl(uint32_t v, uint32_t) { return v; }
static inline uint32_t jmatch(uint32_t) { return 1u; }
static inline uint32_t jmatch64(unsigned long long) { return 1u; }
static inline unsigned long long jshfl64(unsigned long long v, uint32_t) { return v; }
static inline void jsyncwarp() {}
static inline uint32_t jpopc(uint32_t x) { return (uint32_t)__builtin_popcount(x); }
static inline uint32_t jffs(uint32_t x) { return (uint32_t)__builtin_ffs((int)x); }
#else
#define JD __device__ __forceinline__
#define JDM __device__ __forceinline__
#define JNI __device__ __noinline__
#define JKQ_UNROLL _Pragma("unroll")
#define JKQ_WARP 32u
#define JKQ_DMUL(a, b) __dmul_rn((a), (b))
#define JKQ_DADD(a, b) __dadd_rn((a), (b))
#define JKQ_DSUB(a, b) __dsub_rn((a), (b))
#define JKQ_DDIV(a, b) __ddiv_rn((a), (b))
template <class T> JD T jatomic_add(T *p, T v) { return atomicAdd(p, v); }
template <class T> JD T jatomic_cas(T *p, T cmp, T v) { return atomicCAS(p, cmp, v); }
JD void jatomic_min(unsigned long long *p, unsigned long long v) { atomicMin(p, v); }
// (the grid path only ever publishes with it: record, fence, CAS / store of the word that names the record.  A/B builds:
// -DJKQ_FENCE_ACQREL=1 issues fence.acq_rel.gpu, which orders exactly that, instead of the sequentially consistent
// membar.gl of __threadfence())
#ifndef JKQ_FENCE_ACQREL
#define JKQ_FENCE_ACQREL 0
#endif
template <int MODE> JD void jfence() {
  if (MODE == MODE_CTA) __threadfence_block();
  else if (JKQ_FENCE_ACQREL) asm volatile("fence.acq_rel.gpu;" ::: "memory");
  else __threadfence();
}
JD unsigned long long jd2u(double x) { return (unsigned long long)__double_as_longlong(x); }
JD double ju2d(unsigned long long u) { return __longlong_as_double((long long)u); }
JD uint32_t jballot(bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
JD uint32_t jshfl(uint32_t v, uint32_t src) { return __shfl_sync(0xFFFFFFFFu, v, (int)src); }
JD uint32_t jmatch(uint32_t v) { return __match_any_sync(0xFFFFFFFFu, v); }
JD uint32_t jmatch64(unsigned long long v) { return __match_any_sync(0xFFFFFFFFu, v); }
JD unsigned long long jshfl64(unsigned long long v, uint32_t src) { return __shfl_sync(0xFFFFFFFFu, v, (int)src); }
JD void jsyncwarp() { __syncwarp(); }
JD uint32_t jpopc(uint32_t x) { return (uint32_t)__popc(x); }
JD uint32_t jffs(uint32_t x) { return (uint32_t)__ffs((int)x); }
#endif

// ------------------------------------------------------------------------------------ MODE_CTA state
#define JKQ_SITEMS 512u      // work items of one gate
#define JKQ_SDT_SLOTS 2048u  // F/S de-dup slots (generation << 16 | item index + 1)
#define JKQ_SNODES 2048u     // node ids held in shared memory; the collector keeps ids below this
#define JKQ_SUT_SLOTS 4096u  // unique-table slots (u32 node id)
#define JKQ_SVALS 1024u      // reals
#define JKQ_SVT_SLOTS 2048u  // value-table slots
#define JKQ_SPEND 512u       // value-table slots claimed in one phase
#define JKQ_SMATS 128u       // cached canonical ids of gate matrices
#ifndef JKQ_SMEMO
#define JKQ_SMEMO 1024u      // entries of the complex-operation memo (16 bytes each)
#endif
#define JKQ_SCHAIN 512u      // words of chain-compression scratch per gate (a chain of k nodes takes k + 1)
#define JKQ_SCMEMO 256u      // entries of the chain-rebuild memo (8 bytes each)
#define JKQ_SDEFER 8u        // items per level that were created for a level further down
struct CtaState {
  uint32_t item_count, node_count, val_count, item_limit;
  uint32_t node_limit, val_limit, chain_dirty, solo_now;  // solo_now: the cursor warp is the only one at work (real_lookup)
  uint32_t val_lock, pad_l[3];  // pool sizes the gate code may use (a reserve stays for cursor_to_root)
  uint32_t node_hard, val_hard, node_soft, val_soft;
  JHand hand[2];
  uint32_t pend_n[3], err;
  uint32_t want_grid, gen, pad2, pad3;
  uint32_t lsw, root, flushed_nodes, flushed_vals;  // [flushed, count): not yet copied to the HBM arena
  uint32_t nofit;  // the collector found the state too large for the shared pools: it lives in the arena only
  uint64_t rootw;
  uint32_t st[JST_N + 1];
  int32_t next_gate, pad0;
  unsigned long long prof[16];
  uint32_t bflags[4];
  double tol, inv_tol;
  uint64_t cur_u[4];
  uint32_t lb[JKQ_MAXQ + 1], le[JKQ_MAXQ + 1];
  uint32_t defer_n[JKQ_MAXQ];               // chain compression: items waiting for level L ...
  uint16_t defer[JKQ_MAXQ][JKQ_SDEFER];     // ... their indices (0xFFFF: slot reserved but not needed)
  uint32_t chain_n;
  int32_t min_def;                          // lowest level with a waiting item
  unsigned long long chain[JKQ_SCHAIN];     // per chain: bit mask of norm flags, then the k weights
  JGate gate[2];
  uint32_t mtag[JKQ_SMATS];
  uint64_t mu[JKQ_SMATS][4];
  unsigned long long memo[JKQ_SMEMO][2];    // (op, a, b) -> result id of cx_arith, see there
  unsigned long long cmemo[JKQ_SCMEMO];     // (item node, skipped levels, result below) -> rebuilt chain top, see reduce_item
  uint32_t pend[JKQ_SPEND];
  uint32_t mdir[JKQ_SNODES];
  uint32_t dtab[JKQ_SDT_SLOTS];
  uint32_t utab[JKQ_SUT_SLOTS];
  double vals[JKQ_SVALS];
  JVSlot vtab[JKQ_SVT_SLOTS];
  JNode nodes[JKQ_SNODES];
  JItem items[JKQ_SITEMS];
};

// ------------------------------------------------------------------------------------ thin path state
// Levels above the gate's highest control/target only forward a node's two edges.  While such a level
// holds at most one warp's worth of nodes, ONE warp walks it with the frontier in registers and these
// per-level records in shared memory (jkq_thin.cuh) instead of the generic item phases.
struct __attribute__((aligned(8))) ThinEnt {
  uint32_t pk;   // s0 | s1 << 8 | nz0 << 16 | nz1 << 17 | norm flag << 18
  uint32_t pad;
  uint64_t w;    // weight of the child that does not carry ONE
  uint64_t f;    // weight the cursor edge carried when the cursor moved below it: re-applied on the way up
};
struct __attribute__((aligned(16))) ThinRes {
  uint64_t rw;
  uint32_t rnode, pad;
};
struct ThinState {
  ThinRes res[2][JKQ_WARP];          // the cursor edges (level parity)
  ThinEnt ent[JKQ_MAXQ][JKQ_WARP];   // levels above the cursor (levels at and below it: scratch records of thin_gate)
  uint32_t cnt[JKQ_MAXQ + 1];
  int32_t M;                         // cursor level
  uint32_t walked;                   // nodes passed on the way down since the last cursor_to_root (statistics)
  // thin_gate (jkq_thin.cuh): a gate applied below the cursor edges without work items
  ThinRes sres[2][JKQ_WARP];         // on the way down: the entries of the next level; on the way up: the results of a level
  uint64_t vp[2 * JKQ_WARP];         // variants created at the target level: factor on the subspace "all controls
  uint64_t vq[2 * JKQ_WARP];         //   below satisfied" (vp) and on its complement (vq), normalised by the larger one
  uint32_t thin_off, pad_t;          // copy of JEng.thin_off (testing aid), set by cursor_reset
};

struct Ctx {
  JEng *E;
  CtaState *sh;         // MODE_CTA only
  ThinState *T;
  JItem *items;         // frontier pool: shared memory in MODE_CTA, HBM in MODE_GRID
  uint32_t tid, nth;    // index / count over all participating threads
  uint32_t phase;       // barrier counter, identical in every thread
  uint32_t ring;        // phase % 3
  uint32_t block0;      // 1 if this thread belongs to the CTA that runs the serial sections
  uint32_t ltid, lnth;
  uint32_t gprof;       // JKQ_PROFILE builds: this phase is narrow enough for the per-stage cycle counters (g_gprof)
};

// JKQ_PROFILE builds, grid path: where the cycles of a NARROW level phase go (per-thread clock64 deltas, summed; slot + 32 = count)
#if defined(JKQ_PROFILE) && !defined(JKQ_EMU)
__device__ unsigned long long g_gprof[64];
#define JGP_T(v) long long v = clock64()
#define JGP_ADD(c, slot, a, b) do { if ((c).gprof) { atomicAdd(&g_gprof[slot], (unsigned long long)((b) - (a))); atomicAdd(&g_gprof[32 + (slot)], 1ull); } } while (0)
#define JGP_SET(c, v) (c).gprof = (v)
#else
#define JGP_T(v) do {} while (0)
#define JGP_ADD(c, slot, a, b) do {} while (0)
#define JGP_SET(c, v) do {} while (0)
#endif

// Warp-aggregated increment of a counter that every lane of the warp may be bumping at this point: ONE atomic
// per warp instead of up to 32 on the same address (the bump allocators and statistics of the grid path are
// otherwise serialised by the L2 atomic unit).  Returns the lane's own slot.
#ifdef JKQ_EMU
JD uint32_t jagg_inc(uint32_t *p) { return (*p)++; }
JD void jagg_stat(unsigned long long *p, uint32_t v) { *p += v; }
#else
JD uint32_t jagg_inc(uint32_t *p) {
  const unsigned m = __match_any_sync(__activemask(), (unsigned long long)(uintptr_t)p);  // lanes bumping this very counter
  const int leader = __ffs(m) - 1;
  const unsigned lane = threadIdx.x & 31u;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(p, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}
JD void jagg_stat(unsigned long long *p, uint32_t v) {
  const unsigned m = __match_any_sync(__activemask(), (unsigned long long)(uintptr_t)p);
  const uint32_t sum = __reduce_add_sync(m, v);
  if ((int)(threadIdx.x & 31u) == __ffs(m) - 1) atomicAdd(p, (unsigned long long)sum);
}
#endif

// counters nobody waits for: atomics whose result is unused compile to fire-and-forget RED instructions,
// where a load-add-store would stall the (in-order) thread for an L2 round trip
// (MODE_CTA counts in shared memory: a global memory operation in flight makes the next
// __threadfence_block of the same thread wait for an L2 round trip)
#ifdef JKQ_EMU
template <int MODE> JD void jstat(JEng *E, CtaState *S, int which, uint32_t v) {
  if (MODE == MODE_CTA) S->st[which] += v;
  else E->st[which] += v;
}
template <int MODE> JD void jstat_max(JEng *E, CtaState *S, uint32_t v) {
  if (MODE == MODE_CTA) { if (v > S->st[JST_MAXF]) S->st[JST_MAXF] = v; }
  else if (v > E->st[JST_MAXF]) E->st[JST_MAXF] = v;
}
#else
template <int MODE> JD void jstat(JEng *E, CtaState *S, int which, uint32_t v) {
  if (MODE == MODE_CTA) atomicAdd(&S->st[which], v);
  else jagg_stat(&E->st[which], v);
}
template <int MODE> JD void jstat_max(JEng *E, CtaState *S, uint32_t v) {
  if (MODE == MODE_CTA) atomicMax(&S->st[JST_MAXF], v);
  else atomicMax(&E->st[JST_MAXF], (unsigned long long)v);
}
#endif

template <int MODE> JD uint32_t *cnt_items(const Ctx &c) { return MODE == MODE_CTA ? &c.sh->item_count : &c.E->item_count; }
template <int MODE> JD uint32_t *cnt_nodes(const Ctx &c) { return MODE == MODE_CTA ? &c.sh->node_count : &c.E->node_count; }
template <int MODE> JD uint32_t *cnt_vals(const Ctx &c) { return MODE == MODE_CTA ? &c.sh->val_count : &c.E->val_count; }
template <int MODE> JD uint32_t *cnt_pend(const Ctx &c, uint32_t r) { return MODE == MODE_CTA ? &c.sh->pend_n[r] : &c.E->pend_n[r]; }
template <int MODE> JD double jtol(const Ctx &c) { return MODE == MODE_CTA ? c.sh->tol : c.E->tol; }
template <int MODE> JD double jinv_tol(const Ctx &c) { return MODE == MODE_CTA ? c.sh->inv_tol : c.E->inv_tol; }

// ------------------------------------------------------------------------------------ loads
// ld_s: data not written since the last barrier (MODE_GRID: ld.global.cg, L1 may be stale).
// ld_r: data that may be written concurrently (HBM: ld.global.cg; shared memory: volatile).
#ifdef JKQ_EMU
template <int MODE, class T> JD T ld_s(const T *p) { return *p; }
template <int MODE, class T> JD T ld_r(const T *p) { return *p; }
#else
// 16-byte volatile load from shared memory (MODE_CTA tables that other threads write concurrently)
JD uint4 ld_vol16_shared(const void *p) {
  uint4 v;
  asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}
template <int MODE> JD uint32_t ld_s(const uint32_t *p) { return MODE == MODE_GRID ? __ldcg(p) : *p; }
template <int MODE> JD int32_t ld_s(const int32_t *p) { return MODE == MODE_GRID ? __ldcg(p) : *p; }
template <int MODE> JD unsigned long long ld_s(const unsigned long long *p) { return MODE == MODE_GRID ? __ldcg(p) : *p; }
template <int MODE> JD uint64_t ld_s(const uint64_t *p) {
  return MODE == MODE_GRID ? (uint64_t)__ldcg((const unsigned long long *)p) : *p;
}
template <int MODE> JD double ld_s(const double *p) { return MODE == MODE_GRID ? __ldcg(p) : *p; }
template <int MODE> JD uint8_t ld_s(const uint8_t *p) { return MODE == MODE_GRID ? __ldcg(p) : *p; }
template <int MODE> JD JNode ld_s(const JNode *p) {
  uint4 v = MODE == MODE_GRID ? __ldcg((const uint4 *)p) : *(const uint4 *)p;
  JNode n;
  n.c0 = v.x; n.c1 = v.y; n.w = ((uint64_t)v.w << 32) | v.z;
  return n;
}
template <int MODE> JD JVSlot ld_s(const JVSlot *p) {
  uint4 v = MODE == MODE_GRID ? __ldcg((const uint4 *)p) : ld_vol16_shared(p);
  JVSlot s;
  s.vbits = ((unsigned long long)v.y << 32) | v.x; s.keylo = v.z; s.id1 = v.w;
  return s;
}
template <int MODE> JD uint32_t ld_r(const uint32_t *p) { return MODE == MODE_GRID ? __ldcg(p) : *(const volatile uint32_t *)p; }
template <int MODE> JD int32_t ld_r(const int32_t *p) { return MODE == MODE_GRID ? __ldcg(p) : *(const volatile int32_t *)p; }
template <int MODE> JD unsigned long long ld_r(const unsigned long long *p) {
  return MODE == MODE_GRID ? __ldcg(p) : *(const volatile unsigned long long *)p;
}
template <int MODE> JD uint64_t ld_r(const uint64_t *p) {
  return MODE == MODE_GRID ? (uint64_t)__ldcg((const unsigned long long *)p) : (uint64_t)*(const volatile unsigned long long *)p;
}
template <int MODE> JD double ld_r(const double *p) { return MODE == MODE_GRID ? __ldcg(p) : *(const volatile double *)p; }
template <int MODE> JD JNode ld_r(const JNode *p) {
  uint4 v = MODE == MODE_GRID ? __ldcg((const uint4 *)p) : ld_vol16_shared(p);
  JNode n;
  n.c0 = v.x; n.c1 = v.y; n.w = ((uint64_t)v.w << 32) | v.z;
  return n;
}
#endif

template <int MODE> JD uint32_t rd_items(const Ctx &c) { return ld_r<MODE>(cnt_items<MODE>(c)); }
template <int MODE> JD uint32_t rd_nodes(const Ctx &c) { return ld_r<MODE>(cnt_nodes<MODE>(c)); }
template <int MODE> JD uint32_t rd_vals(const Ctx &c) { return ld_r<MODE>(cnt_vals<MODE>(c)); }
template <int MODE> JD bool jhas_error(const Ctx &c) { return MODE == MODE_CTA ? ld_r<MODE>(&c.sh->err) != 0 : ld_r<MODE>(&c.E->status) != 0; }

JD uint64_t jmix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
// The table hashes sit on the latency path of every item: 32-bit multiplies side by side, one
// xor-shift-multiply round to finish.
JD uint32_t hfin(uint32_t x) {
  x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 13;
  return x;
}
JD uint32_t h32(uint32_t x) { return hfin(x * 0x9E3779B1u); }
JD uint32_t hw64(uint64_t w) { return (uint32_t)w * 0xC2B2AE3Du + (uint32_t)(w >> 32) * 0x27D4EB2Fu; }

// Errors and the "does not fit the single-CTA pools" condition are sticky words (JEng.status,
// CtaState.err / want_grid) that every barrier folds into its flag word, so the out-of-line table
// functions below need no per-thread state.
template <int MODE> JD void jset_error(JEng *E, CtaState *S, int code) {
  jatomic_cas(&E->status, 0, code);
  if (MODE == MODE_CTA) S->err = (uint32_t)code;
}
// MODE_CTA: the gate does not fit the shared-memory pools; it is abandoned and redone on the grid path
JD void jwant_grid(CtaState *S) { S->want_grid = 1; }
template <int MODE> JD void jset_error_c(Ctx &c, int code) { jset_error<MODE>(c.E, c.sh, code); }

// ------------------------------------------------------------------------------------ value table
// Semantics of the reference's Clookup (QMDDcomplexD.cpp:372-466) with its tolerance comparator
// (:85-98): table entries are pairwise more than tol apart, a lookup returns the entry within tol of
// x, a miss appends x.  One bucket of width tol holds at most one entry, so a lookup probes the
// buckets k-1, k, k+1.  Anything within tol of 0 is index 0.
//
// Misses of a phase are inserted in parallel AND schedule-independently:
//   propose  the thread claims the slot of bucket k (64-bit CAS on (keylo, id1)) and atomicMin's its value
//            into it: after the phase every bucket holds the smallest value proposed for it;
//   commit   (next phase) a candidate becomes an entry unless the bucket below holds a candidate within
//            tol of it -- then it is dropped, and its proposers find the lower one on retry.
// Which values enter the table therefore depends only on the SET of values proposed in the phase.
//
// Inlining policy (measured on B200, QFT-32): a call costs the single warp ~100 cycles, so the table
// functions that run once or twice per item (unique table, item de-dup, the real lookups inside cx_arith)
// are inlined; the complex arithmetic itself (cx_arith, ~55 call sites) is not: inlining it as well buys
// 1 % and takes the kernels from 25 k to 90 k instructions and the build from 1 to 5 minutes.  Out-of-line
// functions take (E, S, ring) by value so that the caller's context stays in registers.
#define JKQ_VAL_MAX 1.0e8
#define JV_MISS 0xFFFFFFFFu
#define JW_MISS 0xFFFFFFFFFFFFFFFFull
template <int MODE> JD double jtol2(const JEng *E, const CtaState *S) { return MODE == MODE_CTA ? S->tol : E->tol; }
template <int MODE> JD double jinv_tol2(const JEng *E, const CtaState *S) { return MODE == MODE_CTA ? S->inv_tol : E->inv_tol; }
template <int MODE> JD double jval2(const JEng *E, const CtaState *S, uint32_t half) {
  uint32_t idx = half & 0x7FFFFFFFu;
  return MODE == MODE_CTA ? S->vals[idx] : ld_s<MODE>(E->vals + idx);
}
template <int MODE> JD double jval(const Ctx &c, uint32_t half) { return jval2<MODE>(c.E, c.sh, half); }
template <int MODE> JD double jre(const Ctx &c, uint64_t w) {
  double v = jval<MODE>(c, (uint32_t)(w >> 32));
  return (w >> 63) ? -v : v;
}
template <int MODE> JD double jim(const Ctx &c, uint64_t w) {
  double v = jval<MODE>(c, (uint32_t)w);
  return ((w >> 31) & 1) ? -v : v;
}
// squared magnitude of the canonical value (the reference compares the cached Cmag, :452-464)
template <int MODE> JD double jmag2(const Ctx &c, uint64_t w) {
  double r = jval<MODE>(c, (uint32_t)(w >> 32)), i = jval<MODE>(c, (uint32_t)w);
  return JKQ_DADD(JKQ_DMUL(r, r), JKQ_DMUL(i, i));
}
JD uint64_t jneg(uint64_t a) {  // Cnegative, QMDDcomplexD.cpp:483-496
  uint64_t r = a >> 32, i = a & 0xFFFFFFFFull;
  if (r) r ^= 0x80000000ull;
  if (i) i ^= 0x80000000ull;
  return (r << 32) | i;
}

// (JEng.vt_cur == vtab_mask from jkq_create on: a plain cached load -- ld.cg here put an L2 round trip in front of every probe)
template <int MODE> JD uint32_t vt_mask2(const JEng *E) { return MODE == MODE_CTA ? JKQ_SVT_SLOTS - 1 : E->vtab_mask; }
// the HBM value table has four slots per real of the value array
JD bool vt_room(const JEng *E, uint32_t idx) { return idx < E->val_cap && (uint64_t)idx * 2 < (uint64_t)E->vtab_mask; }
template <int MODE> JD JVSlot *vt_tab2(JEng *E, CtaState *S) { return MODE == MODE_CTA ? S->vtab : E->vtab; }
JD long long vt_bucket(double x, double inv_tol) { return (long long)floor(JKQ_DMUL(x, inv_tol)); }
// eight consecutive buckets share one 128-byte line (8 x 16-byte slots): the three probes of a lookup
// mostly touch one line; collisions step by whole lines
JD uint32_t vt_slot0(long long k, uint32_t mask) {
  uint64_t g = (uint64_t)k >> 3;
  uint32_t m = hfin((uint32_t)g * 0x9E3779B1u + (uint32_t)(g >> 32) * 0x85EBCA6Bu);
  return ((m << 3) | ((uint32_t)k & 7u)) & mask;
}
JD uint32_t vt_next(uint32_t h, uint32_t mask) { return (h + 8u) & mask; }
JD unsigned long long *vt_meta(JVSlot *s) { return reinterpret_cast<unsigned long long *>(&s->keylo); }
JD unsigned long long vt_meta_of(uint32_t keylo, uint32_t id1) { return (unsigned long long)keylo | ((unsigned long long)id1 << 32); }

// HBM table: remember that the line of slot h holds something (see JEng.vbm)
template <int MODE> JD void vt_mark(JEng *E, uint32_t h) {
  if (MODE != MODE_GRID) return;
  const uint32_t line = h >> 3, bit = 1u << (line & 31u);
  uint32_t *w = E->vbm + (line >> 5);
#ifdef JKQ_EMU
  *w |= bit;
#else
  atomicOr(w, bit);  // (result unused: a RED, nobody waits for it)
#endif
}

template <int MODE> JD void pend_push(JEng *E, CtaState *S, uint32_t ring, uint32_t slot) {
  uint32_t i = MODE == MODE_CTA ? jatomic_add(&S->pend_n[ring], 1u) : jagg_inc(&E->pend_n[ring]);
  if (MODE == MODE_CTA) {
    if (i < JKQ_SPEND) S->pend[i] = slot;
    else jwant_grid(S);
  } else {
    if (i < E->pend_cap) E->pend[i] = slot;
    else jset_error<MODE>(E, S, JKQ_ERR_PEND_FULL);
  }
}

// a value that missed the table: leave it as a candidate of its bucket
template <int MODE> JD void real_propose(JEng *E, CtaState *S, uint32_t ring, double x, long long k) {
  const uint32_t klo = (uint32_t)k, mask = vt_mask2<MODE>(E);
  const unsigned long long xb = jd2u(x), claim = vt_meta_of(klo, JV_PENDING);
  JVSlot *tab = vt_tab2<MODE>(E, S);
  uint32_t h = vt_slot0(k, mask);
  for (uint32_t probes = 0; probes <= (mask >> 3);) {
    unsigned long long *mp = vt_meta(tab + h);
    unsigned long long meta = ld_r<MODE>(mp);
    if (meta == 0) {
      meta = jatomic_cas(mp, 0ull, claim);
      if (meta == 0) { vt_mark<MODE>(E, h); jatomic_min(&tab[h].vbits, xb); pend_push<MODE>(E, S, ring, h); return; }
    }
    if ((uint32_t)meta == klo) {
      const uint32_t id1 = (uint32_t)(meta >> 32);
      if (id1 == JV_PENDING) { jatomic_min(&tab[h].vbits, xb); return; }
      if (id1 == JV_DROPPED) {
        if (jatomic_cas(mp, meta, claim) == meta) { jatomic_min(&tab[h].vbits, xb); pend_push<MODE>(E, S, ring, h); return; }
        continue;  // somebody else re-claimed it: look at the slot again
      }
    }
    h = vt_next(h, mask);
    probes++;
  }
  if (MODE == MODE_CTA) jwant_grid(S);
  else jset_error<MODE>(E, S, JKQ_ERR_VAL_FULL);
}

// where a value that missed the table goes at once instead of through propose / commit (A/B builds: -DJKQ_INPLACE_CTA=0 ...)
#ifndef JKQ_INPLACE_CTA
#define JKQ_INPLACE_CTA 0
#endif
#ifndef JKQ_INPLACE_GRID
#define JKQ_INPLACE_GRID 1
#endif
template <int MODE> JNI uint32_t real_insert_solo(JEng *E, CtaState *S, double x);
template <int MODE> JNI uint32_t real_insert_par(JEng *E, CtaState *S, double x, long long k);
// index of the entry within tol of x (x >= 0), or JV_MISS; with `propose` a miss leaves x as a candidate
template <int MODE> JD uint32_t real_lookup(JEng *E, CtaState *S, uint32_t ring, double x, bool propose) {
  const double tol = jtol2<MODE>(E, S);
  if (x <= tol) return 0;
  if (!(x < JKQ_VAL_MAX)) {  // absurd magnitude (or NaN)
    if (propose) jset_error<MODE>(E, S, JKQ_ERR_VAL_RANGE);
    return JV_MISS;
  }
  const long long k = vt_bucket(x, jinv_tol2<MODE>(E, S));
  const uint32_t mask = vt_mask2<MODE>(E);
  const JVSlot *tab = vt_tab2<MODE>(E, S);
  uint32_t h[3];
  JVSlot s[3];
  JKQ_UNROLL
  for (int d = 0; d < 3; d++) { h[d] = vt_slot0(k - 1 + d, mask); s[d] = ld_s<MODE>(tab + h[d]); }  // three loads in flight
  JKQ_UNROLL
  for (int d = 0; d < 3; d++) {
    const uint32_t klo = (uint32_t)(k - 1 + d);
    uint32_t hh = h[d];
    JVSlot sl = s[d];
    for (uint32_t probes = 0; sl.id1 != 0 && probes <= (mask >> 3); probes++) {
      if (sl.keylo == klo && sl.id1 < JV_WAITING && fabs(JKQ_DSUB(ju2d(sl.vbits), x)) <= tol) return sl.id1 - 1;
      hh = vt_next(hh, mask);
      sl = ld_s<MODE>(tab + hh);
    }
  }
  if (propose) {
    // in place (real_insert_par below: the item carries on) or as a candidate for the commit round behind the next barrier
    const bool in_place = MODE == MODE_CTA ? (JKQ_INPLACE_CTA || S->solo_now) : JKQ_INPLACE_GRID;
    if (in_place) return real_insert_par<MODE>(E, S, x, k);
    real_propose<MODE>(E, S, ring, x, k);
  }
  return JV_MISS;
}

// commit the candidates collected under pend counter `par` (the ring index of the phase that proposed
// them); threads tid, tid+nth, ...
template <int MODE> JNI void vals_commit(JEng *E, CtaState *S, uint32_t par, uint32_t tid, uint32_t nth) {
  uint32_t n = ld_s<MODE>(MODE == MODE_CTA ? &S->pend_n[par] : &E->pend_n[par]);
  const uint32_t cap = MODE == MODE_CTA ? JKQ_SPEND : E->pend_cap;
  if (n > cap) n = cap;
  const uint32_t *pend = MODE == MODE_CTA ? S->pend : E->pend;
  JVSlot *tab = vt_tab2<MODE>(E, S);
  const uint32_t mask = vt_mask2<MODE>(E);
  const double tol = jtol2<MODE>(E, S), inv_tol = jinv_tol2<MODE>(E, S);
  for (uint32_t i = tid; i < n; i += nth) {
    const uint32_t h = ld_s<MODE>(pend + i);
    const unsigned long long vb = ld_r<MODE>(&tab[h].vbits);
    const double v = ju2d(vb);
    const long long kk = vt_bucket(v, inv_tol) - 1;
    bool drop = false;
    uint32_t hh = vt_slot0(kk, mask);
    for (uint32_t probes = 0; probes <= (mask >> 3); probes++) {
      JVSlot sl = ld_s<MODE>(tab + hh);  // (a neighbour being committed right now reads as present or as gone: same final table)
      if (sl.id1 == 0) break;
      if (sl.keylo == (uint32_t)kk && sl.id1 != JV_DROPPED) {
        double d = JKQ_DSUB(v, ju2d(sl.vbits));
        if (d > 0.0 && d <= tol) { drop = true; break; }
      }
      hh = vt_next(hh, mask);
    }
    uint32_t id1 = JV_DROPPED;
    if (!drop) {
      uint32_t idx = MODE == MODE_CTA ? jatomic_add(&S->val_count, 1u) : jagg_inc(&E->val_count);
      if (MODE == MODE_CTA) {
        if (idx >= S->val_limit) jwant_grid(S);  // the shared table is full: grid path
        else { S->vals[idx] = v; id1 = idx + 1; jstat<MODE>(E, S, JST_VALS, 1u); }  // (copied to HBM by cta_flush)
      } else {
        if (!vt_room(E, idx)) jset_error<MODE>(E, S, JKQ_ERR_VAL_FULL);
        else { E->vals[idx] = v; id1 = idx + 1; jstat<MODE>(E, S, JST_VALS, 1u); }
      }
    }
    if (id1 == JV_DROPPED) tab[h].vbits = JV_INF;
    tab[h].id1 = id1;
  }
}

// single-thread insert while nobody else touches the table (init, gate set-up, root update, collector)
template <int MODE> JNI uint32_t real_insert_direct(JEng *E, CtaState *S, double x) {
  uint32_t idx;
  if (MODE == MODE_CTA) {
    if (S->val_count >= S->val_limit) { jwant_grid(S); return 0; }
    idx = S->val_count++;
    S->vals[idx] = x;  // (copied to HBM by cta_flush)
    jstat<MODE>(E, S, JST_VALS, 1u);
  } else {
    if (!vt_room(E, E->val_count)) { jset_error<MODE>(E, S, JKQ_ERR_VAL_FULL); return 0; }
    idx = E->val_count++;
    E->vals[idx] = x;
    jstat<MODE>(E, S, JST_VALS, 1u);
  }
  const long long k = vt_bucket(x, jinv_tol2<MODE>(E, S));
  const uint32_t mask = vt_mask2<MODE>(E);
  JVSlot *tab = vt_tab2<MODE>(E, S);
  uint32_t h = vt_slot0(k, mask);
  for (uint32_t probes = 0;; probes++) {
    JVSlot sl = ld_s<MODE>(tab + h);
    if (sl.id1 == 0 || (sl.id1 == JV_DROPPED && sl.keylo == (uint32_t)k)) break;
    if (probes > (mask >> 3)) {
      // every slot this bucket can probe is taken (entries plus the tombstones of dropped candidates): the real is
      // in the array but cannot be hashed -- give the gate up; the next collection rebuilds the table from the array
      if (MODE == MODE_CTA) jwant_grid(S);
      else jset_error<MODE>(E, S, JKQ_ERR_VAL_FULL);
      return idx;
    }
    h = vt_next(h, mask);
  }
  JVSlot sl;
  sl.vbits = jd2u(x); sl.keylo = (uint32_t)k; sl.id1 = idx + 1;
  jfence<MODE>();  // the real before the slot that names it (real_insert_solo: other lanes look things up meanwhile)
  tab[h] = sl;
  vt_mark<MODE>(E, h);
  return idx;
}
// A miss while the cursor warp works alone (solo levels, thin gates, cursor moves: every gate of a narrow diagram).
// Nobody but the lanes of this one warp touches the table, so the value goes in at once instead of through the
// propose / commit round that exists for concurrent phases -- and the item that needed it is not run a second time
// (Grover n=20 spent a fifth of its time in those second runs).  The lanes take turns under a lock in shared memory;
// under it a lane looks again (another lane may just have inserted a value within tol: first come, first kept, as
// in the reference, QMDDcomplexD.cpp:402-418) and appends only if the value is still missing.  JV_MISS only when the
// shared pools are exhausted (sticky want_grid: the gate is abandoned).
template <int MODE> JNI uint32_t real_insert_solo(JEng *E, CtaState *S, double x) {
  uint32_t r = JV_MISS;
#ifdef JKQ_EMU
  r = real_lookup<MODE>(E, S, 0, x, false);
  if (r == JV_MISS) { r = real_insert_direct<MODE>(E, S, x); if (S->want_grid) r = JV_MISS; }
#else
  bool done = false;
  while (!done) {
    if (atomicCAS(&S->val_lock, 0u, 1u) == 0u) {
      __threadfence_block();
      r = real_lookup<MODE>(E, S, 0, x, false);
      if (r == JV_MISS) {
        r = real_insert_direct<MODE>(E, S, x);
        if (*(volatile uint32_t *)&S->want_grid) r = JV_MISS;
      }
      __threadfence_block();
      atomicExch(&S->val_lock, 0u);
      done = true;
    }
  }
#endif
  return r;
}

// A miss during a phase that many threads run at once (levels wider than the cursor warp; every phase of the grid path).
// Round 1 collected the misses of a phase as candidates and committed them behind a barrier (real_propose / vals_commit):
// schedule-independent, but an item that met a new value was run again after the commit round -- and again for every
// value that depended on the one before (r160, profile build: 2.4 reruns per item with a miss; the reruns were 47 % of
// all item time, and the first run of such an item took 3.7 times as long as one without a miss).  With JKQ_INPLACE_GRID
// the value goes in at once, first come, first kept like the reference (QMDDcomplexD.cpp:402-418), without locks:
//   1. the thread claims the slot of bucket k by CAS (PENDING).  One slot per bucket: whoever loses waits for the owner's
//      id (same bucket = within tol) or, if the owner gives the slot up, looks the value up again;
//   2. the owner writes its value into the slot and looks at the buckets k-1 and k+1 (it could only issue these loads
//      after its CAS came back, so of two owners of adjacent buckets at least one sees the other).  A committed entry
//      within tol: the owner gives its slot up (DROPPED) and returns that entry.  An owner of k-1 still at work: this
//      one marks itself WAITING and waits for it; an owner of k+1 still PENDING: wait until it is WAITING (it defers) or
//      done.  Waits only go to the bucket below or to a transient state of the one above: no cycle;
//   3. the real is appended, and the id published last (fence in between).
// Lookups take no lock: PENDING / WAITING read as "no match, keep probing".  Which of two values within tol of each other
// becomes the entry depends on timing (they differ by rounding noise: amplitudes agree to ~1e-16 between runs, not
// bit for bit).
#ifndef JKQ_EMU
#define JKQ_SPIN_LIMIT 4000000u
#endif
// state of bucket kb's slot as seen now: returns its meta (0 = no slot), *hp = where (in: its first slot, whose meta the
// caller has loaded already)
JD unsigned long long vt_find_bucket(JEng *E, long long kb, uint32_t mask, uint32_t *hp, unsigned long long meta) {
  const uint32_t klo = (uint32_t)kb;
  uint32_t h = *hp;
  for (uint32_t probes = 0; probes <= (mask >> 3); probes++) {
    if (meta == 0) return 0;
    if ((uint32_t)meta == klo) { *hp = h; return meta; }
    h = vt_next(h, mask);
    meta = ld_r<MODE_GRID>(vt_meta(E->vtab + h));
  }
  return 0;
}
// wait until the slot at h (bucket key klo) leaves the states in `while_mask` (bit 0: PENDING, bit 1: WAITING)
JD unsigned long long vt_wait(JEng *E, uint32_t h, uint32_t while_mask) {
  unsigned long long meta = ld_r<MODE_GRID>(vt_meta(E->vtab + h));
#ifndef JKQ_EMU
  for (uint32_t spins = 0;; spins++) {
    const uint32_t id1 = (uint32_t)(meta >> 32);
    const bool busy = (id1 == JV_PENDING && (while_mask & 1u)) || (id1 == JV_WAITING && (while_mask & 2u));
    if (!busy) break;
    if (spins > JKQ_SPIN_LIMIT) { jset_error<MODE_GRID>(E, nullptr, JKQ_ERR_INTERNAL); break; }
    __nanosleep(40);
    meta = ld_r<MODE_GRID>(vt_meta(E->vtab + h));
  }
#endif
  return meta;
}
template <int MODE> JNI uint32_t real_insert_par(JEng *E, CtaState *S, double x, long long k) {
  if (MODE == MODE_CTA) return real_insert_solo<MODE>(E, S, x);
  const uint32_t klo = (uint32_t)k, mask = vt_mask2<MODE_GRID>(E);
  const double tol = E->tol;
  JVSlot *tab = E->vtab;
  for (uint32_t attempt = 0; attempt < 64u; attempt++) {
    if (attempt) {  // the owner of my bucket gave its slot up: the neighbourhood has changed, look again
      const uint32_t r = real_lookup<MODE>(E, S, 0, x, false);
      if (r != JV_MISS) return r;
    }
    // ---- 1. the slot of bucket k: mine, or somebody else's
    uint32_t h = vt_slot0(k, mask);
    bool mine = false, again = false;
    for (uint32_t probes = 0; probes <= (mask >> 3);) {
      unsigned long long *mp = vt_meta(tab + h);
      unsigned long long meta = ld_r<MODE_GRID>(mp);
      if (meta == 0 || ((uint32_t)meta == klo && (uint32_t)(meta >> 32) == JV_DROPPED)) {
        const unsigned long long old = jatomic_cas(mp, meta, vt_meta_of(klo, JV_PENDING));
        if (old == meta) { mine = true; break; }
        meta = old;
      }
      if ((uint32_t)meta == klo) {
        meta = vt_wait(E, h, 3u);
        const uint32_t id1 = (uint32_t)(meta >> 32);
        if (id1 < JV_WAITING && id1 != 0) return id1 - 1;  // same bucket: within tol by construction
        again = true;                                      // given up (or an error): start over
        break;
      }
      h = vt_next(h, mask);
      probes++;
    }
    if (again) continue;
    if (!mine) { jset_error<MODE_GRID>(E, nullptr, JKQ_ERR_VAL_FULL); return JV_MISS; }
    // ---- 2. I own bucket k.  My value first (an owner next door compares against it once I am done), then the neighbours
    unsigned long long *my_meta = vt_meta(tab + h);
    tab[h].vbits = jd2u(x);
    uint32_t found = JV_MISS;
    uint32_t hl = vt_slot0(k - 1, mask), hu = vt_slot0(k + 1, mask);
    unsigned long long ml = ld_r<MODE_GRID>(vt_meta(tab + hl)), mu = ld_r<MODE_GRID>(vt_meta(tab + hu));  // both in flight ...
    uint32_t idx = jagg_inc(&E->val_count);  // ... and so is the allocation (an index is lost if the slot is given up)
    ml = vt_find_bucket(E, k - 1, mask, &hl, ml);
    mu = vt_find_bucket(E, k + 1, mask, &hu, mu);
    if (ml != 0) {
      uint32_t id1 = (uint32_t)(ml >> 32);
      if (id1 == JV_PENDING || id1 == JV_WAITING) {  // the owner below is at work: it goes first
#ifdef JKQ_EMU
        *my_meta = vt_meta_of(klo, JV_WAITING);
#else
        *(volatile unsigned long long *)my_meta = vt_meta_of(klo, JV_WAITING);
#endif
        ml = vt_wait(E, hl, 3u);
        id1 = (uint32_t)(ml >> 32);
      }
      if (id1 < JV_WAITING && id1 != 0 && fabs(JKQ_DSUB(ju2d(ld_r<MODE_GRID>(&tab[hl].vbits)), x)) <= tol) found = id1 - 1;
    }
    if (found == JV_MISS && mu != 0) {
      uint32_t id1 = (uint32_t)(mu >> 32);
      if (id1 == JV_PENDING) {  // the owner above has not looked at me yet, or is about to: wait until it defers or is done
        mu = vt_wait(E, hu, 1u);
        id1 = (uint32_t)(mu >> 32);
      }
      if (id1 < JV_WAITING && id1 != 0 && fabs(JKQ_DSUB(ju2d(ld_r<MODE_GRID>(&tab[hu].vbits)), x)) <= tol) found = id1 - 1;
    }
    if (found == JV_MISS && !vt_room(E, idx)) { jset_error<MODE_GRID>(E, nullptr, JKQ_ERR_VAL_FULL); found = 0; idx = 0xFFFFFFFFu; }
    if (found != JV_MISS) {  // somebody within tol got there first (or the array is full): give the slot up
      tab[h].vbits = JV_INF;
      jfence<MODE_GRID>();
#ifdef JKQ_EMU
      *my_meta = vt_meta_of(klo, JV_DROPPED);
#else
      *(volatile unsigned long long *)my_meta = vt_meta_of(klo, JV_DROPPED);
#endif
      return idx == 0xFFFFFFFFu ? JV_MISS : found;
    }
    // ---- 3. append, publish
    E->vals[idx] = x;
    jstat<MODE_GRID>(E, nullptr, JST_VALS, 1u);
    jfence<MODE_GRID>();  // the real and the value before the id that names them
#ifdef JKQ_EMU
    *my_meta = vt_meta_of(klo, idx + 1);
#else
    *(volatile unsigned long long *)my_meta = vt_meta_of(klo, idx + 1);
#endif
    vt_mark<MODE_GRID>(E, h);
    return idx;
  }
  jset_error<MODE_GRID>(E, nullptr, JKQ_ERR_INTERNAL);
  return JV_MISS;
}

// parallel hash-only insert of an existing real (collector / table reload): slot claimed by CAS on id1
template <int MODE> JD void vt_put(JEng *E, CtaState *S, uint32_t idx, double x) {
  const long long k = vt_bucket(x, jinv_tol2<MODE>(E, S));
  const uint32_t mask = vt_mask2<MODE>(E);
  JVSlot *tab = vt_tab2<MODE>(E, S);
  uint32_t h = vt_slot0(k, mask);
  for (uint32_t probes = 0; jatomic_cas(&tab[h].id1, 0u, idx + 1) != 0; probes++) {
    if (probes > (mask >> 3)) {  // (cannot happen in a table that was just cleared and is twice the size of the array)
      jset_error<MODE>(E, S, JKQ_ERR_VAL_FULL);
      return;
    }
    h = vt_next(h, mask);
  }
  tab[h].vbits = jd2u(x);
  tab[h].keylo = (uint32_t)k;
  vt_mark<MODE>(E, h);
}

// Clookup for one complex value.  JW_MISS = at least one part is not in the table yet; the missing
// parts have been proposed for the next commit.
template <int MODE> JD uint64_t cx_lookup(JEng *E, CtaState *S, uint32_t ring, double re, double im) {
  const uint32_t sr = re < 0.0, si = im < 0.0;
  uint32_t r = real_lookup<MODE>(E, S, ring, fabs(re), true), i = real_lookup<MODE>(E, S, ring, fabs(im), true);
  if (r == JV_MISS || i == JV_MISS) return JW_MISS;
  if (sr && r) r |= 0x80000000u;
  if (si && i) i |= 0x80000000u;
  return ((uint64_t)r << 32) | (uint64_t)i;
}

// Cmul / Cadd / Cdiv on operands that passed the id short-cuts of QMDDcomplexD.cpp:498-777: arithmetic
// on the canonical values in FP64 (no op caches: recomputing is cheaper than a memo probe on a GPU).
#define JOP_MUL 0u
#define JOP_ADD 1u
#define JOP_DIV 2u
static_assert(JKQ_SVALS <= 1024u, "jpack22 holds 10-bit value indices");
static_assert(JKQ_SNODES <= 2048u, "the chain-rebuild memo holds 11-bit node ids");
// weight id <-> 22 bits (value indices below 1024: single-CTA path)
JD uint64_t jpack22(uint64_t w) { return ((w >> 63) << 21) | (((w >> 32) & 0x3FFull) << 11) | (((w >> 31) & 1ull) << 10) | (w & 0x3FFull); }
JD uint64_t junpack22(uint64_t p) { return ((p >> 21) << 63) | (((p >> 11) & 0x3FFull) << 32) | (((p >> 10) & 1ull) << 31) | (p & 0x3FFull); }
// MODE_GRID experiments of round 2, measured WITHOUT gain and therefore off (scripts/grid_ab.py
// times builds side by side): -DJKQ_GRID_ILP=1 stages an item so that its independent loads are in flight together
// (jkq_grid.cuh), and with it -DJKQ_GRID_WARP=1 puts a whole warp on one item while there are more warps than items
// (cx_batch_w: one real part per lane).  random n=18/160: 76.0 ms as built, 78.6 ms staged, 74.8 ms staged + warp;
// random n=22/200: 423 / 537 / 516 ms.
#ifndef JKQ_GRID_ILP
#define JKQ_GRID_ILP 0
#endif
#ifndef JKQ_GRID_WARP
#define JKQ_GRID_WARP 1
#endif
// ------------------------------------------------------------------------------------ batched complex arithmetic
// one Cmul / Cadd / Cdiv with the id short-cuts of QMDDcomplexD.cpp:498-777 taken at once; what needs arithmetic
// stays pending until the batch runs
struct QOp {
  uint64_t a, b, r;
  uint32_t op, pend;
};
JD void q_set(QOp &o, uint64_t r) { o.r = r; o.pend = 0; }
JD void q_mul(QOp &o, uint64_t a, uint64_t b) {
  o.pend = 0;
  if (a == JW_ONE) o.r = b;
  else if (b == JW_ONE) o.r = a;
  else if (a == 0 || b == 0) o.r = 0;
  else if (a == JW_MONE) o.r = jneg(b);
  else if (b == JW_MONE) o.r = jneg(a);
  else { o.op = JOP_MUL; o.a = a; o.b = b; o.pend = 1; }
}
JD void q_add(QOp &o, uint64_t a, uint64_t b) {
  o.pend = 0;
  if (a == 0) o.r = b;
  else if (b == 0) o.r = a;
  else if (a == jneg(b)) o.r = 0;
  else { o.op = JOP_ADD; o.a = a; o.b = b; o.pend = 1; }
}
JD void q_div(QOp &o, uint64_t a, uint64_t b) {
  o.pend = 0;
  if (a == b) o.r = JW_ONE;
  else if (a == 0) o.r = 0;
  else if (b == JW_ONE) o.r = a;
  else if (b == JW_MONE) o.r = jneg(a);
  else if (a == jneg(b)) o.r = JW_MONE;
  else { o.op = JOP_DIV; o.a = a; o.b = b; o.pend = 1; }
}

#define JQ_PARTS 6  // real lookups whose probes are in flight together (three 16-byte slots each)

// Runs the pending operations of q[0..N): false if a value could not be entered (error raised, or -- builds without
// in-place insertion -- proposed for the commit round: the item is run again).
template <int N> JNI bool cx_batch(Ctx &c, QOp *q) {
  uint32_t any = 0;
  JKQ_UNROLL
  for (int i = 0; i < N; i++) any |= q[i].pend;
  if (!any) return true;
  JEng *E = c.E;
  JGP_T(b0);
  const double *vals = E->vals;
  double ar[N], ai[N], br[N], bi[N];
  JKQ_UNROLL
  for (int i = 0; i < N; i++) {
    if (q[i].pend) {  // all operand loads first
      ar[i] = ld_s<MODE_GRID>(vals + ((uint32_t)(q[i].a >> 32) & 0x7FFFFFFFu)); ai[i] = ld_s<MODE_GRID>(vals + ((uint32_t)q[i].a & 0x7FFFFFFFu));
      br[i] = ld_s<MODE_GRID>(vals + ((uint32_t)(q[i].b >> 32) & 0x7FFFFFFFu)); bi[i] = ld_s<MODE_GRID>(vals + ((uint32_t)q[i].b & 0x7FFFFFFFu));
    }
  }
  const double tol = E->tol, inv_tol = E->inv_tol;
  double lx[2 * N];       // parts that need the table (indexed at run time: local memory, by intent)
  uint32_t lslot[2 * N];
  uint32_t res[2 * N];
  uint32_t sgn = 0;
  int np = 0;
  JKQ_UNROLL
  for (int i = 0; i < N; i++) {
    if (!q[i].pend) continue;
    const uint64_t a = q[i].a, b = q[i].b;
    double xr = ar[i], xi = ai[i], yr = br[i], yi = bi[i];
    if (a >> 63) xr = -xr;
    if ((a >> 31) & 1) xi = -xi;
    if (b >> 63) yr = -yr;
    if ((b >> 31) & 1) yi = -yi;
    double re, im;  // (same expressions as cx_arith)
    if (q[i].op == JOP_MUL) {
      re = JKQ_DSUB(JKQ_DMUL(xr, yr), JKQ_DMUL(xi, yi));
      im = JKQ_DADD(JKQ_DMUL(xr, yi), JKQ_DMUL(xi, yr));
    } else if (q[i].op == JOP_ADD) {
      re = JKQ_DADD(xr, yr);
      im = JKQ_DADD(xi, yi);
    } else if ((b & 0x7FFFFFFFull) == 0) {
      re = JKQ_DDIV(xr, yr);
      im = JKQ_DDIV(xi, yr);
    } else {
      const double d = JKQ_DADD(JKQ_DMUL(yr, yr), JKQ_DMUL(yi, yi));
      re = JKQ_DDIV(JKQ_DADD(JKQ_DMUL(xr, yr), JKQ_DMUL(xi, yi)), d);
      im = JKQ_DDIV(JKQ_DSUB(JKQ_DMUL(xi, yr), JKQ_DMUL(xr, yi)), d);
    }
    if (re < 0.0) sgn |= 1u << (2 * i);
    if (im < 0.0) sgn |= 1u << (2 * i + 1);
    JKQ_UNROLL
    for (int p = 0; p < 2; p++) {
      const double x = fabs(p ? im : re);
      if (x <= tol) res[2 * i + p] = 0;
      else if (!(x < JKQ_VAL_MAX)) { jset_error<MODE_GRID>(E, c.sh, JKQ_ERR_VAL_RANGE); res[2 * i + p] = JV_MISS; }
      else { lx[np] = x; lslot[np] = (uint32_t)(2 * i + p); np++; }
    }
  }
  const uint32_t mask = vt_mask2<MODE_GRID>(E);
  const JVSlot *tab = E->vtab;
  JGP_T(b1);
  JGP_ADD(c, 20, b0, b1);
  for (int base = 0; base < np; base += JQ_PARTS) {
    double x[JQ_PARTS];
    long long k[JQ_PARTS];
    JVSlot s[JQ_PARTS][3];
    JKQ_UNROLL
    for (int p = 0; p < JQ_PARTS; p++) {
      if (base + p < np) {
        x[p] = lx[base + p];
        k[p] = vt_bucket(x[p], inv_tol);
        JKQ_UNROLL
        for (int d = 0; d < 3; d++) s[p][d] = ld_s<MODE_GRID>(tab + vt_slot0(k[p] - 1 + d, mask));
      }
    }
    JKQ_UNROLL
    for (int p = 0; p < JQ_PARTS; p++) {
      if (base + p < np) {
        uint32_t r = JV_MISS;
        JKQ_UNROLL
        for (int d = 0; d < 3; d++) {
          if (r == JV_MISS) {
            const uint32_t klo = (uint32_t)(k[p] - 1 + d);
            JVSlot sl = s[p][d];
            if (sl.id1 != 0) {
              uint32_t hh = vt_slot0(k[p] - 1 + d, mask);
              for (uint32_t probes = 0; sl.id1 != 0 && probes <= (mask >> 3); probes++) {
                if (sl.keylo == klo && sl.id1 < JV_WAITING && fabs(JKQ_DSUB(ju2d(sl.vbits), x[p])) <= tol) { r = sl.id1 - 1; break; }
                hh = vt_next(hh, mask);
                sl = ld_s<MODE_GRID>(tab + hh);
              }
            }
          }
        }
        if (r == JV_MISS) {
          JGP_T(i0);
          if (JKQ_INPLACE_GRID) r = real_insert_par<MODE_GRID>(E, c.sh, x[p], k[p]);
          else real_propose<MODE_GRID>(E, c.sh, c.ring, x[p], k[p]);
          JGP_T(i1);
          JGP_ADD(c, 22, i0, i1);
        }
        res[lslot[base + p]] = r;
      }
    }
  }
  JGP_T(b2);
  JGP_ADD(c, 21, b1, b2);
  bool ok = true;
  JKQ_UNROLL
  for (int i = 0; i < N; i++) {
    if (!q[i].pend) continue;
    uint32_t r = res[2 * i], im = res[2 * i + 1];
    if (r == JV_MISS || im == JV_MISS) { q[i].r = JW_MISS; ok = false; continue; }
    if (((sgn >> (2 * i)) & 1u) && r) r |= 0x80000000u;
    if (((sgn >> (2 * i + 1)) & 1u) && im) im |= 0x80000000u;
    q[i].r = ((uint64_t)r << 32) | (uint64_t)im;
    q[i].pend = 0;
  }
  return ok;
}

// The same batch run by ALL 32 lanes of a warp that works on one item together (narrow levels: there are more warps than
// items).  Every lane holds the same q[]; lane p < 2N takes ONE real part of one operation -- operand loads, arithmetic,
// the three bucket probes, the insertion of a missing value -- and the results come back by shuffle, so the dependent
// instruction chain of a batch is that of one part instead of 2N of them (one thread runs ~7 cycles per instruction on
// this path: narrow level phases are bound by the instructions of one item, not by its memory round trips).  Lanes
// that computed the very same value elect one to do the table work.  Called convergently by the whole warp.
#ifndef JKQ_EMU
template <int N> JNI bool cx_batch_w(Ctx &c, QOp *q) {
  static_assert(2 * N <= 32, "one lane per real part");
  uint32_t any = 0;
  JKQ_UNROLL
  for (int i = 0; i < N; i++) any |= q[i].pend;
  if (!any) return true;
  JEng *E = c.E;
  JGP_T(b0);
  const uint32_t lane = threadIdx.x & 31u, oi = lane >> 1, part = lane & 1u;
  const bool act = lane < 2u * N && q[oi < (uint32_t)N ? oi : 0u].pend != 0;
  uint32_t r = 0;
  bool neg = false, look = false;
  double x = 0.0;
  if (act) {
    const uint64_t a = q[oi].a, b = q[oi].b;
    const uint32_t op = q[oi].op;
    const double *vals = E->vals;
    double xr = ld_s<MODE_GRID>(vals + ((uint32_t)(a >> 32) & 0x7FFFFFFFu)), xi = ld_s<MODE_GRID>(vals + ((uint32_t)a & 0x7FFFFFFFu));
    double yr = ld_s<MODE_GRID>(vals + ((uint32_t)(b >> 32) & 0x7FFFFFFFu)), yi = ld_s<MODE_GRID>(vals + ((uint32_t)b & 0x7FFFFFFFu));
    if (a >> 63) xr = -xr;
    if ((a >> 31) & 1) xi = -xi;
    if (b >> 63) yr = -yr;
    if ((b >> 31) & 1) yi = -yi;
    double v;  // (same expressions as cx_arith, one part per lane)
    if (op == JOP_MUL) v = part ? JKQ_DADD(JKQ_DMUL(xr, yi), JKQ_DMUL(xi, yr)) : JKQ_DSUB(JKQ_DMUL(xr, yr), JKQ_DMUL(xi, yi));
    else if (op == JOP_ADD) v = part ? JKQ_DADD(xi, yi) : JKQ_DADD(xr, yr);
    else if ((b & 0x7FFFFFFFull) == 0) v = part ? JKQ_DDIV(xi, yr) : JKQ_DDIV(xr, yr);
    else {
      const double d = JKQ_DADD(JKQ_DMUL(yr, yr), JKQ_DMUL(yi, yi));
      v = part ? JKQ_DDIV(JKQ_DSUB(JKQ_DMUL(xi, yr), JKQ_DMUL(xr, yi)), d) : JKQ_DDIV(JKQ_DADD(JKQ_DMUL(xr, yr), JKQ_DMUL(xi, yi)), d);
    }
    neg = v < 0.0;
    x = fabs(v);
    if (x <= E->tol) r = 0;
    else if (!(x < JKQ_VAL_MAX)) { jset_error<MODE_GRID>(E, c.sh, JKQ_ERR_VAL_RANGE); r = JV_MISS; }
    else look = true;
  }
  JGP_T(b1);
  JGP_ADD(c, 20, b0, b1);
  // one lane per distinct value does the table work
  const unsigned long long xkey = look ? jd2u(x) : 0xFFF8000000000000ull;  // (match.any costs a round per distinct key: the idle lanes share one)
  const unsigned same = __match_any_sync(0xFFFFFFFFu, xkey);
  const int leader = __ffs(same) - 1;
  if (look && (int)lane == leader) {
    const double tol = E->tol;
    const uint32_t mask = vt_mask2<MODE_GRID>(E);
    const JVSlot *tab = E->vtab;
    const long long k = vt_bucket(x, E->inv_tol);
    JVSlot s[3];
    JKQ_UNROLL
    for (int d = 0; d < 3; d++) s[d] = ld_s<MODE_GRID>(tab + vt_slot0(k - 1 + d, mask));
    r = JV_MISS;
    JKQ_UNROLL
    for (int d = 0; d < 3; d++) {
      if (r == JV_MISS) {
        const uint32_t klo = (uint32_t)(k - 1 + d);
        JVSlot sl = s[d];
        if (sl.id1 != 0) {
          uint32_t hh = vt_slot0(k - 1 + d, mask);
          for (uint32_t probes = 0; sl.id1 != 0 && probes <= (mask >> 3); probes++) {
            if (sl.keylo == klo && sl.id1 < JV_WAITING && fabs(JKQ_DSUB(ju2d(sl.vbits), x)) <= tol) { r = sl.id1 - 1; break; }
            hh = vt_next(hh, mask);
            sl = ld_s<MODE_GRID>(tab + hh);
          }
        }
      }
    }
    if (r == JV_MISS) {
      JGP_T(i0);
      if (JKQ_INPLACE_GRID) r = real_insert_par<MODE_GRID>(E, c.sh, x, k);
      else real_propose<MODE_GRID>(E, c.sh, c.ring, x, k);
      JGP_T(i1);
      JGP_ADD(c, 22, i0, i1);
    }
  }
  __syncwarp();
  const uint32_t rl = __shfl_sync(0xFFFFFFFFu, r, leader);
  if (look) r = rl;
  const unsigned negm = __ballot_sync(0xFFFFFFFFu, neg);
  JGP_T(b2);
  JGP_ADD(c, 21, b1, b2);
  bool ok = true;
  JKQ_UNROLL
  for (int i = 0; i < N; i++) {
    uint32_t rr = __shfl_sync(0xFFFFFFFFu, r, 2 * i), im = __shfl_sync(0xFFFFFFFFu, r, 2 * i + 1);
    if (!q[i].pend) continue;
    if (rr == JV_MISS || im == JV_MISS) { q[i].r = JW_MISS; ok = false; continue; }
    if (((negm >> (2 * i)) & 1u) && rr) rr |= 0x80000000u;
    if (((negm >> (2 * i + 1)) & 1u) && im) im |= 0x80000000u;
    q[i].r = ((uint64_t)rr << 32) | (uint64_t)im;
    q[i].pend = 0;
  }
  return ok;
}
#endif
// W: the whole warp works on the item (cx_batch_w); else one thread
template <int N, bool W> JD bool cx_run(Ctx &c, QOp *q) {
#ifndef JKQ_EMU
  if (W) return cx_batch_w<N>(c, q);
#endif
  return cx_batch<N>(c, q);
}

JD double jmag2_g(const double *vals, uint64_t w) {
  const double r = ld_s<MODE_GRID>(vals + ((uint32_t)(w >> 32) & 0x7FFFFFFFu)), i = ld_s<MODE_GRID>(vals + ((uint32_t)w & 0x7FFFFFFFu));
  return JKQ_DADD(JKQ_DMUL(r, r), JKQ_DMUL(i, i));
}

#ifdef JKQ_EMU_MEMOSTAT
#include <unordered_map>
#include <string>
struct JMemoStat { std::unordered_map<std::string, int> seen; unsigned long long calls = 0, repeats = 0, recent_hits = 0; std::string ring[1024]; unsigned pos = 0; };
inline JMemoStat &jmemostat() { static JMemoStat m; return m; }
#endif
template <int MODE> JNI uint64_t cx_arith(JEng *E, CtaState *S, uint32_t ring, uint32_t op, uint64_t a, uint64_t b) {
  if (MODE == MODE_GRID && JKQ_GRID_ILP) {  // operand loads together, then the probes of both parts together
    Ctx c;
    c.E = E; c.sh = S; c.ring = ring; c.gprof = 0;
    QOp q;
    q.op = op; q.a = a; q.b = b; q.r = JW_MISS; q.pend = 1;
    cx_batch<1>(c, &q);
    return q.r;
  }
#ifdef JKQ_EMU_MEMOSTAT
  {
    JMemoStat &m = jmemostat();
    char key[64];
    snprintf(key, sizeof(key), "%u:%llx:%llx", op, (unsigned long long)a, (unsigned long long)b);
    m.calls++;
    if (m.seen[key]++) m.repeats++;
    // direct-mapped 256-entry cache model
    unsigned h = (unsigned)((a * 0x9E3779B97F4A7C15ull ^ b * 0xC2B2AE3D27D4EB4Full ^ op) >> MEMO_SHIFT);
    if (m.ring[h] == key) m.recent_hits++; else m.ring[h] = key;
  }
#endif
  // Operation memo of the single-CTA path (the reference's ctm / cta / ctd caches, QMDDcomplexD.cpp:498-777, as a
  // direct-mapped table in shared memory): 9 of 10 calls repeat an earlier (op, a, b).  Value indices are below
  // 1024 there, so a weight id packs into 22 bits and (op, a, b) into 46; an entry is two words,
  //   word 0 = VALID | key,  word 1 = key[45:4] << 22 | result,
  // and the slot index is hash(key[45:4]) ^ key[3:0]: two entries that agree on key[45:4] never share a slot,
  // so a reader that catches the two words of different writers cannot take them for one entry.
  uint64_t memo_key = 0;
  uint32_t memo_slot = 0;
  if (MODE == MODE_CTA) {
    const uint64_t pa = jpack22(a), pb = jpack22(b);
    memo_key = ((uint64_t)op << 44) | (pa << 22) | pb;
    const uint64_t top = memo_key >> 4;
    memo_slot = (hfin((uint32_t)top * 0x9E3779B1u + (uint32_t)(top >> 32) * 0x85EBCA6Bu) ^ (uint32_t)(memo_key & 15u)) & (JKQ_SMEMO - 1u);
    const unsigned long long w0 = ld_r<MODE>(&S->memo[memo_slot][0]), w1 = ld_r<MODE>(&S->memo[memo_slot][1]);
#ifndef JKQ_NO_MEMO
    if (w0 == (memo_key | (1ull << 63)) && (w1 >> 22) == top) return junpack22(w1 & 0x3FFFFFull);
#endif
  }
  double ar = jval2<MODE>(E, S, (uint32_t)(a >> 32)), ai = jval2<MODE>(E, S, (uint32_t)a);
  double br = jval2<MODE>(E, S, (uint32_t)(b >> 32)), bi = jval2<MODE>(E, S, (uint32_t)b);
  if (a >> 63) ar = -ar;
  if ((a >> 31) & 1) ai = -ai;
  if (b >> 63) br = -br;
  if ((b >> 31) & 1) bi = -bi;
  double re, im;
  if (op == JOP_MUL) {
    re = JKQ_DSUB(JKQ_DMUL(ar, br), JKQ_DMUL(ai, bi));
    im = JKQ_DADD(JKQ_DMUL(ar, bi), JKQ_DMUL(ai, br));
  } else if (op == JOP_ADD) {
    re = JKQ_DADD(ar, br);
    im = JKQ_DADD(ai, bi);
  } else if ((b & 0x7FFFFFFFull) == 0) {  // purely real divisor, :727-736
    re = JKQ_DDIV(ar, br);
    im = JKQ_DDIV(ai, br);
  } else {
    double d = JKQ_DADD(JKQ_DMUL(br, br), JKQ_DMUL(bi, bi));
    re = JKQ_DDIV(JKQ_DADD(JKQ_DMUL(ar, br), JKQ_DMUL(ai, bi)), d);
    im = JKQ_DDIV(JKQ_DSUB(JKQ_DMUL(ai, br), JKQ_DMUL(ar, bi)), d);
  }
  const uint64_t r = cx_lookup<MODE>(E, S, ring, re, im);
  if (MODE == MODE_CTA && r != JW_MISS) {
    S->memo[memo_slot][0] = memo_key | (1ull << 63);
    S->memo[memo_slot][1] = ((memo_key >> 4) << 22) | jpack22(r);
  }
  return r;
}
template <int MODE> JD bool cmul(Ctx &c, uint64_t a, uint64_t b, uint64_t *out) {
  if (a == JW_ONE) { *out = b; return true; }
  if (b == JW_ONE) { *out = a; return true; }
  if (a == 0 || b == 0) { *out = 0; return true; }
  if (a == JW_MONE) { *out = jneg(b); return true; }
  if (b == JW_MONE) { *out = jneg(a); return true; }
  const uint64_t r = cx_arith<MODE>(c.E, c.sh, c.ring, JOP_MUL, a, b);
  *out = r;
  return r != JW_MISS;
}
template <int MODE> JD bool cadd(Ctx &c, uint64_t a, uint64_t b, uint64_t *out) {
  if (a == 0) { *out = b; return true; }
  if (b == 0) { *out = a; return true; }
  if (a == jneg(b)) { *out = 0; return true; }
  const uint64_t r = cx_arith<MODE>(c.E, c.sh, c.ring, JOP_ADD, a, b);
  *out = r;
  return r != JW_MISS;
}
template <int MODE> JD bool cdiv(Ctx &c, uint64_t a, uint64_t b, uint64_t *out) {
  if (a == b) { *out = JW_ONE; return true; }
  if (a == 0) { *out = 0; return true; }
  if (b == JW_ONE) { *out = a; return true; }
  if (b == JW_MONE) { *out = jneg(a); return true; }
  if (a == jneg(b)) { *out = JW_MONE; return true; }
  const uint64_t r = cx_arith<MODE>(c.E, c.sh, c.ring, JOP_DIV, a, b);
  *out = r;
  return r != JW_MISS;
}

// serial canonicalisation (one thread, may insert): gate set-up and root update
template <int MODE> JNI uint64_t cx_make_serial(JEng *E, CtaState *S, double re, double im) {
  const uint32_t sr = re < 0.0, si = im < 0.0;
  re = fabs(re); im = fabs(im);
  uint32_t r = real_lookup<MODE>(E, S, 0, re, false), i;
  if (r == JV_MISS) r = real_insert_direct<MODE>(E, S, re);
  i = real_lookup<MODE>(E, S, 0, im, false);
  if (i == JV_MISS) i = real_insert_direct<MODE>(E, S, im);
  if (sr && r) r |= 0x80000000u;
  if (si && i) i |= 0x80000000u;
  return ((uint64_t)r << 32) | (uint64_t)i;
}

// ------------------------------------------------------------------------------------ node access
#define JN_TOMB 0xFFFFFFFFu  // c1 of an arena slot that was allocated but never published
JD uint32_t nd_flag(const JNode &n) { return n.c0 >> 31; }
JD uint32_t nd_c0(const JNode &n) { return n.c0 & 0x7FFFFFFFu; }
JD uint64_t nd_w0(const JNode &n) { return nd_flag(n) ? n.w : JW_ONE; }
JD uint64_t nd_w1(const JNode &n) { return nd_flag(n) ? JW_ONE : n.w; }

template <int MODE> JD JNode node_load(const Ctx &c, uint32_t id) {
  return MODE == MODE_CTA ? c.sh->nodes[id] : ld_s<MODE>(c.E->nodes + id);
}

// ------------------------------------------------------------------------------------ unique table
// QMDDutLookup (QMDDpackage.cpp:693-763) as a lock-free open-addressing table of node ids.
// Protocol: write the record, fence, publish the id with atomicCAS.  A reader that observes an id
// in a slot therefore always finds a complete record behind it.  Returns 0 when a pool is exhausted
// (sticky error / want_grid raised).
JD uint32_t ut_hash3(uint32_t c0, uint32_t c1, uint64_t w) { return hfin(c0 * 0x9E3779B1u + c1 * 0x85EBCA77u + hw64(w)); }
JD uint32_t ut_hash(const JNode &n) { return ut_hash3(n.c0, n.c1, n.w); }

template <int MODE> JD uint32_t ut_find_or_insert(JEng *E, CtaState *S, uint32_t c0, uint32_t c1, uint64_t w) {
  JNode rec;
  rec.c0 = c0; rec.c1 = c1; rec.w = w;
  if (MODE == MODE_CTA) {
    uint32_t h = ut_hash3(c0, c1, w) & (JKQ_SUT_SLOTS - 1), mine = 0;
    for (uint32_t probes = 0; probes < JKQ_SUT_SLOTS; probes++) {
      uint32_t s = ld_r<MODE>(&S->utab[h]);
      if (s == 0) {
        if (mine == 0) {
          mine = jatomic_add(&S->node_count, 1u);
          if (mine >= S->node_limit) { jwant_grid(S); return 0; }  // gate abandoned, redone on the grid path
          S->nodes[mine] = rec;  // (copied to HBM by cta_flush)
          jfence<MODE>();
        }
        s = jatomic_cas(&S->utab[h], 0u, mine);
        if (s == 0) { jstat<MODE>(E, S, JST_NODES, 1u); return mine; }
      }
      JNode o = ld_r<MODE>(&S->nodes[s]);
      if (o.c0 == c0 && o.c1 == c1 && o.w == w) {
        if (mine) S->nodes[mine].c1 = JN_TOMB;  // allocated above, lost the race: never referenced, must never be hashed
        return s;
      }
      h = (h + 1) & (JKQ_SUT_SLOTS - 1);
    }
    jwant_grid(S);
    return 0;
  }
  // HBM table: 64-bit slots, generation << 32 | node id.  An entry of an older generation reads as empty (the
  // collector starts a new generation instead of clearing the table) and is claimed by CAS from whatever it holds.
  uint32_t h = ut_hash3(c0, c1, w) & E->utab_mask, mine = 0;
  const unsigned long long gen = (unsigned long long)ld_s<MODE>(&E->ut_gen) << 32;
  for (uint32_t probes = 0; probes <= E->utab_mask; probes++) {
    unsigned long long sv = ld_r<MODE>(E->utab + h);
    uint32_t s = (sv >> 32) == (gen >> 32) ? (uint32_t)sv : 0u;
    while (s == 0) {
      if (mine == 0) {
        mine = jagg_inc(&E->node_count);
        if (mine >= E->node_cap) { jset_error<MODE>(E, S, JKQ_ERR_NODE_FULL); return 0; }
        E->nodes[mine] = rec;
        jfence<MODE>();
      }
      const unsigned long long old = jatomic_cas(E->utab + h, sv, gen | mine);
      if (old == sv) { jstat<MODE>(E, S, JST_NODES, 1u); return mine; }
      sv = old;
      s = (sv >> 32) == (gen >> 32) ? (uint32_t)sv : 0u;
    }
    JNode o = ld_r<MODE>(E->nodes + s);
    if (o.c0 == c0 && o.c1 == c1 && o.w == w) {
      if (mine) E->nodes[mine].c1 = JN_TOMB;
      return s;
    }
    h = (h + 1) & E->utab_mask;
  }
  jset_error<MODE>(E, S, JKQ_ERR_UTAB_FULL);
  return 0;
}

// ------------------------------------------------------------------------------------ frontier de-dup
// The compute tables of the reference (CTlookup/CTinsert, QMDDpackage.cpp:986-1072) exist so that a
// (gate,node) or (a,b,ratio) pair is expanded once.  Here the pairs of one gate are the work items of
// the frontier.  M items (gate x node) are keyed by the node alone: a direct-indexed array, no hashing.
// F/S items go through a hash table.  Every slot carries the gate's generation, so nothing is ever
// cleared between gates.  Both return JI_FAIL when a pool is exhausted (sticky error / want_grid).
#define JI_FAIL 0xFFFFFFFFu
JD void item_init(JItem *it, uint32_t a, uint32_t b, uint64_t wa, uint64_t wb, uint64_t wc, uint32_t tag) {
  it->a = a; it->b = b; it->tag = tag; it->flags = JS_NEW;
  it->wa = wa; it->wb = wb; it->wc = wc;
}
JD void item_init_m(JItem *it, uint32_t node) { it->a = node; it->b = 0; it->tag = JK_M; it->flags = JS_NEW; }

template <int MODE> JD uint32_t item_alloc(JEng *E, CtaState *S) {
  if (MODE == MODE_CTA) {
    uint32_t mine = jatomic_add(&S->item_count, 1u);
    if (mine >= JKQ_SITEMS) { jwant_grid(S); return JI_FAIL; }
    return mine;
  }
  uint32_t mine = jagg_inc(&E->item_count);
  if (mine >= E->item_cap) { jset_error<MODE>(E, S, JKQ_ERR_ITEM_FULL); return JI_FAIL; }
  return mine;
}

template <int MODE> JD uint32_t dt_find_or_insert_m(JEng *E, CtaState *S, JItem *items, uint32_t node) {
  uint32_t mine = JI_FAIL;
  if (MODE == MODE_CTA) {
    const uint32_t gen = S->gen << 16;
    uint32_t s = ld_r<MODE>(&S->mdir[node]);
    for (;;) {
      if ((s & 0xFFFF0000u) == gen) {
        if (mine != JI_FAIL) items[mine].flags = JS_DEAD;
        return (s & 0xFFFFu) - 1;
      }
      if (mine == JI_FAIL) {
        mine = item_alloc<MODE>(E, S);
        if (mine == JI_FAIL) return JI_FAIL;
        item_init_m(items + mine, node);
      }
      uint32_t old = jatomic_cas(&S->mdir[node], s, gen | (mine + 1));
      if (old == s) return mine;
      s = old;
    }
  }
  const unsigned long long gen = (unsigned long long)ld_s<MODE>(&E->gen) << 32;
  unsigned long long s = ld_r<MODE>(E->mdir + node);
  for (;;) {
    if ((s >> 32) == (gen >> 32)) {
      if (mine != JI_FAIL) items[mine].flags = JS_DEAD;
      return (uint32_t)s - 1;
    }
    if (mine == JI_FAIL) {
      mine = item_alloc<MODE>(E, S);
      if (mine == JI_FAIL) return JI_FAIL;
      item_init_m(items + mine, node);
    }
    unsigned long long old = jatomic_cas(E->mdir + node, s, gen | (unsigned long long)(mine + 1));
    if (old == s) return mine;
    s = old;
  }
}

JD uint32_t dt_hash(uint32_t a, uint32_t b, uint64_t wa, uint64_t wb, uint64_t wc, uint32_t tag) {
  return hfin(a * 0x9E3779B1u + b * 0x85EBCA77u + tag * 0x165667B1u + hw64(wa) * 3u + hw64(wb) * 5u + hw64(wc) * 7u);
}
template <int MODE> JD bool item_key_eq(const JItem *o, uint32_t a, uint32_t b, uint64_t wa, uint64_t wb, uint64_t wc, uint32_t tag) {
  return ld_r<MODE>(&o->a) == a && ld_r<MODE>(&o->b) == b && ld_r<MODE>(&o->tag) == tag && ld_r<MODE>(&o->wa) == wa &&
         ld_r<MODE>(&o->wb) == wb && ld_r<MODE>(&o->wc) == wc;
}

template <int MODE>
JD uint32_t dt_find_or_insert(JEng *E, CtaState *S, JItem *items, uint32_t a, uint32_t b, uint64_t wa, uint64_t wb, uint64_t wc, uint32_t tag,
                              bool *created = nullptr) {
  if (created) *created = false;
  if ((tag & 15u) == JK_M) return dt_find_or_insert_m<MODE>(E, S, items, a);
  uint32_t mine = JI_FAIL;
  if (MODE == MODE_CTA) {
    const uint32_t gen = S->gen << 16;
    uint32_t h = dt_hash(a, b, wa, wb, wc, tag) & (JKQ_SDT_SLOTS - 1);
    for (uint32_t probes = 0; probes < JKQ_SDT_SLOTS; probes++) {
      uint32_t s = ld_r<MODE>(&S->dtab[h]);
      while ((s & 0xFFFF0000u) != gen) {  // stale or empty: try to claim
        if (mine == JI_FAIL) {
          mine = item_alloc<MODE>(E, S);
          if (mine == JI_FAIL) return JI_FAIL;
          item_init(items + mine, a, b, wa, wb, wc, tag);
          jfence<MODE>();
        }
        uint32_t old = jatomic_cas(&S->dtab[h], s, gen | (mine + 1));
        if (old == s) { if (created) *created = true; return mine; }
        s = old;
      }
      const uint32_t j = (s & 0xFFFFu) - 1;
      if (item_key_eq<MODE>(items + j, a, b, wa, wb, wc, tag)) {
        if (mine != JI_FAIL) items[mine].flags = JS_DEAD;
        return j;
      }
      h = (h + 1) & (JKQ_SDT_SLOTS - 1);
    }
    jwant_grid(S);
    return JI_FAIL;
  }
  uint32_t h = dt_hash(a, b, wa, wb, wc, tag) & E->dtab_mask;
  const unsigned long long gen = (unsigned long long)ld_s<MODE>(&E->gen) << 32;
  for (uint32_t probes = 0; probes <= E->dtab_mask; probes++) {
    unsigned long long s = ld_r<MODE>(E->dtab + h);
    while ((s >> 32) != (gen >> 32)) {  // stale or empty: try to claim
      if (mine == JI_FAIL) {
        mine = item_alloc<MODE>(E, S);
        if (mine == JI_FAIL) return JI_FAIL;
        item_init(items + mine, a, b, wa, wb, wc, tag);
        jfence<MODE>();
      }
      unsigned long long old = jatomic_cas(E->dtab + h, s, gen | (unsigned long long)(mine + 1));
      if (old == s) { if (created) *created = true; return mine; }
      s = old;
    }
    const uint32_t j = (uint32_t)s - 1;
    if (item_key_eq<MODE>(items + j, a, b, wa, wb, wc, tag)) {
      if (mine != JI_FAIL) items[mine].flags = JS_DEAD;
      return j;
    }
    h = (h + 1) & E->dtab_mask;
  }
  jset_error<MODE>(E, S, JKQ_ERR_DTAB_FULL);
  return JI_FAIL;
}

// item field access: shared memory in MODE_CTA, ld.global.cg in MODE_GRID
template <int MODE, class T> JD T ld_i(const T *p) { return ld_r<MODE>(p); }

// ------------------------------------------------------------------------------------ child slots
#define JFLAG_CHAIN 0x100u  // ch[0] = the one child item several levels down, ch[1] = skipped levels | scratch offset << 8
// level of an F / S item: bits 8..15 of its tag hold level + 1 (0: not recorded -- M items)
JD uint32_t tag_level(uint32_t tag, int level) { return (tag & 0xFFu) | ((uint32_t)(level + 1) << 8); }
struct Slot {
  uint32_t kind, target;
  uint64_t w;
  // REF slots are materialised (de-dup insert) only after ALL arithmetic of the item succeeded
  uint32_t ia, ib, itag;
  uint64_t iwa, iwb, iwc;
};
JD void slot_zero(Slot &s) { s.kind = JC_ZERO; s.target = 0; s.w = 0; }
JD void slot_pass(Slot &s, uint32_t node, uint64_t w) {
  if (w == 0) { slot_zero(s); return; }
  s.kind = JC_PASS; s.target = node; s.w = w;
}
JD void slot_ref(Slot &s, uint64_t factor, uint32_t a, uint32_t b, uint64_t wa, uint64_t wb, uint64_t wc, uint32_t tag) {
  s.kind = JC_REF; s.w = factor; s.ia = a; s.ib = b; s.iwa = wa; s.iwb = wb; s.iwc = wc; s.itag = tag;
}

// R = alpha*A + beta*B on child level m, no control at or below m: the operand pair of QMDDadd
// (QMDDpackage.cpp:1319-1414) with its zero / same-node short-cuts (:1337-1349) and the
// (1, y.w/x.w) normalisation of the memo key (:1354-1359).
template <int MODE>
JD bool combine_sum(Ctx &c, int m, uint64_t al, uint32_t A, uint64_t be, uint32_t B, Slot &s) {
  if (m < 0 || A == B) {  // terminals (both node 0) or the same node: add the weights
    uint64_t w;
    if (al == 0) w = be;
    else if (be == 0) w = al;
    else if (!cadd<MODE>(c, al, be, &w)) return false;
    slot_pass(s, (al == 0) ? B : A, w);
    return true;
  }
  if (al == 0) { slot_pass(s, B, be); return true; }
  if (be == 0) { slot_pass(s, A, al); return true; }
  // Operand order of the memo key.  The reference orders by pointer value (QMDDpackage.cpp:1331-1335) and
  // can afford any ratio y.w/x.w in 200-bit reals; here the operand with the larger weight comes first, so the
  // ratio has magnitude <= 1 (a deep random circuit otherwise produces ratios beyond the value table's
  // range); equal magnitudes fall back to the node order.
  bool swap = A > B;
  if (al != be && al != jneg(be)) {
    const double ma = jmag2<MODE>(c, al), mb = jmag2<MODE>(c, be);
    if (ma != mb) swap = mb > ma;
  }
  if (swap) { uint32_t t = A; A = B; B = t; uint64_t tw = al; al = be; be = tw; }
  uint64_t r;
  if (!cdiv<MODE>(c, be, al, &r)) return false;
  slot_ref(s, al, A, B, JW_ONE, r, 0, tag_level(JK_S, m));
  return true;
}

// R = P*(alpha*A + beta*B) + (I-P)*(gamma*C), C = sel ? B : A, on child level m, where P projects on
// "all controls at levels <= m satisfied".  This fuses what the reference does as four
// QMDDmultiply2 calls into the below-target part of the gate DD (QMDDmvlgate, QMDDpackage.cpp:
// 1763-1800) followed by two QMDDadd calls (QMDDpackage.cpp:1488-1509).
template <int MODE>
JD bool combine(Ctx &c, int low, int m, uint64_t al, uint32_t A, uint64_t be, uint32_t B, uint64_t ga, uint32_t sel, Slot &s) {
  if (m < low) return combine_sum<MODE>(c, m, al, A, be, B, s);
  if (al == 0 && be == 0 && ga == 0) { slot_zero(s); return true; }
  if (sel == 0 && be == 0 && al == ga) { slot_pass(s, A, ga); return true; }  // gate row acts as identity
  if (sel == 1 && al == 0 && be == ga) { slot_pass(s, B, ga); return true; }
  // common factor: the largest of the three weights (first one on ties), so the key weights have magnitude <= 1
  uint64_t f = al ? al : (be ? be : ga), wa, wb, wc;
  {
    double mf = jmag2<MODE>(c, f);
    if (be != 0 && be != f) { const double mb = jmag2<MODE>(c, be); if (mb > mf) { f = be; mf = mb; } }
    if (ga != 0 && ga != f) { const double mg = jmag2<MODE>(c, ga); if (mg > mf) { f = ga; mf = mg; } }
  }
  bool o1 = cdiv<MODE>(c, al, f, &wa), o2 = cdiv<MODE>(c, be, f, &wb), o3 = cdiv<MODE>(c, ga, f, &wc);
  if (!(o1 && o2 && o3)) return false;
  uint32_t An = (wa != 0 || (sel == 0 && wc != 0)) ? A : 0;
  uint32_t Bn = (wb != 0 || (sel == 1 && wc != 0)) ? B : 0;
  slot_ref(s, f, An, Bn, wa, wb, wc, tag_level(JK_F | (sel << 4), m));
  return true;
}

JD int role_satisfied(uint32_t role, int i) {
  return role == JR_IDLE || (role == JR_CPOS && i == 1) || (role == JR_CNEG && i == 0);
}

template <int MODE> JD uint64_t gate_u(const Ctx &c, int k) { return MODE == MODE_CTA ? c.sh->cur_u[k] : ld_s<MODE>(&c.E->cur_u[k]); }

// one child slot of an item of level L.  kind M: na = the state node.  kind F/S: na/nb = operand nodes.
template <int MODE>
JD bool expand_child(Ctx &c, const JGate &g, int L, int i, uint32_t kind, uint32_t sel, uint32_t tag, uint32_t role,
                     uint32_t ia, uint32_t ib, uint64_t iwa, uint64_t iwb, uint64_t iwc, const JNode &na, const JNode &nb, Slot &s) {
  if (kind == JK_M) {
    const uint32_t ch = i ? na.c1 : nd_c0(na);
    const uint64_t w = i ? nd_w1(na) : nd_w0(na);
    if (L > g.target) {
      if (w == 0) slot_zero(s);
      else if (role_satisfied(role, i)) slot_ref(s, w, ch, 0, 0, 0, 0, JK_M);
      else slot_pass(s, ch, w);
      return true;
    }
    // target level: row i of the 2x2 matrix, QMDDpackage.cpp:1488-1509
    uint64_t al, be;
    bool o1 = cmul<MODE>(c, gate_u<MODE>(c, 2 * i), nd_w0(na), &al), o2 = cmul<MODE>(c, gate_u<MODE>(c, 2 * i + 1), nd_w1(na), &be);
    if (!(o1 && o2)) return false;
    return combine<MODE>(c, g.low, L - 1, al, nd_c0(na), be, na.c1, w, (uint32_t)i, s);
  }
  const uint32_t ac = ia ? (i ? na.c1 : nd_c0(na)) : 0, bc = ib ? (i ? nb.c1 : nd_c0(nb)) : 0;
  const uint64_t aw = ia ? (i ? nd_w1(na) : nd_w0(na)) : 0, bw = ib ? (i ? nd_w1(nb) : nd_w0(nb)) : 0;
  if (kind == JK_S) {
    uint64_t be;
    if (!cmul<MODE>(c, iwb, bw, &be)) return false;
    return combine_sum<MODE>(c, L - 1, aw, ac, be, bc, s);
  }
  if (ia == ib) {
    // Both operands are the SAME node: (alpha : beta : gamma) of the child is aw * (wa : wb : wc), i.e.
    // the already normalised key, so the child item needs no arithmetic at all (the exact algebra the
    // reference reaches through its Cmul/Cdiv memo tables, QMDDcomplex.h:93).
    if (aw == 0) { slot_zero(s); return true; }
    if (!role_satisfied(role, i)) {
      uint64_t ga;
      if (!cmul<MODE>(c, iwc, aw, &ga)) return false;
      slot_pass(s, ac, ga);
    } else if (L - 1 >= g.low) {
      slot_ref(s, aw, ac, ac, iwa, iwb, iwc, tag_level(tag, L - 1));
    } else {
      uint64_t sum, w;
      if (!cadd<MODE>(c, iwa, iwb, &sum)) return false;
      if (!cmul<MODE>(c, aw, sum, &w)) return false;
      slot_pass(s, ac, w);
    }
    return true;
  }
  uint64_t ga;
  bool o0 = cmul<MODE>(c, iwc, sel ? bw : aw, &ga);
  if (!role_satisfied(role, i)) {
    if (!o0) return false;
    slot_pass(s, sel ? bc : ac, ga);
    return true;
  }
  uint64_t al, be;
  bool o1 = cmul<MODE>(c, iwa, aw, &al), o2 = cmul<MODE>(c, iwb, bw, &be);
  if (!(o0 && o1 && o2)) return false;
  return combine<MODE>(c, g.low, L - 1, al, ac, be, bc, ga, sel, s);
}

// ------------------------------------------------------------------------------------ expand
// Top-down step for one work item of level L: compute its two child slots.  Side effects (de-dup
// inserts) only after all arithmetic succeeded, so a retry simply runs the function again.
JD uint32_t item_flags(uint32_t state, uint32_t k0, uint32_t k1, uint32_t hint) { return state | (k0 << 2) | (k1 << 4) | (hint << 6); }

// all non-zero edges of the node lead to ONE child?  (two edges to the same node, or one edge and the zero edge)
JD bool chain_node(const JNode &n, uint32_t *child) {
  if (n.w != 0) { *child = n.c1; return nd_c0(n) == n.c1; }
  *child = nd_flag(n) ? n.c1 : nd_c0(n);
  return true;
}

#if defined(JKQ_PROFILE2) && !defined(JKQ_EMU)
__device__ unsigned long long g_fine[32];
#define JFINE_T(v) long long v = clock64()
#define JFINE_ADD(c, slot, a, b) do { if (MODE == MODE_CTA && (c).tid == 0) g_fine[slot] += (unsigned long long)((b) - (a)); } while (0)
#else
#define JFINE_T(v) do {} while (0)
#define JFINE_ADD(c, slot, a, b) do {} while (0)
#endif
// Chain compression of a same-node fused item whose node `na` has all its non-zero edges on one child `chain_c`
// (see the comment inside).  Returns 0 when the plain path has to be taken, else the item's new flags | 1 << 31.
template <int MODE>
JD uint32_t chain_try(Ctx &c, const JGate &g, JItem *it, uint32_t tag, uint64_t iwa, uint64_t iwb, uint64_t iwc, const JNode &na, int L, uint32_t chain_c) {
  // Chain compression.  Both operands are the same node, all its non-zero edges go to the same child c
  // (a basis-state node with one zero edge, or a product-state node with two edges to one child), and the
  // level carries no control: the item's only child is the same key on c, and the result is the old node
  // over that child's result with the top weight passed through.  For the states of QFT / GHZ-like
  // circuits that repeats level after level down to the next control: walk the chain of such
  // nodes once, keep their (norm flag, weight) in scratch, and refer directly to the item at its end --
  // reduce_item rebuilds the chain in one go.  The skipped levels cost a node load and a unique-table
  // insert each instead of two phases.
  CtaState *S = c.sh;
  JFINE_T(w0);
  uint32_t cur = chain_c, k = 0;
  int Lc = L - 1;
  while (Lc > g.low && g.role[Lc] == JR_IDLE && k < 60) {
    const JNode nd = node_load<MODE>(c, cur);
    uint32_t nxt;
    if (!chain_node(nd, &nxt)) break;
    cur = nxt; k++; Lc--;
  }
  if (k >= 1) {
    const uint32_t off = jatomic_add(&S->chain_n, k + 1u);
    const uint32_t dn = off + k + 1u <= JKQ_SCHAIN ? jatomic_add(&S->defer_n[Lc], 1u) : JKQ_SDEFER;
    if (dn < JKQ_SDEFER) {
      unsigned long long mask = 0;
      uint32_t cur2 = chain_c;
      for (uint32_t j = 0; j < k; j++) {
        const JNode nd = node_load<MODE>(c, cur2);
        S->chain[off + 1 + j] = nd.w;
        mask |= (unsigned long long)nd_flag(nd) << j;
        chain_node(nd, &cur2);
      }
      S->chain[off] = mask;
      bool created = false;
      const uint32_t child = dt_find_or_insert<MODE>(c.E, S, c.items, cur, cur, iwa, iwb, iwc, tag_level(tag, Lc), &created);
      S->defer[Lc][dn] = (child != JI_FAIL && created) ? (uint16_t)child : (uint16_t)0xFFFFu;
      if (child == JI_FAIL) return JS_DONE | 0x80000000u;  // want_grid flagged; the gate is abandoned
#ifndef JKQ_EMU
      atomicMin(&S->min_def, Lc);
#else
      if (Lc < S->min_def) S->min_def = Lc;
#endif
      it->ch[0] = child; it->ch[1] = k | (off << 8); it->chw[0] = nd_w0(na); it->chw[1] = nd_w1(na);
      JFINE_T(w1);
      JFINE_ADD(c, 2, w0, w1);
      JFINE_ADD(c, 3, 0, k);
      JFINE_ADD(c, 10, 0, 1);
      return item_flags(JS_DONE, JC_REF, JC_REF, 1u + nd_flag(na)) | JFLAG_CHAIN | 0x80000000u;
    }
    // (scratch or waiting list full: the plain path below; readers clamp defer_n to JKQ_SDEFER)
  }
  return 0;
}

template <int MODE> JD uint32_t expand_item_impl(Ctx &c, const JGate &g, uint32_t idx, int L);
// returns the new flags word of the item (state JS_DONE or JS_RETRY)
template <int MODE> JD uint32_t expand_item(Ctx &c, const JGate &g, uint32_t idx, int L) {
  JFINE_T(t0);
  const uint32_t r = expand_item_impl<MODE>(c, g, idx, L);
  JFINE_T(t1);
  JFINE_ADD(c, 0, t0, t1);
  JFINE_ADD(c, 1, 0, 1);
  return r;
}
template <int MODE> JD uint32_t expand_item_impl(Ctx &c, const JGate &g, uint32_t idx, int L) {
  JItem *it = c.items + idx;
  const uint32_t tag = ld_i<MODE>(&it->tag), ia = ld_i<MODE>(&it->a);
  const uint32_t kind = tag & 15u, sel = (tag >> 4) & 1u;
  const uint32_t role = g.role[L];
  if (kind == JK_M && L > g.target) {
    // Fast path: a level above the target only forwards the node's two edges (QMDDmultiply2 with a
    // block-diagonal gate node, QMDDpackage.cpp:1488-1509) -- no arithmetic.
    const JNode p = node_load<MODE>(c, ia);
    const uint32_t ch0 = nd_c0(p), ch1 = p.c1;
    const uint64_t w0 = nd_w0(p), w1 = nd_w1(p);
    uint32_t t0 = ch0, t1 = ch1, k0, k1;
    bool ok = true;
    if (w0 == 0) { k0 = JC_ZERO; t0 = 0; }
    else if (role_satisfied(role, 0)) { k0 = JC_REF; t0 = dt_find_or_insert_m<MODE>(c.E, c.sh, c.items, ch0); ok = t0 != JI_FAIL; }
    else k0 = JC_PASS;
    if (w1 == 0) { k1 = JC_ZERO; t1 = 0; }
    else if (role_satisfied(role, 1)) {
      k1 = JC_REF;
      if (k0 == JC_REF && ch1 == ch0) t1 = t0;
      else if (ok) { t1 = dt_find_or_insert_m<MODE>(c.E, c.sh, c.items, ch1); ok = t1 != JI_FAIL; }
    } else k1 = JC_PASS;
    if (!ok) return JS_DONE;  // error / want_grid flagged; the gate is abandoned
    it->ch[0] = t0; it->ch[1] = t1; it->chw[0] = w0; it->chw[1] = w1;
    // (chw0, chw1) are the weights of a normalised node: hint = 1 + index of the ONE
    return item_flags(JS_DONE, k0, k1, 1u + nd_flag(p));
  }
  const uint32_t ib = ld_i<MODE>(&it->b);
  uint64_t iwa = 0, iwb = 0, iwc = 0;
  JNode na, nb;
  na.c0 = na.c1 = 0; na.w = 0; nb = na;
  if (ia) na = node_load<MODE>(c, ia);
  if (kind != JK_M) {
    iwa = ld_i<MODE>(&it->wa); iwb = ld_i<MODE>(&it->wb); iwc = ld_i<MODE>(&it->wc);
    if (ib) nb = (ib == ia) ? na : node_load<MODE>(c, ib);
  }
  uint32_t chain_c = 0;
  if (MODE == MODE_CTA && kind == JK_F && ia == ib && ia != 0 && role == JR_IDLE && L - 1 >= g.low && chain_node(na, &chain_c)) {
    const uint32_t cr = chain_try<MODE>(c, g, it, tag, iwa, iwb, iwc, na, L, chain_c);
    if (cr) return cr & 0x7FFFFFFFu;
  }
  Slot s0, s1;
  JFINE_T(x0);
  bool e0 = expand_child<MODE>(c, g, L, 0, kind, sel, tag, role, ia, ib, iwa, iwb, iwc, na, nb, s0);
  bool e1 = expand_child<MODE>(c, g, L, 1, kind, sel, tag, role, ia, ib, iwa, iwb, iwc, na, nb, s1);
  JFINE_T(x1);
  JFINE_ADD(c, 4, x0, x1);
  JFINE_ADD(c, 11, 0, 1);
  if (kind == JK_M) { JFINE_ADD(c, 12, x0, x1); JFINE_ADD(c, 13, 0, 1); }
  else if (kind == JK_F && ia == ib) { JFINE_ADD(c, 14, x0, x1); JFINE_ADD(c, 15, 0, 1); }
  if (!(e0 && e1)) return JS_RETRY;  // both children were tried: all their missing values are proposed
  bool ok = true;
  if (s0.kind == JC_REF) { s0.target = dt_find_or_insert<MODE>(c.E, c.sh, c.items, s0.ia, s0.ib, s0.iwa, s0.iwb, s0.iwc, s0.itag); ok = s0.target != JI_FAIL; }
  if (ok && s1.kind == JC_REF) { s1.target = dt_find_or_insert<MODE>(c.E, c.sh, c.items, s1.ia, s1.ib, s1.iwa, s1.iwb, s1.iwc, s1.itag); ok = s1.target != JI_FAIL; }
  if (!ok) return JS_DONE;  // error / want_grid already flagged; the gate is abandoned
  JFINE_T(x2);
  JFINE_ADD(c, 5, x1, x2);
  it->ch[0] = s0.target; it->ch[1] = s1.target; it->chw[0] = s0.w; it->chw[1] = s1.w;
  // same-node fused item whose two children carry the node's own weights: normalised pair again
  uint32_t hint = 0;
  if (kind == JK_F && ia == ib && s0.w == nd_w0(na) && s1.w == nd_w1(na) && s0.kind != JC_ZERO && s1.kind != JC_ZERO) hint = 1u + nd_flag(na);
  return item_flags(JS_DONE, s0.kind, s1.kind, hint);
}

#ifndef JKQ_EMU
// expand_item for a PAIR of lanes: lane `child` (0 / 1) of the pair computes child slot `child` of the item -- the
// two slots are independent (own arithmetic, own de-dup insert), which halves the dependent chain of a level that
// holds at most 16 items (the usual case on the single-warp path).  `pm` = mask of the two lanes; both lanes take
// every branch below together (the branches depend on the item only) and return the same flags word.
template <int MODE> JD uint32_t expand_item_pair(Ctx &c, const JGate &g, uint32_t idx, int L, uint32_t child, uint32_t pm) {
  JItem *it = c.items + idx;
  const uint32_t tag = ld_i<MODE>(&it->tag), ia = ld_i<MODE>(&it->a);
  const uint32_t kind = tag & 15u, sel = (tag >> 4) & 1u;
  const uint32_t role = g.role[L];
  if (kind == JK_M && L > g.target) {  // a level above the target forwards the node's two edges: one edge per lane
    const JNode p = node_load<MODE>(c, ia);
    const uint32_t ch = child ? p.c1 : nd_c0(p);
    const uint64_t w = child ? nd_w1(p) : nd_w0(p);
    uint32_t k, t = ch;
    bool ok = true;
    if (w == 0) { k = JC_ZERO; t = 0; }
    else if (role_satisfied(role, (int)child)) { k = JC_REF; t = dt_find_or_insert_m<MODE>(c.E, c.sh, c.items, ch); ok = t != JI_FAIL; }
    else k = JC_PASS;
    const uint32_t ko = __shfl_xor_sync(pm, k, 1);
    const bool oko = __shfl_xor_sync(pm, (uint32_t)ok, 1) != 0;
    if (!(ok && oko)) return JS_DONE;  // error / want_grid flagged; the gate is abandoned
    it->ch[child] = t; it->chw[child] = w;
    return item_flags(JS_DONE, child ? ko : k, child ? k : ko, 1u + nd_flag(p));
  }
  const uint32_t ib = ld_i<MODE>(&it->b);
  uint64_t iwa = 0, iwb = 0, iwc = 0;
  JNode na, nb;
  na.c0 = na.c1 = 0; na.w = 0; nb = na;
  if (ia) na = node_load<MODE>(c, ia);
  if (kind != JK_M) {
    iwa = ld_i<MODE>(&it->wa); iwb = ld_i<MODE>(&it->wb); iwc = ld_i<MODE>(&it->wc);
    if (ib) nb = (ib == ia) ? na : node_load<MODE>(c, ib);
  }
  uint32_t chain_c = 0;
  if (MODE == MODE_CTA && kind == JK_F && ia == ib && ia != 0 && role == JR_IDLE && L - 1 >= g.low && chain_node(na, &chain_c)) {
    uint32_t cr = 0;  // the chain walk has nothing to split: lane 0 of the pair does it
    if (child == 0) cr = chain_try<MODE>(c, g, it, tag, iwa, iwb, iwc, na, L, chain_c);
    cr = __shfl_sync(pm, cr, (threadIdx.x & 31u) & ~1u);
    if (cr) return cr & 0x7FFFFFFFu;
  }
  Slot s;
  const bool e = expand_child<MODE>(c, g, L, (int)child, kind, sel, tag, role, ia, ib, iwa, iwb, iwc, na, nb, s);
  const bool eo = __shfl_xor_sync(pm, (uint32_t)e, 1) != 0;
  if (!(e && eo)) return JS_RETRY;  // both children were tried: all their missing values are proposed
  bool ok = true;
  if (s.kind == JC_REF) { s.target = dt_find_or_insert<MODE>(c.E, c.sh, c.items, s.ia, s.ib, s.iwa, s.iwb, s.iwc, s.itag); ok = s.target != JI_FAIL; }
  const bool oko = __shfl_xor_sync(pm, (uint32_t)ok, 1) != 0;
  if (!(ok && oko)) return JS_DONE;  // error / want_grid already flagged; the gate is abandoned
  it->ch[child] = s.target; it->chw[child] = s.w;
  // same-node fused item whose two children carry the node's own weights: normalised pair again
  const bool keeps = s.kind != JC_ZERO && s.w == (child ? nd_w1(na) : nd_w0(na));
  const bool keeps_o = __shfl_xor_sync(pm, (uint32_t)keeps, 1) != 0;
  const uint32_t ko = __shfl_xor_sync(pm, s.kind, 1);
  const uint32_t hint = (kind == JK_F && ia == ib && keeps && keeps_o) ? 1u + nd_flag(na) : 0u;
  return item_flags(JS_DONE, child ? ko : s.kind, child ? s.kind : ko, hint);
}
#endif

// ------------------------------------------------------------------------------------ reduce
// Bottom-up step: QMDDmakeNonterminal (QMDDpackage.cpp:1204-1233) = zero test, NormC normalisation
// (QMDDnormalize :532-595: divide by the largest-magnitude child weight, first index wins ties),
// hash-cons.  Result edge = (node, top weight), to be scaled by the parent's factor.

// |a1| > |a0| beyond the snapping error of the canonical parts?  (index 0 wins ties, :565)
template <int MODE> JD bool mag_greater(const Ctx &c, uint64_t a1, uint64_t a0) {
  if (a1 == a0 || a1 == jneg(a0)) return false;
  double m0 = jmag2<MODE>(c, a0), m1 = jmag2<MODE>(c, a1);
#ifdef JKQ_EXACT_TIES
  return m1 > m0;  // the reference's rule: exact compare, decided by rounding noise on true ties
#else
  // Index 0 also wins when the two magnitudes differ by less than the snapping error of their
  // canonical parts (each real is only defined up to tol): H-type gates produce equal magnitudes all
  // the time, and deciding such a tie by rounding noise gives two normal forms of one vector that
  // never merge again (DESIGN.md "normalisation ties").
  return m1 > JKQ_DADD(m0, JKQ_DMUL(JKQ_DMUL(4.0, jtol<MODE>(c)), sqrt(m1 > m0 ? m1 : m0)));
#endif
}

// child edge i = f[i] * a[i] -> rn[i]; hint = 1 + index of a ONE among (a0, a1) known to be a normalised pair
template <int MODE>
JD bool reduce_core(Ctx &c, uint32_t rn[2], const uint64_t f[2], const uint64_t a[2], uint32_t hint, uint32_t *out_node, uint64_t *out_w) {
  if (a[0] == 0 && a[1] == 0) { *out_node = 0; *out_w = 0; return true; }
  uint32_t mx;
  uint64_t other, top;
  if (hint && a[hint - 1] == JW_ONE && (f[0] == f[1] || a[2 - hint] == 0)) {
    // (a0, a1) are the weights of an already normalised node and both children scale by the same factor:
    // same normal form, the factor is the new top weight -- no arithmetic, no value-table access
    mx = hint - 1;
    other = a[1 - mx];
    top = f[mx];
  } else if (a[0] == 0 || a[1] == 0 || f[0] == f[1]) {
    // Common factor: (F a0, F a1) normalises exactly like (a0, a1) and F moves to the top weight.
    // For a product state or a basis-like state this removes ALL arithmetic from the levels above the
    // gate (a_max is ONE there); the reference gets the same effect from its complex op caches.
    mx = (a[0] == 0) ? 1u : (a[1] == 0 ? 0u : (mag_greater<MODE>(c, a[1], a[0]) ? 1u : 0u));
    bool o1 = cdiv<MODE>(c, mx ? a[0] : a[1], mx ? a[1] : a[0], &other);
    bool o2 = cmul<MODE>(c, mx ? f[1] : f[0], mx ? a[1] : a[0], &top);
    if (!(o1 && o2)) return false;
  } else {
    uint64_t rw0, rw1;
    bool o1 = cmul<MODE>(c, f[0], a[0], &rw0), o2 = cmul<MODE>(c, f[1], a[1], &rw1);
    if (!(o1 && o2)) return false;
    if (rw0 == 0 && rw1 == 0) { *out_node = 0; *out_w = 0; return true; }
    mx = (rw0 == 0) ? 1u : (rw1 == 0 ? 0u : (mag_greater<MODE>(c, rw1, rw0) ? 1u : 0u));
    if (!cdiv<MODE>(c, mx ? rw0 : rw1, mx ? rw1 : rw0, &other)) return false;
    top = mx ? rw1 : rw0;
  }
  if (top == 0) { *out_node = 0; *out_w = 0; return true; }
  *out_node = ut_find_or_insert<MODE>(c.E, c.sh, ((mx == 1 && other == 0) ? 0u : rn[0]) | (mx << 31), (mx == 0 && other == 0) ? 0u : rn[1], other);
  *out_w = top;
  return true;
}

template <int MODE> JD bool reduce_item_impl(Ctx &c, uint32_t idx, uint32_t flags);
template <int MODE> JD bool reduce_item(Ctx &c, uint32_t idx, uint32_t flags) {
  JFINE_T(t0);
  const bool r = reduce_item_impl<MODE>(c, idx, flags);
  JFINE_T(t1);
  JFINE_ADD(c, 6, t0, t1);
  JFINE_ADD(c, 7, 0, 1);
  return r;
}
template <int MODE> JD bool reduce_item_impl(Ctx &c, uint32_t idx, uint32_t flags) {
  JItem *it = c.items + idx;
  if (MODE == MODE_CTA && (flags & JFLAG_CHAIN)) {
    // compressed chain (expand_item): rebuild the skipped nodes over the child's result, bottom-up
    const JItem *ci = c.items + ld_i<MODE>(&it->ch[0]);
    const uint64_t f0 = ld_i<MODE>(&ci->rw);
    uint32_t r = ld_i<MODE>(&ci->rnode);
    if (f0 == 0) { it->rw = 0; it->rnode = 0; return true; }
    const uint32_t info = ld_i<MODE>(&it->ch[1]), k = info & 0xFFu, off = info >> 8;
    const unsigned long long mask = c.sh->chain[off];
    JFINE_T(y0);
    // The rebuilt chain is a function of (the item's node, which fixes the chain below it; the number of skipped
    // levels; the result node under the chain) -- and three rebuilds out of four repeat an earlier one (QFT-32), so
    // they are memoised until the collector renumbers the nodes: ids are below 2048 on this path, key and value
    // share one 64-bit word.
    const unsigned long long ckey = (1ull << 28) | ((unsigned long long)ld_i<MODE>(&it->a) << 17) | ((unsigned long long)k << 11) | r;
    const uint32_t cslot = hfin((uint32_t)ckey * 0x9E3779B1u) & (JKQ_SCMEMO - 1u);
    const unsigned long long cw = ld_r<MODE>(&c.sh->cmemo[cslot]);
#ifdef JKQ_NO_CMEMO
    if (false) r = 0;
#else
    if ((cw >> 11) == ckey) r = (uint32_t)(cw & 0x7FFull);
#endif
    else {
      for (uint32_t j = k; j-- > 0 && r != 0;) {
        const uint32_t fl = (uint32_t)((mask >> j) & 1ull);
        const uint64_t w = c.sh->chain[off + 1 + j];  // 0: the node's other edge is the zero edge
        r = ut_find_or_insert<MODE>(c.E, c.sh, ((fl == 1 && w == 0) ? 0u : r) | (fl << 31), (fl == 0 && w == 0) ? 0u : r, w);
      }
      if (r != 0) c.sh->cmemo[cslot] = (ckey << 11) | r;
    }
    JFINE_T(y1);
    JFINE_ADD(c, 8, y0, y1);
    uint32_t rn[2] = {r, r};
    const uint64_t f[2] = {f0, f0}, a[2] = {ld_i<MODE>(&it->chw[0]), ld_i<MODE>(&it->chw[1])};
    uint32_t rnode;
    uint64_t rw;
    if (r == 0) { it->rw = f0; it->rnode = 0; return true; }  // a pool ran out: the sticky word is set
    if (!reduce_core<MODE>(c, rn, f, a, (flags >> 6) & 3u, &rnode, &rw)) return false;
    it->rw = rw; it->rnode = rnode;
    return true;
  }
  // each child edge is f_i * a_i: f_i = top weight of the child item's result (ONE for finished
  // edges), a_i = the factor / edge weight recorded at expansion
  uint32_t rn[2];
  uint64_t f[2], a[2];
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    const uint32_t k = (flags >> (2 + 2 * i)) & 3u;
    const uint32_t cht = ld_i<MODE>(&it->ch[i]);
    const uint64_t chw = ld_i<MODE>(&it->chw[i]);
    if (k == JC_ZERO) { rn[i] = 0; f[i] = JW_ONE; a[i] = 0; }
    else if (k == JC_PASS) { rn[i] = cht; f[i] = JW_ONE; a[i] = chw; }
    else {
      const JItem *ci = c.items + cht;
      f[i] = ld_i<MODE>(&ci->rw);
      rn[i] = ld_i<MODE>(&ci->rnode);
      a[i] = f[i] ? chw : 0;
    }
    if (a[i] == 0) rn[i] = 0;
  }
  uint32_t rnode;
  uint64_t rw;
  if (!reduce_core<MODE>(c, rn, f, a, (flags >> 6) & 3u, &rnode, &rw)) return false;
  it->rw = rw; it->rnode = rnode;
  return true;
}

#endif  // JKQ_ENGINE_CUH
