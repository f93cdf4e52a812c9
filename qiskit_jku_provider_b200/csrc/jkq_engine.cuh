// Device side of the B200 QMDD engine: value table, unique table, frontier de-dup, level-synchronous
// gate application.  Hand-written CUDA for sm_100a.
//
// The same code runs in two launch shapes (template parameter MODE):
//   MODE_CTA   one CTA, __syncthreads between phases.  The whole working set of a SMALL decision
//              diagram lives in that SM's shared memory (160 KB): node records, unique table, value
//              table, frontier pool, de-dup table, the current gate.  A gate on a 32-node DD is a
//              dependent chain of ~2n level phases, so its cost is latency: an L2 round trip is ~0.23 us
//              on B200, a shared-memory access ~15 ns.  The HBM arena is written through (nodes, reals)
//              so the state is always complete in HBM; the HBM hash tables go stale and are rebuilt by
//              the collector when the grid path takes over.
//   MODE_GRID  cooperative grid, software grid barrier, every table in HBM, ld.global.cg everywhere --
//              the throughput path for blown-up DDs.
// With JKQ_EMU defined the header compiles as plain C++ and runs as ONE emulated thread.  That build
// exists only under tests/emu (logic debugging in a GPU-less container); the product never loads it.
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
#define JKQ_UNROLL
#define JKQ_DMUL(a, b) ((a) * (b))
#define JKQ_DADD(a, b) ((a) + (b))
#define JKQ_DSUB(a, b) ((a) - (b))
#define JKQ_DDIV(a, b) ((a) / (b))
template <class T> static inline T jatomic_add(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T jatomic_cas(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <int MODE> static inline void jfence() {}
#else
#define JD __device__ __forceinline__
#define JDM __device__ __forceinline__
#define JKQ_UNROLL _Pragma("unroll")
#define JKQ_DMUL(a, b) __dmul_rn((a), (b))
#define JKQ_DADD(a, b) __dadd_rn((a), (b))
#define JKQ_DSUB(a, b) __dsub_rn((a), (b))
#define JKQ_DDIV(a, b) __ddiv_rn((a), (b))
template <class T> JD T jatomic_add(T *p, T v) { return atomicAdd(p, v); }
template <class T> JD T jatomic_cas(T *p, T cmp, T v) { return atomicCAS(p, cmp, v); }
template <int MODE> JD void jfence() {
  if (MODE == MODE_CTA) __threadfence_block();
  else __threadfence();
}
#endif

// ------------------------------------------------------------------------------------ MODE_CTA state
#define JKQ_SITEMS 768u      // work items of one gate
#define JKQ_SDT_SLOTS 2048u  // de-dup slots (u32 item index + 1)
#define JKQ_SNODES 2048u     // node ids held in shared memory; the collector keeps ids below this
#define JKQ_SUT_SLOTS 4096u  // unique-table slots (u32 node id)
#define JKQ_SVALS 1024u      // reals
#define JKQ_SVT_SLOTS 2048u  // value-table slots
struct CtaState {
  uint32_t item_count, node_count, pend_n[2];
  uint32_t nodes_created, err, want_grid, val_count;
  uint32_t val_inserts, pad0, pad1, pad2;
  double tol, inv_tol;
  uint64_t cur_u[4];
  uint32_t lb[JKQ_MAXQ + 1], le[JKQ_MAXQ + 1];
  JGate gate;
  uint32_t dtab[JKQ_SDT_SLOTS];
  uint32_t utab[JKQ_SUT_SLOTS];
  double vals[JKQ_SVALS];
  JVSlot vtab[JKQ_SVT_SLOTS];
  JNode nodes[JKQ_SNODES];
  JItem items[JKQ_SITEMS];
};

struct Ctx {
  JEng *E;
  CtaState *sh;         // MODE_CTA only
  JItem *items;         // frontier pool: shared memory in MODE_CTA, HBM in MODE_GRID
  uint32_t spare_node;  // node id allocated by this thread and not used yet (the node already existed)
  uint32_t tid, nth;    // index / count over all participating threads
  uint32_t phase;       // barrier counter, identical in every thread
  uint32_t block0;      // 1 if this thread belongs to the CTA that runs the serial sections
  uint32_t ltid, lnth;
};

template <int MODE> JD uint32_t *cnt_items(const Ctx &c) { return MODE == MODE_CTA ? &c.sh->item_count : &c.E->item_count; }
template <int MODE> JD uint32_t *cnt_nodes(const Ctx &c) { return MODE == MODE_CTA ? &c.sh->node_count : &c.E->node_count; }
template <int MODE> JD uint32_t *cnt_vals(const Ctx &c) { return MODE == MODE_CTA ? &c.sh->val_count : &c.E->val_count; }
template <int MODE> JD uint32_t *cnt_pend(const Ctx &c, uint32_t par) { return MODE == MODE_CTA ? &c.sh->pend_n[par] : &c.E->pend_n[par]; }
template <int MODE> JD double jtol(const Ctx &c) { return MODE == MODE_CTA ? c.sh->tol : c.E->tol; }
template <int MODE> JD double jinv_tol(const Ctx &c) { return MODE == MODE_CTA ? c.sh->inv_tol : c.E->inv_tol; }

// ------------------------------------------------------------------------------------ loads
// ld_s: HBM data not written since the last barrier (MODE_GRID: ld.global.cg, L1 may be stale).
// ld_r: HBM data that may be written concurrently.
#ifdef JKQ_EMU
template <int MODE, class T> JD T ld_s(const T *p) { return *p; }
template <class T> JD T ld_r(const T *p) { return *p; }
#else
template <int MODE> JD uint32_t ld_s(const uint32_t *p) { return MODE == MODE_GRID ? __ldcg(p) : *p; }
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
  uint4 v = MODE == MODE_GRID ? __ldcg((const uint4 *)p) : *(const uint4 *)p;
  JVSlot s;
  s.value = __hiloint2double((int)v.y, (int)v.x); s.id1 = v.z; s.keylo = v.w;
  return s;
}
JD uint32_t ld_r(const uint32_t *p) { return __ldcg(p); }
JD unsigned long long ld_r(const unsigned long long *p) { return __ldcg(p); }
JD uint64_t ld_r(const uint64_t *p) { return (uint64_t)__ldcg((const unsigned long long *)p); }
JD uint8_t ld_r(const uint8_t *p) { return __ldcg(p); }
JD int32_t ld_r(const int32_t *p) { return __ldcg(p); }
JD double ld_r(const double *p) { return __ldcg(p); }
JD JNode ld_r(const JNode *p) {
  uint4 v = __ldcg((const uint4 *)p);
  JNode n;
  n.c0 = v.x; n.c1 = v.y; n.w = ((uint64_t)v.w << 32) | v.z;
  return n;
}
#endif

#ifdef JKQ_EMU
template <int MODE> JD uint32_t rd_items(const Ctx &c) { return *cnt_items<MODE>(c); }
template <int MODE> JD uint32_t rd_nodes(const Ctx &c) { return *cnt_nodes<MODE>(c); }
template <int MODE> JD uint32_t rd_vals(const Ctx &c) { return *cnt_vals<MODE>(c); }
template <int MODE> JD bool jhas_error(const Ctx &c) { return c.E->status != 0; }
#else
template <int MODE> JD uint32_t rd_items(const Ctx &c) { return MODE == MODE_CTA ? c.sh->item_count : __ldcg(&c.E->item_count); }
template <int MODE> JD uint32_t rd_nodes(const Ctx &c) { return MODE == MODE_CTA ? c.sh->node_count : __ldcg(&c.E->node_count); }
template <int MODE> JD uint32_t rd_vals(const Ctx &c) { return MODE == MODE_CTA ? c.sh->val_count : __ldcg(&c.E->val_count); }
template <int MODE> JD bool jhas_error(const Ctx &c) { return MODE == MODE_CTA ? c.sh->err != 0 : __ldcg(&c.E->status) != 0; }
#endif

JD uint64_t jmix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
// 32-bit finaliser (Wellons' lowbias32): the table hashes sit on the latency path of every item, and a
// 64-bit multiply is four 32-bit IMADs on the GPU
JD uint32_t h32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
JD uint32_t h64(uint64_t x) { return h32((uint32_t)x ^ h32((uint32_t)(x >> 32) + 0x9E3779B9u)); }

// MODE_CTA polls a shared-memory copy of the error flag instead of an L2 round trip per phase
template <int MODE> JD void jset_error_c(Ctx &c, int code) {
  jatomic_cas(&c.E->status, 0, code);
  if (MODE == MODE_CTA) c.sh->err = (uint32_t)code;
}

// ------------------------------------------------------------------------------------ value table
// Semantics of the reference's Clookup (QMDDcomplexD.cpp:372-466) with its tolerance comparator
// (:85-98): table entries are pairwise more than tol apart, a lookup returns the SMALLEST entry in
// [x-tol, x+tol], a miss appends x.  One bucket of width tol holds at most one entry, so a lookup
// probes the buckets k-1, k, k+1 in that order.  Anything within tol of 0 is index 0.
template <int MODE> JD double jval(const Ctx &c, uint32_t half) {
  uint32_t idx = half & 0x7FFFFFFFu;
  return MODE == MODE_CTA ? c.sh->vals[idx] : ld_s<MODE>(c.E->vals + idx);
}
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

template <int MODE> JD uint32_t vt_mask(const Ctx &c) { return MODE == MODE_CTA ? JKQ_SVT_SLOTS - 1 : c.E->vtab_mask; }
template <int MODE> JD JVSlot vt_load(const Ctx &c, uint32_t h) { return MODE == MODE_CTA ? c.sh->vtab[h] : ld_s<MODE>(c.E->vtab + h); }
JD long long vt_bucket(double x, double inv_tol) { return (long long)floor(JKQ_DMUL(x, inv_tol)); }

template <int MODE> JD bool real_find(const Ctx &c, double x, uint32_t *idx) {
  const double tol = jtol<MODE>(c);
  if (x <= tol) { *idx = 0; return true; }
  if (!(x < 1.0e8)) return false;  // absurd magnitude: treated as a miss, caught as VAL_FULL later
  const long long k = vt_bucket(x, jinv_tol<MODE>(c));
  const uint32_t mask = vt_mask<MODE>(c);
  for (int d = -1; d <= 1; d++) {
    long long kk = k + d;
    uint32_t h = h64((uint64_t)kk) & mask;
    for (uint32_t probes = 0; probes <= mask; probes++) {
      JVSlot s = vt_load<MODE>(c, h);
      if (s.id1 == 0) break;
      if (s.keylo == (uint32_t)kk && fabs(JKQ_DSUB(s.value, x)) <= tol) { *idx = s.id1 - 1; return true; }
      h = (h + 1) & mask;
    }
  }
  return false;
}

// hash-only insert of an existing real (collector / table reload): slot claimed by CAS on id1
template <int MODE> JD void vt_put(Ctx &c, uint32_t idx, double x) {
  const long long k = vt_bucket(x, jinv_tol<MODE>(c));
  const uint32_t mask = vt_mask<MODE>(c);
  JVSlot *tab = MODE == MODE_CTA ? c.sh->vtab : c.E->vtab;
  uint32_t h = h64((uint64_t)k) & mask;
  while (jatomic_cas(&tab[h].id1, 0u, idx + 1) != 0) h = (h + 1) & mask;
  tab[h].value = x;
  tab[h].keylo = (uint32_t)k;
}

// serial section only (one thread): append x, QMDDcomplexD.cpp:409-418
template <int MODE> JD uint32_t real_insert_serial(Ctx &c, double x) {
  JEng *E = c.E;
  if (MODE == MODE_CTA) {
    CtaState *S = c.sh;
    if (S->val_count >= JKQ_SVALS) { S->want_grid = 1; return 0; }  // the shared table is full: grid path
    uint32_t idx = S->val_count++;
    S->vals[idx] = x;
    E->vals[idx] = x;  // write-through: the HBM array is always complete
    vt_put<MODE>(c, idx, x);
    S->val_inserts++;
    return idx;
  }
  if (E->val_count >= E->val_cap || (uint64_t)E->val_count * 2 >= E->vtab_mask) { jset_error_c<MODE>(c, JKQ_ERR_VAL_FULL); return 0; }
  uint32_t idx = E->val_count++;
  E->vals[idx] = x;
  const long long k = vt_bucket(x, E->inv_tol);
  uint32_t h = h64((uint64_t)k) & E->vtab_mask;
  while (ld_s<MODE>(E->vtab + h).id1 != 0) h = (h + 1) & E->vtab_mask;
  JVSlot s;
  s.value = x; s.id1 = idx + 1; s.keylo = (uint32_t)k;
  E->vtab[h] = s;
  E->st_val_inserts++;
  return idx;
}

template <int MODE> JD void pend_push(Ctx &c, double x) {
  JEng *E = c.E;
  uint32_t i = jatomic_add(cnt_pend<MODE>(c, c.phase & 1), 1u);
  if (i < E->pend_cap) E->pend[i] = x;
  else jset_error_c<MODE>(c, JKQ_ERR_PEND_FULL);
}

// Clookup for one complex value.  false = at least one part is not in the table yet; the missing
// parts have been queued for the next serial insert section.
template <int MODE> JD bool cx_find(Ctx &c, double re, double im, uint64_t *out) {
  uint32_t sr = re < 0.0, si = im < 0.0;
  re = fabs(re); im = fabs(im);
  uint32_t r = 0, i = 0;
  bool okr = real_find<MODE>(c, re, &r), oki = real_find<MODE>(c, im, &i);
  if (!okr) pend_push<MODE>(c, re);
  if (!oki) pend_push<MODE>(c, im);
  if (!(okr && oki)) return false;
  if (sr && r) r |= 0x80000000u;
  if (si && i) i |= 0x80000000u;
  *out = ((uint64_t)r << 32) | (uint64_t)i;
  return true;
}

// Cmul / Cadd / Cdiv: the id short-cuts of QMDDcomplexD.cpp:498-777, arithmetic on the canonical
// values in FP64 (no op caches: recomputing is cheaper than a memo probe on a GPU).
template <int MODE> JD bool cmul(Ctx &c, uint64_t a, uint64_t b, uint64_t *out) {
  if (a == JW_ONE) { *out = b; return true; }
  if (b == JW_ONE) { *out = a; return true; }
  if (a == 0 || b == 0) { *out = 0; return true; }
  if (a == JW_MONE) { *out = jneg(b); return true; }
  if (b == JW_MONE) { *out = jneg(a); return true; }
  double ar = jre<MODE>(c, a), ai = jim<MODE>(c, a), br = jre<MODE>(c, b), bi = jim<MODE>(c, b);
  double re = JKQ_DSUB(JKQ_DMUL(ar, br), JKQ_DMUL(ai, bi));
  double im = JKQ_DADD(JKQ_DMUL(ar, bi), JKQ_DMUL(ai, br));
  return cx_find<MODE>(c, re, im, out);
}
template <int MODE> JD bool cadd(Ctx &c, uint64_t a, uint64_t b, uint64_t *out) {
  if (a == 0) { *out = b; return true; }
  if (b == 0) { *out = a; return true; }
  if (a == jneg(b)) { *out = 0; return true; }
  double re = JKQ_DADD(jre<MODE>(c, a), jre<MODE>(c, b));
  double im = JKQ_DADD(jim<MODE>(c, a), jim<MODE>(c, b));
  return cx_find<MODE>(c, re, im, out);
}
template <int MODE> JD bool cdiv(Ctx &c, uint64_t a, uint64_t b, uint64_t *out) {
  if (a == b) { *out = JW_ONE; return true; }
  if (a == 0) { *out = 0; return true; }
  if (b == JW_ONE) { *out = a; return true; }
  if (b == JW_MONE) { *out = jneg(a); return true; }
  if (a == jneg(b)) { *out = JW_MONE; return true; }
  double ar = jre<MODE>(c, a), ai = jim<MODE>(c, a), br = jre<MODE>(c, b), bi = jim<MODE>(c, b);
  double re, im;
  if ((b & 0x7FFFFFFFull) == 0) {  // purely real divisor, :727-736
    re = JKQ_DDIV(ar, br);
    im = JKQ_DDIV(ai, br);
  } else {
    double d = JKQ_DADD(JKQ_DMUL(br, br), JKQ_DMUL(bi, bi));
    re = JKQ_DDIV(JKQ_DADD(JKQ_DMUL(ar, br), JKQ_DMUL(ai, bi)), d);
    im = JKQ_DDIV(JKQ_DSUB(JKQ_DMUL(ai, br), JKQ_DMUL(ar, bi)), d);
  }
  return cx_find<MODE>(c, re, im, out);
}

// serial canonicalisation (one thread, may insert): gate setup and root update
template <int MODE> JD uint64_t cx_make_serial(Ctx &c, double re, double im) {
  uint32_t sr = re < 0.0, si = im < 0.0, r, i;
  re = fabs(re); im = fabs(im);
  if (!real_find<MODE>(c, re, &r)) r = real_insert_serial<MODE>(c, re);
  if (!real_find<MODE>(c, im, &i)) i = real_insert_serial<MODE>(c, im);
  if (sr && r) r |= 0x80000000u;
  if (si && i) i |= 0x80000000u;
  return ((uint64_t)r << 32) | (uint64_t)i;
}

// ------------------------------------------------------------------------------------ node access
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
// in a slot therefore always finds a complete record behind it.
JD uint32_t ut_hash(const JNode &n) { return h32(n.c0 ^ h32(n.c1 + 0x9E3779B9u) ^ h64(n.w)); }

template <int MODE> JD uint32_t ut_find_or_insert(Ctx &c, const JNode &rec) {
  JEng *E = c.E;
  if (MODE == MODE_CTA) {
    CtaState *S = c.sh;
    uint32_t h = ut_hash(rec) & (JKQ_SUT_SLOTS - 1), mine = 0;
    for (uint32_t probes = 0; probes < JKQ_SUT_SLOTS; probes++) {
      uint32_t s = S->utab[h];
      if (s == 0) {
        if (mine == 0) {
          if (c.spare_node) { mine = c.spare_node; c.spare_node = 0; }
          else mine = jatomic_add(&S->node_count, 1u);
          if (mine >= JKQ_SNODES) { S->want_grid = 1; return 0; }  // gate abandoned, redone on the grid path
          S->nodes[mine] = rec;
          E->nodes[mine] = rec;  // write-through
          jfence<MODE>();
        }
        s = jatomic_cas(&S->utab[h], 0u, mine);
        if (s == 0) { jatomic_add(&S->nodes_created, 1u); return mine; }
      }
      JNode o = S->nodes[s];
      if (o.c0 == rec.c0 && o.c1 == rec.c1 && o.w == rec.w) { c.spare_node = mine; return s; }
      h = (h + 1) & (JKQ_SUT_SLOTS - 1);
    }
    S->want_grid = 1;
    return 0;
  }
  uint32_t h = ut_hash(rec) & E->utab_mask, mine = 0;
  for (uint32_t probes = 0; probes <= E->utab_mask; probes++) {
    uint32_t s = ld_r(E->utab + h);
    if (s == 0) {
      if (mine == 0) {
        if (c.spare_node) { mine = c.spare_node; c.spare_node = 0; }
        else mine = jatomic_add(&E->node_count, 1u);
        if (mine >= E->node_cap) { jset_error_c<MODE>(c, JKQ_ERR_NODE_FULL); return 0; }
        E->nodes[mine] = rec;
        jfence<MODE>();
      }
      s = jatomic_cas(E->utab + h, 0u, mine);
      if (s == 0) { jatomic_add(&E->st_nodes_created, 1ull); return mine; }
    }
    JNode o = ld_r(E->nodes + s);
    if (o.c0 == rec.c0 && o.c1 == rec.c1 && o.w == rec.w) { c.spare_node = mine; return s; }
    h = (h + 1) & E->utab_mask;
  }
  jset_error_c<MODE>(c, JKQ_ERR_UTAB_FULL);
  return 0;
}

// ------------------------------------------------------------------------------------ frontier de-dup
// The compute tables of the reference (CTlookup/CTinsert, QMDDpackage.cpp:986-1072) exist so that a
// (gate,node) or (a,b,ratio) pair is expanded once.  Here the pairs of one gate are the work items of
// the frontier; this table maps an item key to its index.  MODE_GRID: HBM slots carry the gate
// generation in the high word, so nothing is ever cleared.  MODE_CTA: shared slots, cleared per gate.
JD uint32_t dt_hash(uint32_t a, uint32_t b, uint64_t wa, uint64_t wb, uint64_t wc, uint32_t tag) {
  if ((tag & 15u) == JK_M) return h32(a);
  return h32(a ^ h32(b + tag * 0x9E3779B9u) ^ h64(wa) * 3u ^ h64(wb) * 5u ^ h64(wc) * 7u);
}
JD bool item_key_eq(const JItem *o, uint32_t a, uint32_t b, uint64_t wa, uint64_t wb, uint64_t wc, uint32_t tag) {
  return o->a == a && o->b == b && o->tag == tag && o->wa == wa && o->wb == wb && o->wc == wc;
}

// returns the item index; *ok = false: the table could not take the key (error flagged, or -- MODE_CTA --
// the shared pool is full and the gate has to be redone on the grid path: c.sh->want_grid)
template <int MODE>
JD uint32_t dt_find_or_insert(Ctx &c, uint32_t a, uint32_t b, uint64_t wa, uint64_t wb, uint64_t wc, uint32_t tag, bool *ok) {
  JEng *E = c.E;
  uint32_t mine = 0xFFFFFFFFu;
  if (MODE == MODE_CTA) {
    CtaState *S = c.sh;
    uint32_t h = dt_hash(a, b, wa, wb, wc, tag) & (JKQ_SDT_SLOTS - 1);
    for (uint32_t probes = 0; probes < JKQ_SDT_SLOTS; probes++) {
      uint32_t s = S->dtab[h];
      if (s == 0) {
        if (mine == 0xFFFFFFFFu) {
          mine = jatomic_add(&S->item_count, 1u);
          if (mine >= JKQ_SITEMS) { S->want_grid = 1; *ok = false; return 0; }
          JItem *it = c.items + mine;
          it->a = a; it->b = b; it->wa = wa; it->wb = wb; it->wc = wc; it->tag = tag;
          it->state = JS_NEW;
          jfence<MODE>();
        }
        s = jatomic_cas(&S->dtab[h], 0u, mine + 1);
        if (s == 0) return mine;
      }
      if (item_key_eq(c.items + (s - 1), a, b, wa, wb, wc, tag)) {
        if (mine != 0xFFFFFFFFu) c.items[mine].state = JS_DEAD;
        return s - 1;
      }
      h = (h + 1) & (JKQ_SDT_SLOTS - 1);
    }
    S->want_grid = 1;
    *ok = false;
    return 0;
  }
  uint32_t h = dt_hash(a, b, wa, wb, wc, tag) & E->dtab_mask;
  unsigned long long gen = (unsigned long long)ld_s<MODE>(&E->gen) << 32;
  for (uint32_t probes = 0; probes <= E->dtab_mask; probes++) {
    unsigned long long s = ld_r(E->dtab + h);
    while ((s >> 32) != (gen >> 32)) {  // stale or empty: try to claim
      if (mine == 0xFFFFFFFFu) {
        mine = jatomic_add(&E->item_count, 1u);
        if (mine >= E->item_cap) { jset_error_c<MODE>(c, JKQ_ERR_ITEM_FULL); *ok = false; return 0; }
        JItem *it = c.items + mine;
        it->a = a; it->b = b; it->wa = wa; it->wb = wb; it->wc = wc; it->tag = tag;
        it->state = JS_NEW;
        jfence<MODE>();
      }
      unsigned long long old = jatomic_cas(E->dtab + h, s, gen | (unsigned long long)(mine + 1));
      if (old == s) return mine;
      s = old;
    }
    uint32_t j = (uint32_t)s - 1;
    const JItem *o = c.items + j;
    if (ld_r(&o->a) == a && ld_r(&o->b) == b && ld_r(&o->tag) == tag && ld_r(&o->wa) == wa && ld_r(&o->wb) == wb &&
        ld_r(&o->wc) == wc) {
      if (mine != 0xFFFFFFFFu) c.items[mine].state = JS_DEAD;
      return j;
    }
    h = (h + 1) & E->dtab_mask;
  }
  jset_error_c<MODE>(c, JKQ_ERR_DTAB_FULL);
  *ok = false;
  return 0;
}

// item field access: shared memory (plain) in MODE_CTA, ld.global.cg in MODE_GRID
#ifdef JKQ_EMU
template <int MODE, class T> JD T ld_i(const T *p) { return *p; }
#else
template <int MODE, class T> JD T ld_i(const T *p) { return MODE == MODE_CTA ? *p : ld_r(p); }
#endif

// ------------------------------------------------------------------------------------ child slots
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
  if (A > B) { uint32_t t = A; A = B; B = t; uint64_t tw = al; al = be; be = tw; }
  uint64_t r;
  if (!cdiv<MODE>(c, be, al, &r)) return false;
  slot_ref(s, al, A, B, JW_ONE, r, 0, JK_S);
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
  uint64_t f = al ? al : (be ? be : ga), wa, wb, wc;
  if (!cdiv<MODE>(c, al, f, &wa)) return false;
  if (!cdiv<MODE>(c, be, f, &wb)) return false;
  if (!cdiv<MODE>(c, ga, f, &wc)) return false;
  uint32_t An = (wa != 0 || (sel == 0 && wc != 0)) ? A : 0;
  uint32_t Bn = (wb != 0 || (sel == 1 && wc != 0)) ? B : 0;
  slot_ref(s, f, An, Bn, wa, wb, wc, JK_F | (sel << 4));
  return true;
}

JD int role_satisfied(uint32_t role, int i) {
  return role == JR_IDLE || (role == JR_CPOS && i == 1) || (role == JR_CNEG && i == 0);
}

template <int MODE> JD uint64_t gate_u(const Ctx &c, int k) { return MODE == MODE_CTA ? c.sh->cur_u[k] : ld_s<MODE>(&c.E->cur_u[k]); }

// one child slot of an item of level L.  kind M: p = the state node.  kind F/S: na/nb = operand nodes.
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
    if (!cmul<MODE>(c, gate_u<MODE>(c, 2 * i), nd_w0(na), &al)) return false;
    if (!cmul<MODE>(c, gate_u<MODE>(c, 2 * i + 1), nd_w1(na), &be)) return false;
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
      slot_ref(s, aw, ac, ac, iwa, iwb, iwc, tag);
    } else {
      uint64_t sum, w;
      if (!cadd<MODE>(c, iwa, iwb, &sum)) return false;
      if (!cmul<MODE>(c, aw, sum, &w)) return false;
      slot_pass(s, ac, w);
    }
    return true;
  }
  uint64_t ga;
  if (!cmul<MODE>(c, iwc, sel ? bw : aw, &ga)) return false;
  if (!role_satisfied(role, i)) { slot_pass(s, sel ? bc : ac, ga); return true; }
  uint64_t al, be;
  if (!cmul<MODE>(c, iwa, aw, &al)) return false;
  if (!cmul<MODE>(c, iwb, bw, &be)) return false;
  return combine<MODE>(c, g.low, L - 1, al, ac, be, bc, ga, sel, s);
}

// ------------------------------------------------------------------------------------ expand
// Top-down step for one work item of level L: compute its two child slots.  Side effects (de-dup
// inserts) only after all arithmetic succeeded, so a retry simply runs the function again.
// M item key = the state node alone: a leaner probe than the generic 6-field key
template <int MODE> JD uint32_t dt_find_or_insert_m(Ctx &c, uint32_t node, bool *ok) {
  if (MODE == MODE_CTA) {
    CtaState *S = c.sh;
    uint32_t h = h32(node) & (JKQ_SDT_SLOTS - 1), mine = 0xFFFFFFFFu;
    for (uint32_t probes = 0; probes < JKQ_SDT_SLOTS; probes++) {
      uint32_t s = S->dtab[h];
      if (s == 0) {
        if (mine == 0xFFFFFFFFu) {
          mine = jatomic_add(&S->item_count, 1u);
          if (mine >= JKQ_SITEMS) { S->want_grid = 1; *ok = false; return 0; }
          JItem *it = c.items + mine;
          it->a = node; it->b = 0; it->wa = 0; it->wb = 0; it->wc = 0; it->tag = JK_M;
          it->state = JS_NEW;
          jfence<MODE>();
        }
        s = jatomic_cas(&S->dtab[h], 0u, mine + 1);
        if (s == 0) return mine;
      }
      const JItem *o = c.items + (s - 1);
      if (o->a == node && o->tag == JK_M) {
        if (mine != 0xFFFFFFFFu) c.items[mine].state = JS_DEAD;
        return s - 1;
      }
      h = (h + 1) & (JKQ_SDT_SLOTS - 1);
    }
    S->want_grid = 1;
    *ok = false;
    return 0;
  }
  return dt_find_or_insert<MODE>(c, node, 0, 0, 0, 0, JK_M, ok);
}

template <int MODE> JD bool expand_item(Ctx &c, const JGate &g, uint32_t idx, int L) {
  JItem *it = c.items + idx;
  const uint32_t tag = ld_i<MODE>(&it->tag), ia = ld_i<MODE>(&it->a), ib = ld_i<MODE>(&it->b);
  const uint32_t kind = tag & 15u, sel = (tag >> 4) & 1u;
  const uint32_t role = g.role[L];
  if (kind == JK_M && L > g.target) {
    // Fast path, the bulk of every gate on a thin DD: a level above the target only forwards the node's
    // two edges (QMDDmultiply2 with a block-diagonal gate node, QMDDpackage.cpp:1488-1509) -- no arithmetic.
    const JNode p = node_load<MODE>(c, ia);
    const uint32_t ch0 = nd_c0(p), ch1 = p.c1;
    const uint64_t w0 = nd_w0(p), w1 = nd_w1(p);
    uint32_t t0 = ch0, t1 = ch1, k0, k1;
    bool ok = true;
    if (w0 == 0) { k0 = JC_ZERO; t0 = 0; }
    else if (role_satisfied(role, 0)) { k0 = JC_REF; t0 = dt_find_or_insert_m<MODE>(c, ch0, &ok); }
    else k0 = JC_PASS;
    if (w1 == 0) { k1 = JC_ZERO; t1 = 0; }
    else if (role_satisfied(role, 1)) {
      k1 = JC_REF;
      if (k0 == JC_REF && ch1 == ch0) t1 = t0;
      else if (ok) t1 = dt_find_or_insert_m<MODE>(c, ch1, &ok);
    } else k1 = JC_PASS;
    if (!ok) return true;  // error / want_grid flagged; the gate is abandoned
    it->ch[0] = t0; it->chw[0] = w0; it->chk[0] = (uint8_t)k0;
    it->ch[1] = t1; it->chw[1] = w1; it->chk[1] = (uint8_t)k1;
    it->pad0 = (uint8_t)(1u + nd_flag(p));  // (chw0, chw1) are the weights of a normalised node: max = this index
    return true;
  }
  it->pad0 = 0;
  uint64_t iwa = 0, iwb = 0, iwc = 0;
  JNode na, nb;
  na.c0 = na.c1 = 0; na.w = 0; nb = na;
  if (ia) na = node_load<MODE>(c, ia);
  if (kind != JK_M) {
    iwa = ld_i<MODE>(&it->wa); iwb = ld_i<MODE>(&it->wb); iwc = ld_i<MODE>(&it->wc);
    if (ib) nb = (ib == ia) ? na : node_load<MODE>(c, ib);
  }
  Slot s0, s1;
  if (!expand_child<MODE>(c, g, L, 0, kind, sel, tag, role, ia, ib, iwa, iwb, iwc, na, nb, s0)) return false;
  if (!expand_child<MODE>(c, g, L, 1, kind, sel, tag, role, ia, ib, iwa, iwb, iwc, na, nb, s1)) return false;
  bool ok = true;
  if (s0.kind == JC_REF) s0.target = dt_find_or_insert<MODE>(c, s0.ia, s0.ib, s0.iwa, s0.iwb, s0.iwc, s0.itag, &ok);
  if (ok && s1.kind == JC_REF) s1.target = dt_find_or_insert<MODE>(c, s1.ia, s1.ib, s1.iwa, s1.iwb, s1.iwc, s1.itag, &ok);
  if (!ok) return true;  // error / want_grid already flagged; the gate is abandoned
  it->ch[0] = s0.target; it->chw[0] = s0.w; it->chk[0] = (uint8_t)s0.kind;
  it->ch[1] = s1.target; it->chw[1] = s1.w; it->chk[1] = (uint8_t)s1.kind;
  // same-node fused item whose two children carry the node's own weights: normalised pair again
  if (kind == JK_F && ia == ib && s0.w == nd_w0(na) && s1.w == nd_w1(na) && s0.kind != JC_ZERO && s1.kind != JC_ZERO)
    it->pad0 = (uint8_t)(1u + nd_flag(na));
  return true;
}

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

template <int MODE> JD bool reduce_item(Ctx &c, uint32_t idx) {
  JItem *it = c.items + idx;
  // each child edge is f_i * a_i: f_i = top weight of the child item's result (ONE for finished
  // edges), a_i = the factor / edge weight recorded at expansion
  uint32_t rn[2];
  uint64_t f[2], a[2];
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    uint32_t k = ld_i<MODE>(&it->chk[i]);
    uint32_t cht = ld_i<MODE>(&it->ch[i]);
    uint64_t chw = ld_i<MODE>(&it->chw[i]);
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
  if (a[0] == 0 && a[1] == 0) { it->rnode = 0; it->rw = 0; return true; }
  uint32_t mx;
  uint64_t other, top;
  const uint32_t hint = ld_i<MODE>(&it->pad0);
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
    if (!cdiv<MODE>(c, mx ? a[0] : a[1], mx ? a[1] : a[0], &other)) return false;
    if (!cmul<MODE>(c, mx ? f[1] : f[0], mx ? a[1] : a[0], &top)) return false;
  } else {
    uint64_t rw0, rw1;
    if (!cmul<MODE>(c, f[0], a[0], &rw0)) return false;
    if (!cmul<MODE>(c, f[1], a[1], &rw1)) return false;
    if (rw0 == 0 && rw1 == 0) { it->rnode = 0; it->rw = 0; return true; }
    mx = (rw0 == 0) ? 1u : (rw1 == 0 ? 0u : (mag_greater<MODE>(c, rw1, rw0) ? 1u : 0u));
    if (!cdiv<MODE>(c, mx ? rw0 : rw1, mx ? rw1 : rw0, &other)) return false;
    top = mx ? rw1 : rw0;
  }
  if (top == 0) { it->rnode = 0; it->rw = 0; return true; }
  JNode rec;
  rec.c0 = ((mx == 1 && other == 0) ? 0u : rn[0]) | (mx << 31);
  rec.c1 = (mx == 0 && other == 0) ? 0u : rn[1];
  rec.w = other;
  uint32_t id = ut_find_or_insert<MODE>(c, rec);
  it->rnode = id; it->rw = top;
  return true;
}

#endif  // JKQ_ENGINE_CUH
