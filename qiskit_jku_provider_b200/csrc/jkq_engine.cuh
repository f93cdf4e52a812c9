// Device side of the B200 QMDD engine: value table, unique table, frontier de-dup, level-synchronous
// gate application, copying collector.  Hand-written CUDA for sm_100a.
//
// The same code runs in two launch shapes (template parameter MODE):
//   MODE_CTA   one CTA, __syncthreads between phases, plain (L1-cached) loads for data that is stable
//              since the last barrier -- the latency path for DDs of tens to thousands of nodes;
//   MODE_GRID  cooperative grid, software grid barrier, ld.global.cg everywhere -- the throughput path
//              for blown-up DDs.
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
#define JKQ_DMUL(a, b) ((a) * (b))
#define JKQ_DADD(a, b) ((a) + (b))
#define JKQ_DSUB(a, b) ((a) - (b))
#define JKQ_DDIV(a, b) ((a) / (b))
template <class T> static inline T jatomic_add(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T jatomic_cas(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <class T> static inline T jatomic_max(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <int MODE> static inline void jfence() {}
#else
#define JD __device__ __forceinline__
#define JKQ_DMUL(a, b) __dmul_rn((a), (b))
#define JKQ_DADD(a, b) __dadd_rn((a), (b))
#define JKQ_DSUB(a, b) __dsub_rn((a), (b))
#define JKQ_DDIV(a, b) __ddiv_rn((a), (b))
template <class T> JD T jatomic_add(T *p, T v) { return atomicAdd(p, v); }
template <class T> JD T jatomic_cas(T *p, T cmp, T v) { return atomicCAS(p, cmp, v); }
template <class T> JD T jatomic_max(T *p, T v) { return atomicMax(p, v); }
template <int MODE> JD void jfence() {
  if (MODE == MODE_CTA) __threadfence_block();
  else __threadfence();
}
#endif

// MODE_CTA keeps the hot counters and the per-gate de-dup table in shared memory: a shared-memory
// atomic costs tens of cycles, an HBM/L2 atomic several hundred, and on a 32-node DD the atomics are
// the critical path.  The counters are written back to the JEng block at gate boundaries.
#define JKQ_SITEMS 1536u     // work items of one gate that fit the shared pool; beyond: grid path
#define JKQ_SDT_SLOTS 4096u  // shared de-dup slots (u32 item index + 1)
#define JKQ_SNC_SLOTS 2048u  // shared node cache, direct-mapped by node id
struct CtaState {
  uint32_t item_count, node_count, pend_n[2];
  uint32_t nodes_created, err, want_grid, pad;
  uint32_t dtab[JKQ_SDT_SLOTS];
  uint32_t ntag[JKQ_SNC_SLOTS];
  JNode ncache[JKQ_SNC_SLOTS];
  JItem items[JKQ_SITEMS];
};

struct Ctx {
  JEng *E;
  CtaState *sh;       // MODE_CTA only
  JItem *items;       // frontier pool: shared memory in MODE_CTA, HBM in MODE_GRID
  uint32_t spare_node;  // node id allocated by this thread and not used yet (lost unique-table race)
  uint32_t tid, nth;  // index / count over all participating threads
  uint32_t phase;     // barrier counter, identical in every thread
  uint32_t block0;    // 1 if this thread belongs to the CTA that runs the serial sections
  uint32_t ltid, lnth;
};

template <int MODE> JD uint32_t *cnt_items(const Ctx &c) { return MODE == MODE_CTA ? &c.sh->item_count : &c.E->item_count; }
template <int MODE> JD uint32_t *cnt_nodes(const Ctx &c) { return MODE == MODE_CTA ? &c.sh->node_count : &c.E->node_count; }
template <int MODE> JD uint32_t *cnt_pend(const Ctx &c, uint32_t par) { return MODE == MODE_CTA ? &c.sh->pend_n[par] : &c.E->pend_n[par]; }
template <int MODE> JD uint32_t rd_items(const Ctx &c);
template <int MODE> JD uint32_t rd_nodes(const Ctx &c);

// ------------------------------------------------------------------------------------ loads
// ld_s: data not written since the last barrier.  ld_r: data that may be written concurrently.
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

// ld_p: probe of a slot / record that may be written concurrently.  One CTA = one SM = one coherent L1
// for plain stores, and every claim goes through atomicCAS which returns the truth, so MODE_CTA can
// probe through L1; MODE_GRID must read L2.
#ifdef JKQ_EMU
template <int MODE, class T> JD T ld_p(const T *p) { return *p; }
template <int MODE> JD uint32_t rd_items(const Ctx &c) { return *cnt_items<MODE>(c); }
template <int MODE> JD uint32_t rd_nodes(const Ctx &c) { return *cnt_nodes<MODE>(c); }
template <int MODE> JD bool jhas_error(const Ctx &c) { return c.E->status != 0; }
#else
template <int MODE, class T> JD T ld_p(const T *p) { return MODE == MODE_CTA ? *p : ld_r(p); }
template <int MODE> JD uint32_t rd_items(const Ctx &c) { return MODE == MODE_CTA ? c.sh->item_count : __ldcg(&c.E->item_count); }
template <int MODE> JD uint32_t rd_nodes(const Ctx &c) { return MODE == MODE_CTA ? c.sh->node_count : __ldcg(&c.E->node_count); }
template <int MODE> JD bool jhas_error(const Ctx &c) { return MODE == MODE_CTA ? c.sh->err != 0 : __ldcg(&c.E->status) != 0; }
#endif

JD uint64_t jmix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

JD void jset_error(JEng *E, int code) { jatomic_cas(&E->status, 0, code); }
// MODE_CTA polls a shared-memory copy of the error flag instead of an L2 round trip per phase
template <int MODE> JD void jset_error_c(Ctx &c, int code) {
  jatomic_cas(&c.E->status, 0, code);
  if (MODE == MODE_CTA) c.sh->err = (uint32_t)code;
}
template <int MODE> JD bool jhas_error(const Ctx &c);

// ------------------------------------------------------------------------------------ value table
// Semantics of the reference's Clookup (QMDDcomplexD.cpp:372-466) with its tolerance comparator
// (:85-98): table entries are pairwise more than tol apart, a lookup returns the SMALLEST entry in
// [x-tol, x+tol], a miss appends x.  One bucket of width tol holds at most one entry, so a lookup
// probes the buckets k-1, k, k+1 in that order.  Anything within tol of 0 is index 0.
template <int MODE> JD double jval(const Ctx &c, uint32_t half) { return ld_s<MODE>(c.E->vals + (half & 0x7FFFFFFFu)); }
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

template <int MODE> JD bool real_find(const Ctx &c, double x, uint32_t *idx) {
  const JEng *E = c.E;
  if (x <= E->tol) { *idx = 0; return true; }
  double kd = floor(JKQ_DDIV(x, E->tol));
  if (!(kd < 9.0e18)) return false;  // absurd magnitude: treated as a miss, caught as VAL_FULL later
  long long k = (long long)kd;
  for (int d = -1; d <= 1; d++) {
    long long kk = k + d;
    uint32_t h = (uint32_t)jmix((uint64_t)kk) & E->vtab_mask;
    for (uint32_t probes = 0; probes <= E->vtab_mask; probes++) {
      JVSlot s = ld_s<MODE>(E->vtab + h);
      if (s.id1 == 0) break;
      if (s.keylo == (uint32_t)kk && fabs(JKQ_DSUB(s.value, x)) <= E->tol) { *idx = s.id1 - 1; return true; }
      h = (h + 1) & E->vtab_mask;
    }
  }
  return false;
}

// serial section only (one thread): append x, QMDDcomplexD.cpp:409-418
template <int MODE> JD uint32_t real_insert_serial(JEng *E, double x) {
  if (E->val_count >= E->val_cap || (uint64_t)E->val_count * 2 >= E->vtab_mask) { jset_error(E, JKQ_ERR_VAL_FULL); return 0; }
  uint32_t idx = E->val_count++;
  E->vals[idx] = x;
  long long k = (long long)floor(JKQ_DDIV(x, E->tol));
  uint32_t h = (uint32_t)jmix((uint64_t)k) & E->vtab_mask;
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
  if (!real_find<MODE>(c, re, &r)) r = real_insert_serial<MODE>(c.E, re);
  if (!real_find<MODE>(c, im, &i)) i = real_insert_serial<MODE>(c.E, im);
  if (sr && r) r |= 0x80000000u;
  if (si && i) i |= 0x80000000u;
  return ((uint64_t)r << 32) | (uint64_t)i;
}

// ------------------------------------------------------------------------------------ node access
JD uint32_t nd_flag(const JNode &n) { return n.c0 >> 31; }
JD uint32_t nd_c0(const JNode &n) { return n.c0 & 0x7FFFFFFFu; }
JD uint64_t nd_w0(const JNode &n) { return nd_flag(n) ? n.w : JW_ONE; }
JD uint64_t nd_w1(const JNode &n) { return nd_flag(n) ? JW_ONE : n.w; }

// ------------------------------------------------------------------------------------ unique table
// QMDDutLookup (QMDDpackage.cpp:693-763) as a lock-free open-addressing table of node ids.
// Protocol: write the record, fence, publish the id with atomicCAS.  A reader that observes an id
// in a slot therefore always finds a complete record behind it (read with ld.cg).
JD uint32_t ut_hash(const JNode &n) { return (uint32_t)jmix(((uint64_t)n.c0 << 32 | n.c1) ^ jmix(n.w)); }

// MODE_CTA reads nodes through a direct-mapped shared-memory cache (a plain global store drops the
// line from L1, so a node written by one gate would cost the next gate an L2 round trip).
template <int MODE> JD JNode node_load(Ctx &c, uint32_t id) {
  if (MODE == MODE_CTA) {
    CtaState *S = c.sh;
    uint32_t s = id & (JKQ_SNC_SLOTS - 1);
    if (S->ntag[s] == id) {
      JNode n = S->ncache[s];
      if (S->ntag[s] == id) return n;  // (another thread may be replacing the slot: re-check)
    }
    JNode n = ld_s<MODE>(c.E->nodes + id);
    return n;
  }
  return ld_s<MODE>(c.E->nodes + id);
}
template <int MODE> JD void node_cache_put(Ctx &c, uint32_t id, const JNode &n) {
  if (MODE == MODE_CTA) {
    CtaState *S = c.sh;
    uint32_t s = id & (JKQ_SNC_SLOTS - 1);
    S->ntag[s] = 0;  // invalidate, fill, publish: a reader that sees the new tag twice saw the new record
    jfence<MODE>();
    S->ncache[s] = n;
    jfence<MODE>();
    S->ntag[s] = id;
  }
}

template <int MODE> JD uint32_t ut_find_or_insert(Ctx &c, const JNode &rec) {
  JEng *E = c.E;
  uint32_t h = ut_hash(rec) & E->utab_mask, mine = 0;
  for (uint32_t probes = 0; probes <= E->utab_mask; probes++) {
    // MODE_CTA goes straight for the CAS (one L2 round trip when the slot is free, the common case on
    // the latency path); MODE_GRID looks first so that a hit costs no node id
    uint32_t s = MODE == MODE_CTA ? 0u : ld_r(E->utab + h);
    if (s == 0) {
      if (mine == 0) {
        if (c.spare_node) { mine = c.spare_node; c.spare_node = 0; }
        else mine = jatomic_add(cnt_nodes<MODE>(c), 1u);
        if (mine >= E->node_cap) { jset_error_c<MODE>(c, JKQ_ERR_NODE_FULL); return 0; }
        E->nodes[mine] = rec;
        jfence<MODE>();
      }
      s = jatomic_cas(E->utab + h, 0u, mine);
      if (s == 0) {
        if (MODE == MODE_CTA) jatomic_add(&c.sh->nodes_created, 1u);
        else jatomic_add(&E->st_nodes_created, 1ull);
        node_cache_put<MODE>(c, mine, rec);
        return mine;
      }
    }
    JNode o = MODE == MODE_CTA ? node_load<MODE>(c, s) : ld_r(E->nodes + s);
    if (o.c0 == rec.c0 && o.c1 == rec.c1 && o.w == rec.w) {
      c.spare_node = mine;  // keep the id for this thread's next insert
      return s;
    }
    h = (h + 1) & E->utab_mask;
  }
  jset_error_c<MODE>(c, JKQ_ERR_UTAB_FULL);
  return 0;
}

// ------------------------------------------------------------------------------------ frontier de-dup
// The compute tables of the reference (CTlookup/CTinsert, QMDDpackage.cpp:986-1072) exist so that a
// (gate,node) or (a,b,ratio) pair is expanded once.  Here the pairs of one gate are the work items of
// the frontier; this table maps an item key to its index.  Slots carry the gate generation in the
// high word, so nothing is ever cleared.
JD uint32_t dt_hash(uint32_t a, uint32_t b, uint64_t wa, uint64_t wb, uint64_t wc, uint32_t tag) {
  return (uint32_t)jmix(((uint64_t)a << 32 | b) ^ jmix(wa + 0x9E3779B97F4A7C15ull * tag) ^ jmix(wb * 3 + 1) ^ jmix(wc * 5 + 2));
}

JD bool item_key_eq(const JItem *o, uint32_t a, uint32_t b, uint64_t wa, uint64_t wb, uint64_t wc, uint32_t tag) {
  return o->a == a && o->b == b && o->tag == tag && o->wa == wa && o->wb == wb && o->wc == wc;
}

// returns the item index; *ok = false: the table could not take the key (error flagged, or -- MODE_CTA --
// the shared table is crowded and the gate has to be redone on the grid path: c.sh->want_grid)
template <int MODE>
JD uint32_t dt_find_or_insert(Ctx &c, uint32_t a, uint32_t b, uint64_t wa, uint64_t wb, uint64_t wc, uint32_t tag, bool *ok) {
  JEng *E = c.E;
  uint32_t mine = 0xFFFFFFFFu;
  if (MODE == MODE_CTA) {
    // shared-memory table of item indices, cleared at gate setup
    CtaState *S = c.sh;
    uint32_t h = dt_hash(a, b, wa, wb, wc, tag) & (JKQ_SDT_SLOTS - 1);
    for (uint32_t probes = 0; probes < 64; probes++) {
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
JD bool combine(Ctx &c, const JGate &g, int m, uint64_t al, uint32_t A, uint64_t be, uint32_t B, uint64_t ga, uint32_t sel, Slot &s) {
  if (m < g.low) return combine_sum<MODE>(c, m, al, A, be, B, s);
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

// ------------------------------------------------------------------------------------ expand
// Top-down step for one work item of level L: compute its two child slots.  Side effects (de-dup
// inserts) only after all arithmetic succeeded, so a retry simply runs the function again.
template <int MODE> JD bool expand_item(Ctx &c, const JGate &g, uint32_t idx, int L) {
  JEng *E = c.E;
  JItem *it = c.items + idx;
  const uint32_t tag = ld_s<MODE>(&it->tag), ia = ld_s<MODE>(&it->a), ib = ld_s<MODE>(&it->b);
  const uint32_t kind = tag & 15u, sel = (tag >> 4) & 1u;
  const uint32_t role = g.role[L];
  Slot s[2];
  if (kind == JK_M) {
    JNode p = node_load<MODE>(c, ia);
    uint32_t ch[2] = {nd_c0(p), p.c1};
    uint64_t w[2] = {nd_w0(p), nd_w1(p)};
    if (L > g.target) {
      for (int i = 0; i < 2; i++) {
        if (w[i] == 0) slot_zero(s[i]);
        else if (role_satisfied(role, i)) slot_ref(s[i], w[i], ch[i], 0, 0, 0, 0, JK_M);
        else slot_pass(s[i], ch[i], w[i]);
      }
    } else {  // target level: rows of the 2x2 matrix, QMDDpackage.cpp:1488-1509
      for (int i = 0; i < 2; i++) {
        uint64_t al, be;
        if (!cmul<MODE>(c, ld_s<MODE>(&E->cur_u[2 * i]), w[0], &al)) return false;
        if (!cmul<MODE>(c, ld_s<MODE>(&E->cur_u[2 * i + 1]), w[1], &be)) return false;
        if (!combine<MODE>(c, g, L - 1, al, ch[0], be, ch[1], w[i], (uint32_t)i, s[i])) return false;
      }
    }
  } else {
    JNode na, nb;
    na.c0 = na.c1 = 0; na.w = 0; nb = na;
    uint64_t aw[2] = {0, 0}, bw[2] = {0, 0};
    uint32_t ac[2] = {0, 0}, bc[2] = {0, 0};
    const uint64_t iwa = ld_s<MODE>(&it->wa), iwb = ld_s<MODE>(&it->wb), iwc = ld_s<MODE>(&it->wc);
    if (ia) { na = node_load<MODE>(c, ia); ac[0] = nd_c0(na); ac[1] = na.c1; aw[0] = nd_w0(na); aw[1] = nd_w1(na); }
    if (ib) { nb = node_load<MODE>(c, ib); bc[0] = nd_c0(nb); bc[1] = nb.c1; bw[0] = nd_w0(nb); bw[1] = nd_w1(nb); }
    for (int i = 0; i < 2; i++) {
      if (kind == JK_S) {
        uint64_t be;
        if (!cmul<MODE>(c, iwb, bw[i], &be)) return false;
        if (!combine_sum<MODE>(c, L - 1, aw[i], ac[i], be, bc[i], s[i])) return false;
      } else if (ia == ib) {
        // Both operands are the SAME node: (alpha : beta : gamma) of the child is aw_i * (wa : wb : wc),
        // i.e. the already normalised key, so the child item needs no arithmetic at all (the exact
        // algebra the reference reaches through its Cmul/Cdiv memo tables, QMDDcomplex.h:93).
        if (aw[i] == 0) { slot_zero(s[i]); continue; }
        if (!role_satisfied(role, i)) {
          uint64_t ga;
          if (!cmul<MODE>(c, iwc, aw[i], &ga)) return false;
          slot_pass(s[i], ac[i], ga);
        } else if (L - 1 >= g.low) {
          slot_ref(s[i], aw[i], ac[i], ac[i], iwa, iwb, iwc, tag);
        } else {
          uint64_t sum, w;
          if (!cadd<MODE>(c, iwa, iwb, &sum)) return false;
          if (!cmul<MODE>(c, aw[i], sum, &w)) return false;
          slot_pass(s[i], ac[i], w);
        }
      } else {
        uint64_t ga;
        if (!cmul<MODE>(c, iwc, sel ? bw[i] : aw[i], &ga)) return false;
        if (!role_satisfied(role, i)) { slot_pass(s[i], sel ? bc[i] : ac[i], ga); continue; }
        uint64_t al, be;
        if (!cmul<MODE>(c, iwa, aw[i], &al)) return false;
        if (!cmul<MODE>(c, iwb, bw[i], &be)) return false;
        if (!combine<MODE>(c, g, L - 1, al, ac[i], be, bc[i], ga, sel, s[i])) return false;
      }
    }
  }
  for (int i = 0; i < 2; i++) {
    if (s[i].kind == JC_REF) {
      bool ok = true;
      uint32_t j = dt_find_or_insert<MODE>(c, s[i].ia, s[i].ib, s[i].iwa, s[i].iwb, s[i].iwc, s[i].itag, &ok);
      if (!ok) return true;  // error / want_grid already flagged; the gate is abandoned
      s[i].target = j;
    }
    it->ch[i] = s[i].target; it->chw[i] = s[i].w; it->chk[i] = (uint8_t)s[i].kind;
  }
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
  return m1 > JKQ_DADD(m0, JKQ_DMUL(JKQ_DMUL(4.0, c.E->tol), sqrt(m1 > m0 ? m1 : m0)));
#endif
}

template <int MODE> JD bool reduce_item(Ctx &c, uint32_t idx) {
  JEng *E = c.E;
  JItem *it = c.items + idx;
  // each child edge is f_i * a_i: f_i = top weight of the child item's result (ONE for finished
  // edges), a_i = the factor / edge weight recorded at expansion
  uint32_t rn[2];
  uint64_t f[2], a[2];
  for (int i = 0; i < 2; i++) {
    uint32_t k = ld_s<MODE>(&it->chk[i]);
    uint32_t cht = ld_s<MODE>(&it->ch[i]);
    uint64_t chw = ld_s<MODE>(&it->chw[i]);
    if (k == JC_ZERO) { rn[i] = 0; f[i] = JW_ONE; a[i] = 0; }
    else if (k == JC_PASS) { rn[i] = cht; f[i] = JW_ONE; a[i] = chw; }
    else {
      const JItem *ci = c.items + cht;
      f[i] = ld_s<MODE>(&ci->rw);
      rn[i] = ld_s<MODE>(&ci->rnode);
      a[i] = f[i] ? chw : 0;
    }
    if (a[i] == 0) rn[i] = 0;
  }
  if (a[0] == 0 && a[1] == 0) { it->rnode = 0; it->rw = 0; return true; }
  uint32_t mx;
  uint64_t other, top;
  if (a[0] == 0 || a[1] == 0 || f[0] == f[1]) {
    // Common factor: (F a0, F a1) normalises exactly like (a0, a1) and F moves to the top weight.
    // For a product state or a basis-like state this removes ALL arithmetic from the levels above the
    // gate (a_max is ONE there); the reference gets the same effect from its complex op caches.
    mx = (a[0] == 0) ? 1u : (a[1] == 0 ? 0u : (mag_greater<MODE>(c, a[1], a[0]) ? 1u : 0u));
    if (!cdiv<MODE>(c, a[1 - mx], a[mx], &other)) return false;
    if (!cmul<MODE>(c, f[mx], a[mx], &top)) return false;
  } else {
    uint64_t rw[2];
    if (!cmul<MODE>(c, f[0], a[0], &rw[0])) return false;
    if (!cmul<MODE>(c, f[1], a[1], &rw[1])) return false;
    if (rw[0] == 0 && rw[1] == 0) { it->rnode = 0; it->rw = 0; return true; }
    mx = (rw[0] == 0) ? 1u : (rw[1] == 0 ? 0u : (mag_greater<MODE>(c, rw[1], rw[0]) ? 1u : 0u));
    if (!cdiv<MODE>(c, rw[1 - mx], rw[mx], &other)) return false;
    top = rw[mx];
    if (rw[1 - mx] == 0) rn[1 - mx] = 0;
  }
  if (top == 0) { it->rnode = 0; it->rw = 0; return true; }
  if (other == 0) rn[1 - mx] = 0;
  JNode rec;
  rec.c0 = rn[0] | (mx << 31); rec.c1 = rn[1]; rec.w = other;
  uint32_t id = ut_find_or_insert<MODE>(c, rec);
  it->rnode = id; it->rw = top;
  return true;
}

#endif  // JKQ_ENGINE_CUH
