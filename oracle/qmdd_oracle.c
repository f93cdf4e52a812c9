/* ORACLE -- test infrastructure, not product code.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build or run this file.
 *
 * CPU restatement, in plain C and FP64, of the reference's QMDD hot path
 * (qiskit-community/qiskit-jku-provider, /root/reference/src).  Same algorithm, same order of
 * operations, no MPFR: every real is an IEEE double, canonicalised through a tolerance table
 * exactly like the reference's 200-bit values are.  Each function cites the reference lines it
 * restates.  PARITY PINNING: this restatement is validated against the UNMODIFIED reference
 * engine compiled from /root/reference/src (oracle/build_ref.sh -> oracle/_ref/jku_ref) on
 * per-gate per-level node counts and final amplitudes (tests/test_oracle.py,
 * golden outputs of the reference committed under tests/golden/).
 *
 * Differences from the reference that are deliberate and documented:
 *   - reals are FP64 instead of 200-bit MPFR (arithmetic error 1e-16 << Ctol = 1e-10);
 *   - node "addresses" are allocation indices (slab emulation below) -- the reference orders the
 *     operands of QMDDadd by raw pointer value (QMDDpackage.cpp:1331);
 *   - shot sampling can use a counter-based Philox-4x32-10 stream (the product's RNG) instead of
 *     glibc rand(), so counts can be compared exactly with the CUDA sampler.
 *
 * Circuit text format: see oracle/ref_harness.cpp.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ complex value table */
/* Weight id = (R<<32)|I, each half = sign bit 31 + 31-bit index into a table of NON-NEGATIVE
 * reals (QMDDcomplexD.cpp:75-83,450).  Index 0 = 0.0, index 1 = 1.0 (:346-352). */
typedef uint64_t cw;
#define C_ZERO 0x0ull
#define C_ONE 0x0000000100000000ull
#define C_MONE 0x8000000100000000ull

static double Ctol = 1e-10; /* QMDDcomplexD.cpp:336 */
static int tie_rule = 0;    /* --tie-tol: NOT the reference's rule (see normalize) */
static long lossy_snaps_now = 0; /* Clookup hits of the current gate that identified two values more than 1e-13 apart */
static long near_ties_now = 0; /* normalisations of the current gate that met two DIFFERENT magnitudes closer than the
                                * snapping error (the reference decides those by 200-bit rounding noise, :565) */
static double *cval;        /* index -> value          (Ctable)  */
static uint8_t *calive;     /* erased by cleanCtable?            */
static uint32_t Ctentries, ccap;
static uint32_t clive;      /* Ctable.size()                     */

/* Ctable2: value -> index with comparator "|x-y| <= Ctol => equal" (QMDDcomplexD.cpp:85-98).
 * Because an insert only happens when no entry lies within Ctol, entries are pairwise more than
 * Ctol apart, std::map::find(x) returns the SMALLEST entry in [x-Ctol, x+Ctol] whatever the tree
 * shape, and a bucket of width Ctol holds at most one entry.  So: hash of floor(x/Ctol), probe
 * buckets k-1, k, k+1 in that order. */
typedef struct { int64_t key; uint32_t idx; uint8_t used; } VSlot;
static VSlot *vtab;
static uint64_t vmask;
static uint32_t vcount;

static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
static void vtab_put(int64_t key, uint32_t idx) {
  uint64_t h = mix64((uint64_t)key) & vmask;
  while (vtab[h].used) h = (h + 1) & vmask;
  vtab[h].used = 1; vtab[h].key = key; vtab[h].idx = idx;
  vcount++;
}
static void vtab_rebuild(uint64_t nslots) {
  free(vtab);
  vtab = (VSlot *)calloc(nslots, sizeof(VSlot));
  vmask = nslots - 1; vcount = 0;
  for (uint32_t i = 0; i < Ctentries; i++)
    if (calive[i]) vtab_put((int64_t)floor(cval[i] / Ctol), i);
}
static int vtab_get(int64_t key, uint32_t *idx) {
  uint64_t h = mix64((uint64_t)key) & vmask;
  while (vtab[h].used) {
    if (vtab[h].key == key) { *idx = vtab[h].idx; return 1; }
    h = (h + 1) & vmask;
  }
  return 0;
}
/* one real: QMDDcomplexD.cpp:402-418 (find, else append with the next index) */
static uint32_t real_lookup(double x) {
  int64_t k = (int64_t)floor(x / Ctol);
  uint32_t idx;
  for (int d = -1; d <= 1; d++)
    if (vtab_get(k + d, &idx) && fabs(cval[idx] - x) <= Ctol) {
      /* a DIFFERENT value (further away than any FP64 rounding of the same quantity) identified with an entry: from
       * here on the diagram depends on which values happened to be inserted first, i.e. on the exact sequence of
       * intermediate results of the reference's recursion */
      if (fabs(cval[idx] - x) > 1e-13 * (x > 1.0 ? x : 1.0)) lossy_snaps_now++;
      return idx;
    }
  if (Ctentries == ccap) {
    ccap *= 2;
    cval = (double *)realloc(cval, ccap * sizeof(double));
    calive = (uint8_t *)realloc(calive, ccap);
  }
  idx = Ctentries++;
  cval[idx] = x; calive[idx] = 1; clive++;
  if ((uint64_t)(vcount + 1) * 2 > vmask) vtab_rebuild((vmask + 1) * 2);
  else vtab_put(k, idx);
  return idx;
}
/* Clookup, QMDDcomplexD.cpp:372-466: signs split off, |part| canonicalised, a part that snaps to
 * index 0 loses its sign (:420,445). */
static cw Clookup(double re, double im) {
  uint32_t sr = signbit(re) ? 1 : 0, si = signbit(im) ? 1 : 0;
  re = fabs(re); im = fabs(im);
  if (re == 0.0) sr = 0;
  if (im == 0.0) si = 0;
  uint32_t r = real_lookup(re);
  if (sr && r) r |= 0x80000000u;
  uint32_t i = real_lookup(im);
  if (si && i) i |= 0x80000000u;
  return ((cw)r << 32) | (cw)i;
}
static inline double c_re(cw a) { double v = cval[(a >> 32) & 0x7FFFFFFFu]; return (a >> 63) ? -v : v; }
static inline double c_im(cw a) { double v = cval[a & 0x7FFFFFFFu]; return ((a >> 31) & 1) ? -v : v; }
/* Cmag (QMDDcomplexD.cpp:452-464) -- compared squared; sqrt is monotone */
static inline double c_mag2(cw a) {
  double r = cval[(a >> 32) & 0x7FFFFFFFu], i = cval[a & 0x7FFFFFFFu];
  return r * r + i * i;
}
static inline cw Cnegative(cw a) { /* QMDDcomplexD.cpp:483-496 */
  uint64_t r = a >> 32, i = a & 0xFFFFFFFFull;
  if (r) r ^= 0x80000000ull;
  if (i) i ^= 0x80000000ull;
  return (r << 32) | i;
}

/* memo tables cta/ctm/ctd (QMDDcomplex.h:93); cleared by cleanCtable (QMDDcomplexD.cpp:975-979) */
typedef struct { cw a, b, r; uint8_t used; } OpSlot;
typedef struct { OpSlot *s; uint64_t mask, count; } OpMap;
static OpMap cta, ctm, ctd;
static void op_init(OpMap *m, uint64_t n) { free(m->s); m->s = (OpSlot *)calloc(n, sizeof(OpSlot)); m->mask = n - 1; m->count = 0; }
static int op_get(OpMap *m, cw a, cw b, cw *r) {
  uint64_t h = mix64(a * 0x9E3779B97F4A7C15ull ^ mix64(b)) & m->mask;
  while (m->s[h].used) {
    if (m->s[h].a == a && m->s[h].b == b) { *r = m->s[h].r; return 1; }
    h = (h + 1) & m->mask;
  }
  return 0;
}
static void op_put(OpMap *m, cw a, cw b, cw r) {
  if ((m->count + 1) * 2 > m->mask) {
    OpMap old = *m;
    m->s = NULL; op_init(m, (old.mask + 1) * 2);
    for (uint64_t i = 0; i <= old.mask; i++)
      if (old.s[i].used) op_put(m, old.s[i].a, old.s[i].b, old.s[i].r);
    free(old.s);
  }
  uint64_t h = mix64(a * 0x9E3779B97F4A7C15ull ^ mix64(b)) & m->mask;
  while (m->s[h].used) {
    if (m->s[h].a == a && m->s[h].b == b) return; /* unordered_map::insert keeps the first */
    h = (h + 1) & m->mask;
  }
  m->s[h].used = 1; m->s[h].a = a; m->s[h].b = b; m->s[h].r = r; m->count++;
}

static cw Cadd(cw a, cw b) { /* QMDDcomplexD.cpp:498-550 */
  cw t;
  if (a == 0) return b;
  if (b == 0) return a;
  if (a == Cnegative(b)) return 0;
  if (op_get(&cta, a, b, &t)) return t;
  t = Clookup(c_re(a) + c_re(b), c_im(a) + c_im(b));
  op_put(&cta, a, b, t); op_put(&cta, b, a, t);
  return t;
}
static cw Cmul(cw a, cw b) { /* QMDDcomplexD.cpp:600-670 */
  cw t;
  if (a == C_ONE) return b;
  if (b == C_ONE) return a;
  if (a == 0 || b == 0) return 0;
  if (a == C_MONE) return Cnegative(b);
  if (b == C_MONE) return Cnegative(a);
  if (op_get(&ctm, a, b, &t)) return t;
  double ar = c_re(a), ai = c_im(a), br = c_re(b), bi = c_im(b);
  t = Clookup(ar * br - ai * bi, ar * bi + ai * br);
  op_put(&ctm, a, b, t); op_put(&ctm, b, a, t);
  return t;
}
static cw Cdiv(cw a, cw b) { /* QMDDcomplexD.cpp:693-777 */
  cw t;
  if (a == b) return C_ONE;
  if (a == 0) return 0;
  if (b == C_ONE) return a;
  if (b == C_MONE) return Cnegative(a);
  if (a == Cnegative(b)) return C_MONE;
  if (op_get(&ctd, a, b, &t)) return t;
  double ar = c_re(a), ai = c_im(a), br = c_re(b), bi = c_im(b);
  if ((b & 0x7FFFFFFFu) == 0) { /* mpfr_zero_p(b.i), :727-736 */
    t = Clookup(ar / br, ai / br);
  } else {
    double d = br * br + bi * bi;
    t = Clookup((ar * br + ai * bi) / d, (ai * br - ar * bi) / d);
  }
  op_put(&ctd, a, b, t);
  return t;
}

/* ------------------------------------------------------------------ nodes, unique table */
#define NULLP (-1)
#define TNODE 0
typedef struct { int32_t p; cw w; } Edge;
typedef struct {
  int32_t next;  /* unique-table chain / Avail chain */
  uint32_t ref;
  int16_t v;
  uint8_t ident, diag;
  Edge e[4];
} Node;

static Node *nodes;
static int64_t *naddr; /* emulated address rank of a node (see slab emulation) */
static int32_t ncap, nalloc, Avail = NULLP;
static int slab_descending = 1;
static int32_t nslabs;

#define MAXN 300
#define NBUCKET 32768
#define GCLIMIT1 25000
static int32_t *Unique; /* [MAXN][NBUCKET] heads */
static int64_t QMDDnodecount, QMDDpeaknodecount, UTlookups, UTmatch;
static int64_t CTlook_mult, CThit_mult, CTlook_add, CThit_add, Nop_mult, Nop_add;
static int Active[MAXN], ActiveNodeCount;
static const Edge QMDDzero = {TNODE, C_ZERO}, QMDDone = {TNODE, C_ONE};

/* QMDDgetNode, QMDDpackage.cpp:833-860: LIFO Avail chain refilled 2000 nodes at a time.  A slab
 * is one malloc of 2000*96 B > glibc's mmap threshold, and successive anonymous mmaps are placed
 * at DEcreasing addresses, so a later slab compares "smaller" (slab_descending). */
static int32_t getNode(void) {
  int32_t r;
  if (Avail != NULLP) { r = Avail; Avail = nodes[Avail].next; }
  else {
    if (nalloc + 2000 > ncap) {
      ncap = ncap * 2 + 2000;
      nodes = (Node *)realloc(nodes, (size_t)ncap * sizeof(Node));
      naddr = (int64_t *)realloc(naddr, (size_t)ncap * sizeof(int64_t));
    }
    int32_t base = nalloc;
    nalloc += 2000;
    int64_t abase = slab_descending ? -(int64_t)(nslabs + 1) * 4000 : (int64_t)nslabs * 4000;
    nslabs++;
    for (int i = 0; i < 2000; i++) naddr[base + i] = abase + i;
    r = base;
    Avail = base + 1;
    for (int i = 1; i < 1999; i++) nodes[base + i].next = base + i + 1;
    nodes[base + 1999].next = NULLP;
  }
  nodes[r].next = NULLP; nodes[r].ref = 0; nodes[r].ident = nodes[r].diag = 0;
  return r;
}

static inline int terminal(Edge e) { return e.p == TNODE; }

/* QMDDcheckSpecialMatrices, QMDDpackage.cpp:598-691 -- only block/diag/ident are consumed */
static void checkSpecial(int32_t p) {
  Node *n = &nodes[p];
  n->ident = 0; n->diag = 0;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++)
      if (n->e[i * 2 + j].p == NULLP || (i != j && n->e[i * 2 + j].w != C_ZERO)) return;
  n->diag = 1;
  for (int i = 0; i < 2; i++) {
    int j = 3 * i;
    n->diag = nodes[n->e[j].p].diag;
    if (terminal(n->e[j]) && n->e[j].w != C_ZERO && n->v != 0) n->diag = 0;
    if (!terminal(n->e[j]) && nodes[n->e[j].p].v != n->v - 1) n->diag = 0;
    if (!n->diag) return;
  }
  for (int i = 0; i < 2; i++) {
    int j = 3 * i;
    if (n->e[j].w != C_ONE || nodes[n->e[j].p].ident == 0) return;
  }
  n->ident = 1;
}

/* QMDDnormalize (NormC), QMDDpackage.cpp:532-595 */
static Edge normalize(Edge e) {
  Node *n = &nodes[e.p];
  int max = -1;
  double best = 0;
  e.w = C_ONE;
  for (int i = 0; i < 4; i++) {
    if (n->e[i].p == NULLP || n->e[i].w == C_ZERO) continue;
    double m = c_mag2(n->e[i].w);
    /* strict >, first index wins ties (:565).  With --tie-tol (study instrument, off by default) two
     * magnitudes closer than the snapping error of their canonical parts count as a tie, which is
     * what the CUDA engine does (DESIGN.md "normalisation ties"). */
    double near = 4.0 * Ctol * sqrt(m > best ? m : best);
    double slack = tie_rule ? near : 0.0;
    if (max != -1 && m != best && fabs(m - best) <= near) near_ties_now++;
    if (max == -1 || m > best + slack) { best = m; max = i; }
  }
  if (max == -1) { e.w = C_ZERO; return e; }
  if (n->e[max].w == C_ONE) return e;
  e.w = n->e[max].w;
  for (int j = 0; j < 4; j++) {
    if (j == max) n->e[j].w = C_ONE;
    else if (n->e[j].p != NULLP && n->e[j].w != C_ZERO) n->e[j].w = Cdiv(n->e[j].w, e.w);
  }
  return e;
}

/* QMDDutLookup, QMDDpackage.cpp:693-763 (hash-cons; exact compare of the 4 edges) */
static Edge utLookup(Edge e) {
  if (terminal(e)) return e;
  UTlookups++;
  Node *n = &nodes[e.p];
  uint64_t key = 0;
  for (int i = 0; i < 4; i++) key += ((uint64_t)(uint32_t)n->e[i].p >> i) + (n->e[i].w >> 32) + n->e[i].w;
  key = mix64(key) & (NBUCKET - 1);
  int32_t *head = &Unique[(size_t)n->v * NBUCKET + key];
  for (int32_t p = *head; p != NULLP; p = nodes[p].next) {
    if (memcmp(n->e, nodes[p].e, sizeof(n->e)) == 0) {
      n->next = Avail; Avail = e.p; /* fresh node back on the free list (:726-727) */
      UTmatch++;
      e.p = p;
      return e;
    }
  }
  n->next = *head; *head = e.p;
  QMDDnodecount++;
  if (QMDDnodecount > QMDDpeaknodecount) QMDDpeaknodecount = QMDDnodecount;
  checkSpecial(e.p);
  return e;
}

/* QMDDmakeNonterminal, QMDDpackage.cpp:1204-1233 */
static Edge makeNonterminal(int v, Edge edge[4]) {
  int redundant = 1;
  Edge e = edge[0];
  for (int i = 1; redundant && i < 4; i++)
    redundant = (edge[i].p == NULLP) || (edge[i].w == e.w && edge[i].p == e.p);
  if (redundant) return edge[0];
  e.p = getNode();
  e.w = C_ONE;
  nodes[e.p].v = (int16_t)v;
  /* struct padding must compare equal under memcmp */
  memset(nodes[e.p].e, 0, sizeof(nodes[e.p].e));
  for (int i = 0; i < 4; i++) { nodes[e.p].e[i].p = edge[i].p; nodes[e.p].e[i].w = edge[i].w; }
  e = normalize(e);
  e = utLookup(e);
  return e;
}

/* compute tables, QMDDpackage.cpp:986-1072 (exact memo, unbounded until GC / cleanCtable) */
typedef struct { int32_t ap, bp; cw aw, bw; Edge r; uint8_t used; } CTSlot;
typedef struct { CTSlot *s; uint64_t mask, count; } CTMap;
static CTMap CT_add, CT_mult;
static void ct_init(CTMap *m, uint64_t n) { free(m->s); m->s = (CTSlot *)calloc(n, sizeof(CTSlot)); m->mask = n - 1; m->count = 0; }
static uint64_t ct_hash(Edge a, Edge b) { return mix64(((uint64_t)(uint32_t)a.p << 32 | (uint32_t)b.p) ^ mix64(a.w) ^ mix64(b.w * 3)); }
static CTSlot *ct_find(CTMap *m, Edge a, Edge b) {
  uint64_t h = ct_hash(a, b) & m->mask;
  while (m->s[h].used) {
    CTSlot *s = &m->s[h];
    if (s->ap == a.p && s->bp == b.p && s->aw == a.w && s->bw == b.w) return s;
    h = (h + 1) & m->mask;
  }
  return &m->s[h];
}
static void ct_insert(CTMap *m, Edge a, Edge b, Edge r) {
  if ((m->count + 1) * 2 > m->mask) {
    CTMap old = *m;
    m->s = NULL; ct_init(m, (old.mask + 1) * 2);
    for (uint64_t i = 0; i <= old.mask; i++)
      if (old.s[i].used) {
        Edge oa = {old.s[i].ap, old.s[i].aw}, ob = {old.s[i].bp, old.s[i].bw};
        ct_insert(m, oa, ob, old.s[i].r);
      }
    free(old.s);
  }
  CTSlot *s = ct_find(m, a, b);
  if (!s->used) { s->used = 1; s->ap = a.p; s->bp = b.p; s->aw = a.w; s->bw = b.w; m->count++; }
  s->r = r;
}

static Edge QMDDid[MAXN];
static void initComputeTable(void) { /* QMDDpackage.cpp:765-789 */
  ct_init(&CT_add, 1 << 12); ct_init(&CT_mult, 1 << 12);
  for (int i = 0; i < MAXN; i++) QMDDid[i].p = NULLP;
}

/* QMDDadd, QMDDpackage.cpp:1319-1414 (MultMode == 0) */
static Edge QMDDadd(Edge x, Edge y) {
  Edge e1, e2, e[4], r;
  if (x.p == NULLP) return y;
  if (y.p == NULLP) return x;
  Nop_add++;
  if (terminal(y) || naddr[x.p] > naddr[y.p]) { e1 = x; x = y; y = e1; }
  if (x.w == C_ZERO) return y;
  if (y.w == C_ZERO) return x;
  if (x.p == y.p) {
    r = y;
    r.w = Cadd(x.w, y.w);
    if (r.w == C_ZERO) r = QMDDzero;
    return r;
  }
  cw xweight = x.w;
  x.w = C_ONE;
  y.w = Cdiv(y.w, xweight);
  CTlook_add++;
  CTSlot *s = ct_find(&CT_add, x, y);
  if (s->used) {
    CThit_add++;
    r = s->r;
    r.w = Cmul(r.w, xweight);
    return r;
  }
  int w;
  if (terminal(x)) w = nodes[y.p].v;
  else {
    w = nodes[x.p].v;
    if (!terminal(y) && nodes[y.p].v > w) w = nodes[y.p].v;
  }
  for (int i = 0; i < 4; i++) {
    if (!terminal(x) && nodes[x.p].v == w) {
      e1 = nodes[x.p].e[i];
      e1.w = Cmul(e1.w, x.w);
    } else {
      e1 = x;
      if (nodes[y.p].e[i].p == NULLP) { e1.p = NULLP; e1.w = C_ONE; }
    }
    if (!terminal(y) && nodes[y.p].v == w) {
      e2 = nodes[y.p].e[i];
      e2.w = Cmul(e2.w, y.w);
    } else {
      e2 = y;
      if (nodes[x.p].e[i].p == NULLP) { e2.p = NULLP; e2.w = C_ONE; }
    }
    e[i] = QMDDadd(e1, e2);
  }
  r = makeNonterminal(w, e);
  ct_insert(&CT_add, x, y, r);
  r.w = Cmul(r.w, xweight);
  return r;
}

/* QMDDmultiply2, QMDDpackage.cpp:1416-1517 */
static Edge QMDDmultiply2(Edge x, Edge y, int var) {
  Edge e1, e2, e[4], r;
  if (x.p == NULLP) return x;
  if (y.p == NULLP) return y;
  Nop_mult++;
  if (x.w == C_ZERO || y.w == C_ZERO) return QMDDzero;
  if (var == 0) { r.p = TNODE; r.w = Cmul(x.w, y.w); return r; }
  cw xweight = x.w, yweight = y.w;
  x.w = C_ONE; y.w = C_ONE;
  CTlook_mult++;
  CTSlot *s = ct_find(&CT_mult, x, y);
  if (s->used) {
    CThit_mult++;
    r = s->r;
    r.w = Cmul(r.w, xweight);
    r.w = Cmul(r.w, yweight);
    return r;
  }
  int w = var - 1;
  if (nodes[x.p].v == w && nodes[x.p].v == nodes[y.p].v) {
    if (nodes[x.p].ident) {
      r = y;
      ct_insert(&CT_mult, x, y, r);
      r.w = Cmul(xweight, yweight);
      return r;
    }
    if (nodes[y.p].ident) {
      r = x;
      ct_insert(&CT_mult, x, y, r);
      r.w = Cmul(xweight, yweight);
      return r;
    }
  }
  for (int i = 0; i < 4; i += 2)
    for (int j = 0; j < 2; j++) {
      e[i + j].p = NULLP; e[i + j].w = C_ZERO;
      for (int k = 0; k < 2; k++) {
        if (!terminal(x) && nodes[x.p].v == w) { e1 = nodes[x.p].e[i + k]; e1.w = Cmul(e1.w, x.w); }
        else e1 = x;
        if (!terminal(y) && nodes[y.p].v == w) { e2 = nodes[y.p].e[j + 2 * k]; e2.w = Cmul(e2.w, y.w); }
        else e2 = y;
        e[i + j] = QMDDadd(e[i + j], QMDDmultiply2(e1, e2, var - 1));
      }
    }
  r = makeNonterminal(w, e);
  ct_insert(&CT_mult, x, y, r);
  r.w = Cmul(r.w, xweight);
  r.w = Cmul(r.w, yweight);
  return r;
}
static Edge QMDDmultiply(Edge x, Edge y) { /* QMDDpackage.cpp:1519-1529 */
  int var = 0;
  if (!terminal(x) && nodes[x.p].v + 1 > var) var = nodes[x.p].v + 1;
  if (!terminal(y) && nodes[y.p].v + 1 > var) var = nodes[y.p].v + 1;
  return QMDDmultiply2(x, y, var);
}

/* QMDDincref / QMDDdecref, QMDDpackage.cpp:862-947 */
static void incref(Edge e) {
  if (terminal(e)) return;
  Node *n = &nodes[e.p];
  n->ref++;
  if (n->ref == 1) {
    for (int i = 0; i < 4; i++)
      if (nodes[e.p].e[i].p != NULLP) incref(nodes[e.p].e[i]);
    Active[nodes[e.p].v]++;
    ActiveNodeCount++;
  }
}
static void decref(Edge e) {
  if (terminal(e)) return;
  Node *n = &nodes[e.p];
  if (n->ref == 0) { fprintf(stderr, "error in decref\n"); exit(8); }
  n->ref--;
  if (n->ref == 0) {
    for (int i = 0; i < 4; i++)
      if (nodes[e.p].e[i].p != NULLP) decref(nodes[e.p].e[i]);
    Active[nodes[e.p].v]--;
    ActiveNodeCount--;
  }
}
/* QMDDgarbageCollect, QMDDpackage.cpp:791-831 */
static int GCcurrentLimit = GCLIMIT1;
static int64_t gc_runs;
static void garbageCollect(void) {
  if (QMDDnodecount < GCcurrentLimit) return;
  int64_t counta = 0;
  gc_runs++;
  for (int i = 0; i < MAXN; i++)
    for (int j = 0; j < NBUCKET; j++) {
      int32_t lastp = NULLP, p = Unique[(size_t)i * NBUCKET + j];
      while (p != NULLP) {
        int32_t nextp = nodes[p].next;
        if (nodes[p].ref == 0) {
          if (lastp == NULLP) Unique[(size_t)i * NBUCKET + j] = nextp;
          else nodes[lastp].next = nextp;
          nodes[p].next = Avail; Avail = p;
        } else { lastp = p; counta++; }
        p = nextp;
      }
    }
  QMDDnodecount = counta;
  initComputeTable();
}

/* QMDDident, QMDDpackage.cpp:1701-1744 */
static Edge QMDDident(int x, int y) {
  Edge e, f, edge[4];
  if (y < 0) return QMDDone;
  if (x == 0 && QMDDid[y].p != NULLP) return QMDDid[y];
  if (y >= 1 && (f = QMDDid[y - 1]).p != NULLP) {
    edge[0] = f; edge[1] = QMDDzero; edge[2] = QMDDzero; edge[3] = f;
    e = makeNonterminal(y, edge);
    QMDDid[y] = e;
    return e;
  }
  edge[0] = QMDDone; edge[1] = QMDDzero; edge[2] = QMDDzero; edge[3] = QMDDone;
  e = makeNonterminal(x, edge);
  for (int k = x + 1; k <= y; k++) {
    edge[0] = e; edge[1] = QMDDzero; edge[2] = QMDDzero; edge[3] = e;
    e = makeNonterminal(k, edge);
  }
  if (x == 0) QMDDid[y] = e;
  return e;
}

/* QMDDmvlgate, QMDDpackage.cpp:1746-1828.  line[v]: -1 idle, 0/1 control value, 2 target */
static Edge QMDDmvlgate(cw mat[2][2], int n, const int line[]) {
  Edge e, f, em[4], fm[4], temp;
  int z;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++) { em[i * 2 + j].p = TNODE; em[i * 2 + j].w = mat[i][j]; }
  e = QMDDone;
  for (z = 0; line[z] < 2; z++) {
    if (line[z] >= 0) {
      for (int i1 = 0; i1 < 2; i1++)
        for (int i2 = 0; i2 < 2; i2++) {
          int i = i1 * 2 + i2;
          f = (i1 == i2) ? e : QMDDzero;
          for (int k = 0; k < 2; k++)
            for (int j = 0; j < 2; j++) {
              int t = k * 2 + j;
              if (k == j) fm[t] = (k == line[z]) ? em[i] : f;
              else fm[t] = QMDDzero;
            }
          em[i] = makeNonterminal(z, fm);
        }
    } else {
      for (int i = 0; i < 4; i++) {
        fm[0] = em[i]; fm[1] = QMDDzero; fm[2] = QMDDzero; fm[3] = em[i];
        em[i] = makeNonterminal(z, fm);
      }
    }
    e = QMDDident(0, z);
  }
  e = makeNonterminal(z, em);
  for (z++; z < n; z++) {
    if (line[z] >= 0) {
      temp = QMDDident(0, z - 1);
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) {
          if (i == j) em[i * 2 + j] = (i == line[z]) ? e : temp;
          else em[i * 2 + j] = QMDDzero;
        }
      e = makeNonterminal(z, em);
    } else {
      fm[0] = e; fm[1] = QMDDzero; fm[2] = QMDDzero; fm[3] = e;
      e = makeNonterminal(z, fm);
    }
  }
  return e;
}

/* ------------------------------------------------------------------ simulator core */
static Edge circ_e;
static int nqubits;
static uint32_t complex_limit = 10000; /* Simulator.h:91 */
static cw Hm00, Hm11, Vm00, Vm01;
static int64_t clean_runs;

/* cleanCtable + addToComplexTable, QMDDcomplexD.cpp:843-980 */
static uint8_t *keep_mark;
static uint8_t *visit_mark;
static void mark_w(cw w) {
  keep_mark[(w >> 32) & 0x7FFFFFFFu] = 1;
  keep_mark[w & 0x7FFFFFFFu] = 1;
}
static void addToComplexTable(Edge e) {
  if (terminal(e) || visit_mark[e.p]) return;
  visit_mark[e.p] = 1;
  for (int i = 0; i < 4; i++) mark_w(nodes[e.p].e[i].w);
  for (int i = 0; i < 4; i++) addToComplexTable(nodes[e.p].e[i]);
}
static void cleanCtable(Edge *roots, int nroots) {
  clean_runs++;
  keep_mark = (uint8_t *)calloc(Ctentries, 1);
  visit_mark = (uint8_t *)calloc(nalloc, 1);
  for (int i = 0; i < nroots; i++) { mark_w(roots[i].w); addToComplexTable(roots[i]); }
  keep_mark[0] = keep_mark[1] = 1;
  mark_w(Vm00); mark_w(Vm01); mark_w(Hm00); mark_w(Hm11);
  for (uint32_t i = 0; i < Ctentries; i++)
    if (calive[i] && !keep_mark[i]) { calive[i] = 0; clive--; }
  free(keep_mark); free(visit_mark);
  vtab_rebuild(vmask + 1);
  initComputeTable();
  op_init(&cta, 1 << 12); op_init(&ctm, 1 << 12); op_init(&ctd, 1 << 12);
}

/* Simulator::ApplyGate, Simulator.cpp:466-503 (beforeMeasurement == circ.e before any measure) */
static int max_active;
static void ApplyGate(Edge gate) {
  Edge tmp = QMDDmultiply(gate, circ_e);
  incref(tmp);
  decref(circ_e);
  circ_e = tmp;
  garbageCollect();
  if (ActiveNodeCount > max_active) max_active = ActiveNodeCount;
  if (clive > complex_limit) {
    Edge v[1] = {circ_e};
    cleanCtable(v, 1);
    if (complex_limit < 2 * clive) complex_limit *= 2;
  }
}

/* Simulator::AddVariables, Simulator.cpp:63-96 (fresh register only) */
static void AddVariables(int add) {
  Edge f = QMDDone, edges[4];
  edges[1] = edges[2] = edges[3] = QMDDzero;
  for (int p = 0; p < add; p++) { edges[0] = f; f = makeNonterminal(p, edges); }
  incref(f);
  circ_e = f;
  nqubits = add;
}

/* QMDDinit + QMDDcomplexInit + QMDDinitGateMatrices, QMDDpackage.cpp:1248-1317,37-72,
 * QMDDcomplexD.cpp:322-353.  Qmake uses DOUBLE sqrt(2) (QMDDcomplexD.cpp:297-306). */
static cw Nm[2][2], Zm[2][2], Hm[2][2], Sm[2][2];
static void QMDDinit(void) {
  ccap = 1 << 12;
  cval = (double *)malloc(ccap * sizeof(double));
  calive = (uint8_t *)malloc(ccap);
  Ctentries = 0; clive = 0;
  vtab = NULL; vtab_rebuild(1 << 13);
  op_init(&cta, 1 << 12); op_init(&ctm, 1 << 12); op_init(&ctd, 1 << 12);
  Clookup(0.0, 0.0);
  Clookup(1.0, 0.0);
  initComputeTable();
  Unique = (int32_t *)malloc((size_t)MAXN * NBUCKET * sizeof(int32_t));
  for (size_t i = 0; i < (size_t)MAXN * NBUCKET; i++) Unique[i] = NULLP;
  int32_t t = getNode(); /* terminal = first node of the first slab */
  nodes[t].ident = 1; nodes[t].diag = 1; nodes[t].v = -1;
  for (int i = 0; i < 4; i++) { nodes[t].e[i].p = NULLP; nodes[t].e[i].w = C_ZERO; }
  /* gate constants in the reference's insertion order */
  Vm00 = Clookup(0.5, 0.5);
  Vm01 = Clookup(0.5, -0.5);
  double h = (0 + 1 * sqrt(2.0)) / 2;
  Hm00 = Clookup(h, 0.0);
  Hm11 = Clookup((0 + -1 * sqrt(2.0)) / 2, 0.0);
  Nm[0][0] = Nm[1][1] = C_ZERO; Nm[0][1] = Nm[1][0] = C_ONE;
  Hm[0][0] = Hm[0][1] = Hm[1][0] = Hm00; Hm[1][1] = Hm11;
  Zm[0][0] = C_ONE; Zm[0][1] = Zm[1][0] = C_ZERO; Zm[1][1] = C_MONE;
  Sm[0][0] = C_ONE; Sm[0][1] = Sm[1][0] = C_ZERO; Sm[1][1] = Clookup(0.0, 1.0);
}

/* Simulator::GetElementOfVector, Simulator.cpp:415-431 */
static cw GetElementOfVector(uint64_t element) {
  Edge e = circ_e;
  if (terminal(e)) return 0;
  cw l = C_ONE;
  do {
    l = Cmul(l, e.w);
    uint64_t b = (element >> nodes[e.p].v) & 1;
    e = nodes[e.p].e[2 * b];
  } while (!terminal(e));
  return Cmul(l, e.w);
}

/* Simulator::AssignProbs, Simulator.cpp:127-146 (memoised bottom-up squared norm) */
static double *probs;
static uint8_t *probs_set;
static double AssignProbs(Edge e) {
  if (probs_set[e.p]) return c_mag2(e.w) * probs[e.p];
  double sum;
  if (terminal(e)) sum = 1.0;
  else sum = AssignProbs(nodes[e.p].e[0]) + AssignProbs(nodes[e.p].e[2]);
  probs[e.p] = sum; probs_set[e.p] = 1;
  return c_mag2(e.w) * sum;
}

/* Philox-4x32-10 (Salmon et al., SC'11) -- the product's counter-based stream:
 * key = (seed_lo, seed_hi), counter = (shot_lo, shot_hi, draw_block, 0); word j of block b feeds
 * the (4b+j)-th level of the walk, counted from the root. */
static void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

/* Simulator::MeasureAll(false), Simulator.cpp:148-224 -- probs memo kept across shots (the
 * reference recomputes it every shot, same values); returns key with bit q = qubit q */
static uint64_t MeasureAll(int use_philox, uint64_t seed, uint64_t shot) {
  Edge cur = circ_e;
  uint64_t key = 0;
  uint32_t c[4] = {0, 0, 0, 0};
  int step = 0;
  for (int i = nodes[circ_e.p].v; i >= 0; --i, ++step) {
    Node *n = &nodes[cur.p];
    double p0 = probs[n->e[0].p] * c_mag2(n->e[0].w);
    double p1 = probs[n->e[2].p] * c_mag2(n->e[2].w);
    double tmp = p0 + p1;
    p0 /= tmp;
    double u;
    if (use_philox) {
      if ((step & 3) == 0) {
        c[0] = (uint32_t)shot; c[1] = (uint32_t)(shot >> 32); c[2] = (uint32_t)(step >> 2); c[3] = 0;
        philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
      }
      u = ((double)c[step & 3] + 0.5) * (1.0 / 4294967296.0);
    } else {
      u = (double)rand() / RAND_MAX;
    }
    int bit = !(u < p0);
    if (bit) key |= 1ull << (nqubits - 1 - n->v);
    cur = n->e[2 * bit];
  }
  return key;
}

/* ------------------------------------------------------------------ circuit text + driver */
static double parse_angle(const char *tok) {
  const char *p = strstr(tok, "pi");
  if (!p) return atof(tok);
  double v = 3.14159265358979323846;
  if (tok[0] == '-') v = -v;
  const char *star = strchr(tok, '*'), *slash = strchr(tok, '/');
  if (star) v = v * atof(star + 1);
  if (slash) v = v / atof(slash + 1);
  return v;
}

typedef struct { uint64_t key; uint64_t cnt; } HEnt;
static int hent_cmp(const void *a, const void *b) {
  uint64_t x = ((const HEnt *)a)->key, y = ((const HEnt *)b)->key;
  return x < y ? -1 : x > y;
}

/* debugging aid: the live nodes of one level of the final state, on stderr */
static int dump_level = -1;
static void dump_nodes(Edge e, uint8_t *seen) {
  if (e.p == NULLP || e.p == TNODE || seen[e.p]) return;
  seen[e.p] = 1;
  Node *nd = &nodes[e.p];
  if (nd->v == dump_level)
    fprintf(stderr, "node %d: e0 -> %d w=(%.17g, %.17g)  e2 -> %d w=(%.17g, %.17g)\n", e.p, nd->e[0].p, c_re(nd->e[0].w), c_im(nd->e[0].w),
            nd->e[2].p, c_re(nd->e[2].w), c_im(nd->e[2].w));
  for (int i = 0; i < 4; i++) dump_nodes(nd->e[i], seen);
}

int main(int argc, char **argv) {
  const char *circuit = NULL;
  long shots = 0, amp_limit = 0, shot_offset = 0;
  int levels = 0, amps = 0, use_philox = 1, dump_ctable = 0;
  uint64_t seed = 1;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--circuit") && i + 1 < argc) circuit = argv[++i];
    else if (!strcmp(argv[i], "--shots") && i + 1 < argc) shots = atol(argv[++i]);
    else if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = strtoull(argv[++i], NULL, 10);
    else if (!strcmp(argv[i], "--shot-offset") && i + 1 < argc) shot_offset = atol(argv[++i]);
    else if (!strcmp(argv[i], "--levels")) levels = 1;
    else if (!strcmp(argv[i], "--amps")) amps = 1;
    else if (!strcmp(argv[i], "--amp-limit") && i + 1 < argc) amp_limit = atol(argv[++i]);
    else if (!strcmp(argv[i], "--rng") && i + 1 < argc) use_philox = strcmp(argv[++i], "rand") != 0;
    else if (!strcmp(argv[i], "--precision") && i + 1 < argc) Ctol = atof(argv[++i]);
    else if (!strcmp(argv[i], "--slab-ascending")) slab_descending = 0;
    else if (!strcmp(argv[i], "--dump-ctable")) dump_ctable = 1;
    else if (!strcmp(argv[i], "--tie-tol")) tie_rule = 1;
    else if (!strcmp(argv[i], "--dump-level") && i + 1 < argc) dump_level = atoi(argv[++i]);
  }
  FILE *f = circuit ? fopen(circuit, "r") : NULL;
  if (!f) { fprintf(stderr, "usage: qmdd_oracle --circuit FILE [--shots N] [--seed S] [--levels] [--amps] [--rng rand|philox]\n"); return 2; }
  QMDDinit();
  srand((unsigned)seed);
  int n = 0, line[MAXN];
  for (int i = 0; i < MAXN; i++) line[i] = -1;
  char w[64], a1[64], a2[64], a3[64];
  size_t applies = 0, lvcap = 0, lvn = 0, ntcap = 0;
  int *lv = NULL;
  long *nt = NULL, *ls = NULL;
  int peak_active = 0;
  struct timespec t0, t1;
  double t_apply = 0;
  while (fscanf(f, "%63s", w) == 1) {
    if (w[0] == '#') { int ch; while ((ch = fgetc(f)) != EOF && ch != '\n') {} continue; }
    if (!strcmp(w, "qubits")) {
      if (fscanf(f, "%d", &n) != 1) return 2;
      AddVariables(n);
      peak_active = ActiveNodeCount;
      continue;
    }
    cw m[2][2];
    int target = -1, nctl = 0, ctl[MAXN], cval_[MAXN];
    int is_u = 0;
    if (!strcmp(w, "u") || !strcmp(w, "mcu")) {
      if (fscanf(f, "%d %63s %63s %63s", &target, a1, a2, a3) != 4) return 2;
      is_u = 1;
    } else if (!strcmp(w, "cx")) {
      if (fscanf(f, "%d %d", &ctl[0], &target) != 2) return 2;
      cval_[0] = 1; nctl = 1;
      memcpy(m, Nm, sizeof(m));
    } else if (!strcmp(w, "mcg")) {
      if (fscanf(f, "%63s %d", a1, &target) != 2) return 2;
      memcpy(m, a1[0] == 'z' ? Zm : a1[0] == 'h' ? Hm : a1[0] == 's' ? Sm : Nm, sizeof(m));
    } else { fprintf(stderr, "unknown op %s\n", w); return 2; }
    if (!strcmp(w, "mcu") || !strcmp(w, "mcg")) {
      if (fscanf(f, "%d", &nctl) != 1) return 2;
      for (int i = 0; i < nctl; i++) {
        char c[32];
        if (fscanf(f, "%31s", c) != 1) return 2;
        if (c[0] == '~') { ctl[i] = atoi(c + 1); cval_[i] = 0; }
        else { ctl[i] = atoi(c); cval_[i] = 1; }
      }
    }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (is_u) { /* U(theta,phi,lambda), QASMsimulator.cpp:295-298 */
      double th = parse_angle(a1), ph = parse_angle(a2), la = parse_angle(a3);
      m[0][0] = Clookup(cos(-(ph + la) / 2) * cos(th / 2), sin(-(ph + la) / 2) * cos(th / 2));
      m[0][1] = Clookup(-cos(-(ph - la) / 2) * sin(th / 2), -sin(-(ph - la) / 2) * sin(th / 2));
      m[1][0] = Clookup(cos((ph - la) / 2) * sin(th / 2), sin((ph - la) / 2) * sin(th / 2));
      m[1][1] = Clookup(cos((ph + la) / 2) * cos(th / 2), sin((ph + la) / 2) * cos(th / 2));
    }
    for (int i = 0; i < nctl; i++) line[n - 1 - ctl[i]] = cval_[i];
    line[n - 1 - target] = 2;
    Edge g = QMDDmvlgate(m, n, line);
    for (int i = 0; i < nctl; i++) line[n - 1 - ctl[i]] = -1;
    line[n - 1 - target] = -1;
    near_ties_now = 0;
    if (!is_u) lossy_snaps_now = 0;  /* (a U gate's matrix entries were looked up above: they count for this gate) */
    ApplyGate(g);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (levels) {
      if (applies + 1 > ntcap) {
        ntcap = ntcap * 2 + 4096;
        nt = (long *)realloc(nt, ntcap * sizeof(long));
        ls = (long *)realloc(ls, ntcap * sizeof(long));
      }
      nt[applies] = near_ties_now;
      ls[applies] = lossy_snaps_now;
    }
    lossy_snaps_now = 0;
    t_apply += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    applies++;
    if (ActiveNodeCount > peak_active) peak_active = ActiveNodeCount;
    if (levels) {
      if (lvn + n > lvcap) { lvcap = lvcap * 2 + 4096; lv = (int *)realloc(lv, lvcap * sizeof(int)); }
      for (int v = 0; v < n; v++) lv[lvn++] = Active[v];
    }
  }
  fclose(f);
  if (dump_level >= 0) {
    uint8_t *seen = (uint8_t *)calloc(nalloc, 1);
    fprintf(stderr, "root weight (%.17g, %.17g)\n", c_re(circ_e.w), c_im(circ_e.w));
    dump_nodes(circ_e, seen);
    free(seen);
  }
  printf("{\n \"n\": %d,\n \"applies\": %zu,\n \"t_apply_s\": %.9g,\n", n, applies, t_apply);
  printf(" \"nodes_created\": %ld,\n \"ut_lookups\": %ld,\n", (long)(UTlookups - UTmatch), (long)UTlookups);
  printf(" \"ct_look_mult\": %ld, \"ct_hit_mult\": %ld, \"ct_look_add\": %ld, \"ct_hit_add\": %ld,\n",
         (long)CTlook_mult, (long)CThit_mult, (long)CTlook_add, (long)CThit_add);
  printf(" \"nop_mult\": %ld, \"nop_add\": %ld,\n", (long)Nop_mult, (long)Nop_add);
  printf(" \"peak_active\": %d,\n \"ctable_size\": %u,\n \"peak_table_nodes\": %ld,\n \"gc_runs\": %ld, \"clean_runs\": %ld,\n",
         peak_active, clive, (long)QMDDpeaknodecount, (long)gc_runs, (long)clean_runs);
  printf(" \"level_counts\": [");
  for (int v = 0; v < n; v++) printf("%s%d", v ? "," : "", Active[v]);
  printf("]");
  if (levels) {
    printf(",\n \"level_counts_per_gate\": [");
    for (size_t g = 0; g < applies; g++) {
      printf("%s[", g ? "," : "");
      for (int v = 0; v < n; v++) printf("%s%d", v ? "," : "", lv[g * n + v]);
      printf("]");
    }
    printf("]");
    printf(",\n \"near_ties_per_gate\": [");
    for (size_t g = 0; g < applies; g++) printf("%s%ld", g ? "," : "", nt[g]);
    printf("]");
    printf(",\n \"lossy_snaps_per_gate\": [");
    for (size_t g = 0; g < applies; g++) printf("%s%ld", g ? "," : "", ls[g]);
    printf("]");
  }
  if (amps) {
    uint64_t len = 1ull << n;
    if (amp_limit > 0 && (uint64_t)amp_limit < len) len = amp_limit;
    printf(",\n \"amps\": [");
    for (uint64_t i = 0; i < len; i++) {
      uint64_t entry = 0;
      for (int q = 0; q < n; q++) entry |= ((i >> q) & 1ull) << (n - 1 - q);
      cw a = GetElementOfVector(entry);
      printf("%s[%.17g,%.17g]", i ? "," : "", c_re(a), c_im(a));
    }
    printf("]");
  }
  if (dump_ctable) {
    printf(",\n \"ctable\": [");
    int first = 1;
    for (uint32_t i = 0; i < Ctentries; i++)
      if (calive[i]) { printf("%s[%u,%.17g]", first ? "" : ",", i, cval[i]); first = 0; }
    printf("]");
  }
  if (shots > 0) {
    clock_gettime(CLOCK_MONOTONIC, &t0);
    probs = (double *)calloc(nalloc, sizeof(double));
    probs_set = (uint8_t *)calloc(nalloc, 1);
    double p = AssignProbs(circ_e);
    if (fabs(p - 1) > 0.01) fprintf(stderr, "WARNING in measurement: |alpha|^2 + |beta|^2 = %g\n", p);
    HEnt *all = (HEnt *)malloc((size_t)shots * sizeof(HEnt));
    for (long i = 0; i < shots; i++) { all[i].key = MeasureAll(use_philox, seed, (uint64_t)(shot_offset + i)); all[i].cnt = 1; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    qsort(all, shots, sizeof(HEnt), hent_cmp);
    printf(",\n \"shots\": %ld,\n \"t_sample_s\": %.9g,\n \"norm\": %.17g,\n \"counts\": {", shots,
           (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec), p);
    int first = 1;
    for (long i = 0; i < shots;) {
      long j = i;
      while (j < shots && all[j].key == all[i].key) j++;
      char s[65];
      for (int q = 0; q < n; q++) s[q] = (char)('0' + ((all[i].key >> q) & 1)); /* char q = qubit q */
      s[n] = 0;
      printf("%s\"%s\": %ld", first ? "" : ", ", s, j - i);
      first = 0;
      i = j;
    }
    printf("}");
  }
  printf("\n}\n");
  return 0;
}
