// MODE_GRID item steps with the independent memory accesses of an item in flight together.
//
// On the grid path every table lives in HBM / L2 (ld.global.cg, ~290 cycles a round trip on B200, ~760 for an
// atomic, ~800 for a __threadfence) and one CTA of 256 threads per SM leaves two warps per scheduler: an item is a
// chain of dependent round trips, and both the narrow levels (where 38 k threads wait for the chain of ONE item) and
// the wide ones (8 warps per SM cannot hide 300 cycles) cost what that chain costs.  expand_item / reduce_item of
// jkq_engine.cuh walk it one access at a time -- per complex operation: operand values, value-table probes of the real
// part, probes of the imaginary part; three such operations per child slot and stage, two child slots, two de-dup
// inserts with an allocation, a fence and a CAS each: ~45 round trips for a sum item.  Here the same algorithm
// (QMDDmultiply2 / QMDDadd pair expansion, QMDDpackage.cpp:1319-1517; QMDDmakeNonterminal, :1204-1233) is staged:
//   * the complex operations of BOTH child slots that do not depend on each other form one batch (cx_batch): all
//     operand values are loaded together, then the probes of every part that needs the value table (three buckets
//     each, Clookup QMDDcomplexD.cpp:372-466) six parts = 18 sixteen-byte loads at a time;
//   * the two de-dup inserts share one allocation, one fence and have their slot loads and CAS in flight together;
//   * an item's own fields arrive as 16-byte quads.
// Same values, same short-cuts, same normal forms as the one-at-a-time code (which MODE_CTA keeps: shared memory
// answers in 28 cycles and the kernel has no registers to spare); only the order in which one item's independent
// lookups reach the value table differs, and on this path that order is timing-dependent anyway (real_insert_par).
#ifndef JKQ_GRID_CUH
#define JKQ_GRID_CUH

#include "jkq_engine.cuh"

struct JQuad { uint32_t x, y, z, w; };
JD JQuad ldq(const void *p) {
  JQuad q;
#ifdef JKQ_EMU
  memcpy(&q, p, 16);
#else
  const uint4 v = __ldcg((const uint4 *)p);
  q.x = v.x; q.y = v.y; q.z = v.z; q.w = v.w;
#endif
  return q;
}
#ifdef JKQ_EMU
JD bool jlane0() { return true; }
#else
JD bool jlane0() { return (threadIdx.x & 31u) == 0; }
#endif
JD uint64_t jq_lo(const JQuad &q) { return ((uint64_t)q.y << 32) | q.x; }
JD uint64_t jq_hi(const JQuad &q) { return ((uint64_t)q.w << 32) | q.z; }

// ------------------------------------------------------------------------------------ allocation
// warp-aggregated bump of a counter by 1 or 2 per lane (jagg_inc for a pair of items)
#ifdef JKQ_EMU
JD uint32_t jagg_add12(uint32_t *p, uint32_t cnt) { const uint32_t o = *p; *p = o + cnt; return o; }
#else
JD uint32_t jagg_add12(uint32_t *p, uint32_t cnt) {
  const unsigned m = __match_any_sync(__activemask(), (unsigned long long)(uintptr_t)p);
  const unsigned two = __ballot_sync(m, cnt >= 2u) & m;
  const int leader = __ffs(m) - 1;
  const unsigned lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(p, (uint32_t)(__popc(m) + __popc(two)));
  base = __shfl_sync(m, base, leader);
  return base + (uint32_t)(__popc(m & below) + __popc(two & below));
}
#endif

// ------------------------------------------------------------------------------------ expand
// kind M above the target: the node's two edges are forwarded (the fast path of expand_item); the two direct-indexed
// de-dup slots are read together, the two new items share one allocation, the two CAS are in flight together
JNI uint32_t expand_m_above_g(Ctx &c, JItem *it, uint32_t ia, uint32_t role) {
  JEng *E = c.E;
  const JNode p = node_load<MODE_GRID>(c, ia);
  const uint32_t ch[2] = {nd_c0(p), p.c1};
  const uint64_t w[2] = {nd_w0(p), nd_w1(p)};
  uint32_t kd[2], t[2] = {ch[0], ch[1]};
  bool need[2];
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    need[i] = false;
    if (w[i] == 0) { kd[i] = JC_ZERO; t[i] = 0; }
    else if (role_satisfied(role, i)) { kd[i] = JC_REF; need[i] = true; }
    else kd[i] = JC_PASS;
  }
  const bool same = need[0] && need[1] && ch[0] == ch[1];
  if (same) need[1] = false;
  const unsigned long long gen = (unsigned long long)ld_s<MODE_GRID>(&E->gen) << 32;
  unsigned long long sv[2] = {0, 0};
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) if (need[i]) sv[i] = ld_r<MODE_GRID>(E->mdir + ch[i]);
  bool claim[2];
  uint32_t ncl = 0;
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    claim[i] = false;
    if (!need[i]) continue;
    if ((sv[i] >> 32) == (gen >> 32)) t[i] = (uint32_t)sv[i] - 1;
    else { claim[i] = true; ncl++; }
  }
  if (ncl) {
    uint32_t mine = jagg_add12(&E->item_count, ncl);
    if (mine + ncl > E->item_cap) { jset_error<MODE_GRID>(E, c.sh, JKQ_ERR_ITEM_FULL); return JS_DONE; }
    uint32_t id[2];
    unsigned long long old[2];
    JKQ_UNROLL
    for (int i = 0; i < 2; i++) if (claim[i]) { id[i] = mine++; item_init_m(c.items + id[i], ch[i]); }
    JKQ_UNROLL
    for (int i = 0; i < 2; i++) if (claim[i]) old[i] = jatomic_cas(E->mdir + ch[i], sv[i], gen | (unsigned long long)(id[i] + 1));
    JKQ_UNROLL
    for (int i = 0; i < 2; i++) {
      if (!claim[i]) continue;
      if (old[i] == sv[i]) { t[i] = id[i]; continue; }
      c.items[id[i]].flags = JS_DEAD;  // somebody else got there first (the slot may still carry an old generation: the plain loop sorts it out)
      t[i] = dt_find_or_insert_m<MODE_GRID>(E, c.sh, c.items, ch[i]);
      if (t[i] == JI_FAIL) return JS_DONE;
    }
  }
  if (same) t[1] = t[0];
  it->ch[0] = t[0]; it->ch[1] = t[1]; it->chw[0] = w[0]; it->chw[1] = w[1];
  return item_flags(JS_DONE, kd[0], kd[1], 1u + nd_flag(p));
}

// de-dup insert of up to two F / S items (the REF slots of one item): slot loads, allocation, fence and CAS shared
JD bool slot_key_eq(const Slot &x, const Slot &y) {
  return x.ia == y.ia && x.ib == y.ib && x.itag == y.itag && x.iwa == y.iwa && x.iwb == y.iwb && x.iwc == y.iwc;
}
JNI bool dedup2_g(Ctx &c, Slot *s) {
  JEng *E = c.E;
  bool need[2] = {s[0].kind == JC_REF, s[1].kind == JC_REF};
  if (!need[0] && !need[1]) return true;
  const bool same = need[0] && need[1] && slot_key_eq(s[0], s[1]);
  if (same) need[1] = false;
  const unsigned long long gen = (unsigned long long)ld_s<MODE_GRID>(&E->gen) << 32;
  const uint32_t dmask = E->dtab_mask;
  uint32_t h[2] = {0, 0};
  unsigned long long sv[2] = {0, 0};
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    if (need[i]) {
      h[i] = dt_hash(s[i].ia, s[i].ib, s[i].iwa, s[i].iwb, s[i].iwc, s[i].itag) & dmask;
      sv[i] = ld_r<MODE_GRID>(E->dtab + h[i]);
    }
  }
  // a slot of this gate's generation: the same key, or a collision (left to the plain probing loop)
  bool claim[2] = {false, false}, slow[2] = {false, false};
  JQuad k0[2], k1[2];
  uint64_t kwc[2];
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    if (need[i] && (sv[i] >> 32) == (gen >> 32)) {
      const JItem *o = c.items + ((uint32_t)sv[i] - 1);
      k0[i] = ldq(&o->a); k1[i] = ldq(&o->wa); kwc[i] = ld_r<MODE_GRID>(&o->wc);
    }
  }
  uint32_t ncl = 0;
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    if (!need[i]) continue;
    if ((sv[i] >> 32) == (gen >> 32)) {
      if (k0[i].x == s[i].ia && k0[i].y == s[i].ib && k0[i].z == s[i].itag && jq_lo(k1[i]) == s[i].iwa && jq_hi(k1[i]) == s[i].iwb && kwc[i] == s[i].iwc)
        s[i].target = (uint32_t)sv[i] - 1;
      else slow[i] = true;
    } else { claim[i] = true; ncl++; }
  }
  if (ncl) {
    uint32_t mine = jagg_add12(&E->item_count, ncl);
    if (mine + ncl > E->item_cap) { jset_error<MODE_GRID>(E, c.sh, JKQ_ERR_ITEM_FULL); return false; }
    uint32_t id[2];
    unsigned long long old[2];
    JKQ_UNROLL
    for (int i = 0; i < 2; i++) if (claim[i]) { id[i] = mine++; item_init(c.items + id[i], s[i].ia, s[i].ib, s[i].iwa, s[i].iwb, s[i].iwc, s[i].itag); }
    jfence<MODE_GRID>();  // the keys before the slots that name them
    JKQ_UNROLL
    for (int i = 0; i < 2; i++) if (claim[i]) old[i] = jatomic_cas(E->dtab + h[i], sv[i], gen | (unsigned long long)(id[i] + 1));
    JKQ_UNROLL
    for (int i = 0; i < 2; i++) {
      if (!claim[i]) continue;
      if (old[i] == sv[i]) s[i].target = id[i];
      else { c.items[id[i]].flags = JS_DEAD; slow[i] = true; }
    }
  }
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    if (slow[i]) {
      s[i].target = dt_find_or_insert<MODE_GRID>(E, c.sh, c.items, s[i].ia, s[i].ib, s[i].iwa, s[i].iwb, s[i].iwc, s[i].itag);
      if (s[i].target == JI_FAIL) return false;
    }
  }
  if (same) s[1].target = s[0].target;
  return true;
}


#ifndef JKQ_EMU
// ---- the same two steps run by a whole warp on one item: lane i < 2 takes child slot i, the allocation is one atomic of
// lane 0, the results come back by shuffle (every lane holds the same inputs and leaves with the same outputs)
JNI uint32_t expand_m_above_w(Ctx &c, JItem *it, uint32_t ia, uint32_t role) {
  JEng *E = c.E;
  const uint32_t lane = threadIdx.x & 31u;
  const JNode p = node_load<MODE_GRID>(c, ia);
  const uint32_t ch[2] = {nd_c0(p), p.c1};
  const uint64_t w[2] = {nd_w0(p), nd_w1(p)};
  uint32_t kd[2], t[2] = {ch[0], ch[1]};
  bool need[2];
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    need[i] = false;
    if (w[i] == 0) { kd[i] = JC_ZERO; t[i] = 0; }
    else if (role_satisfied(role, i)) { kd[i] = JC_REF; need[i] = true; }
    else kd[i] = JC_PASS;
  }
  const bool same = need[0] && need[1] && ch[0] == ch[1];
  if (same) need[1] = false;
  const unsigned long long gen = (unsigned long long)ld_s<MODE_GRID>(&E->gen) << 32;
  const uint32_t mych = (lane & 1u) ? ch[1] : ch[0];
  const bool mine = lane < 2u && ((lane & 1u) ? need[1] : need[0]);
  unsigned long long sv = 0;
  if (mine) sv = ld_r<MODE_GRID>(E->mdir + mych);
  const bool found = mine && (sv >> 32) == (gen >> 32);
  const bool claim = mine && !found;
  const unsigned cm = __ballot_sync(0xFFFFFFFFu, claim);
  uint32_t base = 0;
  if (cm) {
    if (lane == 0) base = atomicAdd(&E->item_count, (uint32_t)__popc(cm));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base + (uint32_t)__popc(cm) > E->item_cap) { jset_error<MODE_GRID>(E, c.sh, JKQ_ERR_ITEM_FULL); return JS_DONE; }
  }
  uint32_t tt = found ? (uint32_t)sv - 1 : 0;
  if (claim) {
    const uint32_t id = base + (uint32_t)__popc(cm & ((1u << lane) - 1u));
    item_init_m(c.items + id, mych);
    const unsigned long long old = jatomic_cas(E->mdir + mych, sv, gen | (unsigned long long)(id + 1));
    if (old == sv) tt = id;
    else { c.items[id].flags = JS_DEAD; tt = dt_find_or_insert_m<MODE_GRID>(E, c.sh, c.items, mych); }
  }
  __syncwarp();
  const uint32_t t0 = __shfl_sync(0xFFFFFFFFu, tt, 0), t1 = __shfl_sync(0xFFFFFFFFu, tt, 1);
  if (need[0]) t[0] = t0;
  if (need[1]) t[1] = t1;
  if (same) t[1] = t[0];
  if ((need[0] && t[0] == JI_FAIL) || (need[1] && t[1] == JI_FAIL)) return JS_DONE;
  if (lane == 0) { it->ch[0] = t[0]; it->ch[1] = t[1]; it->chw[0] = w[0]; it->chw[1] = w[1]; }
  return item_flags(JS_DONE, kd[0], kd[1], 1u + nd_flag(p));
}

JNI bool dedup2_w(Ctx &c, Slot *s) {
  JEng *E = c.E;
  const uint32_t lane = threadIdx.x & 31u;
  bool need[2] = {s[0].kind == JC_REF, s[1].kind == JC_REF};
  if (!need[0] && !need[1]) return true;
  const bool same = need[0] && need[1] && slot_key_eq(s[0], s[1]);
  if (same) need[1] = false;
  const unsigned long long gen = (unsigned long long)ld_s<MODE_GRID>(&E->gen) << 32;
  const Slot &ms = s[lane & 1u];
  const bool mine = lane < 2u && ((lane & 1u) ? need[1] : need[0]);
  uint32_t h = 0, tt = 0;
  unsigned long long sv = 0;
  bool claim = false, slow = false;
  if (mine) {
    h = dt_hash(ms.ia, ms.ib, ms.iwa, ms.iwb, ms.iwc, ms.itag) & E->dtab_mask;
    sv = ld_r<MODE_GRID>(E->dtab + h);
    if ((sv >> 32) == (gen >> 32)) {
      const JItem *o = c.items + ((uint32_t)sv - 1);
      const JQuad k0 = ldq(&o->a), k1 = ldq(&o->wa);
      const uint64_t kwc = ld_r<MODE_GRID>(&o->wc);
      if (k0.x == ms.ia && k0.y == ms.ib && k0.z == ms.itag && jq_lo(k1) == ms.iwa && jq_hi(k1) == ms.iwb && kwc == ms.iwc) tt = (uint32_t)sv - 1;
      else slow = true;
    } else claim = true;
  }
  const unsigned cm = __ballot_sync(0xFFFFFFFFu, claim);
  uint32_t base = 0;
  if (cm) {
    if (lane == 0) base = atomicAdd(&E->item_count, (uint32_t)__popc(cm));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base + (uint32_t)__popc(cm) > E->item_cap) { jset_error<MODE_GRID>(E, c.sh, JKQ_ERR_ITEM_FULL); return false; }
  }
  if (claim) {
    const uint32_t id = base + (uint32_t)__popc(cm & ((1u << lane) - 1u));
    item_init(c.items + id, ms.ia, ms.ib, ms.iwa, ms.iwb, ms.iwc, ms.itag);
    jfence<MODE_GRID>();  // the key before the slot that names it
    const unsigned long long old = jatomic_cas(E->dtab + h, sv, gen | (unsigned long long)(id + 1));
    if (old == sv) tt = id;
    else { c.items[id].flags = JS_DEAD; slow = true; }
  }
  if (slow) tt = dt_find_or_insert<MODE_GRID>(E, c.sh, c.items, ms.ia, ms.ib, ms.iwa, ms.iwb, ms.iwc, ms.itag);
  __syncwarp();
  const uint32_t t0 = __shfl_sync(0xFFFFFFFFu, tt, 0), t1 = __shfl_sync(0xFFFFFFFFu, tt, 1);
  if (need[0]) s[0].target = t0;
  if (need[1]) s[1].target = t1;
  if (same) s[1].target = s[0].target;
  if ((need[0] && t0 == JI_FAIL) || (need[1] && t1 == JI_FAIL)) return false;
  return true;
}
#endif

// what is left to do for one child slot after a stage
#define GS_DONE 0u
#define GS_PASS_Q2 1u    // slot = (tgt, q[2])
#define GS_COMBINE 2u    // combine(al = q[0], be = q[1], ga = q[2])
#define GS_SUM 3u        // combine_sum(al = q[0], be = q[1])
#define GS_ADDMUL 4u     // same-node fused item below the lowest control: w = aw * (wa + wb)
#define GS_COMBINE2 5u   // magnitudes -> three divisions
#define GS_SUM2 6u       // magnitudes -> operand order -> one division
#define GS_COMBINE3 7u
#define GS_SUM3 8u
#define GS_SUMADD 9u     // terminals / same node: the weights are added
#define GS_ADDMUL2 10u
#define GS_ADDMUL3 11u
struct GChild {
  Slot s;
  uint64_t al, be, ga, f, aw;
  uint32_t A, B, sel, st, tgt;
  int m;
};

// expand_item for MODE_GRID (see the head of this file).  ia / ib / tag: the item's first quad, already loaded.
// W: called convergently by all 32 lanes of a warp with the same arguments (see cx_batch_w); else by one thread
template <bool W> JNI uint32_t expand_item_g(Ctx &c, const JGate &g, uint32_t idx, int L, uint32_t ia, uint32_t ib, uint32_t tag) {
  JEng *E = c.E;
  JItem *it = c.items + idx;
  const uint32_t kind = tag & 15u, sel = (tag >> 4) & 1u, role = g.role[L];
  JGP_T(e0);
  if (kind == JK_M && L > g.target) {
#ifndef JKQ_EMU
    const uint32_t r = W ? expand_m_above_w(c, it, ia, role) : expand_m_above_g(c, it, ia, role);
#else
    const uint32_t r = expand_m_above_g(c, it, ia, role);
#endif
    JGP_T(ea);
    JGP_ADD(c, 13, e0, ea);
    return r;
  }
  // ---- everything the item refers to, in flight together
  uint64_t iwa = 0, iwb = 0, iwc = 0, u[4] = {0, 0, 0, 0};
  JNode na, nb;
  na.c0 = na.c1 = 0; na.w = 0; nb = na;
  if (kind != JK_M) {
    const JQuad k1 = ldq(&it->wa);
    const uint64_t wc = ld_r<MODE_GRID>(&it->wc);
    if (ia) na = node_load<MODE_GRID>(c, ia);
    if (ib) nb = (ib == ia) ? na : node_load<MODE_GRID>(c, ib);
    iwa = jq_lo(k1); iwb = jq_hi(k1); iwc = wc;
  } else {
    JKQ_UNROLL
    for (int k = 0; k < 4; k++) u[k] = ld_s<MODE_GRID>(&E->cur_u[k]);
    if (ia) na = node_load<MODE_GRID>(c, ia);
  }
  const int low = g.low;
  QOp q[6];
  GChild ch[2];
  // ---- stage 1: the products of expand_child
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    GChild &x = ch[i];
    QOp *o = q + 3 * i;
    q_set(o[0], 0); q_set(o[1], 0); q_set(o[2], 0);
    x.m = L - 1; x.sel = sel; x.st = GS_DONE; x.tgt = 0; x.aw = 0; x.A = x.B = 0;
    slot_zero(x.s);
    if (kind == JK_M) {  // target level: row i of the matrix
      const uint64_t w = i ? nd_w1(na) : nd_w0(na);
      q_mul(o[0], u[2 * i], nd_w0(na)); q_mul(o[1], u[2 * i + 1], nd_w1(na)); q_set(o[2], w);
      x.A = nd_c0(na); x.B = na.c1; x.sel = (uint32_t)i; x.st = GS_COMBINE;
      continue;
    }
    const uint32_t ac = ia ? (i ? na.c1 : nd_c0(na)) : 0, bc = ib ? (i ? nb.c1 : nd_c0(nb)) : 0;
    const uint64_t aw = ia ? (i ? nd_w1(na) : nd_w0(na)) : 0, bw = ib ? (i ? nd_w1(nb) : nd_w0(nb)) : 0;
    if (kind == JK_S) {
      q_set(o[0], aw); q_mul(o[1], iwb, bw);
      x.A = ac; x.B = bc; x.st = GS_SUM;
    } else if (ia == ib) {
      if (aw == 0) continue;  // zero slot
      if (!role_satisfied(role, i)) { q_mul(o[2], iwc, aw); x.tgt = ac; x.st = GS_PASS_Q2; }
      else if (L - 1 >= low) slot_ref(x.s, aw, ac, ac, iwa, iwb, iwc, tag_level(tag, L - 1));
      else { x.aw = aw; x.tgt = ac; x.st = GS_ADDMUL; }
    } else {
      q_mul(o[2], iwc, sel ? bw : aw);
      if (!role_satisfied(role, i)) { x.tgt = sel ? bc : ac; x.st = GS_PASS_Q2; }
      else { q_mul(o[0], iwa, aw); q_mul(o[1], iwb, bw); x.A = ac; x.B = bc; x.st = GS_COMBINE; }
    }
  }
  JGP_T(e1);
  JGP_ADD(c, 8, e0, e1);
  bool ok = cx_run<6, W>(c, q);
  JGP_T(e2);
  JGP_ADD(c, 9, e1, e2);
  if (!ok) return JS_RETRY;  // (with in-place insertion a failed lookup is an error: the barrier sees it)
  // ---- stage 2: what the products decide; the magnitudes that order operands / pick the common factor
  const double *vals = E->vals;
  bool need_mag = false;
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    GChild &x = ch[i];
    QOp *o = q + 3 * i;
    if (x.st == GS_PASS_Q2) { slot_pass(x.s, x.tgt, o[2].r); x.st = GS_DONE; continue; }
    if (x.st == GS_ADDMUL) { q_add(o[0], iwa, iwb); x.st = GS_ADDMUL2; continue; }
    if (x.st != GS_COMBINE && x.st != GS_SUM) continue;
    x.al = o[0].r; x.be = o[1].r; x.ga = o[2].r;
    if (x.st == GS_COMBINE && x.m >= low) {
      const uint64_t al = x.al, be = x.be, ga = x.ga;
      if (al == 0 && be == 0 && ga == 0) { slot_zero(x.s); x.st = GS_DONE; }
      else if (x.sel == 0 && be == 0 && al == ga) { slot_pass(x.s, x.A, ga); x.st = GS_DONE; }
      else if (x.sel == 1 && al == 0 && be == ga) { slot_pass(x.s, x.B, ga); x.st = GS_DONE; }
      else { x.st = GS_COMBINE2; need_mag = true; }
      continue;
    }
    // combine_sum
    const uint64_t al = x.al, be = x.be;
    if (x.m < 0 || x.A == x.B) {
      if (al == 0) { slot_pass(x.s, x.B, be); x.st = GS_DONE; }
      else if (be == 0) { slot_pass(x.s, x.A, al); x.st = GS_DONE; }
      else { q_add(o[0], al, be); x.st = GS_SUMADD; }
    } else if (al == 0) { slot_pass(x.s, x.B, be); x.st = GS_DONE; }
    else if (be == 0) { slot_pass(x.s, x.A, al); x.st = GS_DONE; }
    else { x.st = GS_SUM2; need_mag = true; }
  }
  if (need_mag) {
    double mg[2][3];
    JKQ_UNROLL
    for (int i = 0; i < 2; i++) {  // loads of both slots together
      if (ch[i].st == GS_COMBINE2 || ch[i].st == GS_SUM2) {
        mg[i][0] = jmag2_g(vals, ch[i].al); mg[i][1] = jmag2_g(vals, ch[i].be);
        mg[i][2] = ch[i].st == GS_COMBINE2 ? jmag2_g(vals, ch[i].ga) : 0.0;
      }
    }
    JKQ_UNROLL
    for (int i = 0; i < 2; i++) {
      GChild &x = ch[i];
      QOp *o = q + 3 * i;
      if (x.st == GS_COMBINE2) {
        // common factor: the largest of the three weights (first one on ties), as in combine()
        const uint64_t al = x.al, be = x.be, ga = x.ga;
        uint64_t f = al ? al : (be ? be : ga);
        double mf = al ? mg[i][0] : (be ? mg[i][1] : mg[i][2]);
        if (be != 0 && be != f) { if (mg[i][1] > mf) { f = be; mf = mg[i][1]; } }
        if (ga != 0 && ga != f) { if (mg[i][2] > mf) { f = ga; mf = mg[i][2]; } }
        x.f = f;
        q_div(o[0], al, f); q_div(o[1], be, f); q_div(o[2], ga, f);
        x.st = GS_COMBINE3;
      } else if (x.st == GS_SUM2) {
        // operand order of the memo key, as in combine_sum()
        bool swap = x.A > x.B;
        if (x.al != x.be && x.al != jneg(x.be)) {
          if (mg[i][0] != mg[i][1]) swap = mg[i][1] > mg[i][0];
        }
        if (swap) { const uint32_t t = x.A; x.A = x.B; x.B = t; const uint64_t tw = x.al; x.al = x.be; x.be = tw; }
        q_div(o[0], x.be, x.al);
        x.st = GS_SUM3;
      }
    }
  }
  // ---- stage 3: divisions / additions
  JGP_T(e3);
  JGP_ADD(c, 10, e2, e3);
  ok = cx_run<6, W>(c, q);
  JGP_T(e4);
  JGP_ADD(c, 11, e3, e4);
  if (!ok) return JS_RETRY;
  bool more = false;
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    GChild &x = ch[i];
    QOp *o = q + 3 * i;
    if (x.st == GS_COMBINE3) {
      const uint64_t wa = o[0].r, wb = o[1].r, wc = o[2].r;
      const uint32_t An = (wa != 0 || (x.sel == 0 && wc != 0)) ? x.A : 0;
      const uint32_t Bn = (wb != 0 || (x.sel == 1 && wc != 0)) ? x.B : 0;
      slot_ref(x.s, x.f, An, Bn, wa, wb, wc, tag_level(JK_F | (x.sel << 4), x.m));
      x.st = GS_DONE;
    } else if (x.st == GS_SUM3) {
      slot_ref(x.s, x.al, x.A, x.B, JW_ONE, o[0].r, 0, tag_level(JK_S, x.m));
      x.st = GS_DONE;
    } else if (x.st == GS_SUMADD) {
      slot_pass(x.s, x.A, o[0].r);  // (al != 0 here)
      x.st = GS_DONE;
    } else if (x.st == GS_ADDMUL2) {
      q_mul(o[0], x.aw, o[0].r);
      x.st = GS_ADDMUL3;
      more = true;
    }
  }
  if (more) {
    ok = cx_run<6, W>(c, q);
    if (!ok) return JS_RETRY;
    JKQ_UNROLL
    for (int i = 0; i < 2; i++)
      if (ch[i].st == GS_ADDMUL3) { slot_pass(ch[i].s, ch[i].tgt, q[3 * i].r); ch[i].st = GS_DONE; }
  }
  // ---- de-dup inserts, only now that all arithmetic succeeded
  Slot s[2] = {ch[0].s, ch[1].s};
  JGP_T(e5);
#ifndef JKQ_EMU
  if (!(W ? dedup2_w(c, s) : dedup2_g(c, s))) return JS_DONE;  // error flagged; the gate is abandoned
#else
  if (!dedup2_g(c, s)) return JS_DONE;
#endif
  JGP_T(e6);
  JGP_ADD(c, 12, e5, e6);
  if (!W || jlane0()) { it->ch[0] = s[0].target; it->ch[1] = s[1].target; it->chw[0] = s[0].w; it->chw[1] = s[1].w; }
  uint32_t hint = 0;
  if (kind == JK_F && ia == ib && s[0].w == nd_w0(na) && s[1].w == nd_w1(na) && s[0].kind != JC_ZERO && s[1].kind != JC_ZERO) hint = 1u + nd_flag(na);
  return item_flags(JS_DONE, s[0].kind, s[1].kind, hint);
}

// ------------------------------------------------------------------------------------ reduce
// |a1| > |a0| beyond the snapping error, from magnitudes already loaded (mag_greater of jkq_engine.cuh)
JD bool mag_greater_v(uint64_t a1, uint64_t a0, double m1, double m0, double tol) {
  if (a1 == a0 || a1 == jneg(a0)) return false;
#ifdef JKQ_EXACT_TIES
  return m1 > m0;
#else
  return m1 > JKQ_DADD(m0, JKQ_DMUL(JKQ_DMUL(4.0, tol), sqrt(m1 > m0 ? m1 : m0)));
#endif
}

// reduce_core with the two independent operations of a normalisation in one batch
template <bool W> JD bool reduce_core_g(Ctx &c, uint32_t rn[2], const uint64_t f[2], const uint64_t a[2], uint32_t hint, uint32_t *out_node, uint64_t *out_w) {
  JEng *E = c.E;
  if (a[0] == 0 && a[1] == 0) { *out_node = 0; *out_w = 0; return true; }
  uint32_t mx;
  uint64_t other, top;
  QOp q[2];
  if (hint && a[hint - 1] == JW_ONE && (f[0] == f[1] || a[2 - hint] == 0)) {
    mx = hint - 1;
    other = a[1 - mx];
    top = f[mx];
  } else if (a[0] == 0 || a[1] == 0 || f[0] == f[1]) {
    if (a[0] == 0) mx = 1u;
    else if (a[1] == 0) mx = 0u;
    else {
      const double m0 = jmag2_g(E->vals, a[0]), m1 = jmag2_g(E->vals, a[1]);
      mx = mag_greater_v(a[1], a[0], m1, m0, E->tol) ? 1u : 0u;
    }
    q_div(q[0], mx ? a[0] : a[1], mx ? a[1] : a[0]);
    q_mul(q[1], mx ? f[1] : f[0], mx ? a[1] : a[0]);
    if (!cx_run<2, W>(c, q)) return false;
    other = q[0].r; top = q[1].r;
  } else {
    q_mul(q[0], f[0], a[0]); q_mul(q[1], f[1], a[1]);
    if (!cx_run<2, W>(c, q)) return false;
    const uint64_t rw0 = q[0].r, rw1 = q[1].r;
    if (rw0 == 0 && rw1 == 0) { *out_node = 0; *out_w = 0; return true; }
    if (rw0 == 0) mx = 1u;
    else if (rw1 == 0) mx = 0u;
    else {
      const double m0 = jmag2_g(E->vals, rw0), m1 = jmag2_g(E->vals, rw1);
      mx = mag_greater_v(rw1, rw0, m1, m0, E->tol) ? 1u : 0u;
    }
    q_div(q[0], mx ? rw0 : rw1, mx ? rw1 : rw0);
    q_set(q[1], 0);
    if (!cx_run<2, W>(c, q)) return false;
    other = q[0].r;
    top = mx ? rw1 : rw0;
  }
  if (top == 0) { *out_node = 0; *out_w = 0; return true; }
  JGP_T(u0);
  uint32_t node = 0;
  if (!W || jlane0()) node = ut_find_or_insert<MODE_GRID>(E, c.sh, ((mx == 1 && other == 0) ? 0u : rn[0]) | (mx << 31), (mx == 0 && other == 0) ? 0u : rn[1], other);
#ifndef JKQ_EMU
  if (W) { __syncwarp(); node = __shfl_sync(0xFFFFFFFFu, node, 0); }
#endif
  *out_node = node;
  JGP_T(u1);
  JGP_ADD(c, 18, u0, u1);
  *out_w = top;
  return true;
}

template <bool W> JNI bool reduce_item_g(Ctx &c, uint32_t idx, uint32_t flags) {
  JItem *it = c.items + idx;
  JGP_T(r0);
  const JQuad q2 = ldq(&it->wc), q3 = ldq(&it->chw[0]);  // (wc, ch[0], ch[1]) and (chw[0], chw[1])
  const uint32_t cht[2] = {q2.z, q2.w};
  const uint64_t chw[2] = {jq_lo(q3), jq_hi(q3)};
  JQuad r4[2];
  JKQ_UNROLL
  for (int i = 0; i < 2; i++)
    if (((flags >> (2 + 2 * i)) & 3u) == JC_REF) r4[i] = ldq(&c.items[cht[i]].rw);  // (rw, rnode, pad) of both child items together
  uint32_t rn[2];
  uint64_t f[2], a[2];
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {
    const uint32_t k = (flags >> (2 + 2 * i)) & 3u;
    if (k == JC_ZERO) { rn[i] = 0; f[i] = JW_ONE; a[i] = 0; }
    else if (k == JC_PASS) { rn[i] = cht[i]; f[i] = JW_ONE; a[i] = chw[i]; }
    else { f[i] = jq_lo(r4[i]); rn[i] = r4[i].z; a[i] = f[i] ? chw[i] : 0; }
    if (a[i] == 0) rn[i] = 0;
  }
  uint32_t rnode;
  uint64_t rw;
  JGP_T(r1);
  JGP_ADD(c, 16, r0, r1);
  if (!reduce_core_g<W>(c, rn, f, a, (flags >> 6) & 3u, &rnode, &rw)) return false;
  JGP_T(r2);
  JGP_ADD(c, 17, r1, r2);
  if (!W || jlane0()) { it->rw = rw; it->rnode = rnode; }
  return true;
}

#endif  // JKQ_GRID_CUH
