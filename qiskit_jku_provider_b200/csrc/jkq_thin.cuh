// The cursor: lazy treatment of the levels ABOVE the gates being applied.
//
// In the reference every ApplyGate starts at the root: QMDDmultiply2 (QMDDpackage.cpp:1416-1517) walks
// down to the gate's highest control/target through levels where it only forwards the two edges of each
// state node (block-diagonal gate node), and QMDDmakeNonterminal (:1204-1233) rebuilds those levels on
// the way back.  For the "small DD, deep circuit" shapes (QFT, GHZ, Grover) that walk IS the cost of a
// gate: a chain of ~2(n - top) dependent steps with no parallelism.  Consecutive gates mostly work in
// the same region of the diagram, so the walk is not redone per gate.  The engine keeps a cursor level M:
//   * levels above M hold the structure of the state as it was when the cursor passed them: one
//     ThinEnt per node and level (child slots, edge weights, and a factor to re-apply);
//   * at level M the state is a list of up to 32 (node, weight) edges, one per lane of ONE warp.
// A gate whose highest level `top` is below M moves the cursor down (a lane loads its node, the warp
// removes duplicate children with match/shuffle, children are handed to lanes by rank); a gate above M
// moves it up (reduce_core: the same normalisation and unique-table code as the generic path); a gate
// at M just replaces the edges at M by its results.  Only a collection, sampling, a snapshot or the end
// of the kernel rebuild everything up to the root (cursor_to_root).  Where a level outgrows the warp
// the cursor stops, and the generic item phases (jkq_apply.cuh) start one level below it.
// Runs in block 0 / warp 0 of either launch shape; everybody else waits at the next barrier.
#ifndef JKQ_THIN_CUH
#define JKQ_THIN_CUH

#include "jkq_engine.cuh"

#if defined(JKQ_PROFILE) && !defined(JKQ_EMU)
#define JTHIN_T(var) long long var = clock64()
#define JTHIN_ADD(c, slot, t0, t1) do { if ((c).tid == 0) prof_add<MODE>((c).E, (c).sh, slot, (unsigned long long)((t1) - (t0))); } while (0)
#else
#define JTHIN_T(var) do {} while (0)
#define JTHIN_ADD(c, slot, t0, t1) do {} while (0)
#endif

template <int MODE> JD void prof_add(JEng *E, CtaState *S, int slot, unsigned long long v) {
#ifndef JKQ_EMU
  if (MODE == MODE_CTA) S->prof[slot] += v;  // one thread; folded into JEng.prof by cta_flush
  else atomicAdd(&E->prof[slot], v);
#endif
}

JD uint32_t lanes_below(uint32_t lane) { return (1u << lane) - 1u; }

// sticky error / want_grid words as barrier flags.  Every lane of the cursor warp calls this (always from uniform
// control flow) and every lane gets the SAME answer: the words are written by whichever lane ran out of a pool, and a
// ballot orders nothing in memory -- two lanes reading the word right after one could disagree, part of the warp
// would leave through the error path while the rest waited at the next __syncwarp (seen on B200 as a hang when the
// shared value table filled up inside cursor_to_root).  The OR over the warp makes the decision collective.
template <int MODE> JD uint32_t thin_trouble(const Ctx &c) {
  uint32_t fl = 0;
  if (MODE == MODE_CTA ? ld_r<MODE>(&c.sh->err) != 0 : ld_r<MODE>(&c.E->status) != 0) fl |= JF_ERROR;
  if (MODE == MODE_CTA && ld_r<MODE>(&c.sh->want_grid)) fl |= JF_GRID;
#ifndef JKQ_EMU
  fl = __reduce_or_sync(0xFFFFFFFFu, fl);
#endif
  return fl;
}
// some lane met a value that is not in the table yet: commit the candidates (warp-local: nobody else runs)
template <int MODE> JD void thin_commit(Ctx &c) {
  jsyncwarp();
  vals_commit<MODE>(c.E, c.sh, c.ring, c.ltid, JKQ_WARP);
  jsyncwarp();
  if (c.ltid == 0) { *cnt_pend<MODE>(c, c.ring) = 0; jstat<MODE>(c.E, c.sh, JST_RETRY, 1u); }
  jsyncwarp();
}

// the cursor sits on the root edge (kernel start, after a collection)
template <int MODE> JD void cursor_reset(Ctx &c, int n) {
  ThinState *T = c.T;
  if (c.ltid == 0) {
    ThinRes r;
    r.rnode = MODE == MODE_CTA ? c.sh->root : ld_s<MODE>(&c.E->root);
    r.rw = MODE == MODE_CTA ? c.sh->rootw : ld_s<MODE>(&c.E->rootw);
    r.pad = 0;
    T->res[(n - 1) & 1][0] = r;
    T->cnt[n - 1] = 1;
    T->M = n - 1;
    T->walked = 0;
  }
  jsyncwarp();
}

// Children of the cursor edges at level L (>= 1), duplicates removed.  A child is owned by the first
// edge (all c0 edges by lane, then all c1 edges by lane) that names it; owners are ranked 0..newcnt-1.
struct ThinKids {
  bool nz0, nz1, own0, own1;
  uint32_t c0, c1, s0, s1, newcnt;
  ThinEnt ent;
};
template <int MODE> JD ThinKids thin_children(Ctx &c, int L, uint32_t cnt) {
  ThinState *T = c.T;
  const uint32_t lane = c.ltid;
  ThinKids k;
  ThinRes r;
  r.rw = 0; r.rnode = 0; r.pad = 0;
  if (lane < cnt) r = T->res[L & 1][lane];
  const bool act = lane < cnt && r.rw != 0;  // (a zero edge has no children)
  JNode nd;
  nd.c0 = 0; nd.c1 = 0; nd.w = 0;
  if (act) nd = node_load<MODE>(c, r.rnode);
  k.c0 = nd_c0(nd); k.c1 = nd.c1;
  k.nz0 = act && nd_w0(nd) != 0; k.nz1 = act && nd_w1(nd) != 0;
  if (cnt == 1) {  // a chain (product state): nothing to de-duplicate but the node's own two edges
    const bool same = k.nz0 && k.nz1 && k.c0 == k.c1;
    k.own0 = k.nz0; k.own1 = k.nz1 && !same;
    k.s0 = 0; k.s1 = (k.nz0 && !same) ? 1u : 0u;
    k.newcnt = jshfl((k.own0 ? 1u : 0u) + (k.own1 ? 1u : 0u), 0);
  } else {
    const uint32_t v0 = k.nz0 ? k.c0 : (0x80000000u | lane);
    const uint32_t lead0 = jffs(jmatch(v0)) - 1u;
    k.own0 = k.nz0 && lead0 == lane;
    bool dup1 = false;
    uint32_t src = 0;
    for (uint32_t j = 0; j < cnt; j++) {
      const uint32_t vj = jshfl(v0, j);
      if (k.nz1 && vj == k.c1) { dup1 = true; src = j; }
    }
    const uint32_t ldsrc = jshfl(lead0, src);
    const uint32_t v1 = (k.nz1 && !dup1) ? k.c1 : (0x80000000u | lane);
    const uint32_t lead1 = jffs(jmatch(v1)) - 1u;
    k.own1 = k.nz1 && !dup1 && lead1 == lane;
    const uint32_t b0 = jballot(k.own0), b1 = jballot(k.own1), n0 = jpopc(b0);
    k.newcnt = n0 + jpopc(b1);
    k.s0 = jpopc(b0 & lanes_below(lead0));
    k.s1 = dup1 ? jpopc(b0 & lanes_below(ldsrc)) : n0 + jpopc(b1 & lanes_below(lead1));
  }
  k.ent.pk = k.s0 | (k.s1 << 8) | ((k.nz0 ? 1u : 0u) << 16) | ((k.nz1 ? 1u : 0u) << 17) | (nd_flag(nd) << 18);
  k.ent.pad = 0; k.ent.w = nd.w; k.ent.f = r.rw;
  return k;
}

// One level up: the edges of level L from its ThinEnt records and the edges below (cursor edges of level
// L-1, or -- from_items -- the results of the generic items [0, ..) of level L-1).  Returns 0 or barrier
// flags (JF_ERROR / JF_GRID, uniform); on failure nothing has been written.
template <int MODE> JD uint32_t cursor_up_step(Ctx &c, int L, bool from_items) {
  ThinState *T = c.T;
  const uint32_t lane = c.ltid;
  const uint32_t cnt = T->cnt[L];
  const bool act = lane < cnt;
  uint32_t rn[2] = {0, 0}, hint = 0;
  uint64_t f[2] = {JW_ONE, JW_ONE}, a[2] = {0, 0}, ef = 0;
  if (act) {
    const ThinEnt e = T->ent[L][lane];
    const uint32_t flag = (e.pk >> 18) & 1u;
    ef = e.f;
    JKQ_UNROLL
    for (int i = 0; i < 2; i++) {
      if (!((e.pk >> (16 + i)) & 1u)) continue;
      const uint32_t s = (e.pk >> (8 * i)) & 0xFFu;
      if (from_items) {
        const JItem *it = c.items + s;
        f[i] = ld_i<MODE>(&it->rw); rn[i] = ld_i<MODE>(&it->rnode);
      } else {
        const ThinRes r = T->res[(L - 1) & 1][s];
        f[i] = r.rw; rn[i] = r.rnode;
      }
      a[i] = (flag == (uint32_t)i) ? JW_ONE : e.w;
      if (f[i] == 0) a[i] = 0;
      if (a[i] == 0) rn[i] = 0;
    }
    hint = 1u + flag;
  }
  uint32_t rnode = 0, rounds = 0;
  uint64_t rw = 0;
  for (;;) {
    bool ok = true;
    if (act && ef != 0) {
      ok = reduce_core<MODE>(c, rn, f, a, hint, &rnode, &rw);
      if (ok && rw != 0 && rnode != 0) ok = cmul<MODE>(c, ef, rw, &rw);  // (rnode == 0: a pool ran out, seen below)
    }
    const uint32_t bad = jballot(!ok);
    if (bad || jballot(act && ok && rw != 0 && rnode == 0)) {
      const uint32_t fl = thin_trouble<MODE>(c);
      if (fl) return fl;
    }
    if (!bad) break;
    if (++rounds > JKQ_RETRY_LIMIT) jset_error<MODE>(c.E, c.sh, JKQ_ERR_RETRY_LIMIT);
    thin_commit<MODE>(c);
  }
  if (act) {
    ThinRes r;
    r.rw = rw; r.rnode = rw ? rnode : 0; r.pad = 0;
    T->res[L & 1][lane] = r;
  }
  jsyncwarp();
  return 0;
}

// rebuild everything up to the root and publish the root edge (before a collection, sampling, the end of
// the kernel)
template <int MODE> JD uint32_t cursor_to_root(Ctx &c, int n) {
  ThinState *T = c.T;
  int M = T->M;
  if (MODE == MODE_CTA) {  // this walk may dip into the reserve of the shared-memory pools
    if (c.ltid == 0) { c.sh->node_limit = c.sh->node_hard; c.sh->val_limit = c.sh->val_hard; }
    jsyncwarp();
  }
  for (; M < n - 1; M++) {
    const uint32_t f = cursor_up_step<MODE>(c, M + 1, false);
    if (f) {
      if (c.ltid == 0) { T->M = M; if (MODE == MODE_CTA) { c.sh->node_limit = c.sh->node_soft; c.sh->val_limit = c.sh->val_soft; } }
      jsyncwarp();
      return f;
    }
  }
  if (c.ltid == 0) {
    if (MODE == MODE_CTA) { c.sh->node_limit = c.sh->node_soft; c.sh->val_limit = c.sh->val_soft; }
    T->M = n - 1;
    const ThinRes r = T->res[(n - 1) & 1][0];
    if (r.rw == 0 || r.rnode == 0) jset_error<MODE>(c.E, c.sh, JKQ_ERR_ZERO_STATE);
    if (MODE == MODE_CTA) { c.sh->root = r.rnode; c.sh->rootw = r.rw; }  // (copied to HBM by cta_flush)
    else { c.E->root = r.rnode; c.E->rootw = r.rw; }
    jstat<MODE>(c.E, c.sh, JST_ITEMS, T->walked);
    T->walked = 0;
  }
  jsyncwarp();
  return 0;
}

// Bring the cursor to the gate: up to `top` if it is below, then down as far as the warp width allows.
// Leaves the M items [0, *icount) of the first generic level *lsw:
//   *blocked = 0   cursor level == top == *lsw: item s works on cursor edge s
//   *blocked = 1   the level below the cursor is too wide: the cursor stays at *lsw + 1, its ThinEnt
//                  records are written, the items are its (de-duplicated) children
template <int MODE> JD uint32_t cursor_prepare(Ctx &c, int top, int *lsw, uint32_t *icount, uint32_t *blocked) {
  ThinState *T = c.T;
  const uint32_t lane = c.ltid;
  int M = T->M;
  uint32_t walked = 0;
  *blocked = 0; *lsw = M; *icount = 0;
  JFINE_T(u0);
  for (; M < top; M++) {
    const uint32_t f = cursor_up_step<MODE>(c, M + 1, false);
    JFINE_ADD(c, 17, 0, 1);
    if (f) { if (lane == 0) T->M = M; jsyncwarp(); return f; }
  }
  JFINE_T(u1);
  JFINE_ADD(c, 16, u0, u1);
  uint32_t cnt = T->cnt[M];
  while (M > top) {
    JFINE_ADD(c, 19, 0, 1);
    const ThinKids k = thin_children<MODE>(c, M, cnt);
    if (lane < cnt) T->ent[M][lane] = k.ent;
    walked += cnt;
    if (k.newcnt > JKQ_WARP) {  // too wide below: generic M items for the children, the cursor stays
      if (k.own0) item_init_m(c.items + k.s0, k.c0);
      if (k.own1) item_init_m(c.items + k.s1, k.c1);
      *blocked = 1; *lsw = M - 1; *icount = k.newcnt;
      if (lane == 0) { T->M = M; T->walked += walked; }
      jsyncwarp();
      return 0;
    }
    ThinRes r;
    r.rw = JW_ONE; r.pad = 0;
    if (k.own0) { r.rnode = k.c0; T->res[(M - 1) & 1][k.s0] = r; }
    if (k.own1) { r.rnode = k.c1; T->res[(M - 1) & 1][k.s1] = r; }
    M--;
    cnt = k.newcnt;
    if (lane == 0) T->cnt[M] = cnt;
    jsyncwarp();
  }
  JFINE_T(u2);
  JFINE_ADD(c, 18, u1, u2);
  // the cursor is at the gate's top level: one M item per cursor edge (a zero edge gets a dead item)
  if (lane < cnt) {
    const ThinRes r = T->res[M & 1][lane];
    item_init_m(c.items + lane, r.rnode);
    if (r.rw == 0) c.items[lane].flags = JS_DEAD;
  }
  *lsw = M; *icount = cnt;
  if (lane == 0) { T->M = M; T->walked += walked; }
  jsyncwarp();
  return 0;
}

// After the generic phases of a gate: the new edges at the cursor level.
template <int MODE> JD uint32_t cursor_finish(Ctx &c, int lsw, uint32_t blocked) {
  ThinState *T = c.T;
  const uint32_t lane = c.ltid;
  if (blocked) return cursor_up_step<MODE>(c, lsw + 1, true);
  const uint32_t cnt = T->cnt[lsw];
  const bool act = lane < cnt;
  ThinRes r;
  r.rw = 0; r.rnode = 0; r.pad = 0;
  uint64_t irw = 0;
  uint32_t irn = 0;
  if (act) {
    r = T->res[lsw & 1][lane];
    if (r.rw != 0) { irw = ld_i<MODE>(&c.items[lane].rw); irn = ld_i<MODE>(&c.items[lane].rnode); }
  }
  uint64_t nw = 0;
  uint32_t rounds = 0;
  for (;;) {
    bool ok = true;
    if (act && irw != 0) ok = cmul<MODE>(c, r.rw, irw, &nw);
    const uint32_t bad = jballot(!ok);
    if (!bad) break;
    const uint32_t fl = thin_trouble<MODE>(c);
    if (fl) return fl;
    if (++rounds > JKQ_RETRY_LIMIT) jset_error<MODE>(c.E, c.sh, JKQ_ERR_RETRY_LIMIT);
    thin_commit<MODE>(c);
  }
  if (act) {
    r.rw = irw ? nw : 0; r.rnode = r.rw ? irn : 0;
    T->res[lsw & 1][lane] = r;
  }
  jsyncwarp();
  return 0;
}

#endif  // JKQ_THIN_CUH
