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
    T->thin_off = c.E->thin_off;
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
  // the cursor is at the gate's top level (cursor_items makes the work items if thin_gate does not take the gate)
  *lsw = M; *icount = cnt;
  if (lane == 0) { T->M = M; T->walked += walked; }
  jsyncwarp();
  return 0;
}
// cursor on the gate's top level: one M item per cursor edge (a zero edge gets a dead item)
template <int MODE> JD void cursor_items(Ctx &c, int M, uint32_t cnt) {
  ThinState *T = c.T;
  const uint32_t lane = c.ltid;
  if (lane < cnt) {
    const ThinRes r = T->res[M & 1][lane];
    item_init_m(c.items + lane, r.rnode);
    if (r.rw == 0) c.items[lane].flags = JS_DEAD;
  }
  jsyncwarp();
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


// ------------------------------------------------------------------------------------ thin gates
// A gate applied directly below the cursor edges, without work items, hash de-duplication or barriers.
//
// With the cursor on the gate's top level, QMDDmultiply2 (QMDDpackage.cpp:1416-1517) only ever meets pairs
// (gate node, ONE state node) as long as, at the target level, every product u_ij * w_j * [c_j] that is not zero
// refers to one and the same child: a diagonal or anti-diagonal matrix (u1, rz, z, s, t, x, cx, controlled phases),
// any matrix on a node whose two edges share their child or that has a zero edge (product and basis-like states: every
// gate of QFT / GHZ).  QMDDadd (:1319-1414) then never has two different operands -- it adds weights.  Such a gate is a
// PARTIAL DESCENT: from the cursor edges down to the gate's lowest level, following only the edges that satisfy the
// controls, one lane per (node, variant), duplicates removed with match/shuffle exactly like the cursor does, then back
// up through reduce_core (same normalisation, same unique table as everything else).
//   above the target   entry = node; a control level forwards only its satisfied edge, the other edge is kept as it is
//   target level       child i = P (u_i0 w0 + u_i1 w1) [c] + (I - P) w_i [c_i]   with P = "all controls below satisfied";
//                      no control below: a finished edge.  Otherwise the pair (p, q) of factors on P and I - P,
//                      divided by the larger one, is a VARIANT (T->vp / T->vq), the child an entry (c, variant)
//   below the target   entry = (node, variant); idle level: both edges forward; control level: the unsatisfied edge
//                      takes q, the satisfied one forwards (or takes p on the gate's lowest level)
// A result that has the children and the weight ratio of the node it came from IS that node (no table access): on a
// basis-state control the whole way up costs a compare per level.
// Anything else (two different operands at the target: H on an entangled node; more than a warp of entries on a
// level) is left to the generic item phases: *handled = 0 and nothing has changed.
#define TGK_S1 5
#define TGK_K0 10
#define TGK_K1 12
#define TGK_PC 14
#define TGK_VALID (1u << 16)
#define TGK_SKIP (1u << 17)  // fast-forwarded level: the entry's child is the entry of the same lane one level down
#define TGK_RUN 18           // ... and the level its run of fast-forwarded levels started on (6 bits)
#define TG_VAR_NONE 0xFFu

// ranks of the distinct REF children of a level (key = node | variant << 32); same scheme as thin_children
struct ThinDedup {
  bool own0, own1;
  uint32_t s0, s1, newcnt;
};
template <int MODE>
JD ThinDedup thin_dedup(const Ctx &c, uint32_t cnt, bool v0ok, unsigned long long key0, bool v1ok, unsigned long long key1) {
  const uint32_t lane = c.ltid;
  ThinDedup d;
  if (cnt == 1) {
    const bool same = v0ok && v1ok && key0 == key1;
    d.own0 = v0ok; d.own1 = v1ok && !same;
    d.s0 = 0; d.s1 = (v0ok && !same) ? 1u : 0u;
    d.newcnt = jshfl((d.own0 ? 1u : 0u) + (d.own1 ? 1u : 0u), 0);
    return d;
  }
  const unsigned long long v0 = v0ok ? key0 : (0x8000000000000000ull | lane);
  const uint32_t lead0 = jffs(jmatch64(v0)) - 1u;
  d.own0 = v0ok && lead0 == lane;
  bool dup1 = false;
  uint32_t src = 0;
  for (uint32_t j = 0; j < cnt; j++) {
    const unsigned long long vj = jshfl64(v0, j);
    if (v1ok && vj == key1) { dup1 = true; src = j; }
  }
  const uint32_t ldsrc = jshfl(lead0, src);
  const unsigned long long v1 = (v1ok && !dup1) ? key1 : (0xC000000000000000ull | lane);
  const uint32_t lead1 = jffs(jmatch64(v1)) - 1u;
  d.own1 = v1ok && !dup1 && lead1 == lane;
  const uint32_t b0 = jballot(d.own0), b1 = jballot(d.own1), n0 = jpopc(b0);
  d.newcnt = n0 + jpopc(b1);
  d.s0 = jpopc(b0 & lanes_below(lead0));
  d.s1 = dup1 ? jpopc(b0 & lanes_below(ldsrc)) : n0 + jpopc(b1 & lanes_below(lead1));
  return d;
}

// reduce_core with the "this is the node it came from" short-cut: `nd` / `node` = the node the two child edges
// (rn, f * a) were derived from
template <int MODE>
JD bool thin_reduce(Ctx &c, uint32_t node, const JNode &nd, uint32_t rn[2], const uint64_t f[2], const uint64_t a[2], uint32_t *out_node,
                    uint64_t *out_w) {
  const uint32_t flag = nd_flag(nd);
  // (a0, a1) still the node's own normalised pair (or its ONE next to a child that became zero)?
  const uint32_t hint = (a[flag] == JW_ONE && (a[1u - flag] == nd.w || a[1u - flag] == 0)) ? 1u + flag : 0u;
  if (rn[0] == nd_c0(nd) && rn[1] == nd.c1) {
    if (hint && a[1u - flag] == nd.w && (f[0] == f[1] || nd.w == 0)) { *out_node = node; *out_w = f[flag]; return true; }
    if (nd.w == 0 && a[1u - flag] == 0 && a[flag] != 0) {  // one edge before and after: the node stays, its factor moves up
      *out_node = node;
      return cmul<MODE>(c, f[flag], a[flag], out_w);
    }
  }
  return reduce_core<MODE>(c, rn, f, a, hint, out_node, out_w);
}

// ------------------------------------------------------------------------------------ linear sums
// al * [A] + be * [B] for two DIFFERENT sub-diagrams on level m, as one lane's loop instead of work items -- for the case
// where QMDDadd (QMDDpackage.cpp:1319-1414) stays a chain: on every level at most ONE of the two children is again a sum
// of two different nodes, the other one is a zero edge, an edge of one operand only, or the same node on both sides (the
// weights are added).  That is H on the node that carries the marked path of a Grover state: |marked> + b |uniform rest>
// shares everything but one path.  The steps are those of the generic path, in its order along the chain -- combine_sum
// (operand order by magnitude, ratio = the (1, y.w / x.w) key normalisation of :1354-1359), the S item's two children
// (expand_child), reduce_core on the way up -- so the values that reach the value table and the nodes that reach the
// unique table are the same; what is gone are the items, their de-dup inserts and two level phases per level.
// Records of the levels passed on the way down live in the cursor's scratch rows (T->ent[level][slot], free below the
// cursor while thin_probe runs).  Result: the edge (rn, f * a) as the parent's reduce_core wants it.  false: not a
// chain, or some value / node could not be entered (sticky words raised) -- nothing of consequence has been written
// (nodes and values that nobody refers to).
#ifndef JKQ_SUMS_ATTR
#define JKQ_SUMS_ATTR JD
#endif
#ifndef JKQ_LIN_ATTR
#define JKQ_LIN_ATTR JD
#endif
#define TL_REF1 (1u << 31)  // ThinEnt.pk of a chain record: index of the child that is the sum further down; pad = node of the other
template <int MODE>
JD bool lin_child(Ctx &c, int mc, uint64_t aw, uint32_t ac, uint64_t bew, uint32_t bc, uint32_t *kind, uint32_t *node, uint64_t *w) {
  // combine_sum up to the point where it would create an item: JC_ZERO / JC_PASS (node, w) / JC_REF
  if (mc < 0 || ac == bc) {
    uint64_t s;
    if (aw == 0) s = bew;
    else if (bew == 0) s = aw;
    else if (!cadd<MODE>(c, aw, bew, &s)) return false;
    *node = (aw == 0) ? bc : ac; *w = s; *kind = s ? JC_PASS : JC_ZERO;
    return true;
  }
  if (aw == 0) { *node = bc; *w = bew; *kind = bew ? JC_PASS : JC_ZERO; return true; }
  if (bew == 0) { *node = ac; *w = aw; *kind = JC_PASS; return true; }
  *kind = JC_REF;
  return true;
}
template <int MODE>
JKQ_LIN_ATTR bool lin_add(Ctx &c, uint32_t slot, int m, uint64_t al, uint32_t A, uint64_t be, uint32_t B, uint32_t *out_rn, uint64_t *out_f, uint64_t *out_a,
                uint32_t *levels) {
  ThinState *T = c.T;
  const int m_top = m;
  uint32_t rnode = 0;
  uint64_t rw = 0, factor = 0;
  for (;;) {  // ---- down: (al, A, be, B) is a sum of two different nodes with both weights non-zero, on level m >= 0
    bool swap = A > B;  // operand order of the memo key, as in combine_sum
    if (al != be && al != jneg(be)) {
      const double ma = jmag2<MODE>(c, al), mb = jmag2<MODE>(c, be);
      if (ma != mb) swap = mb > ma;
    }
    if (swap) { const uint32_t tn = A; A = B; B = tn; const uint64_t tw = al; al = be; be = tw; }
    uint64_t r;
    if (!cdiv<MODE>(c, be, al, &r)) return false;
    // the S item (A, B, 1, r) on level m; its edge carries the factor al
    const JNode na = node_load<MODE>(c, A), nb = node_load<MODE>(c, B);
    uint32_t kd[2], nd[2] = {0, 0};
    uint64_t wd[2] = {0, 0}, bew[2];
    JKQ_UNROLL
    for (int i = 0; i < 2; i++) {
      if (!cmul<MODE>(c, r, i ? nd_w1(nb) : nd_w0(nb), &bew[i])) return false;
      if (!lin_child<MODE>(c, m - 1, i ? nd_w1(na) : nd_w0(na), i ? na.c1 : nd_c0(na), bew[i], i ? nb.c1 : nd_c0(nb), &kd[i], &nd[i], &wd[i])) return false;
    }
    if (kd[0] == JC_REF && kd[1] == JC_REF) return false;  // the sum branches: work items
    if (kd[0] != JC_REF && kd[1] != JC_REF) {  // the chain ends here: this item's reduction
      uint32_t rn[2] = {kd[0] == JC_ZERO ? 0u : nd[0], kd[1] == JC_ZERO ? 0u : nd[1]};
      const uint64_t f[2] = {JW_ONE, JW_ONE}, a[2] = {kd[0] == JC_ZERO ? 0ull : wd[0], kd[1] == JC_ZERO ? 0ull : wd[1]};
      if (!reduce_core<MODE>(c, rn, f, a, 0, &rnode, &rw)) return false;
      if (rw != 0 && rnode == 0 && m >= 0) return false;  // a pool ran out
      factor = al;
      break;
    }
    const uint32_t ri = kd[1] == JC_REF ? 1u : 0u, oi = 1u - ri;
    ThinEnt e;
    e.pk = (ri ? TL_REF1 : 0u) | (kd[oi] == JC_ZERO ? 0u : 1u);
    e.pad = kd[oi] == JC_ZERO ? 0u : nd[oi];
    e.w = kd[oi] == JC_ZERO ? 0ull : wd[oi];
    e.f = al;
    T->ent[m][slot] = e;
    // the child that is a sum again, one level down
    const uint64_t naw = ri ? nd_w1(na) : nd_w0(na);
    const uint32_t nac = ri ? na.c1 : nd_c0(na), nbc = ri ? nb.c1 : nd_c0(nb);
    al = naw; A = nac; be = bew[ri]; B = nbc;
    m--;
  }
  *levels = (uint32_t)(m_top - m + 1);
  for (int L = m + 1; L <= m_top; L++) {  // ---- up: the item of level L over the result of level L - 1
    const ThinEnt e = T->ent[L][slot];
    const uint32_t ri = (e.pk & TL_REF1) ? 1u : 0u, oi = 1u - ri;
    uint32_t rn[2];
    uint64_t f[2], a[2];
    f[ri] = rw; a[ri] = rw ? factor : 0ull; rn[ri] = a[ri] ? rnode : 0u;  // (reduce_item: the child item's result times the slot's factor)
    f[oi] = JW_ONE; a[oi] = e.w; rn[oi] = e.w ? e.pad : 0u;
    if (!reduce_core<MODE>(c, rn, f, a, 0, &rnode, &rw)) return false;
    if (rw != 0 && rnode == 0) return false;  // a pool ran out
    factor = e.f;
  }
  *out_rn = rw ? rnode : 0u; *out_f = rw; *out_a = rw ? factor : 0ull;
  return true;
}

// thin_probe's tail when a row of the target node is such a sum (no control below the target, the cursor on the target
// level: act lanes hold the target nodes).  Out of line: the gates of QFT / GHZ never get here, and their path keeps its
// constant factors and its registers.  Called by the whole warp.
template <int MODE> JKQ_SUMS_ATTR uint32_t thin_probe_sums(Ctx &c, const JGate &g, uint32_t *handled, uint32_t *entries) {
  ThinState *T = c.T;
  const uint32_t lane = c.ltid;
  const int tgt = g.target, start = tgt;  // (thin_probe asks for this only with the cursor on the target level)
  *handled = 0; *entries = 0;
  if (T->M != tgt) return 0;
  ThinRes edge;
  edge.rw = 0; edge.rnode = 0; edge.pad = 0;
  if (lane < T->cnt[start]) edge = T->res[start & 1][lane];
  const bool act = edge.rw != 0;
  JNode nd;
  nd.c0 = 0; nd.c1 = 0; nd.w = 0;
  if (act) nd = node_load<MODE>(c, edge.rnode);
  struct { bool sum[2]; uint64_t sal[2], sbe[2]; uint32_t c0, c1; } ps;
  ps.c0 = nd_c0(nd); ps.c1 = nd.c1;
  uint32_t rn[2] = {0, 0}, sum_levels = 0;
  uint64_t f[2] = {JW_ONE, JW_ONE}, a[2] = {0, 0};
  bool ok_rows = true;
  JKQ_UNROLL
  for (int i = 0; i < 2; i++) {  // the target node's two rows, as in thin_probe
    ps.sum[i] = false; ps.sal[i] = 0; ps.sbe[i] = 0;
    if (!act) continue;
    uint64_t al, be;
    const bool o1 = cmul<MODE>(c, gate_u<MODE>(c, 2 * i), nd_w0(nd), &al), o2 = cmul<MODE>(c, gate_u<MODE>(c, 2 * i + 1), nd_w1(nd), &be);
    if (!(o1 && o2)) { ok_rows = false; continue; }
    if (al != 0 && be != 0 && ps.c0 != ps.c1) { ps.sum[i] = true; ps.sal[i] = al; ps.sbe[i] = be; continue; }
    if (!cadd<MODE>(c, al, be, &a[i])) { ok_rows = false; continue; }
    rn[i] = al != 0 ? ps.c0 : ps.c1;
  }
  if (jballot(!ok_rows)) return 0;
  bool sok = true;
  // one lane per (cursor edge, row) -- lane l + 16 takes row 1 of edge l -- each walking its chain by itself (values go
  // in under the lock of real_insert_solo, nodes through the lock-free unique table)
#ifdef JKQ_EMU
  for (int i = 0; i < 2; i++)
    if (act && ps.sum[i]) {
      uint32_t lv = 0;
      if (!lin_add<MODE>(c, lane, tgt - 1, ps.sal[i], ps.c0, ps.sbe[i], ps.c1, &rn[i], &f[i], &a[i], &lv)) sok = false;
      sum_levels += lv;
    }
#else
  {
    const uint32_t src = lane & (JKQ_WARP / 2 - 1u), row = lane / (JKQ_WARP / 2);
    const bool go0 = act && ps.sum[0], go1 = act && ps.sum[1];
    const uint32_t g0 = jshfl(go0 ? 1u : 0u, src), g1 = jshfl(go1 ? 1u : 0u, src);
    const unsigned long long xa0 = jshfl64(ps.sal[0], src), xa1 = jshfl64(ps.sal[1], src), xb0 = jshfl64(ps.sbe[0], src), xb1 = jshfl64(ps.sbe[1], src);
    const uint32_t xc0 = jshfl(ps.c0, src), xc1 = jshfl(ps.c1, src);
    uint32_t res_rn = 0, lv = 0;
    uint64_t res_f = 0, res_a = 0;
    bool mine_ok = true;
    if (row ? g1 != 0 : g0 != 0) mine_ok = lin_add<MODE>(c, lane, tgt - 1, row ? xa1 : xa0, xc0, row ? xb1 : xb0, xc1, &res_rn, &res_f, &res_a, &lv);
    jsyncwarp();
    const uint32_t up = src + JKQ_WARP / 2;  // the lane that ran my row 1
    const uint32_t rn1 = jshfl(res_rn, up), ok1 = jshfl(mine_ok ? 1u : 0u, up), lv1 = jshfl(lv, up);
    const unsigned long long f1 = jshfl64(res_f, up), a1 = jshfl64(res_a, up);
    if (lane < JKQ_WARP / 2) {
      if (go0) { rn[0] = res_rn; f[0] = res_f; a[0] = res_a; sok = mine_ok; sum_levels += lv; }
      if (go1) { rn[1] = rn1; f[1] = f1; a[1] = a1; sok = sok && ok1 != 0; sum_levels += lv1; }
    }
  }
#endif
  {
    const uint32_t fl = thin_trouble<MODE>(c);  // a pool ran out inside a chain: the gate is abandoned as usual
    if (fl) return fl;
  }
  if (jballot(!sok)) return 0;  // not a chain after all: work items
  if (a[0] == 0) rn[0] = 0;
  if (a[1] == 0) rn[1] = 0;
  uint32_t rnode = 0;
  uint64_t rw = 0;
  bool fail = false, ok = true;
  if (act) {
    const bool lane_sum = ps.sum[0] || ps.sum[1];
    if (lane_sum) ok = reduce_core<MODE>(c, rn, f, a, 0, &rnode, &rw);  // (the M item's reduction: no hint)
    else {
      ok = thin_reduce<MODE>(c, edge.rnode, nd, rn, f, a, &rnode, &rw);  // (a lane without a sum next to one with)
    }
    if (ok && !fail && rw != 0) {
      if (rnode == 0) {}  // (a pool ran out: seen below)
      else ok = cmul<MODE>(c, edge.rw, rw, &rw);
    }
  }
  {
    const uint32_t fl = thin_trouble<MODE>(c);
    if (fl) return fl;
  }
  if (jballot(!ok || fail)) return 0;  // (a value that is neither there nor entered at once: builds without in-place insertion)
  if (act) {
    ThinRes r;
    r.rw = rw; r.pad = 0;
    r.rnode = rw == 0 ? 0u : rnode;
    T->res[start & 1][lane] = r;
  }
  jsyncwarp();
  *handled = 1;
#ifndef JKQ_EMU
  *entries = jpopc(jballot(act)) + __reduce_add_sync(0xFFFFFFFFu, sum_levels);
#else
  *entries = (act ? 1u : 0u) + sum_levels;
#endif
  return 0;
}

// thin_probe: the cases of thin_gate that need no records at all, every cursor edge in its own lane with nothing
// exchanged between lanes:
//   * no control above the target, and from the cursor edge down to the target only nodes whose non-zero edges lead
//     to one child (or the cursor already sits on the target level);
//   * the target node local as in thin_gate;
//   * every control below the target on a node with ONE non-zero edge, reached through single-child nodes: the
//     sub-diagram under the target is then multiplied by p (all controls read "satisfied") or q (some control does
//     not) and stays the node it is -- child i of the target becomes (c, p_i) or (c, q_i), no division, no table access.
// The new target node comes from thin_reduce.  With the cursor parked above the target the edge only takes the new top
// weight, which needs the target node to come out unchanged (a phase on a basis-state qubit); otherwise, and whenever
// a condition above fails for any lane, nothing is written and thin_gate / the item phases take the gate.
template <int MODE> JD uint32_t thin_probe(Ctx &c, const JGate &g, uint32_t *handled, uint32_t *entries) {
  ThinState *T = c.T;
  const uint32_t lane = c.ltid;
  const int start = T->M, tgt = g.target, low = g.low;
  *handled = 0; *entries = 0;
  if (g.top != tgt || (T->thin_off & 32u)) return 0;
  const uint64_t idle = g.idle;
  const uint32_t cnt = T->cnt[start];
  ThinRes edge;
  edge.rw = 0; edge.rnode = 0; edge.pad = 0;
  if (lane < cnt) edge = T->res[start & 1][lane];
  const bool act = lane < cnt && edge.rw != 0;
  bool fail = false;
  uint32_t node = edge.rnode;
  JNode nd;
  nd.c0 = 0; nd.c1 = 0; nd.w = 0;
  if (act) nd = node_load<MODE>(c, node);
  for (int L = start; L > tgt; L--) {  // down to the target through single-child nodes
    uint32_t only = 0;
    if (act && !fail) {
      if (!chain_node(nd, &only)) fail = true;
      else { node = only; nd = node_load<MODE>(c, node); }
    }
  }
  if (jballot(fail)) return 0;
  const uint64_t w0 = nd_w0(nd), w1 = nd_w1(nd);
  const uint32_t c0 = nd_c0(nd), c1 = nd.c1;
  uint32_t rn[2] = {0, 0}, rounds = 0;
  uint64_t p[2] = {0, 0}, a[2] = {0, 0};
  const uint64_t f[2] = {JW_ONE, JW_ONE};
  bool sumfail = false;  // why a lane failed: a row is a sum of two different sub-diagrams (looked at on the way out only)
  for (;;) {  // the target node's two rows
    bool ok = true;
    if (act) {
      JKQ_UNROLL
      for (int i = 0; i < 2; i++) {
        uint64_t al, be;
        const bool o1 = cmul<MODE>(c, gate_u<MODE>(c, 2 * i), w0, &al), o2 = cmul<MODE>(c, gate_u<MODE>(c, 2 * i + 1), w1, &be);
        if (!(o1 && o2)) { ok = false; continue; }
        if (al != 0 && be != 0 && c0 != c1) { fail = true; sumfail = true; continue; }  // a sum of two different sub-diagrams
        if (!cadd<MODE>(c, al, be, &p[i])) { ok = false; continue; }
        rn[i] = al != 0 ? c0 : c1;  // the child p sits on
      }
    }
    const uint32_t bad = jballot(!ok);
    if (!bad) break;
    const uint32_t fl = thin_trouble<MODE>(c);
    if (fl) return fl;
    if (++rounds > JKQ_RETRY_LIMIT) jset_error<MODE>(c.E, c.sh, JKQ_ERR_RETRY_LIMIT);
    thin_commit<MODE>(c);
  }
  if (act && !fail) {
    if (low == tgt) { a[0] = p[0]; a[1] = p[1]; }
    else {
      uint32_t seen = 0, seen_sat = 0;  // the walk below child 0, reused for child 1 when it is the same node
      JKQ_UNROLL
      for (int i = 0; i < 2; i++) {
        const uint64_t q = i ? w1 : w0;
        const uint32_t ci = i ? c1 : c0;
        if (p[i] != 0 && q != 0 && rn[i] != ci) { fail = true; continue; }
        if (p[i] == 0) rn[i] = ci;
        if (p[i] == q) { a[i] = q; continue; }  // (also both zero)
        bool sat = true;
        if (i == 1 && seen == rn[1]) sat = seen_sat != 0;
        else {
          uint32_t z = rn[i];
          for (int Lc = tgt - 1; Lc >= low && !fail; Lc--) {
            const JNode zd = node_load<MODE>(c, z);
            if ((idle >> Lc) & 1ull) {
              uint32_t only = 0;
              if (!chain_node(zd, &only)) fail = true;
              z = only;
              continue;
            }
            const uint32_t si = g.role[Lc] == JR_CPOS ? 1u : 0u;
            const uint64_t ws = si ? nd_w1(zd) : nd_w0(zd), wu = si ? nd_w0(zd) : nd_w1(zd);
            if (ws != 0 && wu != 0) { fail = true; break; }  // a control in superposition splits the sub-diagram
            if (ws == 0) { sat = false; break; }
            z = si ? zd.c1 : nd_c0(zd);
          }
          seen = rn[i]; seen_sat = sat ? 1u : 0u;
        }
        a[i] = sat ? p[i] : q;
      }
    }
    if (a[0] == 0) rn[0] = 0;
    if (a[1] == 0) rn[1] = 0;
  }
  if (jballot(fail)) {
    // Not a gate for this function.  If the reason is a sum of two different sub-diagrams with nothing else in the way (no
    // control below the target, the cursor on the target level, a lane to spare per row), thin_probe_sums can try it as a
    // chain: *handled = 2 tells the caller.  Decided out here so that the gates that do pass carry nothing for it.
    if (low == tgt && start == tgt && tgt > 0 && !(T->thin_off & 64u) && (JKQ_WARP == 1u || cnt <= JKQ_WARP / 2) && jballot(sumfail) != 0) *handled = 2;
    return 0;
  }
  uint32_t rnode = 0;
  uint64_t rw = 0;
  rounds = 0;
  for (;;) {
    bool ok = true;
    if (act && !fail) {
      ok = thin_reduce<MODE>(c, node, nd, rn, f, a, &rnode, &rw);
      if (ok && rw != 0) {
        if (rnode == 0) {}  // (a pool ran out: seen below)
        else if (start > tgt && rnode != node) fail = true;
        else ok = cmul<MODE>(c, edge.rw, rw, &rw);
      }
    }
    const uint32_t bad = jballot(!ok);
    if (bad || jballot(act && ok && !fail && rw != 0 && rnode == 0)) {
      const uint32_t fl = thin_trouble<MODE>(c);
      if (fl) return fl;
    }
    if (!bad) break;
    if (++rounds > JKQ_RETRY_LIMIT) jset_error<MODE>(c.E, c.sh, JKQ_ERR_RETRY_LIMIT);
    thin_commit<MODE>(c);
  }
  if (jballot(fail)) return 0;
  if (act) {
    ThinRes r;
    r.rw = rw; r.pad = 0;
    r.rnode = rw == 0 ? 0u : (start > tgt ? edge.rnode : rnode);
    T->res[start & 1][lane] = r;
  }
  jsyncwarp();
  *handled = 1; *entries = jpopc(jballot(act)) * (uint32_t)(start - low + 1);
  return 0;
}

template <int MODE> JD uint32_t thin_gate(Ctx &c, const JGate &g, uint32_t *handled, uint32_t *entries) {
  ThinState *T = c.T;
  const uint32_t lane = c.ltid;
  const int start = T->M, top = g.top, tgt = g.target, low = g.low;  // the cursor may be parked above the gate: start >= top
  const uint64_t idle = g.idle;
  *handled = 0; *entries = 0;
  const uint32_t off = T->thin_off;
  if (off && ((off & 8u) || ((off & 1u) && top > tgt) || ((off & 2u) && low < tgt))) return 0;
  uint32_t cnt = T->cnt[start], total = 0;
  if ((off & 4u) && cnt > 1) return 0;
  ThinRes edge;
  edge.rw = 0; edge.rnode = 0; edge.pad = 0;
  if (lane < cnt) edge = T->res[start & 1][lane];
  bool act = lane < cnt && edge.rw != 0;
  uint32_t node = edge.rnode, var = TG_VAR_NONE;
  int L = start, lowest = start;
  JNode nd;
  nd.c0 = 0; nd.c1 = 0; nd.w = 0;
  if (act) nd = node_load<MODE>(c, node);
  // ---- down.  Every lane leaves a record on every level (no VALID bit: no entry), so the way up needs no counts.
  for (;;) {  // nd = the node of this lane's entry on level L
    if (L < start && ((idle >> L) & 1ull)) {
      // Fast-forward over idle levels (uninvolved levels between the cursor, the gate's top and its lowest level)
      // where the non-zero edges of EVERY entry's node lead to one child -- product and basis-like states, level
      // after level between a target and its control.  Every entry steps to that child in its own lane: no
      // arithmetic, no de-duplication (entries that meet are merged on the next regular level); the record is the
      // node id and the level the run started on.
      const int Lt = L;
      for (;;) {
        uint32_t only = 0;
        const bool chain = !act || chain_node(nd, &only);
        if (jballot(!chain)) break;
        T->ent[L][lane].pk = (act ? (TGK_VALID | TGK_SKIP) : 0u) | ((uint32_t)Lt << TGK_RUN);
        if (act) { T->ent[L][lane].pad = node; node = only; }
        L--;
        if (act) nd = node_load<MODE>(c, node);
        if (!((idle >> L) & 1ull)) break;
      }
      total += (uint32_t)(Lt - L) * jpopc(jballot(act));
    }
    const uint32_t role = g.role[L];
    const uint64_t w0 = nd_w0(nd), w1 = nd_w1(nd);
    const uint32_t c0 = nd_c0(nd), c1 = nd.c1;
    uint32_t k[2], cn[2] = {0, 0}, cv[2] = {TG_VAR_NONE, TG_VAR_NONE}, pc = 0, rounds = 0;
    uint64_t cw[2];
    bool nonlocal = false;
    for (;;) {
      bool ok = true;
      k[0] = k[1] = JC_ZERO; cw[0] = cw[1] = 0; pc = 0;
      if (act && L > tgt) {
        JKQ_UNROLL
        for (int i = 0; i < 2; i++) {
          const uint64_t wi = i ? w1 : w0;
          if (wi == 0) continue;
          cw[i] = wi;
          if (role_satisfied(role, i)) { k[i] = JC_REF; cn[i] = i ? c1 : c0; }
          else { k[i] = JC_PASS; pc |= (uint32_t)i << i; }
        }
      } else if (act && L == tgt) {
        JKQ_UNROLL
        for (int i = 0; i < 2; i++) {
          uint64_t al, be, p = 0;
          const bool o1 = cmul<MODE>(c, gate_u<MODE>(c, 2 * i), w0, &al), o2 = cmul<MODE>(c, gate_u<MODE>(c, 2 * i + 1), w1, &be);
          if (!(o1 && o2)) { ok = false; continue; }
          if (al != 0 && be != 0 && c0 != c1) { nonlocal = true; continue; }  // a sum of two different sub-diagrams
          if (!cadd<MODE>(c, al, be, &p)) { ok = false; continue; }
          const uint32_t xi = al != 0 ? 0u : 1u;  // the child p sits on
          if (low == tgt) {
            if (p != 0) { k[i] = JC_PASS; pc |= xi << i; cw[i] = p; }
            continue;
          }
          const uint64_t q = i ? w1 : w0;  // on child i
          if (p != 0 && q != 0 && (xi ? c1 : c0) != (i ? c1 : c0)) { nonlocal = true; continue; }
          if (p == 0 && q == 0) continue;
          if (p == q) { k[i] = JC_PASS; pc |= (uint32_t)i << i; cw[i] = q; continue; }  // this row acts as the identity
          uint64_t fac = p, kp, kq;  // common factor: the larger magnitude, p on ties
          if (p == 0 || (q != 0 && q != jneg(p) && jmag2<MODE>(c, q) > jmag2<MODE>(c, p))) fac = q;
          const bool o3 = cdiv<MODE>(c, p, fac, &kp), o4 = cdiv<MODE>(c, q, fac, &kq);
          if (!(o3 && o4)) { ok = false; continue; }
          k[i] = JC_REF; cn[i] = p != 0 ? (xi ? c1 : c0) : (i ? c1 : c0); cv[i] = 2u * lane + (uint32_t)i; cw[i] = fac;
          T->vp[cv[i]] = kp; T->vq[cv[i]] = kq;
        }
      } else if (act) {
        if (role == JR_IDLE) {
          if (w0 != 0) { k[0] = JC_REF; cn[0] = c0; cv[0] = var; cw[0] = w0; }
          if (w1 != 0) { k[1] = JC_REF; cn[1] = c1; cv[1] = var; cw[1] = w1; }
        } else {
          const uint32_t si = role == JR_CPOS ? 1u : 0u, ui = 1u - si;
          const uint64_t ws = si ? w1 : w0, wu = si ? w0 : w1;
          if (wu != 0) {
            uint64_t t;
            if (!cmul<MODE>(c, T->vq[var], wu, &t)) ok = false;
            else if (t != 0) { k[ui] = JC_PASS; pc |= ui << ui; cw[ui] = t; }
          }
          if (ws != 0) {
            if (L > low) { k[si] = JC_REF; cn[si] = si ? c1 : c0; cv[si] = var; cw[si] = ws; }
            else {
              uint64_t t;
              if (!cmul<MODE>(c, T->vp[var], ws, &t)) ok = false;
              else if (t != 0) { k[si] = JC_PASS; pc |= si << si; cw[si] = t; }
            }
          }
        }
      }
      const uint32_t bad = jballot(!ok);
      if (!bad) break;
      const uint32_t fl = thin_trouble<MODE>(c);
      if (fl) return fl;
      if (++rounds > JKQ_RETRY_LIMIT) jset_error<MODE>(c.E, c.sh, JKQ_ERR_RETRY_LIMIT);
      thin_commit<MODE>(c);
    }
    if (jballot(nonlocal)) return 0;  // the generic phases take the gate; nothing has been changed
    const ThinDedup d = thin_dedup<MODE>(c, cnt, act && k[0] == JC_REF, (unsigned long long)cn[0] | ((unsigned long long)cv[0] << 32),
                                         act && k[1] == JC_REF, (unsigned long long)cn[1] | ((unsigned long long)cv[1] << 32));
    if (d.newcnt > JKQ_WARP || ((off & 4u) && d.newcnt > 1)) return 0;
    if (act) {
      ThinEnt e;
      // (the ranks of children that are not entries are meaningless and may exceed five bits)
      e.pk = (k[0] == JC_REF ? d.s0 : 0u) | ((k[1] == JC_REF ? d.s1 : 0u) << TGK_S1) | (k[0] << TGK_K0) | (k[1] << TGK_K1) | (pc << TGK_PC) | TGK_VALID;
      e.pad = node; e.w = cw[0]; e.f = cw[1];
      T->ent[L][lane] = e;
    } else T->ent[L][lane].pk = 0;
    total += jpopc(jballot(act));
    if (d.newcnt == 0) { lowest = L; break; }
    ThinRes nx;
    nx.rw = 0;
    if (d.own0) { nx.rnode = cn[0]; nx.pad = cv[0]; T->sres[(L - 1) & 1][d.s0] = nx; }
    if (d.own1) { nx.rnode = cn[1]; nx.pad = cv[1]; T->sres[(L - 1) & 1][d.s1] = nx; }
    jsyncwarp();
    cnt = d.newcnt;
    act = lane < cnt;
    L--;
    if (act) { nx = T->sres[L & 1][lane]; node = nx.rnode; var = nx.pad; nd = node_load<MODE>(c, node); }
    jsyncwarp();
  }
  // ---- up.  (my_node, my_w) = this lane's result one level down; my_same: it is the node the lane's entry there started from
  uint32_t my_node = 0;
  uint64_t my_w = 0;
  bool my_same = false, unpublished = false;
  for (L = lowest; L <= start; L++) {
    const uint32_t pk = T->ent[L][lane].pk;
    const bool on = (pk & TGK_VALID) != 0;
    if (jballot(on && !(pk & TGK_SKIP)) == 0) {
      // Fast-forwarded level: the lane's own result from below under the old node -- which stays the old node while
      // nothing below it changed: then the whole run is stepped over at once.  Otherwise the node is rebuilt over the
      // new child (both edges take the same factor, so the normal form is the node's own and the factor moves up).
      unpublished = true;
      if (jballot(on && my_w != 0 && !my_same) == 0) {
        const int Lt = (int)((pk >> TGK_RUN) & 63u);  // (every lane holds it, entry or not)
        if (on) my_node = my_w ? T->ent[Lt][lane].pad : 0;
        L = Lt;
        continue;
      }
      if (on) {
        const uint32_t old = T->ent[L][lane].pad;
        if (my_w == 0) my_node = 0;
        else if (my_same) my_node = old;
        else {
          const JNode od = node_load<MODE>(c, old);
          const uint32_t fl = nd_flag(od);
          my_node = ut_find_or_insert<MODE>(c.E, c.sh, ((fl == 1 && od.w == 0) ? 0u : my_node) | (fl << 31), (fl == 0 && od.w == 0) ? 0u : my_node, od.w);
          my_same = my_node == old;
        }
      }
      if (jballot(on && my_w != 0 && my_node == 0)) {  // a pool ran out
        const uint32_t fl = thin_trouble<MODE>(c);
        if (fl) return fl;
      }
      continue;
    }
    if (unpublished) {  // the results of the fast-forwarded level below go through the ring like those of any level
      ThinRes r;
      r.rw = my_w; r.rnode = my_node; r.pad = 0;
      T->sres[(L - 1) & 1][lane] = r;
      jsyncwarp();
      unpublished = false;
    }
    ThinEnt e;
    e.pk = pk; e.pad = 0; e.w = 0; e.f = 0;
    nd.c0 = 0; nd.c1 = 0; nd.w = 0;
    uint32_t rn[2] = {0, 0};
    uint64_t f[2] = {JW_ONE, JW_ONE}, a[2] = {0, 0};
    if (on) {
      e = T->ent[L][lane];
      nd = node_load<MODE>(c, e.pad);
      JKQ_UNROLL
      for (int i = 0; i < 2; i++) {
        const uint32_t kind = (e.pk >> (TGK_K0 + 2 * i)) & 3u;
        const uint64_t cwi = i ? e.f : e.w;
        if (kind == JC_PASS) { rn[i] = ((e.pk >> (TGK_PC + i)) & 1u) ? nd.c1 : nd_c0(nd); a[i] = cwi; }
        else if (kind == JC_REF) {
          const ThinRes r = T->sres[(L - 1) & 1][(e.pk >> (TGK_S1 * i)) & 31u];
          f[i] = r.rw; rn[i] = r.rnode;
          a[i] = f[i] ? cwi : 0;
        }
        if (a[i] == 0) rn[i] = 0;
      }
    }
    uint32_t rnode = 0, rounds = 0;
    uint64_t rw = 0;
    for (;;) {
      bool ok = true;
      if (on) {
        ok = thin_reduce<MODE>(c, e.pad, nd, rn, f, a, &rnode, &rw);
        if (ok && L == start && rw != 0 && rnode != 0) ok = cmul<MODE>(c, edge.rw, rw, &rw);
      }
      const uint32_t bad = jballot(!ok);
      if (bad || jballot(on && ok && rw != 0 && rnode == 0)) {  // (a node of 0 under a weight: a pool ran out)
        const uint32_t fl = thin_trouble<MODE>(c);
        if (fl) return fl;
      }
      if (!bad) break;
      if (++rounds > JKQ_RETRY_LIMIT) jset_error<MODE>(c.E, c.sh, JKQ_ERR_RETRY_LIMIT);
      thin_commit<MODE>(c);
    }
    ThinRes r;
    r.rw = rw; r.rnode = rw ? rnode : 0; r.pad = 0;
    my_node = r.rnode; my_w = rw; my_same = on && r.rnode == e.pad;
    if (L == start) { if (on) T->res[start & 1][lane] = r; }
    else T->sres[L & 1][lane] = r;
    jsyncwarp();
  }
  *handled = 1; *entries = total;
  return 0;
}

#endif  // JKQ_THIN_CUH
