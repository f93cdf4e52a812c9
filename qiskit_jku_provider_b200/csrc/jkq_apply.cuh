// Phase drivers of the B200 QMDD engine: barriers, level-synchronous gate application (thin warp
// walk above the gate, top-down expand / bottom-up reduce of work items below), copying collector,
// state initialisation and the bottom-up probability pass that prepares shot sampling.
// See jkq_engine.cuh for the MODE / JKQ_EMU conventions.
#ifndef JKQ_APPLY_CUH
#define JKQ_APPLY_CUH

#include "jkq_engine.cuh"
#include "jkq_thin.cuh"
#include "jkq_grid.cuh"

#if defined(JKQ_PROFILE) && !defined(JKQ_EMU)
#define JPROF_T(var) long long var = clock64()
#define JPROF_ADD(c, slot, t0, t1) do { if ((c).tid == 0) prof_add<MODE>((c).E, (c).sh, slot, (unsigned long long)((t1) - (t0))); } while (0)
#define JPROF_INC(c, slot) do { if ((c).tid == 0) prof_add<MODE>((c).E, (c).sh, slot, 1); } while (0)
#else
#define JPROF_T(var) do {} while (0)
#define JPROF_ADD(c, slot, t0, t1) do {} while (0)
#define JPROF_INC(c, slot) do {} while (0)
#endif
// prof slots: 0 thin descent (+ gate set-up)  1 thin ascent (+ root update)  2 expand work  3 reduce work
//             4 barrier wait  5 value commit  6 gates  7 generic phases  8 gate total  9 collector

// ------------------------------------------------------------------------------------ barriers
// Barrier that also OR-reduces a small flag word, so every thread takes the same branch afterwards.
// One flag word per phase in a ring of three: word r is written during phase p (p % 3 == r), read
// right after the barrier, and cleared -- together with the pending-value counter of the same ring
// slot -- by thread 0 during phase p+2, when nobody can be looking at it any more.
#ifdef JKQ_EMU
template <int MODE> JD uint32_t sync_or(Ctx &c, uint32_t flags) {
  const uint32_t nxt = c.ring == 2 ? 0 : c.ring + 1;
  *cnt_pend<MODE>(c, nxt) = 0;
  c.phase++; c.ring = nxt;
  if (c.E->status) flags |= JF_ERROR;
  if (MODE == MODE_CTA && c.sh->want_grid) flags |= JF_GRID;
  return flags;
}
JD void block_sync() {}
#else
JD void block_sync() { __syncthreads(); }

// MODE_CTA: one __syncthreads.  MODE_GRID: single-word arrival counter in HBM (see below); the kernel is
// launched cooperatively (all CTAs co-resident).
template <int MODE> JD uint32_t sync_or(Ctx &c, uint32_t flags) {
  const uint32_t slot = c.ring, nxt = slot == 2 ? 0 : slot + 1;
  uint32_t r;
  if (MODE == MODE_CTA) {
    CtaState *S = c.sh;
    // sticky words raised by the table functions: folded in BEFORE the barrier (a thread sees its own
    // failure), never read after it -- the cursor warp may already be raising them for the next gate
    if (*(volatile uint32_t *)&S->err) flags |= JF_ERROR;
    if (*(volatile uint32_t *)&S->want_grid) flags |= JF_GRID;
    if (flags) atomicOr(&S->bflags[slot], flags);
    if (c.ltid == 0) { S->bflags[nxt] = 0; S->pend_n[nxt] = 0; }
    __syncthreads();
    r = *(volatile uint32_t *)&S->bflags[slot];
  } else {
    __shared__ uint32_t s_flags;
    JEng *E = c.E;
    r = 0;
    if (__syncthreads_or(flags != 0)) {
      if (__syncthreads_or(flags & JF_RETRY)) r |= JF_RETRY;
      if (__syncthreads_or(flags & JF_ERROR)) r |= JF_ERROR;
      if (__syncthreads_or(flags & JF_GRID)) r |= JF_GRID;
    }
    if (threadIdx.x == 0) {
      if (blockIdx.x < 1024) E->dbg_phase[blockIdx.x] = c.phase;
      if (r) atomicOr(&E->bar_flags[slot], r);
      // One atomic per CTA on one word (the scheme of cooperative groups' grid.sync): CTA 0 adds what is
      // missing to flip the top bit once everybody has arrived; everybody spins until the bit differs from
      // the value its own arrival saw.  2.3k cycles on B200 against 4.1k for a count + generation pair.
      if (blockIdx.x == 0) { E->bar_flags[nxt] = 0; E->pend_n[nxt] = 0; }  // next phase's word and counter: nobody uses them now
      const unsigned int add = blockIdx.x == 0 ? 0x80000000u - (gridDim.x - 1u) : 1u;
      __threadfence();
      const unsigned int old = atomicAdd(&E->bar_count, add);
      long long t0 = 0;
      unsigned int spins = 0;
      while (((old ^ *(volatile unsigned int *)&E->bar_count) & 0x80000000u) == 0) {
        if ((++spins & 1023u) == 0) {
          if (t0 == 0) t0 = clock64();
          else if (clock64() - t0 > 8000000000ll) {
            if (atomicCAS(&E->status, 0, JKQ_ERR_BARRIER_TIMEOUT) == 0) E->dbg_timeout_block = blockIdx.x;
            break;
          }
        }
      }
      __threadfence();
      s_flags = *(volatile unsigned int *)&E->bar_flags[slot];
      if (*(volatile int *)&E->status != 0) s_flags |= JF_ERROR;  // sticky: raised by the table functions or a timeout
    }
    __syncthreads();
    r = s_flags;
    __syncthreads();
  }
  c.phase++; c.ring = nxt;
  return r;
}
#endif

// ------------------------------------------------------------------------------------ level phases
// Work of level L = the items created for it by the level above ([b, e), minus those that belong further
// down) plus the items that chain compression (expand_item) parked for it.
template <int MODE> JD uint32_t defer_count(const Ctx &c, int L) {
  if (MODE != MODE_CTA) return 0;
  const uint32_t n = ld_r<MODE>(&c.sh->defer_n[L]);
  return n < JKQ_SDEFER ? n : JKQ_SDEFER;
}
// item index of work slot j of level L (JI_FAIL: nothing to do in this slot)
template <int MODE> JD uint32_t level_item(const Ctx &c, int L, uint32_t b, uint32_t e, uint32_t j) {
  uint32_t i;
  if (j < e - b) i = b + j;
  else {
    if (MODE != MODE_CTA) return JI_FAIL;
    i = c.sh->defer[L][j - (e - b)];
    if (i == 0xFFFFu) return JI_FAIL;
  }
  const uint32_t lvl = (ld_i<MODE>(&c.items[i].tag) >> 8) & 0xFFu;
  return (lvl == 0 || lvl == (uint32_t)L + 1u) ? i : JI_FAIL;  // (an item of [b, e) parked for a lower level)
}

// thread index in which consecutive warps lie on different CTAs (the grid path: different SMs)
template <int MODE> JD uint32_t spread_id(const Ctx &c) {
#ifdef JKQ_EMU
  return c.tid;
#else
  if (MODE == MODE_CTA) return c.ltid;
  return ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * JKQ_WARP + (threadIdx.x & (JKQ_WARP - 1u));
#endif
}

// One level of the generic path: run every item of [b, e), and while some item met a value that is
// not in the table yet, commit the proposed values (vals_commit) and run those items again.
template <int MODE, bool expand>
JD uint32_t run_level(Ctx &c, const JGate &g, uint32_t b, uint32_t e, int L) {
#if defined(JKQ_PROFILE) && !defined(JKQ_EMU)
  struct WProf {  // level phases of the grid path by width class
    Ctx &c; long long t0; uint32_t n, ph0;
    __device__ WProf(Ctx &c_, uint32_t n_) : c(c_), t0(clock64()), n(n_), ph0(c_.phase) {}
    __device__ ~WProf() {
      if (MODE == MODE_GRID && c.tid == 0) {
        int cls = 0;
        for (uint32_t x = n; x > 4 && cls < 7; x >>= 2) cls++;
        c.E->prof_w[cls] += (unsigned long long)(clock64() - t0); c.E->prof_w[8 + cls] += 1;
        c.E->prof_w[16 + cls] += c.phase - ph0; c.E->prof_w[24 + cls] += n;
      }
    }
  } wprof(c, e - b);
#endif
  for (uint32_t round = 0;; round++) {
    uint32_t my = 0;
    JPROF_T(pt0);
    const uint32_t nwork = e - b + defer_count<MODE>(c, L);
    // Lanes of one warp that work on different items take different paths through expand_item / reduce_item and run one
    // after the other (r160, profile build: a level of at most 64 items cost 48 k cycles with item j on thread j).  A
    // level narrower than the grid is therefore spread out: one item per warp while there are enough warps, consecutive
    // warps on different SMs.
    uint32_t stride = 1;
    while (stride < JKQ_WARP && nwork * stride * 2u <= c.nth) stride <<= 1;
    const uint32_t sid = spread_id<MODE>(c);
    // more warps than items: a whole warp per item (jkq_grid.cuh: the parts of the item's arithmetic and its two de-dup
    // inserts on different lanes) -- a narrow level costs the instruction chain of one item, which this cuts
    const bool wmode = MODE == MODE_GRID && JKQ_GRID_ILP && JKQ_GRID_WARP && JKQ_WARP > 1u && stride == JKQ_WARP;
    JGP_SET(c, MODE == MODE_GRID && nwork <= 32u && (!wmode || (sid & (JKQ_WARP - 1u)) == 0));
    for (uint32_t j = (!wmode && (sid & (stride - 1u))) ? nwork : sid / stride; j < nwork; j += c.nth / stride) {
      if (MODE == MODE_GRID && JKQ_GRID_ILP) {  // jkq_grid.cuh: the item's first quad in one load, the staged item steps
        JItem *it = c.items + b + j;
        JGP_T(gp0);
        const JQuad q0 = ldq(it);  // a, b, tag, flags
        const uint32_t lvl = (q0.z >> 8) & 0xFFu;
        if (lvl != 0 && lvl != (uint32_t)L + 1u) continue;  // (parked for a lower level)
        uint32_t fl = q0.w;
        const uint32_t st = fl & 3u;
        if (st == JS_DEAD) continue;
        if (round > 0 && st != JS_RETRY) continue;
        JGP_T(gp1);
        JGP_ADD(c, 0, gp0, gp1);
#ifndef JKQ_EMU
        if (wmode) {
          if (expand) fl = expand_item_g<true>(c, g, b + j, L, q0.x, q0.y, q0.z);
          else fl = (fl & ~3u) | (reduce_item_g<true>(c, b + j, fl) ? JS_DONE : JS_RETRY);
        } else
#endif
        if (expand) fl = expand_item_g<false>(c, g, b + j, L, q0.x, q0.y, q0.z);
        else fl = (fl & ~3u) | (reduce_item_g<false>(c, b + j, fl) ? JS_DONE : JS_RETRY);
        JGP_T(gp2);
        JGP_ADD(c, expand ? 1 : 2, gp1, gp2);
        if (!wmode || (sid & (JKQ_WARP - 1u)) == 0) it->flags = fl;
        if ((fl & 3u) == JS_RETRY) my |= JF_RETRY;
        continue;
      }
      const uint32_t i = level_item<MODE>(c, L, b, e, j);
      if (i == JI_FAIL) continue;
      JItem *it = c.items + i;
      uint32_t fl = ld_i<MODE>(&it->flags);
      const uint32_t st = fl & 3u;
      if (st == JS_DEAD) continue;
      if (round > 0 && st != JS_RETRY) continue;
      if (expand) fl = expand_item<MODE>(c, g, i, L);
      else fl = (fl & ~3u) | (reduce_item<MODE>(c, i, fl) ? JS_DONE : JS_RETRY);
      it->flags = fl;
      if ((fl & 3u) == JS_RETRY) my |= JF_RETRY;
    }
    JPROF_T(pt1);
#if defined(JKQ_PROFILE) && !defined(JKQ_EMU)
    if (MODE == MODE_GRID && (sid & (stride - 1u)) == 0 && sid / stride < nwork) {  // this thread had at least one item
      const int slot = round > 0 ? 4 : (my ? 2 : 0);
      atomicAdd(&c.E->prof_x[slot], (unsigned long long)(pt1 - pt0));
      atomicAdd(&c.E->prof_x[slot + 1], 1ull);
    }
#endif
    uint32_t flags = sync_or<MODE>(c, my);
    JPROF_T(pt2);
#if defined(JKQ_PROFILE) && !defined(JKQ_EMU)
    if (MODE == MODE_GRID && c.gprof) {
      if ((sid & (stride - 1u)) == 0 && sid / stride < nwork) { JGP_ADD(c, 3, pt1, pt2); JGP_ADD(c, 5, pt0, pt1); }  // (lane 0 of a warp in wmode: c.gprof)
      if (c.tid == 0) { JGP_ADD(c, 4, pt1, pt2); JGP_ADD(c, 6, pt0, pt2); }
    }
#endif
    JGP_SET(c, 0);
    JPROF_ADD(c, expand ? 2 : 3, pt0, pt1);
    JPROF_ADD(c, 4, pt1, pt2);
    JPROF_INC(c, 7);
    if (flags & JF_ERROR) return JF_ERROR;
    if (flags & JF_GRID) return JF_GRID;
    if (!(flags & JF_RETRY)) return 0;
    if (round > JKQ_RETRY_LIMIT && c.tid == 0) jset_error<MODE>(c.E, c.sh, JKQ_ERR_RETRY_LIMIT);  // (seen at the next barrier)
    vals_commit<MODE>(c.E, c.sh, c.ring == 0 ? 2u : c.ring - 1u, c.tid, c.nth);  // the candidates of the phase just ended
    if (c.tid == 0) jstat<MODE>(c.E, c.sh, JST_RETRY, 1u);
    flags = sync_or<MODE>(c, 0);
    JPROF_T(pt3);
    JPROF_ADD(c, 5, pt2, pt3);
    if (flags & JF_ERROR) return JF_ERROR;
    if (flags & JF_GRID) return JF_GRID;
  }
}

// item ranges of the levels of one gate: shared memory on the CTA path, the control block on the grid
// path (written by the warp that ran the level, read by everybody after a barrier)
template <int MODE> JD void lr_set(Ctx &c, int L, uint32_t b, uint32_t e) {
  if (MODE == MODE_CTA) { c.sh->lb[L] = b; c.sh->le[L] = e; }
  else { c.E->lb[L] = b; c.E->le[L] = e; }
}
template <int MODE> JD uint32_t lr_begin(const Ctx &c, int L) { return MODE == MODE_CTA ? c.sh->lb[L] : ld_s<MODE>(&c.E->lb[L]); }
template <int MODE> JD uint32_t lr_end(const Ctx &c, int L) { return MODE == MODE_CTA ? c.sh->le[L] : ld_s<MODE>(&c.E->le[L]); }

// One level run by the cursor warp alone (at most one item per lane): no CTA / grid barrier, values are
// committed warp-locally.  Returns 0 or barrier flags.
#ifndef JKQ_EMU
// expansion of a level of at most 16 items: two lanes per item, one per child slot (expand_item_pair)
template <int MODE>
JD uint32_t solo_expand_pairs(Ctx &c, const JGate &g, uint32_t b, uint32_t e, int L, uint32_t nwork) {
  const uint32_t lane = c.ltid, slot = lane >> 1, child = lane & 1u, pm = 3u << (lane & ~1u);
  const uint32_t i = slot < nwork ? level_item<MODE>(c, L, b, e, slot) : JI_FAIL;
  JItem *it = c.items + (i == JI_FAIL ? 0 : i);
  bool need = i != JI_FAIL;
  if (need) need = (ld_i<MODE>(&it->flags) & 3u) != JS_DEAD;
  uint32_t rounds = 0;
  for (;;) {
    if (need) {
      const uint32_t fl = expand_item_pair<MODE>(c, g, i, L, child, pm);
      if (child == 0) it->flags = fl;
      need = (fl & 3u) == JS_RETRY;
    }
    const uint32_t bad = jballot(need);
    jsyncwarp();
    const uint32_t tr = thin_trouble<MODE>(c);  // pools exhausted / errors are sticky words
    if (tr) return tr;
    if (!bad) return 0;
    if (++rounds > JKQ_RETRY_LIMIT) jset_error<MODE>(c.E, c.sh, JKQ_ERR_RETRY_LIMIT);  // (seen by thin_trouble next round)
    thin_commit<MODE>(c);
  }
}
#endif

template <int MODE, bool expand>
JD uint32_t solo_level(Ctx &c, const JGate &g, uint32_t b, uint32_t e, int L) {
  const uint32_t nwork = e - b + defer_count<MODE>(c, L);
#if !defined(JKQ_EMU) && !defined(JKQ_NO_PAIRS)
  if (expand && nwork <= JKQ_WARP / 2 && !(MODE == MODE_GRID && JKQ_GRID_ILP)) return solo_expand_pairs<MODE>(c, g, b, e, L, nwork);
#endif
#ifndef JKQ_EMU
  if (MODE == MODE_GRID && JKQ_GRID_ILP && JKQ_GRID_WARP) {
    // the cursor warp alone on the grid path: the items one after the other, all lanes on each (jkq_grid.cuh)
    for (uint32_t j = 0; j < nwork; j++) {
      const uint32_t i = level_item<MODE>(c, L, b, e, j);
      if (i == JI_FAIL) continue;
      JItem *it = c.items + i;
      const JQuad q0 = ldq(it);
      uint32_t fl = q0.w;
      if ((fl & 3u) == JS_DEAD) continue;
      for (uint32_t rounds = 0;; rounds++) {
        if (expand) fl = expand_item_g<true>(c, g, i, L, q0.x, q0.y, q0.z);
        else fl = (fl & ~3u) | (reduce_item_g<true>(c, i, fl) ? JS_DONE : JS_RETRY);
        if (c.ltid == 0) it->flags = fl;
        jsyncwarp();
        const uint32_t tr = thin_trouble<MODE>(c);  // pools exhausted / errors are sticky words
        if (tr) return tr;
        if ((fl & 3u) != JS_RETRY) break;
        if (rounds > JKQ_RETRY_LIMIT) jset_error<MODE>(c.E, c.sh, JKQ_ERR_RETRY_LIMIT);  // (seen by thin_trouble next round)
        thin_commit<MODE>(c);
      }
    }
    return 0;
  }
#endif
  const uint32_t i = c.ltid < nwork ? level_item<MODE>(c, L, b, e, c.ltid) : JI_FAIL;
  JItem *it = c.items + (i == JI_FAIL ? 0 : i);
  bool need = i != JI_FAIL;
  uint32_t fl = 0;
  if (need) { fl = ld_i<MODE>(&it->flags); need = (fl & 3u) != JS_DEAD; }
  uint32_t rounds = 0;
  for (;;) {
    if (need) {
      if (MODE == MODE_GRID && JKQ_GRID_ILP) {
        if (expand) { const JQuad q0 = ldq(it); fl = expand_item_g<false>(c, g, i, L, q0.x, q0.y, q0.z); }
        else fl = (fl & ~3u) | (reduce_item_g<false>(c, i, fl) ? JS_DONE : JS_RETRY);
      } else
      if (expand) fl = expand_item<MODE>(c, g, i, L);
      else fl = (fl & ~3u) | (reduce_item<MODE>(c, i, fl) ? JS_DONE : JS_RETRY);
      it->flags = fl;
      need = (fl & 3u) == JS_RETRY;
    }
    const uint32_t bad = jballot(need);
    jsyncwarp();
    const uint32_t tr = thin_trouble<MODE>(c);  // pools exhausted / errors are sticky words
    if (tr) return tr;
    if (!bad) return 0;
    if (++rounds > JKQ_RETRY_LIMIT) jset_error<MODE>(c.E, c.sh, JKQ_ERR_RETRY_LIMIT);  // (seen by thin_trouble next round)
    thin_commit<MODE>(c);
  }
}

// ------------------------------------------------------------------------------------ gate set-up
// canonical ids of the four matrix entries, cached per distinct matrix of the circuit until the next
// sweep of the reals renumbers them (the reference re-runs Cmake per gate, QASMsimulator.cpp:295-298).
// One thread, nobody else touches the value table.
template <int MODE> JD void gate_setup(Ctx &c, const JGate &g) {
  JEng *E = c.E;
  const uint32_t mid = (uint32_t)g.pad;
  if (MODE == MODE_CTA) {
    CtaState *S = c.sh;
    const uint32_t slot = mid & (JKQ_SMATS - 1);
    if (S->mtag[slot] == mid + 1) {
      for (int k = 0; k < 4; k++) S->cur_u[k] = S->mu[slot][k];
    } else {
      for (int k = 0; k < 4; k++) {
        uint64_t u = cx_make_serial<MODE>(c.E, c.sh, g.m[2 * k], g.m[2 * k + 1]);
        S->cur_u[k] = u; S->mu[slot][k] = u;
      }
      S->mtag[slot] = mid + 1;
    }
    return;  // (the generation was bumped at the end of the previous gate)
  }
  const uint32_t epoch = E->val_epoch;
  const bool cached = E->mat_epoch && ld_s<MODE>(E->mat_epoch + mid) == epoch;
  for (int k = 0; k < 4; k++) {
    uint64_t u = cached ? ld_s<MODE>(E->mat_u + 4 * mid + k) : cx_make_serial<MODE>(c.E, c.sh, g.m[2 * k], g.m[2 * k + 1]);
    if (!cached && E->mat_epoch) E->mat_u[4 * mid + k] = u;
    E->cur_u[k] = u;
  }
  if (!cached && E->mat_epoch) E->mat_epoch[mid] = epoch;
  E->gen++;
}

// ------------------------------------------------------------------------------------ one gate
// Simulator::ApplyGate (Simulator.cpp:466-503) = QMDDmultiply(gate, state) as one top-down and one
// bottom-up sweep over the levels.  Returns 0, JF_ERROR, or JF_GRID (CTA mode only: the gate does not
// fit the shared-memory pools and is left unapplied).
// should the collector run after this gate?  (one thread decides, everybody reads the hand-over word)
template <int MODE> JD uint32_t want_collect(const Ctx &c, bool log) {
  const uint32_t nodes_now = rd_nodes<MODE>(c), vals_now = rd_vals<MODE>(c);
  if (log) return 1;
  if (MODE == MODE_CTA) return nodes_now > c.sh->node_hard / 2 || vals_now > (c.sh->val_hard / 4) * 3;
  return nodes_now > ld_s<MODE>(&c.E->gc_limit);
}

// widest level the cursor warp runs by itself when the option "solo_max" is not set (see JEng.solo_max)
#ifndef JKQ_SOLO_MAX_DEFAULT
#define JKQ_SOLO_MAX_DEFAULT (JKQ_WARP < 8u ? JKQ_WARP : 8u)
#endif
// ... on the grid path, where the warp takes the items of such a level one after the other (solo_level) and a level handed to
// the grid costs one barrier
#ifndef JKQ_SOLO_MAX_GRID
#define JKQ_SOLO_MAX_GRID (JKQ_WARP < 2u ? JKQ_WARP : 2u)
#endif
// SUMS: the kernel carries thin_probe_sums.  The circuits whose cost is the latency of one warp are sensitive to the mere
// size of the code around their path (QFT-32: +3 to +6 % with it), so the single-CTA kernel exists without it as well; that
// one stops once it keeps meeting gates that want it (JEng.sums_ask / want_sums) and the host launches the other.
template <int MODE, bool SUMS> JD uint32_t apply_gate(Ctx &c, const JGate *gates, int gi, int n_gates, int n, bool log) {
  JEng *E = c.E;
  const bool thin_warp = c.block0 && c.ltid < JKQ_WARP;
  JPROF_T(gt0);
  if (MODE == MODE_CTA) {
    CtaState *S = c.sh;
    // stage the NEXT gate's descriptor while this one runs: the warps other than the cursor warp when there are any;
    // a single warp issues asynchronous 16-byte copies (cp.async: no register, no stall) and waits for them at the end
    // of the gate
    if (gi + 1 < n_gates) {
      uint32_t *dst = reinterpret_cast<uint32_t *>(&S->gate[(gi + 1) & 1]);
      const uint32_t *src = reinterpret_cast<const uint32_t *>(gates + gi + 1);
#ifndef JKQ_EMU
      if (c.lnth <= JKQ_WARP) {
        if (c.ltid < sizeof(JGate) / 16)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + 4 * c.ltid)), "l"(src + 4 * c.ltid) : "memory");
      } else
#endif
      if (c.lnth <= JKQ_WARP || c.ltid >= JKQ_WARP) {
        const uint32_t first = c.lnth <= JKQ_WARP ? c.ltid : c.ltid - JKQ_WARP, step = c.lnth <= JKQ_WARP ? c.lnth : c.lnth - JKQ_WARP;
        for (uint32_t i = first; i < sizeof(JGate) / 4; i += step) dst[i] = src[i];
      }
    }
  }
  const JGate &g = MODE == MODE_CTA ? c.sh->gate[gi & 1] : gates[gi];
  // ---- the cursor warp: set-up, cursor move, and -- while no level holds more than one warp of items --
  // the whole gate by itself (solo), without a barrier
  JHand *hand = MODE == MODE_CTA ? &c.sh->hand[gi & 1] : &E->hand[gi & 1];
  if (thin_warp) {
    if (MODE == MODE_CTA && c.ltid == 0) c.sh->solo_now = 1;  // (everybody else is at the gate barrier)
    if (MODE == MODE_CTA && c.sh->gen >= 0xFFF0u) {  // generation tags about to wrap: clear the de-dup tables (nobody else is using them)
      for (uint32_t i = c.ltid; i < JKQ_SNODES; i += JKQ_WARP) c.sh->mdir[i] = 0;
      for (uint32_t i = c.ltid; i < JKQ_SDT_SLOTS; i += JKQ_WARP) c.sh->dtab[i] = 0;
      jsyncwarp();
      if (c.ltid == 0) c.sh->gen = 1;
      jsyncwarp();
    }
    if (MODE == MODE_CTA && c.sh->chain_dirty) {  // chain-compression scratch of the last gate that went through the item phases
      for (uint32_t L = c.ltid; L < (uint32_t)n; L += JKQ_WARP) c.sh->defer_n[L] = 0;
      if (c.ltid == 0) { c.sh->chain_n = 0; c.sh->min_def = 0x7FFFFFFF; c.sh->chain_dirty = 0; }
    }
    if (c.ltid == 0) gate_setup<MODE>(c, g);
    jsyncwarp();
    int lsw = 0, wide_L = -1;
    uint32_t icount = 0, blocked = 0;
    const uint32_t solo_opt = ld_s<MODE>(&E->solo_max);
    const uint32_t solo_max = solo_opt && solo_opt < JKQ_WARP ? solo_opt : (MODE == MODE_GRID && JKQ_GRID_ILP && JKQ_GRID_WARP ? JKQ_SOLO_MAX_GRID : JKQ_SOLO_MAX_DEFAULT);
    JPROF_T(sp0);
    // Where the cursor goes for this gate: up to the gate's top level if it is below; down only as far as the next
    // few gates need it (JGate.park) when the gate is one thin_gate is bound to take (diagonal / anti-diagonal matrix):
    // u1 on a control far below, then cx back up at the target -- the cursor stays at the target.
    int goal = g.top;
    if ((g.rsv & 1) && !(c.T->thin_off & 16u)) {
      const int M0 = c.T->M;
      if (M0 > goal) goal = M0 < g.park ? M0 : g.park;
    }
    uint32_t f = cursor_prepare<MODE>(c, goal, &lsw, &icount, &blocked);
    JPROF_T(sp1);
    JPROF_ADD(c, 11, sp0, sp1);
    uint32_t thin_done = 0, thin_entries = 0;
#ifndef JKQ_NO_THIN_GATES
    if (!f && !blocked) f = thin_probe<MODE>(c, g, &thin_done, &thin_entries);  // single-child paths, controls on basis-like nodes: no records
    if (!f && thin_done == 2) {  // a row of the target node is a sum of two different sub-diagrams (Grover's H)
      if (SUMS) f = thin_probe_sums<MODE>(c, g, &thin_done, &thin_entries);
      else {
        thin_done = 0;
        // sums_ask counts down: a circuit with a few such gates lets the work items have them (a relaunch costs ~0.15 ms);
        // one that keeps meeting them (Grover: 40 per iteration) gives this gate up, untouched, and the host launches
        // the kernel that has the code
        const uint32_t ask = MODE == MODE_CTA ? ld_r<MODE>(&E->sums_ask) : 0u;
        if (ask == 1u) {
          if (c.ltid == 0) { E->want_sums = 1; jwant_grid(c.sh); }
          jsyncwarp();
          f = JF_GRID;
        } else if (ask > 1u) {
          jsyncwarp();
          if (c.ltid == 0) E->sums_ask = ask - 1u;
        }
      }
    }
    if (!f && !blocked && !thin_done) f = thin_gate<MODE>(c, g, &thin_done, &thin_entries);  // the whole gate below the cursor edges, if it is local
#endif
    if (!f && !thin_done && !blocked && goal > g.top) f = cursor_prepare<MODE>(c, g.top, &lsw, &icount, &blocked);  // the item phases start at the top level
    uint32_t b = 0, e = icount;
    int lowest = lsw + 1;
    if (!f && !thin_done && !blocked) cursor_items<MODE>(c, lsw, icount);
    if (!f && !thin_done) {
      if (c.ltid == 0) { *cnt_items<MODE>(c) = icount; if (MODE == MODE_CTA) c.sh->chain_dirty = 1; }
      jsyncwarp();
      for (int L = lsw; L >= 0 && (e > b || (MODE == MODE_CTA && ld_r<MODE>(&c.sh->min_def) <= L)); L--) {
        const uint32_t nwork = e - b + defer_count<MODE>(c, L);
        if (nwork > solo_max) {
          if (MODE == MODE_CTA && c.lnth <= JKQ_WARP && E->cta_wide_ok) {
            // launched with the cursor warp alone (all a narrow diagram ever needs): give the gate up like one that
            // does not fit, and let the host launch the wide shape from here
            if (c.ltid == 0) { E->want_wide = 1; jwant_grid(c.sh); }
            jsyncwarp();
            f = JF_GRID;
            break;
          }
          wide_L = L;
          break;
        }
        if (c.ltid == 0) lr_set<MODE>(c, L, b, e);
        lowest = L;
        if (nwork == 0) continue;  // a level skipped by chain compression
        f = solo_level<MODE, true>(c, g, b, e, L);
        JPROF_INC(c, 15);
        if (f) break;
        b = e; e = rd_items<MODE>(c);
      }
    }
    JPROF_T(sp2);
    JPROF_ADD(c, 12, sp1, sp2);
    if (!f && wide_L < 0 && !thin_done) {
      for (int L = lowest; L <= lsw && !f; L++) {
        const uint32_t lb = lr_begin<MODE>(c, L), le = lr_end<MODE>(c, L);
        if (le - lb + defer_count<MODE>(c, L) == 0) continue;
        f = solo_level<MODE, false>(c, g, lb, le, L);
      }
      JPROF_T(sp3);
      JPROF_ADD(c, 13, sp2, sp3);
      if (!f) f = cursor_finish<MODE>(c, lsw, blocked);
      JPROF_T(sp4);
      JPROF_ADD(c, 14, sp3, sp4);
    }
    if (c.ltid == 0) {
      if (MODE == MODE_CTA && wide_L >= 0) c.sh->solo_now = 0;  // every warp joins in: values go through propose / commit
      hand->lsw = (uint32_t)lsw; hand->blocked = blocked; hand->wide_L = wide_L; hand->wide_b = b; hand->wide_low = lowest;
      hand->collect = want_collect<MODE>(c, log);
      if (!f && wide_L < 0) {
        const uint32_t cnt = thin_done ? thin_entries : rd_items<MODE>(c);
        jstat<MODE>(E, c.sh, JST_ITEMS, cnt);
        jstat_max<MODE>(E, c.sh, cnt);
        jstat<MODE>(E, c.sh, JST_GATES, 1u);
        if (thin_done) jstat<MODE>(E, c.sh, JST_THIN, 1u);
        if (MODE == MODE_CTA && !thin_done) c.sh->gen++;  // read by everybody only after the barrier below
      }
    }
  }
  {
#ifndef JKQ_EMU
    if (MODE == MODE_CTA && c.lnth <= JKQ_WARP) asm volatile("cp.async.wait_all;" ::: "memory");  // the next gate's descriptor
#endif
    uint32_t fl = sync_or<MODE>(c, 0);
    JPROF_T(gt1);
    JPROF_ADD(c, 0, gt0, gt1);
    if (fl & JF_ERROR) return JF_ERROR;
    if (fl & JF_GRID) return JF_GRID;
  }
  const int wide_L = ld_s<MODE>(&hand->wide_L);
  if (wide_L < 0) {  // the gate is done
    JPROF_INC(c, 6);
    return 0;
  }
  // ---- a level outgrew the warp: everybody continues from there
  JPROF_T(gt1b);
  const int lsw = (int)ld_s<MODE>(&hand->lsw);
  const uint32_t blocked = ld_s<MODE>(&hand->blocked);
  const uint32_t item_limit = MODE == MODE_CTA ? c.sh->item_limit : 0xFFFFFFFFu;
  uint32_t b = ld_s<MODE>(&hand->wide_b), e = rd_items<MODE>(c);
  int lowest = ld_s<MODE>(&hand->wide_low);
  if (MODE == MODE_CTA && e - b > item_limit) return JF_GRID;
  for (int L = wide_L; L >= 0 && (e > b || (MODE == MODE_CTA && c.sh->min_def <= L)); L--) {
    if (c.block0 && c.ltid == 0) lr_set<MODE>(c, L, b, e);
    lowest = L;
    uint32_t f = run_level<MODE, true>(c, g, b, e, L);
    if (f) return f;
    b = e; e = rd_items<MODE>(c);  // children of level 0 are terminals, never items
    if (MODE == MODE_CTA && e - b > item_limit) return JF_GRID;
  }
  for (int L = lowest; L <= lsw; L++) {
    uint32_t f = run_level<MODE, false>(c, g, lr_begin<MODE>(c, L), lr_end<MODE>(c, L), L);
    if (f) return f;
  }
  JPROF_T(gt2);
  if (thin_warp) {
    const uint32_t f = cursor_finish<MODE>(c, lsw, blocked);
    if (!f && c.ltid == 0) {
      const uint32_t cnt = rd_items<MODE>(c);
      jstat<MODE>(E, c.sh, JST_ITEMS, cnt);
      jstat_max<MODE>(E, c.sh, cnt);
      jstat<MODE>(E, c.sh, JST_GATES, 1u);
      if (MODE == MODE_CTA) c.sh->gen++;  // read by everybody only after the barrier below
    }
    if (c.ltid == 0) hand->collect = want_collect<MODE>(c, log);
  }
  {
    uint32_t fl = sync_or<MODE>(c, 0);
    JPROF_T(gt3);
    JPROF_ADD(c, 1, gt2, gt3);
    JPROF_ADD(c, 8, gt1b, gt3);
    JPROF_INC(c, 6);
    JPROF_INC(c, 10);
    if (fl & JF_ERROR) return JF_ERROR;
    if (fl & JF_GRID) return JF_GRID;
  }
  return 0;
}

// ------------------------------------------------------------------------------------ MODE_CTA write-back
// The single-CTA path works on shared memory only; nodes and reals created since the last write-back,
// the root edge and the counters go to the HBM arena in bulk: before a collection (which works on the
// arena) and when the kernel ends.  All threads; ends with a barrier.
template <int MODE> JD void cta_flush(Ctx &c) {
  if (MODE != MODE_CTA) return;
  CtaState *S = c.sh;
  JEng *E = c.E;
  const uint32_t nn = S->node_count < S->node_hard ? S->node_count : S->node_hard, nv = S->val_count < S->val_hard ? S->val_count : S->val_hard;
  for (uint32_t i = S->flushed_nodes + c.ltid; i < nn; i += c.lnth) E->nodes[i] = S->nodes[i];
  for (uint32_t i = S->flushed_vals + c.ltid; i < nv; i += c.lnth) E->vals[i] = S->vals[i];
  block_sync();
  if (c.ltid == 0) {
    S->flushed_nodes = nn; S->flushed_vals = nv;
    E->root = S->root; E->rootw = S->rootw; E->next_gate = S->next_gate;
    for (int k = 0; k < JST_N; k++) {
      if (k == JST_MAXF) { if (S->st[k] > E->st[k]) E->st[k] = S->st[k]; }
      else E->st[k] += S->st[k];
      S->st[k] = 0;
    }
    for (int k = 0; k < 16; k++) { E->prof[k] += S->prof[k]; S->prof[k] = 0; }
  }
  block_sync();
}

// ------------------------------------------------------------------------------------ collector
// Mark-compact as a level-by-level copy into the second arena (replaces the ref-count sweep
// QMDDgarbageCollect, QMDDpackage.cpp:791-831, and cleanCtable, QMDDcomplexD.cpp:868-980):
// discover live nodes from the root one level at a time, give them dense new ids in discovery order
// (so every level is contiguous: [lvl_off[L], lvl_off[L] + lvl_cnt[L])), drop unreferenced reals,
// rebuild the hash tables.  It always works on the HBM arena (which is complete: MODE_CTA writes
// through); MODE_CTA then reloads its shared-memory copies and leaves the HBM hash tables alone.
// The HBM frontier pool is idle between gates and serves as scratch.
//   reload      MODE_CTA at kernel start: shared memory is empty, (re)load reals and tables
//   force_vals  sweep / rebuild the value table regardless of the limit
JD void vt_clear_slot(JVSlot *s) { s->vbits = JV_INF; s->keylo = 0; s->id1 = 0; }

// CTA-wide sum, and exclusive rank of a flag among the threads of the CTA (+ the CTA's count); all threads call
#ifdef JKQ_EMU
JD uint32_t block_sum(Ctx &, uint32_t v) { return v; }
JD uint32_t block_scan_flag(Ctx &, bool flag, uint32_t *total) { *total = flag ? 1u : 0u; return 0; }
#else
JD uint32_t block_sum(Ctx &c, uint32_t v) {
  __shared__ uint32_t s_sum;
  if (c.ltid == 0) s_sum = 0;
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
  if ((c.ltid & 31u) == 0 && v) atomicAdd(&s_sum, v);
  __syncthreads();
  const uint32_t r = s_sum;
  __syncthreads();
  return r;
}
JD uint32_t block_scan_flag(Ctx &c, bool flag, uint32_t *total) {
  __shared__ uint32_t s_w[32];
  const uint32_t b = __ballot_sync(0xFFFFFFFFu, flag), lane = c.ltid & 31u, w = c.ltid >> 5, nw = (c.lnth + 31u) >> 5;
  if (lane == 0) s_w[w] = __popc(b);
  __syncthreads();
  uint32_t off = 0, tot = 0;
  for (uint32_t k = 0; k < nw; k++) { const uint32_t x = s_w[k]; if (k < w) off += x; tot += x; }
  __syncthreads();
  *total = tot;
  return off + __popc(b & ((1u << lane) - 1u));
}
#endif

// Out of line, context by value (the caller's stays in registers): gc_call() below takes over the ring index.
// collector / table rebuild: node `id` (not in the table yet, nobody looks anything up meanwhile) into the HBM table
JD void ut_put_hbm(JEng *E, unsigned long long gen, const JNode &nn, uint32_t id) {
  uint32_t h = ut_hash(nn) & E->utab_mask;
  for (;;) {
    unsigned long long sv = ld_r<MODE_GRID>(E->utab + h);
    if ((sv >> 32) != (gen >> 32) && jatomic_cas(E->utab + h, sv, gen | id) == sv) return;
    h = (h + 1) & E->utab_mask;
  }
}
// The part of the HBM value table in use.  Sizing it to the reals that survive a sweep (16 slots each) cut the clears
// to a few MB, but a single gate of a blown-up circuit can add more reals than the whole previous state held
// (random n=20, 200 ops: "value table full" at gate 601) and an open-addressing table cannot grow in the middle of
// a phase -- so the part in use is the whole allocation, cleared only by a sweep of the reals (a handful per run:
// the limit doubles, Simulator.cpp:495-497), not by every collection.  Generation tags as in the unique table do not
// work here: a claimed slot takes atomicMin proposals and would have to be reset first.
JD uint32_t vt_mask_for(const JEng *E, uint32_t) { return E->vtab_mask; }

// Clear the HBM value table before a sweep re-fills it: only the 128-byte lines in which a slot was claimed since the
// last clear (JEng.vbm), found by a scan of the 1-bit-per-line map (4 MB for 2^28 slots) -- writes proportional to the
// reals that were in the table, not to its allocation.  All threads of the grid; the caller puts a barrier behind it.
template <int MODE> JD void vt_clear_used(Ctx &c) {
  JEng *E = c.E;
  if (ld_s<MODE>(&E->vt_dirty_all)) {  // first use of this allocation: everything
    for (uint32_t i = c.tid; i <= E->vtab_mask; i += c.nth) vt_clear_slot(E->vtab + i);
    for (uint32_t i = c.tid; i < ((E->vtab_mask + 1) >> 8) + 1; i += c.nth) E->vbm[i] = 0;
    return;
  }
  const uint32_t nw = ((E->vtab_mask + 1) >> 8) + 1;
  for (uint32_t i = c.tid; i < nw; i += c.nth) {
    uint32_t w = ld_s<MODE>(E->vbm + i);
    if (!w) continue;
    E->vbm[i] = 0;
    while (w) {
      const uint32_t b = jffs(w) - 1u;
      w &= w - 1u;
      JVSlot *line = E->vtab + (((size_t)i * 32u + b) << 3);
      JKQ_UNROLL
      for (int k = 0; k < 8; k++) vt_clear_slot(line + k);
    }
  }
}

template <int MODE> JNI uint32_t gc_run(Ctx c, bool reload, bool force_vals) {
  JEng *E = c.E;
  const int n = (int)E->n;
  JPROF_T(ct0);
  cta_flush<MODE>(c);  // MODE_CTA: the HBM arena is complete from here on
  uint32_t *list = (uint32_t *)E->items;  // list[new_id - 1] = old id
  const uint32_t old_count = rd_nodes<MODE>(c), old_vals = rd_vals<MODE>(c);
  // the reals are only swept when the table outgrew its limit, like the reference's
  // `Ctable.size() > complex_limit` (Simulator.cpp:488-498): dropping an unreferenced value changes
  // which representative later values snap to
  uint32_t vlimit = ld_s<MODE>(&E->val_gc_limit);
  if (MODE == MODE_CTA && vlimit > (c.sh->val_hard / 4) * 3) vlimit = (c.sh->val_hard / 4) * 3;
  const bool clean_vals = force_vals || old_vals > vlimit;
  for (uint32_t i = c.tid; i < old_count; i += c.nth) E->fwd[i] = 0;
  // The HBM unique table is NOT cleared: it starts a new generation (below).  The value table is cleared only when the
  // reals are swept.  (Clearing both on every collection wrote 6.4 GB of zeros on random n=18 / 160 ops, ncu r1:
  // dram__bytes_write 11x the algorithmic traffic.)
  if (MODE != MODE_GRID)
    for (uint32_t i = c.ltid; i < JKQ_SUT_SLOTS; i += c.lnth) c.sh->utab[i] = 0;
  if (clean_vals)
    for (uint32_t i = c.tid; i < old_vals; i += c.nth) E->vfwd[i] = 0;
  if (MODE == MODE_CTA)  // node ids are about to change: memoised chain rebuilds are void
    for (uint32_t i = c.ltid; i < JKQ_SCMEMO; i += c.lnth) c.sh->cmemo[i] = 0;
  if (MODE == MODE_CTA && (clean_vals || reload)) {
    for (uint32_t i = c.ltid; i < JKQ_SVT_SLOTS; i += c.lnth) vt_clear_slot(c.sh->vtab + i);
    for (uint32_t i = c.ltid; i < JKQ_SMATS; i += c.lnth) c.sh->mtag[i] = 0;  // cached matrix ids are void
    for (uint32_t i = c.ltid; i < JKQ_SMEMO; i += c.lnth) c.sh->memo[i][0] = 0;  // and so are memoised results
  }
  if (c.block0 && c.ltid == 0) {
    list[0] = E->root;
    E->gc_count = 1;
  }
  sync_or<MODE>(c, 0);
  if (c.block0 && c.ltid == 0) E->fwd[ld_s<MODE>(&E->root)] = 1;
  uint32_t b = 0, e = 1;
  for (int L = n - 1; L >= 0; L--) {
    if (c.block0 && c.ltid == 0) { E->lvl_off[L] = b + 1; E->lvl_cnt[L] = e - b; }
    if (L > 0)
      for (uint32_t i = b + c.tid; i < e; i += c.nth) {
        JNode nd = ld_s<MODE_GRID>(E->nodes + ld_s<MODE_GRID>(list + i));
        uint32_t ch[2] = {nd_c0(nd), nd.c1};
        for (int k = 0; k < 2; k++)
          if (ch[k] != 0 && jatomic_cas(E->fwd + ch[k], 0u, 0xFFFFFFFFu) == 0) {
            uint32_t pos = jagg_inc(&E->gc_count);
            list[pos] = ch[k];
            E->fwd[ch[k]] = pos + 1;
          }
      }
    sync_or<MODE>(c, 0);
    b = e; e = ld_s<MODE_GRID>(&E->gc_count);
  }
  const uint32_t live = b;  // == gc_count
  // copy nodes with forwarded children; mark the reals they use
  for (uint32_t i = c.tid; i < live; i += c.nth) {
    JNode nd = ld_s<MODE_GRID>(E->nodes + ld_s<MODE_GRID>(list + i));
    uint32_t c0 = nd_c0(nd), c1 = nd.c1;
    JNode nn;
    nn.c0 = (c0 ? ld_s<MODE_GRID>(E->fwd + c0) : 0) | (nd.c0 & 0x80000000u);
    nn.c1 = c1 ? ld_s<MODE_GRID>(E->fwd + c1) : 0;
    nn.w = nd.w;
    E->nodes_alt[i + 1] = nn;
    if (clean_vals) {
      E->vfwd[(uint32_t)(nd.w >> 32) & 0x7FFFFFFFu] = 1;
      E->vfwd[(uint32_t)nd.w & 0x7FFFFFFFu] = 1;
    }
  }
  if (clean_vals && c.block0 && c.ltid == 0) {  // the root weight is referenced too
    const uint64_t rw = E->rootw;
    E->vfwd[(uint32_t)(rw >> 32) & 0x7FFFFFFFu] = 1;
    E->vfwd[(uint32_t)rw & 0x7FFFFFFFu] = 1;
  }
  sync_or<MODE>(c, 0);
  if (clean_vals) {
    // Compact the reals, keeping insertion order (constants 0..3 stay): a two-pass prefix sum, one
    // contiguous chunk per CTA -- count, barrier, then every CTA adds up the counts of the chunks before
    // its own and scatters.
    const uint32_t nblk = MODE == MODE_CTA ? 1u : c.nth / c.lnth, blk = MODE == MODE_CTA ? 0u : c.tid / c.lnth;
    const uint32_t chunk = (old_vals + nblk - 1) / nblk, lo = blk * chunk, hi = lo + chunk < old_vals ? lo + chunk : old_vals;
    uint32_t mine = 0;
    for (uint32_t i = lo + c.ltid; i < hi; i += c.lnth) mine += (i < 4 || ld_s<MODE_GRID>(E->vfwd + i)) ? 1u : 0u;
    const uint32_t total = block_sum(c, mine);
    if (c.ltid == 0) E->gc_bsum[blk] = total;
    sync_or<MODE>(c, 0);
    uint32_t base = 0;
    if (MODE != MODE_CTA)
      for (uint32_t k = 0; k < blk; k++) base += ld_s<MODE_GRID>(E->gc_bsum + k);
    for (uint32_t t = lo; t < hi; t += c.lnth) {
      const uint32_t i = t + c.ltid;
      const bool keep = i < hi && (i < 4 || ld_s<MODE_GRID>(E->vfwd + i));
      uint32_t tile_total;
      const uint32_t pos = base + block_scan_flag(c, keep, &tile_total);
      if (keep) { E->vals_alt[pos] = ld_s<MODE_GRID>(E->vals + i); E->vfwd[i] = pos; }
      base += tile_total;
    }
    sync_or<MODE>(c, 0);
  }
  if (c.block0 && c.ltid == 0) {
    if (clean_vals) {
      const uint32_t nblk = MODE == MODE_CTA ? 1u : c.nth / c.lnth;
      uint32_t nv = 0;
      for (uint32_t k = 0; k < nblk; k++) nv += ld_s<MODE_GRID>(E->gc_bsum + k);
      const uint64_t rw = E->rootw;
      *cnt_vals<MODE>(c) = nv;
      E->val_count = nv;
      E->rootw = ((uint64_t)(((uint32_t)(rw >> 32) & 0x80000000u) | ld_s<MODE_GRID>(E->vfwd + ((uint32_t)(rw >> 32) & 0x7FFFFFFFu))) << 32) |
                 (((uint32_t)rw & 0x80000000u) | ld_s<MODE_GRID>(E->vfwd + ((uint32_t)rw & 0x7FFFFFFFu)));
      E->val_epoch++;  // ids renumbered: cached gate-matrix ids are void
      if (E->val_gc_limit < 2 * nv) E->val_gc_limit *= 2;  // Simulator.cpp:495-497
      if (E->val_gc_limit > E->val_cap / 2) E->val_gc_limit = E->val_cap / 2;
      if (MODE == MODE_GRID) E->vt_cur = vt_mask_for(E, nv);
    }
    if (MODE == MODE_GRID) E->ut_gen++;  // every older unique-table entry now reads as empty
    E->root = 1;
    if (MODE == MODE_CTA) { c.sh->root = 1; c.sh->rootw = E->rootw; c.sh->flushed_nodes = live + 1; c.sh->flushed_vals = *cnt_vals<MODE>(c); }
    *cnt_nodes<MODE>(c) = live + 1;
    E->node_count = live + 1;
    uint32_t lim = 4 * (live + 1);
    if (lim < E->gc_min) lim = E->gc_min;
    if (lim > E->node_cap / 2) lim = E->node_cap / 2;
    E->gc_limit = lim;
    E->st[JST_GC_RUNS]++;
    E->st[JST_GC_COPIED] += live;
    if (live + 1 >= E->node_cap / 2) jset_error_c<MODE>(c, JKQ_ERR_NODE_FULL);
    if (MODE == MODE_CTA) {
      // too big for the shared-memory pools: hand over to the grid path
      if (live + 1 > c.sh->node_hard / 4 || *cnt_vals<MODE>(c) > (c.sh->val_hard / 8) * 7) jwant_grid(c.sh);
    }
  }
  sync_or<MODE>(c, 0);  // (a JF_GRID raised here is picked up below through sh->want_grid)
  const uint32_t nv = rd_vals<MODE>(c);
  const bool fits = MODE != MODE_CTA || !c.sh->want_grid;
  const unsigned long long ugen = MODE == MODE_GRID ? (unsigned long long)ld_s<MODE_GRID>(&E->ut_gen) << 32 : 0ull;
  if (MODE == MODE_GRID && clean_vals) {  // the lines of the value table that hold something
    vt_clear_used<MODE>(c);
    sync_or<MODE>(c, 0);
    if (c.block0 && c.ltid == 0) E->vt_dirty_all = 0;
  }
  if (MODE == MODE_CTA && !fits && c.ltid == 0) c.sh->nofit = 1;
  // copy back with remapped weights, rebuild the unique table
  for (uint32_t i = c.tid; i < live; i += c.nth) {
    JNode nn = ld_s<MODE_GRID>(E->nodes_alt + i + 1);
    if (clean_vals) {
      uint32_t hr = (uint32_t)(nn.w >> 32), hi = (uint32_t)nn.w;
      hr = (hr & 0x80000000u) | ld_s<MODE_GRID>(E->vfwd + (hr & 0x7FFFFFFFu));
      hi = (hi & 0x80000000u) | ld_s<MODE_GRID>(E->vfwd + (hi & 0x7FFFFFFFu));
      nn.w = ((uint64_t)hr << 32) | hi;
    }
    E->nodes[i + 1] = nn;
    if (MODE == MODE_GRID) ut_put_hbm(E, ugen, nn, i + 1);
    else if (fits) {
      c.sh->nodes[i + 1] = nn;
      uint32_t h = ut_hash(nn) & (JKQ_SUT_SLOTS - 1);
      while (jatomic_cas(&c.sh->utab[h], 0u, i + 1) != 0) h = (h + 1) & (JKQ_SUT_SLOTS - 1);
    }
  }
  // the reals: copy the compacted array back; rebuild the value table
  if (clean_vals)
    for (uint32_t i = c.tid; i < nv; i += c.nth) E->vals[i] = ld_s<MODE_GRID>(E->vals_alt + i);
  if (MODE == MODE_GRID) {
    if (clean_vals)
      for (uint32_t i = c.tid + 1; i < nv; i += c.nth) vt_put<MODE>(c.E, c.sh, i, ld_s<MODE_GRID>(E->vals_alt + i));  // 0.0 is never hashed
  } else if (fits && (clean_vals || reload)) {
    const double *srcv = clean_vals ? E->vals_alt : E->vals;
    for (uint32_t i = c.ltid; i < nv; i += c.lnth) {
      double x = ld_s<MODE_GRID>(srcv + i);
      c.sh->vals[i] = x;
      if (i > 0) vt_put<MODE>(c.E, c.sh, i, x);
    }
  }
  uint32_t r = sync_or<MODE>(c, 0) & (JF_ERROR | JF_GRID);
  JPROF_T(ct1);
  JPROF_ADD(c, 9, ct0, ct1);
  return r | (c.ring << 8) | (c.phase << 10);
}
template <int MODE> JD uint32_t gc_call(Ctx &c, bool reload, bool force_vals) {
  const uint32_t r = gc_run<MODE>(c, reload, force_vals);
  c.ring = (r >> 8) & 3u; c.phase = r >> 10;
  return r & (JF_ERROR | JF_GRID);  // JF_GRID (MODE_CTA): the live DD no longer fits the shared-memory pools
}

// ------------------------------------------------------------------------------------ init
// QMDDinit + Simulator::AddVariables (QMDDpackage.cpp:1248-1317, Simulator.cpp:63-96): empty tables,
// the reals 0, 1 and the reference's gate constants 1/2 and FP64 sqrt(2)/2 in its insertion order
// (QMDDpackage.cpp:44-53), then the |0...0> chain.  Runs on the grid path (HBM tables).
template <int MODE> JD uint32_t init_state(Ctx &c) {
  // One thread.  The HBM hash tables are NOT touched: they are marked stale (tables_valid = 0) and get
  // rebuilt by the collector if the grid path ever needs them; the de-dup tables are generation-tagged and
  // the generation keeps counting across circuits, so they are never cleared either.
  JEng *E = c.E;
  if (c.block0 && c.ltid == 0) {
    E->status = 0; E->want_grid = 0; E->next_gate = 0;
    E->item_count = 0; E->pend_n[0] = E->pend_n[1] = E->pend_n[2] = 0;
    E->gen += 2;
    for (int k = 0; k < 12; k++) E->st[k] = 0;
    for (int k = 0; k < 16; k++) E->prof[k] = 0;
    for (int k = 0; k < 32; k++) E->prof_w[k] = 0;
    for (int k = 0; k < 8; k++) E->prof_x[k] = 0;
    JNode t; t.c0 = 0; t.c1 = 0; t.w = 0;
    E->nodes[0] = t;
    // the reals 0, 1 and the reference's gate constants 1/2, FP64 sqrt(2)/2 in its insertion order
    E->vals[0] = 0.0; E->vals[1] = 1.0; E->vals[2] = 0.5; E->vals[3] = 0.70710678118654757;
    E->val_count = 4;
    uint32_t prev = 0;
    for (uint32_t L = 0; L < E->n; L++) {
      JNode nd; nd.c0 = prev; nd.c1 = 0; nd.w = 0;  // child 0 carries ONE, child 1 is the zero edge
      E->nodes[L + 1] = nd;
      prev = L + 1;
    }
    E->node_count = E->n + 1;
    E->root = prev; E->rootw = JW_ONE;
    E->gc_limit = E->gc_min > E->node_cap / 2 ? E->node_cap / 2 : E->gc_min;
    E->val_gc_limit = 10000 < E->val_cap / 2 ? 10000 : E->val_cap / 2;  // complex_limit, Simulator.h:91
    E->tables_valid = 0;
    E->val_epoch += 1;
  }
  return 0;
}

// ------------------------------------------------------------------------------------ mode hand-over
// MODE_CTA prologue: counters and constants into shared memory, then one collection that loads the
// live DD (nodes, unique table, reals, value table) into shared memory.  The HBM hash tables are
// stale from here on (tables_valid = 0).  MODE_GRID prologue: if they are stale, one collection
// rebuilds them.  Returns JF_GRID when the state does not fit the CTA path.
template <int MODE> JD uint32_t mode_enter(Ctx &c) {
  JEng *E = c.E;
  if (MODE == MODE_CTA) {
    CtaState *S = c.sh;
    if (c.ltid == 0) {
      S->item_count = E->item_count; S->node_count = E->node_count; S->val_count = E->val_count;
      S->item_limit = E->cta_item_limit;
      S->node_hard = E->node_cap < JKQ_SNODES ? E->node_cap : JKQ_SNODES;
      if (E->cta_node_limit && E->cta_node_limit < S->node_hard) S->node_hard = E->cta_node_limit;
      S->val_hard = E->val_cap < JKQ_SVALS ? E->val_cap : JKQ_SVALS;
      if (E->cta_val_limit && E->cta_val_limit < S->val_hard) S->val_hard = E->cta_val_limit;
      S->node_soft = S->node_hard - S->node_hard / 8;  // the rest is the reserve of cursor_to_root
      S->val_soft = S->val_hard - S->val_hard / 16;
      if (E->cta_no_reserve) { S->node_soft = S->node_hard; S->val_soft = S->val_hard; }
      S->node_limit = S->node_soft; S->val_limit = S->val_soft;
      S->pend_n[0] = S->pend_n[1] = S->pend_n[2] = 0;
      S->bflags[0] = S->bflags[1] = S->bflags[2] = S->bflags[3] = 0;
      S->err = (uint32_t)E->status; S->want_grid = 0; S->nofit = 0;
      S->gen = 1; S->lsw = 0; S->root = E->root; S->rootw = E->rootw; S->chain_dirty = 1; S->solo_now = 0; S->val_lock = 0;
      S->flushed_nodes = S->node_count; S->flushed_vals = S->val_count;  // shared memory holds nothing yet
      S->next_gate = E->next_gate;
      for (int k = 0; k <= JST_N; k++) S->st[k] = 0;
      for (int k = 0; k < 16; k++) S->prof[k] = 0;
      S->tol = E->tol; S->inv_tol = E->inv_tol;
      E->tables_valid = 0;
    }
    for (uint32_t i = c.ltid; i < JKQ_SNODES; i += c.lnth) S->mdir[i] = 0;
    for (uint32_t i = c.ltid; i < JKQ_SDT_SLOTS; i += c.lnth) S->dtab[i] = 0;
    block_sync();
    // reals beyond the shared table: sweep them first; still too many -> grid path
    const uint32_t r = gc_call<MODE>(c, true, c.sh->val_count > (c.sh->val_hard / 4) * 3);
    return (r & JF_ERROR) ? JF_ERROR : (r & JF_GRID);
  }
  if (!ld_s<MODE>(&E->tables_valid)) {
    if (ld_s<MODE>(&E->cursor_saved)) {
      // No root to collect from (the state is a saved cursor): hash every arena node and every real as
      // they are.  Ids stay what they are; slots that lost an insertion race carry a tombstone.
      const uint32_t nn = ld_s<MODE>(&E->node_count), nv = ld_s<MODE>(&E->val_count);
      if (c.block0 && c.ltid == 0) { E->ut_gen++; E->vt_cur = vt_mask_for(E, nv); }
      sync_or<MODE>(c, 0);
      const unsigned long long ugen = (unsigned long long)ld_s<MODE>(&E->ut_gen) << 32;
      vt_clear_used<MODE>(c);
      sync_or<MODE>(c, 0);
      if (c.block0 && c.ltid == 0) E->vt_dirty_all = 0;
      for (uint32_t i = 1 + c.tid; i < nn; i += c.nth) {
        const JNode nd = ld_s<MODE>(E->nodes + i);
        if (nd.c1 == JN_TOMB) continue;
        ut_put_hbm(E, ugen, nd, i);
      }
      for (uint32_t i = 1 + c.tid; i < nv; i += c.nth) vt_put<MODE>(c.E, c.sh, i, ld_s<MODE>(E->vals + i));
    } else if (gc_call<MODE>(c, false, true) & JF_ERROR) return JF_ERROR;
    if (c.block0 && c.ltid == 0) E->tables_valid = 1;
    sync_or<MODE>(c, 0);
  }
  return 0;
}
template <int MODE> JD void mode_leave(Ctx &c) {
  if (MODE != MODE_CTA) return;
  cta_flush<MODE>(c);
  if (c.ltid == 0) {
    CtaState *S = c.sh;
    c.E->item_count = S->item_count;
    // The shared counters are the truth on this path (the arena's are only refreshed here and by the collector).
    // After a failed allocation they overshoot the pools: clamp.  Not when the collector found the state too large
    // for the pools (nofit): nothing was allocated since, and the counts may legitimately exceed the pool sizes (a
    // state that was built on the grid path) -- clamping them cut live reals / nodes off.
    if (S->nofit) {
      c.E->node_count = S->node_count;
      c.E->val_count = S->val_count;
    } else {
      c.E->node_count = S->node_count < S->node_hard ? S->node_count : S->node_hard;
      c.E->val_count = S->val_count < S->val_hard ? S->val_count : S->val_hard;
    }
  }
}

// ------------------------------------------------------------------------------------ circuit driver
// Applies gates [E->next_gate, n_gates).  Collects when the arena passed its limit, or after every
// gate when a level log was requested (parity tests: reference Active[v], QMDDpackage.h:199-201).
// MODE_CTA, enter = false: the state of the previous call is still in this CTA's shared memory (same kernel, same
// CTA, left with leave = false) -- no reload, no collection; leave = false: nothing is written back to the arena.
template <int MODE, bool SUMS = false> JD void run_gates(Ctx &c, const JGate *gates, int n_gates, bool enter = true, bool leave = true) {
  JEng *E = c.E;
  const int n = (int)E->n;
  const bool thin_warp = c.block0 && c.ltid < JKQ_WARP;
  uint32_t f = 0;
  if (MODE != MODE_CTA || enter) f = mode_enter<MODE>(c);
  else {
    if (c.ltid == 0) {  // the barrier ring starts over with the fresh context of this call
      c.sh->pend_n[0] = c.sh->pend_n[1] = c.sh->pend_n[2] = 0;
      c.sh->bflags[0] = c.sh->bflags[1] = c.sh->bflags[2] = c.sh->bflags[3] = 0;
      c.sh->want_grid = 0;
    }
    block_sync();
  }
  int g = MODE == MODE_CTA ? c.sh->next_gate : ld_s<MODE>(&E->next_gate);
  if (MODE == MODE_CTA && !f && g < n_gates) {  // stage the first gate descriptor
    uint32_t *dst = reinterpret_cast<uint32_t *>(&c.sh->gate[g & 1]);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(gates + g);
    for (uint32_t i = c.ltid; i < sizeof(JGate) / 4; i += c.lnth) dst[i] = src[i];
  }
  if (MODE == MODE_GRID && ld_s<MODE>(&E->cursor_saved)) {  // continue from the cursor a single-CTA kernel left behind
    if (c.block0) {
      const uint32_t *src = reinterpret_cast<const uint32_t *>(E->thin_spill);
      uint32_t *dst = reinterpret_cast<uint32_t *>(c.T);
      for (uint32_t i = c.ltid; i < sizeof(ThinState) / 4; i += c.lnth) dst[i] = ld_s<MODE>(src + i);
    }
  } else if (thin_warp) cursor_reset<MODE>(c, n);
  sync_or<MODE>(c, 0);
  if (MODE == MODE_GRID && c.block0 && c.ltid == 0) E->cursor_saved = 0;
  const bool log = E->lvl_log != 0;
  if (MODE == MODE_GRID && log && !f && g > 0 && g < n_gates) {
    // Level log and a start in the middle of the circuit (hand-over from the single-CTA kernel): that kernel may have
    // applied gate g - 1 without getting to the collection that counts its levels (the rebuild up to the root ran out
    // of the shared pools).  Count them here; a row that was already written comes out the same.
    if (thin_warp) cursor_to_root<MODE>(c, n);
    f = sync_or<MODE>(c, 0) & JF_ERROR;
    if (!f) {
      if (gc_call<MODE>(c, false, false) & JF_ERROR) f = JF_ERROR;
      else {
        if (thin_warp) cursor_reset<MODE>(c, n);
        if (c.block0 && c.ltid == 0)
          for (uint32_t L = 0; L < E->n; L++) E->lvl_log[(size_t)(g - 1) * E->n + L] = E->lvl_cnt[L];
      }
    }
  }
  for (; !f && g < n_gates; g++) {
    f = apply_gate<MODE, SUMS>(c, gates, g, n_gates, n, log);
    if (f) break;
    // Gate g is applied from here on, whatever happens to the collection below: if the rebuild up to the root runs out
    // of the shared pools (JF_GRID), the grid path must continue AFTER this gate (it used to redo it: gate applied twice).
    if (c.block0 && c.ltid == 0) { if (MODE == MODE_CTA) c.sh->next_gate = g + 1; else E->next_gate = g + 1; }
    const bool collect = ld_s<MODE>(MODE == MODE_CTA ? &c.sh->hand[g & 1].collect : &E->hand[g & 1].collect) != 0;
    uint32_t gcf = 0;
    if (collect) {
      if (thin_warp) cursor_to_root<MODE>(c, n);
      f = sync_or<MODE>(c, 0) & (JF_ERROR | JF_GRID);
      if (f) break;
      gcf = gc_call<MODE>(c, false, false);
      if (gcf & JF_ERROR) { f = JF_ERROR; break; }
      if (thin_warp) cursor_reset<MODE>(c, n);
      if (log && c.block0 && c.ltid == 0)
        for (uint32_t L = 0; L < E->n; L++) E->lvl_log[(size_t)g * E->n + L] = E->lvl_cnt[L];
    }
    if (gcf & JF_GRID) { f = JF_GRID; g++; break; }  // MODE_CTA: the live DD outgrew the shared-memory pools
  }
  // The state goes back to the arena as a DD with a root: rebuild the levels above the cursor.  An
  // abandoned gate (JF_GRID) left the cursor describing the state BEFORE that gate; the sticky want_grid
  // word is cleared first, or the first value miss of the rebuild would give up.
  if (!(f & JF_ERROR)) {
    if (MODE == MODE_CTA && (f & JF_GRID)) {
      // The abandoned gate may have stopped in the middle of a phase: value candidates that were claimed (PENDING)
      // but never committed would sit in the shared value table for good -- a later value within tol above one is
      // dropped at every commit in favour of a neighbour that never becomes an entry, and a value of the same bucket
      // is never even put on the commit list (the retry loop of cursor_to_root then never ends).  Drop them.
      for (uint32_t i = c.ltid; i < JKQ_SVT_SLOTS; i += c.lnth)
        if (c.sh->vtab[i].id1 == JV_PENDING) { c.sh->vtab[i].vbits = JV_INF; c.sh->vtab[i].id1 = JV_DROPPED; }
      if (c.ltid == 0) { c.sh->want_grid = 0; c.sh->pend_n[0] = c.sh->pend_n[1] = c.sh->pend_n[2] = 0; }
      block_sync();
    }
    const bool force_spill = MODE == MODE_CTA && (f & JF_GRID) && E->cta_no_reserve == 2;  // testing aid
    if (thin_warp && !force_spill) cursor_to_root<MODE>(c, n);
    const uint32_t r = sync_or<MODE>(c, 0) | (force_spill ? JF_GRID : 0u);
    f |= r & JF_ERROR;
    if (MODE == MODE_CTA && !(r & JF_ERROR) && (r & JF_GRID)) {
      // Not even the reserve of the shared pools was enough to rebuild up to the root: hand the cursor
      // itself to the grid path (which has the room).  The arena goes out as it is.
      uint32_t *dst = reinterpret_cast<uint32_t *>(E->thin_spill);
      const uint32_t *src = reinterpret_cast<const uint32_t *>(c.T);
      for (uint32_t i = c.ltid; i < sizeof(ThinState) / 4; i += c.lnth) dst[i] = src[i];
      if (c.ltid == 0) { E->cursor_saved = 1; c.sh->want_grid = 0; }
      block_sync();
      f |= JF_GRID;
    }
  }
  if ((f & JF_GRID) && c.block0 && c.ltid == 0) { E->want_grid = 1; if (MODE == MODE_GRID) E->gen++; }
  sync_or<MODE>(c, 0);
  if (MODE != MODE_CTA || leave) mode_leave<MODE>(c);
}

// ------------------------------------------------------------------------------------ measurement in shared memory
// MODE_CTA, state rooted in shared memory (after run_gates(..., leave = false)): squared norms bottom-up, the marginal
// of one qubit and one MeasureAll walk WITHOUT going through the arena -- the same arithmetic as build_sampling /
// marginal_block / sample_walk (q0 = p0 / (p0 + p1) from the same products in the same order), on a discovery-ordered
// list of the live nodes.  Scratch lives in the item pool, which is idle between gates.  All threads call.
struct ShotScratch {
  uint32_t pos[JKQ_SNODES];  // node id -> 1 + index in list, 0 = not reached
  uint32_t list[1024];
  double prob[1024];         // squared norm of the sub-vector below list[i]
  double q0[1024];           // P(branch 0 | node)
  double mass[1024];
  uint32_t lvl_b[JKQ_MAXQ + 1];
  uint32_t cnt;
  int fail;
  double sum, norm;
};
static_assert(sizeof(ShotScratch) <= sizeof(JItem) * JKQ_SITEMS, "scratch must fit the shared item pool");

#ifndef JKQ_EMU
// returns false when the live DD has more than 1024 nodes (the caller goes through build_sampling instead)
JD bool cta_norms(Ctx &c, ShotScratch *X) {
  CtaState *S = c.sh;
  const int n = (int)c.E->n;
  const uint32_t nn = S->node_count < JKQ_SNODES ? S->node_count : JKQ_SNODES;
  for (uint32_t i = c.ltid; i < nn; i += c.lnth) X->pos[i] = 0;
  block_sync();
  if (c.ltid == 0) { X->list[0] = S->root; X->pos[S->root] = 1; X->cnt = 1; X->fail = 0; }
  block_sync();
  uint32_t b = 0, e = 1;
  for (int L = n - 1; L >= 0; L--) {
    if (c.ltid == 0) X->lvl_b[L] = b;
    if (L > 0)
      for (uint32_t i = b + c.ltid; i < e; i += c.lnth) {
        const JNode nd = S->nodes[X->list[i]];
        const uint32_t ch[2] = {nd_c0(nd), nd.c1};
        for (int k = 0; k < 2; k++)
          if (ch[k] != 0 && atomicCAS(&X->pos[ch[k]], 0u, 0xFFFFFFFFu) == 0u) {
            const uint32_t p = atomicAdd(&X->cnt, 1u);
            if (p < 1024u) { X->list[p] = ch[k]; X->pos[ch[k]] = p + 1u; }
            else X->fail = 1;
          }
      }
    block_sync();
    // every thread takes the SAME snapshot of the list: the counter and the failure flag are read between two
    // barriers, before any warp starts appending the next level (a slower warp would otherwise see entries of the
    // level after next as part of this one)
    const uint32_t cnt_now = X->cnt;
    const int fail_now = X->fail;
    block_sync();
    b = e;
    e = cnt_now < 1024u ? cnt_now : 1024u;
    if (fail_now) return false;
  }
  // level L occupies [lvl_b[L], lvl_b[L - 1]) (level 0 ends at cnt); bottom-up like build_sampling
  for (int L = 0; L < n; L++) {
    const uint32_t lo = X->lvl_b[L], hi = L == 0 ? X->cnt : X->lvl_b[L - 1];
    for (uint32_t i = lo + c.ltid; i < hi; i += c.lnth) {
      const JNode nd = S->nodes[X->list[i]];
      const uint64_t w0 = nd_w0(nd), w1 = nd_w1(nd);
      const uint32_t c0 = nd_c0(nd), c1 = nd.c1;
      const double p0 = w0 ? JKQ_DMUL(c0 ? X->prob[X->pos[c0] - 1u] : 1.0, jmag2<MODE_CTA>(c, w0)) : 0.0;
      const double p1 = w1 ? JKQ_DMUL(c1 ? X->prob[X->pos[c1] - 1u] : 1.0, jmag2<MODE_CTA>(c, w1)) : 0.0;
      const double tot = JKQ_DADD(p0, p1);
      X->prob[i] = tot;
      X->q0[i] = JKQ_DDIV(p0, tot);
    }
    block_sync();
  }
  if (c.ltid == 0) X->norm = JKQ_DMUL(X->prob[0], jmag2<MODE_CTA>(c, S->rootw));
  block_sync();
  return true;
}

// P(the qubit on `level` reads 0): AssignProbsOne (Simulator.cpp:341-413) on the lists of cta_norms
JD double cta_marginal(Ctx &c, ShotScratch *X, int level) {
  CtaState *S = c.sh;
  const int n = (int)c.E->n;
  const uint32_t total = X->cnt;
  for (uint32_t i = c.ltid; i < total; i += c.lnth) X->mass[i] = 0.0;
  if (c.ltid == 0) X->sum = 0.0;
  block_sync();
  if (c.ltid == 0) X->mass[0] = 1.0;
  block_sync();
  for (int L = n - 1; L > level; L--) {
    const uint32_t lo = X->lvl_b[L], hi = X->lvl_b[L - 1];
    for (uint32_t i = lo + c.ltid; i < hi; i += c.lnth) {
      const double m = X->mass[i];
      if (m == 0.0) continue;
      const JNode nd = S->nodes[X->list[i]];
      const double q = X->q0[i];
      if (q > 0.0) atomicAdd(&X->mass[X->pos[nd_c0(nd)] - 1u], m * q);
      if (q < 1.0) atomicAdd(&X->mass[X->pos[nd.c1] - 1u], m * (1.0 - q));
    }
    block_sync();
  }
  const uint32_t lo = X->lvl_b[level], hi = level == 0 ? X->cnt : X->lvl_b[level - 1];
  double acc = 0.0;
  for (uint32_t i = lo + c.ltid; i < hi; i += c.lnth) acc += X->mass[i] * X->q0[i];
  if (acc != 0.0) atomicAdd(&X->sum, acc);
  block_sync();
  const double r = X->sum;
  block_sync();
  return r;
}
#endif

// ------------------------------------------------------------------------------------ sampling prep
// Simulator::AssignProbs (Simulator.cpp:127-146) once per DD instead of once per shot: after a
// collection the arena is level-ordered, so one bottom-up pass gives every node its squared norm;
// the per-node branch probability of MeasureAll (Simulator.cpp:165-187) is stored next to the child
// ids in a 16-byte sampling record.
template <int MODE> JD uint32_t build_sampling(Ctx &c, JSNode *snodes, double *prob) {
  JEng *E = c.E;
  uint32_t f = mode_enter<MODE>(c);  // MODE_CTA: this already is a collection
  if (f) {
    if ((f & JF_GRID) && c.ltid == 0) E->want_grid = 1;
    sync_or<MODE>(c, 0);
    mode_leave<MODE>(c);
    return f;
  }
  if (MODE == MODE_GRID && (gc_call<MODE>(c, false, false) & JF_ERROR)) return JF_ERROR;
  if (c.block0 && c.ltid == 0) prob[0] = 1.0;
  sync_or<MODE>(c, 0);
  for (uint32_t L = 0; L < E->n; L++) {
    uint32_t off = ld_s<MODE_GRID>(&E->lvl_off[L]), cnt = ld_s<MODE_GRID>(&E->lvl_cnt[L]);
    for (uint32_t i = c.tid; i < cnt; i += c.nth) {
      uint32_t id = off + i;
      JNode nd = ld_s<MODE_GRID>(E->nodes + id);
      uint64_t w0 = nd_w0(nd), w1 = nd_w1(nd);
      uint32_t c0 = nd_c0(nd), c1 = nd.c1;
      double p0 = w0 ? JKQ_DMUL(ld_s<MODE_GRID>(prob + c0), jmag2<MODE>(c, w0)) : 0.0;
      double p1 = w1 ? JKQ_DMUL(ld_s<MODE_GRID>(prob + c1), jmag2<MODE>(c, w1)) : 0.0;
      double tot = JKQ_DADD(p0, p1);
      prob[id] = tot;
      JSNode s;
      s.c0 = c0; s.c1 = c1; s.q0 = JKQ_DDIV(p0, tot);
      snodes[id] = s;
    }
    sync_or<MODE>(c, 0);
  }
  if (c.block0 && c.ltid == 0) E->norm = JKQ_DMUL(ld_s<MODE_GRID>(prob + ld_s<MODE_GRID>(&E->root)), jmag2<MODE>(c, E->rootw));
  uint32_t r = sync_or<MODE>(c, 0);
  mode_leave<MODE>(c);
  return r;
}

#endif  // JKQ_APPLY_CUH
