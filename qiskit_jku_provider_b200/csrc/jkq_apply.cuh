// Phase drivers of the B200 QMDD engine: barriers, serial value-insert section, level-synchronous
// gate application (top-down expand, bottom-up reduce), copying collector, state initialisation and
// the bottom-up probability pass that prepares shot sampling.  See jkq_engine.cuh for the MODE /
// JKQ_EMU conventions.
#ifndef JKQ_APPLY_CUH
#define JKQ_APPLY_CUH

#include "jkq_engine.cuh"

// ------------------------------------------------------------------------------------ barriers
#ifdef JKQ_EMU
template <int MODE> JD uint32_t sync_or(Ctx &c, uint32_t flags) { c.phase++; return flags; }
JD void block_sync() {}
#else
JD void block_sync() { __syncthreads(); }

JD uint32_t block_or_bits(uint32_t flags) {
  if (!__syncthreads_or(flags != 0)) return 0;
  uint32_t r = 0;
  if (__syncthreads_or(flags & JF_RETRY)) r |= JF_RETRY;
  if (__syncthreads_or(flags & JF_ERROR)) r |= JF_ERROR;
  if (__syncthreads_or(flags & JF_GRID)) r |= JF_GRID;
  return r;
}

// Barrier that also OR-reduces a small flag word, so every thread takes the same branch afterwards.
// MODE_CTA: __syncthreads_or.  MODE_GRID: sense-reversing counter barrier in HBM; the kernel is
// launched cooperatively (all CTAs co-resident), one flag word per barrier phase in a ring of three.
template <int MODE> JD uint32_t sync_or(Ctx &c, uint32_t flags) {
  uint32_t r = block_or_bits(flags);
  if (MODE == MODE_GRID) {
    __shared__ uint32_t s_flags;
    JEng *E = c.E;
    const uint32_t slot = c.phase % 3u;
    if (threadIdx.x == 0) {
      if (blockIdx.x < 1024) E->dbg_phase[blockIdx.x] = c.phase;
      if (r) atomicOr(&E->bar_flags[slot], r);
      unsigned int gen = *(volatile unsigned int *)&E->bar_gen;
      __threadfence();
      if (atomicAdd(&E->bar_count, 1u) == gridDim.x - 1) {
        E->bar_count = 0;
        E->bar_flags[(slot + 1) % 3u] = 0;  // next phase's word: nobody reads or writes it now
        __threadfence();
        atomicAdd(&E->bar_gen, 1u);
      } else {
        long long t0 = clock64();
        while (*(volatile unsigned int *)&E->bar_gen == gen) {
          __nanosleep(64);
          if (clock64() - t0 > 8000000000ll) {
            if (atomicCAS(&E->status, 0, JKQ_ERR_BARRIER_TIMEOUT) == 0) E->dbg_timeout_block = blockIdx.x;
            break;
          }
        }
      }
      __threadfence();
      s_flags = *(volatile unsigned int *)&E->bar_flags[slot];
      if (*(volatile int *)&E->status == JKQ_ERR_BARRIER_TIMEOUT) s_flags |= JF_ERROR;
    }
    __syncthreads();
    r = s_flags;
    __syncthreads();
  }
  c.phase++;
  return r;
}
#endif

// ------------------------------------------------------------------------------------ serial insert
// All values that missed the table in the phase just finished are inserted by ONE thread, in
// ascending order, re-checking each against the table first.  This reproduces the reference's
// "first inserted wins / pairwise more than tol apart" invariant (QMDDcomplexD.cpp:402-437)
// independently of how the GPU scheduled the threads that produced the values.
template <int MODE> JD void serial_insert(Ctx &c) {
  JEng *E = c.E;
  const uint32_t par = (c.phase - 1) & 1u;  // the phase that produced the values ended at the last barrier
  if (c.block0) {
    uint32_t n = ld_s<MODE>(cnt_pend<MODE>(c, par));
    if (n > E->pend_cap) n = E->pend_cap;
    double *p = E->pend;
    if (n > 32) {  // bitonic sort by the whole CTA, padded with +inf
      uint32_t P = 1;
      while (P < n) P <<= 1;
      for (uint32_t i = n + c.ltid; i < P; i += c.lnth) p[i] = 1.0e300;
      block_sync();
      for (uint32_t k = 2; k <= P; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t i = c.ltid; i < P; i += c.lnth) {
            uint32_t ixj = i ^ j;
            if (ixj > i) {
              double a = ld_r(p + i), b = ld_r(p + ixj);
              if ((a > b) == ((i & k) == 0)) { p[i] = b; p[ixj] = a; }
            }
          }
          block_sync();
        }
    }
    if (c.ltid == 0) {
      if (n <= 32)
        for (uint32_t i = 1; i < n; i++) {  // insertion sort
          double x = ld_r(p + i);
          uint32_t j = i;
          for (; j > 0 && ld_r(p + j - 1) > x; j--) p[j] = ld_r(p + j - 1);
          p[j] = x;
        }
      for (uint32_t i = 0; i < n; i++) {
        double x = ld_r(p + i);
        uint32_t idx;
        if (!real_find<MODE>(c, x, &idx)) real_insert_serial<MODE>(E, x);
      }
      *cnt_pend<MODE>(c, par) = 0;
      E->st_retry_rounds++;
    }
  }
  sync_or<MODE>(c, 0);
}

// ------------------------------------------------------------------------------------ level phases
template <int MODE>
JD uint32_t run_level(Ctx &c, const JGate &g, uint32_t b, uint32_t e, int L, bool expand) {
  JEng *E = c.E;
  for (uint32_t round = 0;; round++) {
    uint32_t my = 0;
    for (uint32_t i = b + c.tid; i < e; i += c.nth) {
      JItem *it = c.items + i;
      uint32_t st = ld_s<MODE>(&it->state);
      if (st == JS_DEAD) continue;
      if (round > 0 && st != JS_RETRY) continue;
      bool ok = expand ? expand_item<MODE>(c, g, i, L) : reduce_item<MODE>(c, i);
      it->state = ok ? JS_DONE : JS_RETRY;
      if (!ok) my |= JF_RETRY;
    }
    if (jhas_error<MODE>(c)) my |= JF_ERROR;
    if (MODE == MODE_CTA && c.sh->want_grid) my |= JF_GRID;
    uint32_t flags = sync_or<MODE>(c, my);
    if (flags & JF_ERROR) return JF_ERROR;
    if (flags & JF_GRID) return JF_GRID;
    if (!(flags & JF_RETRY)) return 0;
    serial_insert<MODE>(c);
  }
}

// ------------------------------------------------------------------------------------ one gate
// Simulator::ApplyGate (Simulator.cpp:466-503) = QMDDmultiply(gate, state) as one top-down and one
// bottom-up sweep over the levels.  Returns 0, JF_ERROR, or JF_GRID (CTA mode only: frontier too
// large, gate left unapplied).
template <int MODE> JD uint32_t apply_gate(Ctx &c, const JGate &g) {
  JEng *E = c.E;
  const int n = (int)E->n;
  if (c.block0 && c.ltid == 0) {
    for (int k = 0; k < 4; k++) E->cur_u[k] = cx_make_serial<MODE>(c, g.m[2 * k], g.m[2 * k + 1]);
    E->gen++;
    JItem *it = c.items;
    it->a = E->root; it->b = 0; it->wa = 0; it->wb = 0; it->wc = 0; it->tag = JK_M; it->state = JS_NEW;
    *cnt_items<MODE>(c) = 1;
  }
  if (MODE == MODE_CTA)
    for (uint32_t i = c.ltid; i < JKQ_SDT_SLOTS; i += c.lnth) c.sh->dtab[i] = 0;
  if (sync_or<MODE>(c, jhas_error<MODE>(c) ? JF_ERROR : 0) & JF_ERROR) return JF_ERROR;

  uint32_t lb[JKQ_MAXQ + 1], le[JKQ_MAXQ + 1];  // items of level L: [lb[L], le[L])
  uint32_t b = 0, e = 1;
  int lowest = n;
  for (int L = n - 1; L >= 0 && e > b; L--) {
    lb[L] = b; le[L] = e; lowest = L;
    uint32_t f = run_level<MODE>(c, g, b, e, L, true);
    if (f) return f;
    b = e; e = rd_items<MODE>(c);  // children of level 0 are terminals, never items
    if (MODE == MODE_CTA && e - b > E->cta_item_limit) return JF_GRID;
  }
  for (int L = lowest; L < n; L++) {
    uint32_t f = run_level<MODE>(c, g, lb[L], le[L], L, false);
    if (f) return f;
  }
  if (c.block0 && c.ltid == 0) {
    const JItem *it = c.items;
    uint64_t rw = ld_s<MODE>(&it->rw), w = 0;
    uint32_t rn = ld_s<MODE>(&it->rnode);
    if (rw == JW_ONE) w = E->rootw;
    else if (E->rootw == JW_ONE) w = rw;
    else if (rw != 0) {
      double ar = jre<MODE>(c, rw), ai = jim<MODE>(c, rw), br = jre<MODE>(c, E->rootw), bi = jim<MODE>(c, E->rootw);
      w = cx_make_serial<MODE>(c, JKQ_DSUB(JKQ_DMUL(ar, br), JKQ_DMUL(ai, bi)), JKQ_DADD(JKQ_DMUL(ar, bi), JKQ_DMUL(ai, br)));
    }
    if (w == 0 || rn == 0) jset_error_c<MODE>(c, JKQ_ERR_ZERO_STATE);
    E->root = rn; E->rootw = w;
    unsigned long long cnt = rd_items<MODE>(c);
    E->st_items += cnt;
    if (cnt > E->st_max_frontier) E->st_max_frontier = cnt;
    E->st_gates++;
  }
  if (sync_or<MODE>(c, jhas_error<MODE>(c) ? JF_ERROR : 0) & JF_ERROR) return JF_ERROR;
  return 0;
}

// ------------------------------------------------------------------------------------ collector
// Mark-compact as a level-by-level copy into the second arena (replaces the ref-count sweep
// QMDDgarbageCollect, QMDDpackage.cpp:791-831, and cleanCtable, QMDDcomplexD.cpp:868-980):
// discover live nodes from the root one level at a time, give them dense new ids in discovery order
// (so every level is contiguous: [lvl_off[L], lvl_off[L] + lvl_cnt[L])), drop unreferenced reals,
// rebuild both hash tables.  The frontier pool is idle between gates and serves as scratch.
template <int MODE> JD uint32_t gc_run(Ctx &c) {
  JEng *E = c.E;
  const int n = (int)E->n;
  uint32_t *list = (uint32_t *)E->items;  // list[new_id - 1] = old id
  c.spare_node = 0;  // ids are renumbered: a spare id from before the collection is meaningless
  const uint32_t old_count = rd_nodes<MODE>(c), old_vals = ld_s<MODE>(&E->val_count);
  // the reals are only swept when the table outgrew its limit, like the reference's
  // `Ctable.size() > complex_limit` (Simulator.cpp:488-498): dropping an unreferenced value changes
  // which representative later values snap to
  const bool clean_vals = old_vals > ld_s<MODE>(&E->val_gc_limit);
  for (uint32_t i = c.tid; i < old_count; i += c.nth) E->fwd[i] = 0;
  for (uint32_t i = c.tid; i <= E->utab_mask; i += c.nth) E->utab[i] = 0;
  if (clean_vals) {
    for (uint32_t i = c.tid; i < old_vals; i += c.nth) E->vfwd[i] = 0;
    for (uint32_t i = c.tid; i <= E->vtab_mask; i += c.nth) E->vtab[i].id1 = 0;
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
        JNode nd = ld_s<MODE>(E->nodes + ld_s<MODE>(list + i));
        uint32_t ch[2] = {nd_c0(nd), nd.c1};
        for (int k = 0; k < 2; k++)
          if (ch[k] != 0 && jatomic_cas(E->fwd + ch[k], 0u, 0xFFFFFFFFu) == 0) {
            uint32_t pos = jatomic_add(&E->gc_count, 1u);
            list[pos] = ch[k];
            E->fwd[ch[k]] = pos + 1;
          }
      }
    sync_or<MODE>(c, 0);
    b = e; e = ld_s<MODE>(&E->gc_count);
  }
  const uint32_t live = b;  // == gc_count
  // copy nodes with forwarded children; mark the reals they use
  for (uint32_t i = c.tid; i < live; i += c.nth) {
    JNode nd = ld_s<MODE>(E->nodes + ld_s<MODE>(list + i));
    uint32_t c0 = nd_c0(nd), c1 = nd.c1;
    JNode nn;
    nn.c0 = (c0 ? ld_s<MODE>(E->fwd + c0) : 0) | (nd.c0 & 0x80000000u);
    nn.c1 = c1 ? ld_s<MODE>(E->fwd + c1) : 0;
    nn.w = nd.w;
    E->nodes_alt[i + 1] = nn;
    if (clean_vals) {
      E->vfwd[(uint32_t)(nd.w >> 32) & 0x7FFFFFFFu] = 1;
      E->vfwd[(uint32_t)nd.w & 0x7FFFFFFFu] = 1;
    }
  }
  sync_or<MODE>(c, 0);
  if (c.block0 && c.ltid == 0) {
    if (clean_vals) {  // compact the reals, keeping insertion order; constants stay
      uint64_t rw = E->rootw;
      E->vfwd[(uint32_t)(rw >> 32) & 0x7FFFFFFFu] = 1;
      E->vfwd[(uint32_t)rw & 0x7FFFFFFFu] = 1;
      uint32_t nv = 0;
      for (uint32_t i = 0; i < old_vals; i++) {
        if (i < 4 || ld_s<MODE>(E->vfwd + i)) {
          E->vals_alt[nv] = ld_s<MODE>(E->vals + i);
          E->vfwd[i] = nv++;
        }
      }
      E->val_count = nv;
      E->rootw = ((uint64_t)(((uint32_t)(rw >> 32) & 0x80000000u) | E->vfwd[(uint32_t)(rw >> 32) & 0x7FFFFFFFu]) << 32) |
                 (((uint32_t)rw & 0x80000000u) | E->vfwd[(uint32_t)rw & 0x7FFFFFFFu]);
      if (E->val_gc_limit < 2 * nv) E->val_gc_limit *= 2;  // Simulator.cpp:495-497
      if (E->val_gc_limit > E->val_cap / 2) E->val_gc_limit = E->val_cap / 2;
    }
    E->root = 1;
    *cnt_nodes<MODE>(c) = live + 1;
    E->node_count = live + 1;
    uint32_t lim = 4 * (live + 1);
    if (lim < E->gc_min) lim = E->gc_min;
    if (lim > E->node_cap / 2) lim = E->node_cap / 2;
    E->gc_limit = lim;
    E->st_gc_runs++;
    E->st_gc_nodes_copied += live;
    if (live + 1 >= E->node_cap / 2) jset_error_c<MODE>(c, JKQ_ERR_NODE_FULL);
  }
  sync_or<MODE>(c, 0);
  const uint32_t nv = ld_s<MODE>(&E->val_count);
  // copy back with remapped weights, rebuild the unique table
  for (uint32_t i = c.tid; i < live; i += c.nth) {
    JNode nn = ld_s<MODE>(E->nodes_alt + i + 1);
    if (clean_vals) {
      uint32_t hr = (uint32_t)(nn.w >> 32), hi = (uint32_t)nn.w;
      hr = (hr & 0x80000000u) | ld_s<MODE>(E->vfwd + (hr & 0x7FFFFFFFu));
      hi = (hi & 0x80000000u) | ld_s<MODE>(E->vfwd + (hi & 0x7FFFFFFFu));
      nn.w = ((uint64_t)hr << 32) | hi;
    }
    E->nodes[i + 1] = nn;
    uint32_t h = ut_hash(nn) & E->utab_mask;
    while (jatomic_cas(E->utab + h, 0u, i + 1) != 0) h = (h + 1) & E->utab_mask;
  }
  if (MODE == MODE_CTA)  // node ids changed: drop the shared node cache
    for (uint32_t i = c.ltid; i < JKQ_SNC_SLOTS; i += c.lnth) c.sh->ntag[i] = 0;
  // copy the reals back, rebuild the value table (distinct buckets: claim a slot, then fill it)
  for (uint32_t i = c.tid; clean_vals && i < nv; i += c.nth) {
    double x = ld_s<MODE>(E->vals_alt + i);
    E->vals[i] = x;
    if (i == 0) continue;  // 0.0 is never hashed: real_find short-cuts x <= tol
    long long k = (long long)floor(JKQ_DDIV(x, E->tol));
    uint32_t h = (uint32_t)jmix((uint64_t)k) & E->vtab_mask;
    while (jatomic_cas(&E->vtab[h].id1, 0u, i + 1) != 0) h = (h + 1) & E->vtab_mask;
    E->vtab[h].value = x;
    E->vtab[h].keylo = (uint32_t)k;
  }
  return sync_or<MODE>(c, jhas_error<MODE>(c) ? JF_ERROR : 0) & JF_ERROR;
}

// ------------------------------------------------------------------------------------ init
// QMDDinit + Simulator::AddVariables (QMDDpackage.cpp:1248-1317, Simulator.cpp:63-96): empty tables,
// the reals 0, 1 and the reference's gate constants 1/2 and FP64 sqrt(2)/2 in its insertion order
// (QMDDpackage.cpp:44-53), then the |0...0> chain.
template <int MODE> JD uint32_t init_state(Ctx &c) {
  JEng *E = c.E;
  for (uint32_t i = c.tid; i <= E->utab_mask; i += c.nth) E->utab[i] = 0;
  for (uint32_t i = c.tid; i <= E->vtab_mask; i += c.nth) { JVSlot z; z.value = 0; z.id1 = 0; z.keylo = 0; E->vtab[i] = z; }
  for (uint32_t i = c.tid; i <= E->dtab_mask; i += c.nth) E->dtab[i] = 0;
  if (c.block0 && c.ltid == 0) {
    E->status = 0; E->want_grid = 0; E->next_gate = 0;
    E->val_count = 0; E->node_count = 1; E->item_count = 0; E->pend_n[0] = E->pend_n[1] = 0;
    E->gen = 1;
    E->st_items = E->st_nodes_created = E->st_val_inserts = E->st_gc_runs = E->st_retry_rounds = 0;
    E->st_ut_probes = E->st_dt_probes = E->st_vt_probes = E->st_max_frontier = E->st_gates = E->st_gc_nodes_copied = 0;
    JNode t; t.c0 = 0; t.c1 = 0; t.w = 0;
    E->nodes[0] = t;
  }
  sync_or<MODE>(c, 0);
  if (c.block0 && c.ltid == 0) {
    E->vals[0] = 0.0; E->val_count = 1;  // index 0 never sits in the hash: real_find short-cuts x <= tol
    real_insert_serial<MODE>(E, 1.0);
    real_insert_serial<MODE>(E, 0.5);
    real_insert_serial<MODE>(E, 0.70710678118654757);
    uint32_t prev = 0;
    for (uint32_t L = 0; L < E->n; L++) {
      JNode nd; nd.c0 = prev; nd.c1 = 0; nd.w = 0;  // child 0 carries ONE, child 1 is the zero edge
      uint32_t id = E->node_count++;
      E->nodes[id] = nd;
      uint32_t h = ut_hash(nd) & E->utab_mask;
      while (ld_s<MODE>(E->utab + h) != 0) h = (h + 1) & E->utab_mask;
      E->utab[h] = id;
      prev = id;
    }
    E->root = prev; E->rootw = JW_ONE;
    E->gc_limit = E->gc_min > E->node_cap / 2 ? E->node_cap / 2 : E->gc_min;
    E->val_gc_limit = 10000 < E->val_cap / 2 ? 10000 : E->val_cap / 2;  // complex_limit, Simulator.h:91
  }
  return sync_or<MODE>(c, 0);
}

// ------------------------------------------------------------------------------------ circuit driver
// Applies gates [E->next_gate, n_gates).  Collects when the arena passed gc_limit, or after every
// gate when a level log was requested (parity tests: reference Active[v], QMDDpackage.h:199-201).
template <int MODE> JD void cta_state_load(Ctx &c) {
  if (MODE != MODE_CTA) return;
  if (c.ltid == 0) {
    CtaState *S = c.sh;
    S->item_count = c.E->item_count; S->node_count = c.E->node_count;
    S->pend_n[0] = S->pend_n[1] = 0;
    S->nodes_created = 0; S->err = (uint32_t)c.E->status; S->want_grid = 0;
  }
  for (uint32_t i = c.ltid; i < JKQ_SNC_SLOTS; i += c.lnth) c.sh->ntag[i] = 0;
  sync_or<MODE>(c, 0);
}
template <int MODE> JD void cta_state_store(Ctx &c) {
  if (MODE != MODE_CTA) return;
  if (c.ltid == 0) {
    CtaState *S = c.sh;
    c.E->item_count = S->item_count; c.E->node_count = S->node_count;
    c.E->st_nodes_created += S->nodes_created;
    S->nodes_created = 0;
  }
}

template <int MODE> JD void run_gates(Ctx &c, const JGate *gates, int n_gates) {
  JEng *E = c.E;
  cta_state_load<MODE>(c);
  int g = ld_s<MODE>((const uint32_t *)&E->next_gate);
  for (; g < n_gates; g++) {
    uint32_t f = apply_gate<MODE>(c, gates[g]);
    if (f & JF_ERROR) break;
    if (f & JF_GRID) {
      if (c.block0 && c.ltid == 0) { E->want_grid = 1; E->gen++; }
      break;
    }
    bool log = E->lvl_log != 0;
    if (log || rd_nodes<MODE>(c) > ld_s<MODE>(&E->gc_limit)) {
      if (gc_run<MODE>(c)) break;
      if (log && c.block0 && c.ltid == 0)
        for (uint32_t L = 0; L < E->n; L++) E->lvl_log[(size_t)g * E->n + L] = E->lvl_cnt[L];
    }
    if (c.block0 && c.ltid == 0) E->next_gate = g + 1;
  }
  sync_or<MODE>(c, 0);
  cta_state_store<MODE>(c);
}

// ------------------------------------------------------------------------------------ sampling prep
// Simulator::AssignProbs (Simulator.cpp:127-146) once per DD instead of once per shot: after a
// collection the arena is level-ordered, so one bottom-up pass gives every node its squared norm;
// the per-node branch probability of MeasureAll (Simulator.cpp:165-187) is stored next to the child
// ids in a 16-byte sampling record.
template <int MODE> JD uint32_t build_sampling(Ctx &c, JSNode *snodes, double *prob) {
  JEng *E = c.E;
  cta_state_load<MODE>(c);
  if (gc_run<MODE>(c)) return JF_ERROR;
  cta_state_store<MODE>(c);
  if (c.block0 && c.ltid == 0) prob[0] = 1.0;
  sync_or<MODE>(c, 0);
  for (uint32_t L = 0; L < E->n; L++) {
    uint32_t off = ld_s<MODE>(&E->lvl_off[L]), cnt = ld_s<MODE>(&E->lvl_cnt[L]);
    for (uint32_t i = c.tid; i < cnt; i += c.nth) {
      uint32_t id = off + i;
      JNode nd = ld_s<MODE>(E->nodes + id);
      uint64_t w0 = nd_w0(nd), w1 = nd_w1(nd);
      uint32_t c0 = nd_c0(nd), c1 = nd.c1;
      double p0 = w0 ? JKQ_DMUL(ld_s<MODE>(prob + c0), jmag2<MODE>(c, w0)) : 0.0;
      double p1 = w1 ? JKQ_DMUL(ld_s<MODE>(prob + c1), jmag2<MODE>(c, w1)) : 0.0;
      double tot = JKQ_DADD(p0, p1);
      prob[id] = tot;
      JSNode s;
      s.c0 = c0; s.c1 = c1; s.q0 = JKQ_DDIV(p0, tot);
      snodes[id] = s;
    }
    sync_or<MODE>(c, 0);
  }
  if (c.block0 && c.ltid == 0) E->norm = JKQ_DMUL(ld_s<MODE>(prob + ld_s<MODE>(&E->root)), jmag2<MODE>(c, E->rootw));
  return sync_or<MODE>(c, 0);
}

#endif  // JKQ_APPLY_CUH
