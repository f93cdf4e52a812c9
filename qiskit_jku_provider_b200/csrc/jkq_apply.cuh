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
        if (!real_find<MODE>(c, x, &idx)) real_insert_serial<MODE>(c, x);
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
  for (uint32_t round = 0;; round++) {
    uint32_t my = 0;
#ifdef JKQ_PROFILE
    long long pt0 = clock64();
#endif
    for (uint32_t i = b + c.tid; i < e; i += c.nth) {
      JItem *it = c.items + i;
      uint32_t st = ld_i<MODE>(&it->state);
      if (st == JS_DEAD) continue;
      if (round > 0 && st != JS_RETRY) continue;
      bool ok = expand ? expand_item<MODE>(c, g, i, L) : reduce_item<MODE>(c, i);
      it->state = ok ? JS_DONE : JS_RETRY;
      if (!ok) my |= JF_RETRY;
    }
    if (jhas_error<MODE>(c)) my |= JF_ERROR;
    if (MODE == MODE_CTA && c.sh->want_grid) my |= JF_GRID;
#ifdef JKQ_PROFILE
    long long pt1 = clock64();
#endif
    uint32_t flags = sync_or<MODE>(c, my);
#ifdef JKQ_PROFILE
    if (c.tid == 0) {
      long long pt2 = clock64();
      c.E->prof[expand ? 0 : 1] += (unsigned long long)(pt1 - pt0);
      c.E->prof[2] += (unsigned long long)(pt2 - pt1);
      c.E->prof[3] += 1;
      {  // classify the phase by its first item: 0 = M above target, 1 = M at target, 2 = F, 3 = S
        uint32_t tg = c.items[b].tag & 15u;
        int cls = tg == JK_M ? (L > g.target ? 0 : 1) : (tg == JK_F ? 2 : 3);
        c.E->prof2[(expand ? 0 : 4) + cls] += (unsigned long long)(pt1 - pt0);
        c.E->prof2[8 + (expand ? 0 : 4) + cls] += 1;
      }
    }
#endif
    if (flags & JF_ERROR) return JF_ERROR;
    if (flags & JF_GRID) return JF_GRID;
    if (!(flags & JF_RETRY)) return 0;
    serial_insert<MODE>(c);
    if (MODE == MODE_CTA && c.sh->want_grid) return JF_GRID;  // shared value table full (uniform: read after the barrier)
  }
}

// item ranges of the levels of one gate: registers / local memory on the grid path, shared memory on
// the CTA path (a local array would live in L1, which the CTA path gives up for shared memory)
template <int MODE> struct LevelRanges {
  uint32_t lb[MODE == MODE_CTA ? 1 : JKQ_MAXQ + 1], le[MODE == MODE_CTA ? 1 : JKQ_MAXQ + 1];
  JDM void set(Ctx &c, int L, uint32_t b, uint32_t e) {
    if (MODE == MODE_CTA) { if (c.ltid == 0) { c.sh->lb[L] = b; c.sh->le[L] = e; } }
    else { lb[L] = b; le[L] = e; }
  }
  JDM uint32_t begin(const Ctx &c, int L) const { return MODE == MODE_CTA ? c.sh->lb[L] : lb[L]; }
  JDM uint32_t end(const Ctx &c, int L) const { return MODE == MODE_CTA ? c.sh->le[L] : le[L]; }
};

// ------------------------------------------------------------------------------------ one gate
// Simulator::ApplyGate (Simulator.cpp:466-503) = QMDDmultiply(gate, state) as one top-down and one
// bottom-up sweep over the levels.  Returns 0, JF_ERROR, or JF_GRID (CTA mode only: the gate does not
// fit the shared-memory pools and is left unapplied).
template <int MODE> JD uint32_t apply_gate(Ctx &c, const JGate &g_hbm) {
  JEng *E = c.E;
  const int n = (int)E->n;
#ifdef JKQ_PROFILE
  long long gt0 = clock64();
#endif
  if (MODE == MODE_CTA) {  // stage the gate descriptor and clear the de-dup table while thread 0 does the setup
    uint32_t *dst = reinterpret_cast<uint32_t *>(&c.sh->gate);
    const uint32_t *srcw = reinterpret_cast<const uint32_t *>(&g_hbm);
    for (uint32_t i = c.ltid; i < sizeof(JGate) / 4; i += c.lnth) dst[i] = srcw[i];
    for (uint32_t i = c.ltid; i < JKQ_SDT_SLOTS; i += c.lnth) c.sh->dtab[i] = 0;
  }
  if (c.block0 && c.ltid == 0) {
    // canonical ids of the four matrix entries, cached per distinct matrix of the circuit until the
    // next sweep of the reals renumbers them (the reference re-runs Cmake per gate, QASMsimulator.cpp:295-298)
    const uint32_t mid = (uint32_t)g_hbm.pad;
    const uint32_t epoch = E->val_epoch;
    const bool cached = E->mat_epoch && ld_s<MODE>(E->mat_epoch + mid) == epoch;
    for (int k = 0; k < 4; k++) {
      uint64_t u = cached ? ld_s<MODE>(E->mat_u + 4 * mid + k) : cx_make_serial<MODE>(c, g_hbm.m[2 * k], g_hbm.m[2 * k + 1]);
      if (!cached && E->mat_epoch) E->mat_u[4 * mid + k] = u;
      if (MODE == MODE_CTA) c.sh->cur_u[k] = u;
      else E->cur_u[k] = u;
    }
    if (!cached && E->mat_epoch) E->mat_epoch[mid] = epoch;
    if (MODE == MODE_GRID) E->gen++;
    JItem *it = c.items;
    it->a = E->root; it->b = 0; it->wa = 0; it->wb = 0; it->wc = 0; it->tag = JK_M; it->state = JS_NEW;
    *cnt_items<MODE>(c) = 1;
  }
  {
    uint32_t f0 = jhas_error<MODE>(c) ? JF_ERROR : 0;
    uint32_t fl = sync_or<MODE>(c, f0);
    if (fl & JF_ERROR) return JF_ERROR;
    if (MODE == MODE_CTA && c.sh->want_grid) return JF_GRID;
  }
  const JGate &g = MODE == MODE_CTA ? c.sh->gate : g_hbm;
#ifdef JKQ_PROFILE
  long long gt1 = clock64();
  if (c.tid == 0) E->prof[4] += (unsigned long long)(gt1 - gt0);
#endif
  LevelRanges<MODE> lr;
  uint32_t b = 0, e = 1;
  int lowest = n;
  for (int L = n - 1; L >= 0 && e > b; L--) {
    lr.set(c, L, b, e);
    lowest = L;
    uint32_t f = run_level<MODE>(c, g, b, e, L, true);
    if (f) return f;
    b = e; e = rd_items<MODE>(c);  // children of level 0 are terminals, never items
    if (MODE == MODE_CTA && e - b > E->cta_item_limit) return JF_GRID;
  }
  for (int L = lowest; L < n; L++) {
    uint32_t f = run_level<MODE>(c, g, lr.begin(c, L), lr.end(c, L), L, false);
    if (f) return f;
  }
#ifdef JKQ_PROFILE
  long long gt2 = clock64();
#endif
  if (c.block0 && c.ltid == 0) {
    const JItem *it = c.items;
    uint64_t rw = ld_i<MODE>(&it->rw), w = 0;
    uint32_t rn = ld_i<MODE>(&it->rnode);
    if (rw == JW_ONE) w = E->rootw;
    else if (E->rootw == JW_ONE) w = rw;
    else if (rw != 0) {
      double ar = jre<MODE>(c, rw), ai = jim<MODE>(c, rw), br = jre<MODE>(c, E->rootw), bi = jim<MODE>(c, E->rootw);
      w = cx_make_serial<MODE>(c, JKQ_DSUB(JKQ_DMUL(ar, br), JKQ_DMUL(ai, bi)), JKQ_DADD(JKQ_DMUL(ar, bi), JKQ_DMUL(ai, br)));
    }
    if (MODE == MODE_CTA && c.sh->want_grid) {
      // the shared value table filled up while the root weight was made: keep the old root, redo on the grid
    } else {
      if (w == 0 || rn == 0) jset_error_c<MODE>(c, JKQ_ERR_ZERO_STATE);
      E->root = rn; E->rootw = w;
      unsigned long long cnt = rd_items<MODE>(c);
      E->st_items += cnt;
      if (cnt > E->st_max_frontier) E->st_max_frontier = cnt;
      E->st_gates++;
    }
  }
  {
    uint32_t fl = sync_or<MODE>(c, jhas_error<MODE>(c) ? JF_ERROR : 0);
    if (fl & JF_ERROR) return JF_ERROR;
    if (MODE == MODE_CTA && c.sh->want_grid) return JF_GRID;
  }
#ifdef JKQ_PROFILE
  if (c.tid == 0) { E->prof[5] += (unsigned long long)(clock64() - gt2); E->prof[6] += (unsigned long long)(clock64() - gt0); }
#endif
  return 0;
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
template <int MODE> JD uint32_t gc_run(Ctx &c, bool reload, bool force_vals) {
  JEng *E = c.E;
  const int n = (int)E->n;
  uint32_t *list = (uint32_t *)E->items;  // list[new_id - 1] = old id
  c.spare_node = 0;  // ids are renumbered: a spare id from before the collection is meaningless
  const uint32_t old_count = rd_nodes<MODE>(c), old_vals = rd_vals<MODE>(c);
  // the reals are only swept when the table outgrew its limit, like the reference's
  // `Ctable.size() > complex_limit` (Simulator.cpp:488-498): dropping an unreferenced value changes
  // which representative later values snap to
  uint32_t vlimit = ld_s<MODE>(&E->val_gc_limit);
  if (MODE == MODE_CTA && vlimit > (JKQ_SVALS / 4) * 3) vlimit = (JKQ_SVALS / 4) * 3;
  const bool clean_vals = force_vals || old_vals > vlimit;
  for (uint32_t i = c.tid; i < old_count; i += c.nth) E->fwd[i] = 0;
  if (MODE == MODE_GRID)
    for (uint32_t i = c.tid; i <= E->utab_mask; i += c.nth) E->utab[i] = 0;
  else
    for (uint32_t i = c.ltid; i < JKQ_SUT_SLOTS; i += c.lnth) c.sh->utab[i] = 0;
  if (clean_vals) {
    for (uint32_t i = c.tid; i < old_vals; i += c.nth) E->vfwd[i] = 0;
    if (MODE == MODE_GRID)
      for (uint32_t i = c.tid; i <= E->vtab_mask; i += c.nth) E->vtab[i].id1 = 0;
  }
  if (MODE == MODE_CTA && (clean_vals || reload))
    for (uint32_t i = c.ltid; i < JKQ_SVT_SLOTS; i += c.lnth) c.sh->vtab[i].id1 = 0;
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
      *cnt_vals<MODE>(c) = nv;
      E->val_count = nv;
      E->rootw = ((uint64_t)(((uint32_t)(rw >> 32) & 0x80000000u) | E->vfwd[(uint32_t)(rw >> 32) & 0x7FFFFFFFu]) << 32) |
                 (((uint32_t)rw & 0x80000000u) | E->vfwd[(uint32_t)rw & 0x7FFFFFFFu]);
      E->val_epoch++;  // ids renumbered: cached gate-matrix ids are void
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
    if (MODE == MODE_CTA) {
      // too big for the shared-memory pools: hand over to the grid path
      if (live + 1 > JKQ_SNODES / 4 || *cnt_vals<MODE>(c) > (JKQ_SVALS / 8) * 7) c.sh->want_grid = 1;
    }
  }
  sync_or<MODE>(c, 0);
  const uint32_t nv = rd_vals<MODE>(c);
  const bool fits = MODE != MODE_CTA || !c.sh->want_grid;
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
    if (MODE == MODE_GRID) {
      uint32_t h = ut_hash(nn) & E->utab_mask;
      while (jatomic_cas(E->utab + h, 0u, i + 1) != 0) h = (h + 1) & E->utab_mask;
    } else if (fits) {
      c.sh->nodes[i + 1] = nn;
      uint32_t h = ut_hash(nn) & (JKQ_SUT_SLOTS - 1);
      while (jatomic_cas(&c.sh->utab[h], 0u, i + 1) != 0) h = (h + 1) & (JKQ_SUT_SLOTS - 1);
    }
  }
  // the reals: copy the compacted array back; rebuild the value table
  if (clean_vals)
    for (uint32_t i = c.tid; i < nv; i += c.nth) E->vals[i] = ld_s<MODE>(E->vals_alt + i);
  if (MODE == MODE_GRID) {
    if (clean_vals)
      for (uint32_t i = c.tid + 1; i < nv; i += c.nth) vt_put<MODE>(c, i, ld_s<MODE>(E->vals_alt + i));  // 0.0 is never hashed
  } else if (fits && (clean_vals || reload)) {
    const double *srcv = clean_vals ? E->vals_alt : E->vals;
    for (uint32_t i = c.ltid; i < nv; i += c.lnth) {
      double x = srcv[i];
      c.sh->vals[i] = x;
      if (i > 0) vt_put<MODE>(c, i, x);
    }
  }
  return sync_or<MODE>(c, jhas_error<MODE>(c) ? JF_ERROR : 0) & JF_ERROR;
}

// ------------------------------------------------------------------------------------ init
// QMDDinit + Simulator::AddVariables (QMDDpackage.cpp:1248-1317, Simulator.cpp:63-96): empty tables,
// the reals 0, 1 and the reference's gate constants 1/2 and FP64 sqrt(2)/2 in its insertion order
// (QMDDpackage.cpp:44-53), then the |0...0> chain.  Runs on the grid path (HBM tables).
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
    for (int k = 0; k < 8; k++) E->prof[k] = 0;
    for (int k = 0; k < 16; k++) E->prof2[k] = 0;
    E->st_ut_probes = E->st_dt_probes = E->st_vt_probes = E->st_max_frontier = E->st_gates = E->st_gc_nodes_copied = 0;
    JNode t; t.c0 = 0; t.c1 = 0; t.w = 0;
    E->nodes[0] = t;
  }
  sync_or<MODE>(c, 0);
  if (c.block0 && c.ltid == 0) {
    E->vals[0] = 0.0; E->val_count = 1;  // index 0 never sits in the hash: real_find short-cuts x <= tol
    real_insert_serial<MODE>(c, 1.0);
    real_insert_serial<MODE>(c, 0.5);
    real_insert_serial<MODE>(c, 0.70710678118654757);
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
    E->tables_valid = 1;
    E->val_epoch = 1;
  }
  return sync_or<MODE>(c, 0);
}

// ------------------------------------------------------------------------------------ mode hand-over
// MODE_CTA prologue: counters and constants into shared memory, then one collection that loads the
// live DD (nodes, unique table, reals, value table) into shared memory.  The HBM hash tables are
// stale from here on (tables_valid = 0).  MODE_GRID prologue: if they are stale, one collection
// rebuilds them.  Returns JF_GRID when the state does not fit the CTA path.
template <int MODE> JD uint32_t mode_enter(Ctx &c) {
  JEng *E = c.E;
  if (MODE == MODE_CTA) {
    if (c.ltid == 0) {
      CtaState *S = c.sh;
      S->item_count = E->item_count; S->node_count = E->node_count; S->val_count = E->val_count;
      S->pend_n[0] = S->pend_n[1] = 0;
      S->nodes_created = 0; S->val_inserts = 0; S->err = (uint32_t)E->status; S->want_grid = 0;
      S->tol = E->tol; S->inv_tol = E->inv_tol;
      E->tables_valid = 0;
    }
    sync_or<MODE>(c, 0);
    // reals beyond the shared table: sweep them first; still too many -> grid path
    if (gc_run<MODE>(c, true, c.sh->val_count > (JKQ_SVALS / 4) * 3)) return JF_ERROR;
    return c.sh->want_grid ? JF_GRID : 0;
  }
  if (!ld_s<MODE>(&E->tables_valid)) {
    if (gc_run<MODE>(c, false, true)) return JF_ERROR;
    if (c.block0 && c.ltid == 0) E->tables_valid = 1;
    sync_or<MODE>(c, 0);
  }
  return 0;
}
template <int MODE> JD void mode_leave(Ctx &c) {
  if (MODE != MODE_CTA) return;
  if (c.ltid == 0) {
    CtaState *S = c.sh;
    c.E->item_count = S->item_count; c.E->node_count = S->node_count; c.E->val_count = S->val_count;
    c.E->st_nodes_created += S->nodes_created;
    c.E->st_val_inserts += S->val_inserts;
    S->nodes_created = 0; S->val_inserts = 0;
  }
}

// ------------------------------------------------------------------------------------ circuit driver
// Applies gates [E->next_gate, n_gates).  Collects when the arena passed its limit, or after every
// gate when a level log was requested (parity tests: reference Active[v], QMDDpackage.h:199-201).
template <int MODE> JD void run_gates(Ctx &c, const JGate *gates, int n_gates) {
  JEng *E = c.E;
  uint32_t f = mode_enter<MODE>(c);
  int g = ld_s<MODE>((const uint32_t *)&E->next_gate);
  for (; !f && g < n_gates; g++) {
    f = apply_gate<MODE>(c, gates[g]);
    if (f) break;
    const bool log = E->lvl_log != 0;
    const uint32_t nodes_now = rd_nodes<MODE>(c), vals_now = rd_vals<MODE>(c);
    bool collect = log || nodes_now > ld_s<MODE>(&E->gc_limit);
    if (MODE == MODE_CTA) collect = collect || nodes_now > JKQ_SNODES / 2 || vals_now > (JKQ_SVALS / 4) * 3;
    if (collect) {
      if (gc_run<MODE>(c, false, false)) { f = JF_ERROR; break; }
      if (log && c.block0 && c.ltid == 0)
        for (uint32_t L = 0; L < E->n; L++) E->lvl_log[(size_t)g * E->n + L] = E->lvl_cnt[L];
    }
    if (c.block0 && c.ltid == 0) E->next_gate = g + 1;
    if (MODE == MODE_CTA && c.sh->want_grid) { f = JF_GRID; g++; break; }  // (uniform: set before gc_run's last barrier)
  }
  if ((f & JF_GRID) && c.block0 && c.ltid == 0) { E->want_grid = 1; if (MODE == MODE_GRID) E->gen++; }
  sync_or<MODE>(c, 0);
  mode_leave<MODE>(c);
}

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
  if (MODE == MODE_GRID && gc_run<MODE>(c, false, false)) return JF_ERROR;
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
  uint32_t r = sync_or<MODE>(c, 0);
  mode_leave<MODE>(c);
  return r;
}

#endif  // JKQ_APPLY_CUH
