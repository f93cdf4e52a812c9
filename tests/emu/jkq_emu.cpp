// TEST INFRASTRUCTURE ONLY -- single-thread host emulation of the device code in
// qiskit_jku_provider_b200/csrc/*.cuh (compiled with -DJKQ_EMU).  It lets the algorithm of the CUDA
// engine (frontier expansion, tables, collector, sampling walk) be checked against the oracle in this
// GPU-less container.  It is never built into, nor loaded by, the product library: the product path
// is CUDA only and fails loudly without a GPU.
//
// Same CLI and JSON as oracle/qmdd_oracle.c.
#define JKQ_EMU 1
#include "../../qiskit_jku_provider_b200/csrc/jkq_apply.cuh"
#include "../../qiskit_jku_provider_b200/csrc/jkq_sample.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

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

static void set_u(JGate &g, double th, double ph, double la) {
  double c = cos(th / 2), s = sin(th / 2);
  g.m[0] = cos(-(ph + la) / 2) * c; g.m[1] = sin(-(ph + la) / 2) * c;
  g.m[2] = -cos(-(ph - la) / 2) * s; g.m[3] = -sin(-(ph - la) / 2) * s;
  g.m[4] = cos((ph - la) / 2) * s; g.m[5] = sin((ph - la) / 2) * s;
  g.m[6] = cos((ph + la) / 2) * c; g.m[7] = sin((ph + la) / 2) * c;
}
static void set_named(JGate &g, char name) {
  memset(g.m, 0, sizeof(g.m));
  double h = (0 + 1 * sqrt(2.0)) / 2;
  if (name == 'x') { g.m[2] = 1; g.m[4] = 1; }
  else if (name == 'z') { g.m[0] = 1; g.m[6] = -1; }
  else if (name == 'h') { g.m[0] = g.m[2] = g.m[4] = h; g.m[6] = -h; }
  else if (name == 's') { g.m[0] = 1; g.m[7] = 1; }
}
static void finish_gate(JGate &g, int n, int target, int nctl, const int *ctl, const int *cv) {
  memset(g.role, JR_IDLE, sizeof(g.role));
  g.target = n - 1 - target;
  g.role[g.target] = JR_TARGET;
  g.low = g.top = g.target;
  for (int i = 0; i < nctl; i++) {
    int L = n - 1 - ctl[i];
    g.role[L] = cv[i] ? JR_CPOS : JR_CNEG;
    if (L < g.low) g.low = L;
    if (L > g.top) g.top = L;
  }
  g.pad = 0;
}

int main(int argc, char **argv) {
  const char *circuit = NULL;
  long shots = 0, amp_limit = 0;
  int levels = 0, amps = 0;
  uint64_t seed = 1;
  uint32_t node_cap = 1u << 22, gc_min = 1u << 16, cta_node_limit = 0, cta_val_limit = 0, cta_no_reserve = 0;
  int dump_level = -1;
  bool force_grid = false;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--circuit") && i + 1 < argc) circuit = argv[++i];
    else if (!strcmp(argv[i], "--shots") && i + 1 < argc) shots = atol(argv[++i]);
    else if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = strtoull(argv[++i], NULL, 10);
    else if (!strcmp(argv[i], "--levels")) levels = 1;
    else if (!strcmp(argv[i], "--amps")) amps = 1;
    else if (!strcmp(argv[i], "--amp-limit") && i + 1 < argc) amp_limit = atol(argv[++i]);
    else if (!strcmp(argv[i], "--node-cap") && i + 1 < argc) node_cap = (uint32_t)atol(argv[++i]);
    else if (!strcmp(argv[i], "--gc-min") && i + 1 < argc) gc_min = (uint32_t)atol(argv[++i]);
    else if (!strcmp(argv[i], "--dump-level") && i + 1 < argc) dump_level = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--cta-node-limit") && i + 1 < argc) cta_node_limit = (uint32_t)atol(argv[++i]);
    else if (!strcmp(argv[i], "--cta-val-limit") && i + 1 < argc) cta_val_limit = (uint32_t)atol(argv[++i]);
    else if (!strcmp(argv[i], "--cta-no-reserve") && i + 1 < argc) cta_no_reserve = (uint32_t)atol(argv[++i]);
    else if (!strcmp(argv[i], "--grid")) force_grid = true;
  }
  FILE *f = circuit ? fopen(circuit, "r") : NULL;
  if (!f) { fprintf(stderr, "usage: jkq_emu --circuit FILE ...\n"); return 2; }
  int n = 0;
  std::vector<JGate> gates, mats;
  char w[64], a1[64], a2[64], a3[64];
  while (fscanf(f, "%63s", w) == 1) {
    if (w[0] == '#') { int ch; while ((ch = fgetc(f)) != EOF && ch != '\n') {} continue; }
    if (!strcmp(w, "qubits")) { if (fscanf(f, "%d", &n) != 1) return 2; continue; }
    JGate g;
    int target = -1, nctl = 0, ctl[JKQ_MAXQ], cv[JKQ_MAXQ];
    if (!strcmp(w, "u") || !strcmp(w, "mcu")) {
      if (fscanf(f, "%d %63s %63s %63s", &target, a1, a2, a3) != 4) return 2;
      set_u(g, parse_angle(a1), parse_angle(a2), parse_angle(a3));
    } else if (!strcmp(w, "cx")) {
      if (fscanf(f, "%d %d", &ctl[0], &target) != 2) return 2;
      cv[0] = 1; nctl = 1;
      set_named(g, 'x');
    } else if (!strcmp(w, "mcg")) {
      if (fscanf(f, "%63s %d", a1, &target) != 2) return 2;
      set_named(g, a1[0]);
    } else { fprintf(stderr, "unknown op %s\n", w); return 2; }
    if (!strcmp(w, "mcu") || !strcmp(w, "mcg")) {
      if (fscanf(f, "%d", &nctl) != 1) return 2;
      for (int i = 0; i < nctl; i++) {
        char c[32];
        if (fscanf(f, "%31s", c) != 1) return 2;
        if (c[0] == '~') { ctl[i] = atoi(c + 1); cv[i] = 0; }
        else { ctl[i] = atoi(c); cv[i] = 1; }
      }
    }
    finish_gate(g, n, target, nctl, ctl, cv);
    {  // index among the distinct matrices of the circuit, as jkq_load_circuit assigns it
      size_t mid = 0;
      for (; mid < mats.size(); mid++)
        if (!memcmp(mats[mid].m, g.m, sizeof(g.m))) break;
      if (mid == mats.size()) mats.push_back(g);
      g.pad = (int32_t)mid;
    }
    gates.push_back(g);
  }
  fclose(f);
  for (size_t i = 0; i < gates.size(); i++) {  // as finish_gates in jkq_b200.cu
    JGate &g = gates[i];
    g.idle = 0;
    for (int L = 0; L < JKQ_MAXQ; L++)
      if (g.role[L] == JR_IDLE) g.idle |= 1ull << L;
    g.park = g.top;
    for (size_t j = i + 1; j < gates.size() && j <= i + JKQ_PARK_WINDOW; j++)
      if (gates[j].top > g.park) g.park = gates[j].top;
    const bool diag = g.m[2] == 0.0 && g.m[3] == 0.0 && g.m[4] == 0.0 && g.m[5] == 0.0;
    const bool anti = g.m[0] == 0.0 && g.m[1] == 0.0 && g.m[6] == 0.0 && g.m[7] == 0.0;
    g.rsv = (diag || anti) ? 1 : 0;
  }

  JEng *E = (JEng *)calloc(1, sizeof(JEng));
  E->n = n; E->tol = 1e-10; E->inv_tol = 1.0 / E->tol;
  E->node_cap = node_cap; E->utab_mask = node_cap * 2 - 1;
  E->val_cap = 1u << 20; E->vtab_mask = (1u << 22) - 1;
  E->pend_cap = 1u << 20; E->item_cap = node_cap; E->dtab_mask = node_cap * 2 - 1;
  E->cta_item_limit = JKQ_SITEMS / 2; E->gc_min = gc_min;
  E->cta_node_limit = cta_node_limit;
  E->cta_val_limit = cta_val_limit;
  E->cta_no_reserve = cta_no_reserve;
  E->nodes = (JNode *)calloc(E->node_cap, sizeof(JNode));
  E->nodes_alt = (JNode *)calloc(E->node_cap, sizeof(JNode));
  E->utab = (unsigned long long *)calloc((size_t)E->utab_mask + 1, 8);
  E->ut_gen = 1; E->vt_cur = E->vtab_mask; E->vt_dirty_all = 1;
  E->vbm = (uint32_t *)calloc((((size_t)E->vtab_mask + 1) >> 8) + 1, 4);
  E->vals = (double *)calloc(E->val_cap, 8);
  E->vals_alt = (double *)calloc(E->val_cap, 8);
  E->vtab = (JVSlot *)calloc((size_t)E->vtab_mask + 1, sizeof(JVSlot));
  E->pend = (uint32_t *)calloc(E->pend_cap, 4);
  E->items = (JItem *)calloc(E->item_cap, sizeof(JItem));
  E->dtab = (unsigned long long *)calloc((size_t)E->dtab_mask + 1, 8);
  E->mdir = (unsigned long long *)calloc(E->node_cap, 8);
  E->thin_spill = calloc(1, sizeof(ThinState));
  E->fwd = (uint32_t *)calloc(E->node_cap, 4);
  E->vfwd = (uint32_t *)calloc(E->val_cap, 4);
  if (levels) E->lvl_log = (uint32_t *)calloc(gates.size() * n + 1, 4);

  Ctx c;
  static CtaState cta_state;
  static ThinState thin_state;
  c.sh = &cta_state;
  c.T = &thin_state;
  c.items = cta_state.items;
  c.E = E; c.tid = 0; c.nth = 1; c.phase = 0; c.ring = 0; c.block0 = 1; c.ltid = 0; c.lnth = 1;
  Ctx gctx = c;
  gctx.sh = nullptr; gctx.items = E->items;
  init_state<MODE_GRID>(gctx);
  int grid_switches = 0;
  bool grid = force_grid;
  // same escalation as the host driver (run_segment in jkq_b200.cu): CTA path until it does not fit
  while (!E->status && (E->next_gate < (int)gates.size() || E->cursor_saved)) {
    E->want_grid = 0;
    if (grid) run_gates<MODE_GRID, true>(gctx, gates.data(), (int)gates.size());
    else run_gates<MODE_CTA, true>(c, gates.data(), (int)gates.size());
    if ((E->want_grid || E->cursor_saved) && !grid) { grid = true; grid_switches++; }
    else if (E->next_gate < (int)gates.size() && !E->status) { fprintf(stderr, "stopped without reason\n"); return 3; }
  }
  if (E->status) { fprintf(stderr, "engine error %d at gate %d\n", E->status, E->next_gate); return 3; }

  std::vector<JSNode> sn(E->node_cap);
  std::vector<double> prob(E->node_cap);
  E->want_grid = 0;
  bool sampling_cta = !force_grid;
  if (sampling_cta) {
    build_sampling<MODE_CTA>(c, sn.data(), prob.data());
    if (E->want_grid) { sampling_cta = false; E->want_grid = 0; }
  }
  if (!sampling_cta) build_sampling<MODE_GRID>(gctx, sn.data(), prob.data());
  if (E->status) { fprintf(stderr, "engine error %d in build_sampling\n", E->status); return 3; }
  if (dump_level >= 0) {
    for (uint32_t i = 0; i < E->lvl_cnt[dump_level]; i++) {
      uint32_t id = E->lvl_off[dump_level] + i;
      JNode nd = E->nodes[id];
      fprintf(stderr, "node %u: flag %u c0 %u c1 %u w=(%.17g, %.17g) ids %llx\n", id, nd_flag(nd), nd_c0(nd), nd.c1,
              jre<MODE_GRID>(gctx, nd.w), jim<MODE_GRID>(gctx, nd.w), (unsigned long long)nd.w);
    }
  }
#ifdef JKQ_EMU_MEMOSTAT
  fprintf(stderr, "cx_arith calls %llu  repeats (unbounded memo) %llu  hits (256-entry direct-mapped) %llu\n", jmemostat().calls, jmemostat().repeats, jmemostat().recent_hits);
#endif
  printf("{\n \"n\": %d,\n \"applies\": %zu,\n", n, gates.size());
  printf(" \"items\": %llu,\n \"nodes_created\": %llu,\n \"val_inserts\": %llu,\n \"gc_runs\": %llu,\n \"retry_rounds\": %llu,\n \"max_frontier\": %llu,\n",
         E->st[JST_ITEMS], E->st[JST_NODES], E->st[JST_VALS], E->st[JST_GC_RUNS], E->st[JST_RETRY], E->st[JST_MAXF]);
  printf(" \"grid_switches\": %d,\n \"thin_gates\": %llu,\n", grid_switches, E->st[JST_THIN]);
  printf(" \"ctable_size\": %u,\n \"norm\": %.17g,\n", E->val_count, E->norm);
  printf(" \"level_counts\": [");
  for (int v = 0; v < n; v++) printf("%s%u", v ? "," : "", E->lvl_cnt[v]);
  printf("]");
  if (levels) {
    printf(",\n \"level_counts_per_gate\": [");
    for (size_t g = 0; g < gates.size(); g++) {
      printf("%s[", g ? "," : "");
      for (int v = 0; v < n; v++) printf("%s%u", v ? "," : "", E->lvl_log[g * n + v]);
      printf("]");
    }
    printf("]");
  }
  if (amps) {
    uint64_t len = 1ull << n;
    if (amp_limit > 0 && (uint64_t)amp_limit < len) len = amp_limit;
    printf(",\n \"amps\": [");
    for (uint64_t i = 0; i < len; i++) {
      double re, im;
      amp_walk<MODE_GRID>(gctx, i, &re, &im);
      printf("%s[%.17g,%.17g]", i ? "," : "", re, im);
    }
    printf("]");
  }
  if (shots > 0) {
    uint8_t bitpos[JKQ_MAXQ];
    for (int L = 0; L < n; L++) bitpos[L] = (uint8_t)(n - 1 - L);  // key bit q = qubit q
    std::map<uint64_t, long> hist;
    for (long i = 0; i < shots; i++) hist[sample_walk(sn.data(), E->root, n, bitpos, seed, (uint64_t)i)]++;
    printf(",\n \"shots\": %ld,\n \"counts\": {", shots);
    bool first = true;
    for (auto &kv : hist) {
      char s[65];
      for (int q = 0; q < n; q++) s[q] = (char)('0' + ((kv.first >> q) & 1));
      s[n] = 0;
      printf("%s\"%s\": %ld", first ? "" : ", ", s, kv.second);
      first = false;
    }
    printf("}");
  }
  printf("\n}\n");
  return 0;
}
