// ORACLE SCAFFOLDING -- test infrastructure, not product code.
//
// Gate-list driver that links the UNMODIFIED reference engine sources
// (/root/reference/src/{Simulator,QMDDpackage,QMDDcomplexD,...}.cpp) through the
// declaration-only shim headers in oracle/shim/.  It drives the reference's own
// QMDDmvlgate + Simulator::ApplyGate + Simulator::MeasureAll exactly the way the
// reference front end does (src/QASMsimulator.cpp:295-305,326-331,487-500) and dumps
// what the parity tests need:
//   * per-level live node counts Active[v] (src/QMDDpackage.h:199-201) after every gate,
//   * final amplitudes at full FP64 (GetElementOfVector + Cvalue, not the 6-digit Cprint),
//   * counts (character i = qubit i, src/QASMsimulator.cpp:491-494),
//   * the engine's own work counters and wall time of the apply loop / shot loop.
// Nothing here is copied from the reference; the U(theta,phi,lambda) matrix is the
// mathematical definition at src/QASMsimulator.cpp:295-298 restated.
//
// Circuit text format ("jkq circuit v1", also read by oracle/qmdd_oracle.c and written by
// qiskit_jku_provider_b200/circuits.py):
//   qubits N
//   u   T THETA PHI LAMBDA
//   cx  C T
//   mcu T THETA PHI LAMBDA K C1 .. CK      (control token "~q" = negative control)
//   mcg NAME T K C1 .. CK                  (NAME in x z h s: the reference's exact Nm Zm Hm Sm)
// angle token: decimal double, or [-]pi[*N][/D] (evaluated in 200-bit like the QASM parser).
#include <Simulator.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

using mpfr::mpreal;

struct GateOp {
  int kind;  // 0 = U-type (angles), 1 = named constant matrix
  int target;
  std::string th, ph, la;  // angle tokens
  char name;
  std::vector<std::pair<int, int>> controls;  // (qubit, value)
};

static mpreal parse_angle(const std::string &tok) {
  size_t p = tok.find("pi");
  if (p == std::string::npos) return mpreal(atof(tok.c_str()));
  mpreal v = mpfr::const_pi();
  if (p > 0 && tok[0] == '-') v = -v;
  size_t star = tok.find('*'), slash = tok.find('/');
  if (star != std::string::npos) v = v * mpreal(atof(tok.c_str() + star + 1));
  if (slash != std::string::npos) v = v / mpreal(atof(tok.c_str() + slash + 1));
  return v;
}

class Harness : public Simulator {
 public:
  void Simulate() override {}
  void Simulate(int) override {}

  int n = 0;
  std::vector<GateOp> ops;

  bool load(const char *fname) {
    std::ifstream in(fname);
    if (!in) return false;
    std::string w;
    while (in >> w) {
      if (w[0] == '#') { std::getline(in, w); continue; }
      if (w == "qubits") { in >> n; continue; }
      GateOp g;
      g.name = 0;
      if (w == "u") {
        g.kind = 0;
        in >> g.target >> g.th >> g.ph >> g.la;
      } else if (w == "cx") {
        int c;
        g.kind = 1; g.name = 'x';
        in >> c >> g.target;
        g.controls.push_back(std::make_pair(c, 1));
      } else if (w == "mcu" || w == "mcg") {
        int k;
        if (w == "mcu") { g.kind = 0; in >> g.target >> g.th >> g.ph >> g.la; }
        else { std::string nm; g.kind = 1; in >> nm >> g.target; g.name = nm[0]; }
        in >> k;
        for (int i = 0; i < k; i++) {
          std::string c;
          in >> c;
          if (c[0] == '~') g.controls.push_back(std::make_pair(atoi(c.c_str() + 1), 0));
          else g.controls.push_back(std::make_pair(atoi(c.c_str()), 1));
        }
      } else {
        fprintf(stderr, "unknown op %s\n", w.c_str());
        return false;
      }
      ops.push_back(g);
    }
    return n > 0;
  }

  void apply(const GateOp &g) {
    QMDD_matrix m;
    if (g.kind == 0) {
      mpreal theta = parse_angle(g.th), phi = parse_angle(g.ph), lambda = parse_angle(g.la);
      // U(theta,phi,lambda), SU(2) phase convention of src/QASMsimulator.cpp:295-298
      m[0][0] = Cmake(cos(-(phi + lambda) / 2) * cos(theta / 2), sin(-(phi + lambda) / 2) * cos(theta / 2));
      m[0][1] = Cmake(-cos(-(phi - lambda) / 2) * sin(theta / 2), -sin(-(phi - lambda) / 2) * sin(theta / 2));
      m[1][0] = Cmake(cos((phi - lambda) / 2) * sin(theta / 2), sin((phi - lambda) / 2) * sin(theta / 2));
      m[1][1] = Cmake(cos((phi + lambda) / 2) * cos(theta / 2), sin((phi + lambda) / 2) * cos(theta / 2));
    } else {
      QMDD_matrix *src = &Nm;
      if (g.name == 'z') src = &Zm;
      else if (g.name == 'h') src = &Hm;
      else if (g.name == 's') src = &Sm;
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) m[i][j] = (*src)[i][j];
    }
    for (auto &c : g.controls) line[nqubits - 1 - c.first] = c.second;
    line[nqubits - 1 - g.target] = 2;
    QMDDedge f = QMDDmvlgate(m, nqubits, line);
    for (auto &c : g.controls) line[nqubits - 1 - c.first] = -1;
    line[nqubits - 1 - g.target] = -1;
    ApplyGate(f);
  }

  int run(int argc, char **argv) {
    const char *circuit = NULL;
    long shots = 0, amp_limit = 0;
    bool levels = false, amps = false, dump_ctable = false;
    unsigned long seed = 1;
    for (int i = 1; i < argc; i++) {
      std::string a = argv[i];
      if (a == "--circuit" && i + 1 < argc) circuit = argv[++i];
      else if (a == "--shots" && i + 1 < argc) shots = atol(argv[++i]);
      else if (a == "--seed" && i + 1 < argc) seed = strtoul(argv[++i], NULL, 10);
      else if (a == "--levels") levels = true;
      else if (a == "--amps") amps = true;
      else if (a == "--amp-limit" && i + 1 < argc) amp_limit = atol(argv[++i]);
      else if (a == "--precision" && i + 1 < argc) Ctol = mpreal(atof(argv[++i]));
      else if (a == "--dump-ctable") dump_ctable = true;
    }
    if (!circuit || !load(circuit)) {
      fprintf(stderr, "usage: jku_ref --circuit FILE [--shots N] [--seed S] [--levels] [--amps]\n");
      return 2;
    }
    srand(seed);
    AddVariables(n, "q");

    std::vector<int> lv;
    if (levels) lv.reserve(ops.size() * n);
    int peak_active = ActiveNodeCount;
    auto t0 = std::chrono::steady_clock::now();
    for (size_t g = 0; g < ops.size(); g++) {
      apply(ops[g]);
      if (ActiveNodeCount > peak_active) peak_active = ActiveNodeCount;
      if (levels)
        for (int v = 0; v < n; v++) lv.push_back(Active[v]);
    }
    auto t1 = std::chrono::steady_clock::now();
    double t_apply = std::chrono::duration<double>(t1 - t0).count();

    printf("{\n \"n\": %d,\n \"applies\": %zu,\n \"t_apply_s\": %.9g,\n", n, ops.size(), t_apply);
    printf(" \"nodes_created\": %ld,\n \"ut_lookups\": %ld,\n", (long)(UTlookups - UTmatch), (long)UTlookups);
    printf(" \"ct_look_mult\": %ld, \"ct_hit_mult\": %ld, \"ct_look_add\": %ld, \"ct_hit_add\": %ld,\n",
           (long)CTlook[mult], (long)CThit[mult], (long)CTlook[add], (long)CThit[add]);
    printf(" \"nop_mult\": %ld, \"nop_add\": %ld,\n", (long)Nop[mult], (long)Nop[add]);
    printf(" \"peak_active\": %d,\n \"ctable_size\": %zu,\n \"peak_table_nodes\": %ld,\n", peak_active, Ctable.size(),
           (long)QMDDpeaknodecount);
    printf(" \"level_counts\": [");
    for (int v = 0; v < n; v++) printf("%s%d", v ? "," : "", Active[v]);
    printf("]");
    if (levels) {
      printf(",\n \"level_counts_per_gate\": [");
      for (size_t g = 0; g < ops.size(); g++) {
        printf("%s[", g ? "," : "");
        for (int v = 0; v < n; v++) printf("%s%d", v ? "," : "", lv[g * n + v]);
        printf("]");
      }
      printf("]");
    }
    if (amps) {
      // amplitude index i: bit q of i = qubit q (little endian); variable = n-1-q
      unsigned long long len = 1ull << n;
      if (amp_limit > 0 && (unsigned long long)amp_limit < len) len = amp_limit;
      printf(",\n \"amps\": [");
      for (unsigned long long i = 0; i < len; i++) {
        unsigned long long entry = 0;
        for (int q = 0; q < n; q++) entry |= ((i >> q) & 1ull) << (n - 1 - q);
        uint64_t w = GetElementOfVector(entry);
        complex c = Cvalue(w);
        double re = mpfr_get_d(c.r, MPFR_RNDN), im = mpfr_get_d(c.i, MPFR_RNDN);
        if ((w >> 63) & 1) re = -re;
        if ((w >> 31) & 1) im = -im;
        printf("%s[%.17g,%.17g]", i ? "," : "", re, im);
      }
      printf("]");
    }
    if (dump_ctable) {
      std::map<uint32_t, double> sorted;
      for (auto &kv : Ctable) sorted[kv.first] = mpfr_get_d(&kv.second, MPFR_RNDN);
      printf(",\n \"ctable\": [");
      bool first = true;
      for (auto &kv : sorted) { printf("%s[%u,%.17g]", first ? "" : ",", kv.first, kv.second); first = false; }
      printf("]");
    }
    if (shots > 0) {
      std::map<std::string, long> result;
      auto s0 = std::chrono::steady_clock::now();
      ResetBeforeMeasurement();
      std::string s(n, '0');
      for (long i = 0; i < shots; i++) {
        MeasureAll(false);
        for (int q = 0; q < n; q++) s[q] = (char)('0' + measurements[n - 1 - q]);
        result[s]++;
      }
      auto s1 = std::chrono::steady_clock::now();
      printf(",\n \"shots\": %ld,\n \"t_sample_s\": %.9g,\n \"counts\": {", shots,
             std::chrono::duration<double>(s1 - s0).count());
      bool first = true;
      for (auto &kv : result) {
        printf("%s\"%s\": %ld", first ? "" : ", ", kv.first.c_str(), kv.second);
        first = false;
      }
      printf("}");
    }
    printf("\n}\n");
    return 0;
  }
};

int main(int argc, char **argv) {
  // order as in src/main.cpp:104-120: srand, QMDDinit, then construct the Simulator
  QMDDinit(0);
  Harness *h = new Harness();
  return h->run(argc, argv);
}
