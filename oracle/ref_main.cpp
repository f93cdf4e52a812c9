// ORACLE SCAFFOLDING -- test infrastructure, not product code.
//
// Stand-in for the reference's CLI entry (src/main.cpp:75-150), which needs
// Boost.program_options (absent in this image).  Same flags, same call order
// (srand -> QMDDinit -> optional Ctol override -> QASMsimulator -> Simulate(shots) -> --ps stats),
// so the UNMODIFIED reference OpenQASM front end + engine can be run here as
// `oracle/_ref/jku_simulator --simulate_qasm[=file] --seed= --shots= ...`.
#include <QASMsimulator.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <string>

using mpfr::mpreal;

static bool flag_value(const char *arg, const char *name, std::string &out) {
  size_t n = strlen(name);
  if (strncmp(arg, name, n) != 0) return false;
  if (arg[n] == '\0') { out = ""; return true; }
  if (arg[n] == '=') { out = arg + n + 1; return true; }
  return false;
}

int main(int argc, char **argv) {
  unsigned long seed = time(NULL);
  unsigned int shots = 1;
  bool qasm = false, ps = false, sv = false, pr = false, have_prec = false;
  double prec = 0;
  std::string fname, v;
  for (int i = 1; i < argc; i++) {
    if (flag_value(argv[i], "--seed", v)) seed = strtoul(v.c_str(), NULL, 10);
    else if (flag_value(argv[i], "--simulate_qasm", v)) { qasm = true; fname = v; }
    else if (flag_value(argv[i], "--shots", v)) shots = (unsigned int)strtoul(v.c_str(), NULL, 10);
    else if (flag_value(argv[i], "--precision", v)) { have_prec = true; prec = atof(v.c_str()); }
    else if (!strcmp(argv[i], "--ps")) ps = true;
    else if (!strcmp(argv[i], "--display_statevector")) sv = true;
    else if (!strcmp(argv[i], "--display_probabilities")) pr = true;
    else { std::cout << "Allowed options: --seed= --simulate_qasm[=file] --shots= --ps --display_statevector --display_probabilities --precision=\n"; return 1; }
  }
  if (!qasm) { std::cout << "need --simulate_qasm\n"; return 1; }
  srand(seed);
  QMDDinit(0);
  if (have_prec) Ctol = mpreal(prec);
  QASMsimulator *sim = fname.empty() ? new QASMsimulator(sv, pr) : new QASMsimulator(fname, sv, pr);
  auto t1 = std::chrono::high_resolution_clock::now();
  sim->Simulate(shots);
  auto t2 = std::chrono::high_resolution_clock::now();
  std::chrono::duration<float> diff = t2 - t1;
  if (ps) {
    std::cout << std::endl << "SIMULATION STATS: " << std::endl;
    std::cout << "  Number of applied gates: " << sim->GetGatecount() << std::endl;
    std::cout << "  Simulation time: " << diff.count() << " seconds" << std::endl;
    std::cout << "  Maximal size of DD (number of nodes) during simulation: " << sim->GetMaxActive() << std::endl;
  }
  return 0;
}
