// jku_simulator -- the executable interface of the reference on the B200 engine.
//
// The reference's Python backend talks to its simulator through exactly this program: argv flags plus OpenQASM on
// stdin, JSON on stdout (qiskit_jku_provider/qasm_simulator_jku.py:82-97 -> src/main.cpp:75-150 ->
// QASMsimulator::Simulate(int shots), src/QASMsimulator.cpp:479-566).  This file is main.cpp + the driver and
// printing part of QASMsimulator::Simulate on top of the C ABI (include/jkq.h): same flags, same JSON schema,
// same key conventions, so the reference package runs on the GPU engine by pointing its executable path here.
//   counts       bitstring over ALL qubits, character i = qubit i (QASMsimulator.cpp:489-500)
//   snapshots    per slot: "probabilities" [2^k], "probabilities_ket" {bits: p > 0}, "statevector" [2^n strings in
//                the format of Cprint, QMDDcomplexD.cpp:100-135]; first snapshot argument = most significant bit
//   --ps         the three SIMULATION STATS lines of main.cpp:142-147
// Shots: when no gate follows a measurement the counts are sampled from the state before the first measurement
// (Simulator::ResetBeforeMeasurement, QASMsimulator.cpp:487-500); otherwise the circuit is re-simulated for every
// shot with fresh measurement outcomes (:501-522).  Random numbers are Philox streams of --seed, not glibc rand().
//
// Resident server.  A fresh process pays the CUDA start-up (context creation, module load, pool allocation: 0.3-0.9 s on
// a B200 box) before its first gate, which is more than most programs take to simulate.  `jku_simulator --serve` is the
// same program as a resident process on a Unix socket ($JKU_SIMULATOR_SOCKET, default /tmp/jku_simulator.<uid>.sock)
// with one CUDA context and one engine handle; every ordinary invocation first looks for it and, when it answers,
// forwards its argv and program text and prints what comes back -- same flags, same stdin, same stdout / stderr / exit
// code, no CUDA in the short-lived process.  With JKU_SIMULATOR_AUTOSERVE=1 an invocation that finds no server starts
// one (detached, exits after --idle seconds without a request, default 600) and serves itself in-process meanwhile.
#include <chrono>
#include <cerrno>
#include <csignal>
#include <fcntl.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/un.h>
#include <unistd.h>
#include <stdexcept>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/jkq.h"

namespace {

const char *USAGE =
    "Allowed options:\n"
    "  --help                      produce help message\n"
    "  --seed arg                  seed for random number generator\n"
    "  --simulate_qasm [=arg(=)]   simulate a quantum circuit given in QPENQASM 2.0 format (if no file is given, the\n"
    "                              circuit is read from stdin)\n"
    "  --shots arg                 number of shots\n"
    "  --ps                        print simulation stats (applied gates, sim. time, and maximal size of the DD)\n"
    "  --display_statevector       adds the state-vector to snapshots\n"
    "  --display_probabilities     adds the probabilities of the basis states to snapshots\n"
    "  --precision arg             two numbers are treated to be equal if their difference is smaller than this value\n"
    "  --device arg                CUDA device ordinal (default 0)\n"
    "  --dump_ops                  parse only: print the op list as JSON (no GPU needed)\n"
    "  --serve                     stay resident and serve invocations over $JKU_SIMULATOR_SOCKET\n"
    "  --idle arg                  with --serve: exit after this many seconds without a request (default 600)\n"
    "  --no_server                 do not look for a resident server\n";

struct Options {
  bool help = false, simulate = false, ps = false, statevector = false, probabilities = false, dump_ops = false;
  bool serve = false, no_server = false;
  int idle = 600;
  bool have_seed = false, have_shots = false, have_precision = false;
  unsigned long seed = 0;
  unsigned long shots = 1;
  double precision = 0;
  int device = 0;
  std::string file;
};

bool parse_args(int argc, char **argv, Options &o, std::string &err) {
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i], val;
    bool has_val = false;
    const size_t eq = a.find('=');
    if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; }
    auto need = [&]() {
      if (has_val) return true;
      if (i + 1 < argc) { val = argv[++i]; return true; }
      err = "the required argument for option '" + a + "' is missing";
      return false;
    };
    if (a == "--help") o.help = true;
    else if (a == "--ps") o.ps = true;
    else if (a == "--display_statevector") o.statevector = true;
    else if (a == "--display_probabilities") o.probabilities = true;
    else if (a == "--dump_ops") o.dump_ops = true;
    else if (a == "--serve") o.serve = true;
    else if (a == "--no_server") o.no_server = true;
    else if (a == "--idle") { if (!need()) return false; o.idle = atoi(val.c_str()); }
    else if (a == "--simulate_qasm") {
      o.simulate = true;
      if (has_val) o.file = val;
      else if (i + 1 < argc && strncmp(argv[i + 1], "--", 2) != 0) o.file = argv[++i];
    } else if (a == "--seed") { if (!need()) return false; o.seed = strtoul(val.c_str(), nullptr, 10); o.have_seed = true; }
    else if (a == "--shots") { if (!need()) return false; o.shots = strtoul(val.c_str(), nullptr, 10); o.have_shots = true; }
    else if (a == "--precision") { if (!need()) return false; o.precision = strtod(val.c_str(), nullptr); o.have_precision = true; }
    else if (a == "--device") { if (!need()) return false; o.device = atoi(val.c_str()); }
    else { err = "unrecognised option '" + a + "'"; return false; }
  }
  return true;
}

struct SnapshotData {
  std::vector<double> probabilities;             // reference index order (first argument = MSB); empty = not recorded
  std::vector<std::string> statevector;          // empty = not recorded
  // a register too wide to hold 2^n strings: the amplitudes are read from the engine range by range while they are printed
  bool stream = false;
  uint32_t slot = 0, k = 0;
  std::vector<uint32_t> qubit_of_bit;            // bit b of the printed index = this qubit
};
// from here on a state vector is streamed (JKQ_STREAM_MIN_QUBITS: testing aid, puts small registers on the streaming path)
static uint32_t stream_min_qubits() {
  const char *e = getenv("JKQ_STREAM_MIN_QUBITS");
  return e ? (uint32_t)atoi(e) : 21u;
}
#define STREAM_MIN_QUBITS stream_min_qubits()
static const uint64_t STREAM_CHUNK = 1ull << 20;

// Cprint (QMDDcomplexD.cpp:100-135): parts that snap to zero are left out, "0" when both do
std::string format_amplitude(double re, double im, double tol) {
  const bool has_re = fabs(re) > tol, has_im = fabs(im) > tol;
  if (!has_re && !has_im) return "0";
  std::ostringstream os;
  if (has_re) {
    if (re < 0) os << "-";
    os << fabs(re);
  }
  if (has_im) os << (im < 0 ? "-" : "+") << fabs(im) << "i";
  return os.str();
}

struct Fatal : std::runtime_error {  // what used to be exit(1) in the middle of a run: a request of the server must not end it
  explicit Fatal(const std::string &m) : std::runtime_error(m) {}
};

struct Runner {
  std::ostream *out = &std::cout, *errs = &std::cerr;
  jkq_handle *h = nullptr;
  const jkq_program *prog = nullptr;
  jkq_program_info info;
  double tol = 1e-10;
  std::map<uint32_t, SnapshotData> snapshots;     // by slot, like the reference's std::map<int, Snapshot*>
  std::vector<std::vector<double>> shown;         // show_probabilities statements, in order
  uint64_t gates = 0, max_nodes = 0;
  bool want_stats = false;

  void die(const char *what, int rc) {
    *errs << "ERROR: " << what << " failed (" << rc << "): " << (h ? jkq_last_error(h) : "") << std::endl;
    throw Fatal(what);
  }

  // load ops[0, n_ops), apply
  void apply(size_t n_ops, uint64_t measure_seed) {
    int rc = jkq_load_circuit(h, info.n_qubits, info.ops, n_ops);
    if (rc) die("jkq_load_circuit", rc);
    jkq_set_option(h, "measure_seed", (int64_t)measure_seed);
    rc = jkq_apply(h);
    if (rc == JKQ_E_ZERO_STATE) {
      *errs << "ERROR: numerical instabilities led to a 0-vector! Abort simulation!" << std::endl;  // Simulator.cpp:239
      throw Fatal("zero vector");
    }
    if (rc) die("jkq_apply", rc);
    jkq_stats st;
    if (!jkq_get_stats(h, &st)) gates = st.gates;
    if (want_stats) {
      size_t n = 0;
      if (jkq_get_level_log(h, nullptr, 0, &n)) n = 0;
      std::vector<uint32_t> log(n ? n : 1);
      if (n && !jkq_get_level_log(h, log.data(), n, &n)) {
        for (size_t g = 0; g + info.n_qubits <= n; g += info.n_qubits) {
          uint64_t s = 0;
          for (uint32_t v = 0; v < info.n_qubits; v++) s += log[g + v];
          if (s > max_nodes) max_nodes = s;
        }
      }
    }
  }

  // read back what the snapshot / show_probabilities statements among ops[0, n_ops) recorded
  void collect(size_t n_ops, bool statevector, bool probabilities) {
    std::vector<uint32_t> qubits(64);
    for (uint32_t s = 0; s < info.n_snapshots; s++) {
      uint32_t slot = 0;
      size_t k = 0;
      jkq_program_get_snapshot(prog, s, &slot, qubits.data(), qubits.size(), &k);
      SnapshotData d;
      if (probabilities) {
        size_t len = 0;
        if (jkq_get_probabilities(h, slot, nullptr, 0, &len)) continue;  // (the statement lies behind n_ops)
        std::vector<double> raw(len);
        int rc = jkq_get_probabilities(h, slot, raw.data(), len, &len);
        if (rc) die("jkq_get_probabilities", rc);
        // engine order: bit j = j-th lowest listed qubit; reference order: bit k-1-a = argument a
        std::vector<uint32_t> rank(k);
        for (size_t a = 0; a < k; a++) {
          uint32_t r = 0;
          for (size_t b = 0; b < k; b++) r += qubits[b] < qubits[a];
          rank[a] = r;
        }
        d.probabilities.resize(len);
        for (size_t i = 0; i < len; i++) {
          size_t e = 0;
          for (size_t a = 0; a < k; a++) e |= ((i >> (k - 1 - a)) & 1) << rank[a];
          d.probabilities[i] = raw[e];
        }
      }
      if (statevector && k == info.n_qubits && k >= STREAM_MIN_QUBITS) {
        double probe[2];
        d.qubit_of_bit.resize(k);
        for (size_t a = 0; a < k; a++) d.qubit_of_bit[k - 1 - a] = qubits[a];  // first argument = most significant bit
        if (jkq_get_snapshot_amplitudes_ordered(h, slot, d.qubit_of_bit.data(), 0, 1, probe) == 0) {  // (else: the statement lies behind n_ops)
          d.stream = true; d.slot = slot; d.k = (uint32_t)k;
        }
      } else if (statevector && k == info.n_qubits) {
        const size_t len = (size_t)1 << k;
        std::vector<double> amps(2 * len);
        if (jkq_get_statevector(h, slot, amps.data(), len) == 0) {
          d.statevector.resize(len);
          for (size_t i = 0; i < len; i++) {
            size_t e = 0;  // engine index: bit q = qubit q
            for (size_t a = 0; a < k; a++) e |= ((i >> (k - 1 - a)) & 1) << qubits[a];
            d.statevector[i] = format_amplitude(amps[2 * e], amps[2 * e + 1], tol);
          }
        }
      } else if (statevector && n_ops == info.n_ops) {
        *errs << "Snapshot must contain all qubits when containing statevector!" << std::endl;  // QASMsimulator.cpp:1014
      }
      snapshots[slot] = d;  // (an entry per snapshot statement even when nothing was asked for, QASMsimulator.cpp:1036)
    }
    shown.clear();
    for (uint32_t i = 0;; i++) {
      size_t len = 0;
      if (jkq_get_probabilities(h, JKQ_SLOT_SHOW_PROBABILITIES + i, nullptr, 0, &len)) break;
      std::vector<double> raw(len);
      if (jkq_get_probabilities(h, JKQ_SLOT_SHOW_PROBABILITIES + i, raw.data(), len, &len)) break;
      shown.push_back(raw);
    }
  }
};

void dump_ops(const jkq_program_info &info, std::ostream &cout_) {
  std::ostream &out = cout_;
  out.precision(17);
  out << "{\"n_qubits\": " << info.n_qubits << ", \"n_clbits\": " << info.n_clbits << ", \"gate_after_measure\": " << info.gate_after_measure
            << ", \"first_measure\": " << info.first_measure << ", \"ops\": [";
  for (size_t i = 0; i < info.n_ops; i++) {
    const jkq_op &o = info.ops[i];
    out << (i ? ",\n " : "\n ") << "{\"kind\": " << o.kind << ", \"target\": " << o.target << ", \"ctrl_pos\": " << o.ctrl_pos
              << ", \"ctrl_neg\": " << o.ctrl_neg << ", \"arg\": " << o.arg << ", \"m\": [";
    for (int k = 0; k < 8; k++) out << (k ? ", " : "") << o.m[k];
    out << "]}";
  }
  out << "\n]}" << std::endl;
}

}  // namespace

// One invocation: options + program text in, what the process would print out.  `cache` keeps the engine handle of
// the previous request alive (resident server); a one-shot process passes an empty one.
struct HandleCache {
  jkq_handle *h = nullptr;
  int device = -1;
  double precision = 0;
};

static int run_request(const Options &opt, const std::string &text, const std::string &dir, std::ostream &cout_, std::ostream &cerr_,
                       HandleCache &cache) {
  std::ostream &out = cout_, &errs = cerr_;
  const unsigned long seed = opt.have_seed ? opt.seed : (unsigned long)time(nullptr);
  if (opt.shots < 1) errs << "Shots have to be greater than 0!" << std::endl;  // QASMsimulator.cpp:480-482

  jkq_program *prog = nullptr;
  const uint32_t flags = (opt.statevector ? JKQ_QASM_STATEVECTOR : 0) | (opt.probabilities ? JKQ_QASM_PROBABILITIES : 0);
  if (jkq_qasm_parse(text.c_str(), dir.empty() ? nullptr : dir.c_str(), flags, &prog)) {
    errs << "ERROR: " << jkq_program_error(prog) << std::endl;
    jkq_program_free(prog);
    return 1;
  }
  Runner r;
  r.out = &out; r.errs = &errs;
  r.prog = prog;
  jkq_program_get_info(prog, &r.info);
  if (opt.dump_ops) {
    dump_ops(r.info, out);
    jkq_program_free(prog);
    return 0;
  }

  jkq_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.device = opt.device;
  if (opt.have_precision) { cfg.tolerance = opt.precision; r.tol = opt.precision; }
  int rc = 0;
  if (cache.h && (cache.device != opt.device || cache.precision != (opt.have_precision ? opt.precision : 0.0))) {
    jkq_destroy(cache.h);
    cache.h = nullptr;
  }
  if (!cache.h) {
    rc = jkq_create(&cache.h, &cfg);
    if (rc) {
      errs << "ERROR: no usable CUDA device (" << rc << "): this simulator has no CPU path" << std::endl;
      if (cache.h) jkq_destroy(cache.h);
      cache.h = nullptr;
      jkq_program_free(prog);
      return 1;
    }
    cache.device = opt.device;
    cache.precision = opt.have_precision ? opt.precision : 0.0;
  }
  r.h = cache.h;
  r.want_stats = opt.ps;
  jkq_set_option(r.h, "level_log", opt.ps ? 1 : 0);
  try {

  const auto t1 = std::chrono::high_resolution_clock::now();
  const jkq_program_info &info = r.info;
  const unsigned long shots = opt.shots < 1 ? 1 : opt.shots;
  std::map<std::string, unsigned long long> result;
  auto add_counts = [&]() {
    size_t n = 0;
    jkq_get_counts(r.h, nullptr, nullptr, 0, &n);
    std::vector<uint64_t> keys(n ? n : 1), cnts(n ? n : 1);
    int rc2 = jkq_get_counts(r.h, keys.data(), cnts.data(), n, &n);
    if (rc2) r.die("jkq_get_counts", rc2);
    for (size_t i = 0; i < n; i++) {
      std::string s(info.n_qubits, '0');
      for (uint32_t q = 0; q < info.n_qubits; q++) s[q] = (char)('0' + ((keys[i] >> q) & 1));
      result[s] += cnts[i];
    }
  };
  if (!info.gate_after_measure) {
    size_t prefix = info.first_measure >= 0 ? (size_t)info.first_measure : info.n_ops;
    // a guarded first measurement (`if (c == 0) measure q -> c;`): the guard belongs to the measurement, the state
    // "before the first measurement" ends in front of it (an IF op with nothing behind it is not a valid op list)
    while (prefix > 0 && prefix < info.n_ops && info.ops[prefix - 1].kind == JKQ_OP_IF) prefix--;
    if (prefix < info.n_ops) {
      // statements behind the first measurement see the collapsed state (Simulate() runs to the end before the
      // shots are drawn from `beforeMeasurement`)
      r.apply(info.n_ops, seed);
      r.collect(info.n_ops, opt.statevector, opt.probabilities);
    }
    r.apply(prefix, seed);
    if (prefix == info.n_ops) r.collect(info.n_ops, opt.statevector, opt.probabilities);
    rc = jkq_sample(r.h, shots, seed, 0, 0);
    if (rc) r.die("jkq_sample", rc);
    add_counts();
  } else {
    // Per-shot re-simulation (QASMsimulator.cpp:501-522) as one kernel launch: every shot is an instance of the
    // circuit on its own SM with its own measurement outcomes.  Shot s: measurement stream `mseed + s`, final
    // MeasureAll = shot s of stream `seed` -- the same draws as the sequential loop below, which serves the shots
    // the single-CTA path could not finish.
    const uint64_t mseed = (uint64_t)seed * 0x9E3779B97F4A7C15ull;
    std::vector<uint64_t> keys(shots);
    std::vector<int32_t> status(shots, 1);
    jkq_batch *batch = nullptr;
    uint64_t batch_gates = 0;
    if (jkq_batch_create(&batch, opt.device, 1 << 16, &cfg) == 0) {
      if (jkq_batch_run_shots(batch, info.n_qubits, info.ops, info.n_ops, shots, seed, 0, mseed, 0, keys.data(), nullptr, status.data(),
                              &batch_gates, nullptr) != 0)
        std::fill(status.begin(), status.end(), 1);
    }
    if (batch) jkq_batch_destroy(batch);
    for (unsigned long s = 0; s < shots; s++) {
      if (status[s]) continue;  // (redone below on the full engine)
      std::string str(info.n_qubits, '0');
      for (uint32_t q = 0; q < info.n_qubits; q++) str[q] = (char)('0' + ((keys[s] >> q) & 1));
      result[str] += 1;
    }
    bool want_last = info.n_snapshots > 0 || opt.ps;  // snapshots, show_probabilities and stats come from the last shot's run
    for (size_t i = 0; i < info.n_ops; i++) want_last = want_last || info.ops[i].kind == JKQ_OP_PROBABILITIES;
    for (unsigned long s = 0; s < shots; s++) {
      if (!status[s] && !(want_last && s == shots - 1)) continue;
      r.apply(info.n_ops, mseed + s);
      if (!status[s]) continue;  // (only re-run for its snapshots)
      rc = jkq_sample(r.h, 1, seed, s, 0);
      if (rc) r.die("jkq_sample", rc);
      add_counts();
    }
    r.collect(info.n_ops, opt.statevector, opt.probabilities);
  }
  const auto t2 = std::chrono::high_resolution_clock::now();

  // show_probabilities statements (QASMsimulator.cpp:1039-1055) print while the reference simulates: before the JSON
  for (const std::vector<double> &p : r.shown) {
    out << "Probabilities of the states |";
    for (int i = (int)info.n_qubits - 1; i >= 0; i--) {
      char name[64] = "q";
      uint32_t off = 0, size = 0;
      for (uint32_t k = 0; k < info.n_qregs; k++) {
        char nm[48];
        jkq_program_get_register(prog, 0, k, nm, sizeof(nm), &off, &size);
        if ((uint32_t)i >= off && (uint32_t)i < off + size) { snprintf(name, sizeof(name), "%s[%u]", nm, (uint32_t)i - off); break; }
      }
      out << name << " ";
    }
    out << ">:" << std::endl;
    for (size_t i = 0; i < p.size(); i++) {
      out << "  |";
      for (int j = (int)info.n_qubits - 1; j >= 0; j--) out << ((i >> j) & 1);
      out << ">: " << p[i] << std::endl;
    }
  }

  out << "{" << std::endl << "  \"counts\": {" << std::endl;
  {
    bool first = true;
    for (const auto &kv : result) {
      if (!first) out << ",\n";
      out << "    \"" << kv.first << "\": " << kv.second;
      first = false;
    }
  }
  out << "  }";
  if (!r.snapshots.empty()) {
    out << "," << std::endl;
    out << "  \"snapshots\": {" << std::endl;
    for (auto it = r.snapshots.begin(); it != r.snapshots.end(); ++it) {
      const SnapshotData &d = it->second;
      out << "    \"" << it->first << "\": {" << std::endl;
      if (!d.probabilities.empty()) {
        const size_t k = (size_t)llround(log2((double)d.probabilities.size()));
        out << "      \"probabilities\": [" << d.probabilities[0];
        for (size_t i = 1; i < d.probabilities.size(); i++) out << ", " << d.probabilities[i];
        out << "]," << std::endl;
        out << "      \"probabilities_ket\": {";
        bool first = true;
        std::map<std::string, double> ket;
        for (size_t i = 0; i < d.probabilities.size(); i++) {
          if (!(d.probabilities[i] > 0.0)) continue;
          std::string s(k, '0');
          for (size_t j = 0; j < k; j++) s[j] = (char)('0' + ((i >> (k - 1 - j)) & 1));
          ket[s] = d.probabilities[i];
        }
        for (const auto &kv : ket) {
          out << (first ? "" : ", ") << "\"" << kv.first << "\": " << kv.second;
          first = false;
        }
        out << "}";
      }
      if (d.stream) {
        if (!d.probabilities.empty()) out << "," << std::endl;
        out << "      \"statevector\": [";
        const uint64_t len = 1ull << d.k;
        std::vector<double> amps(2 * STREAM_CHUNK);
        for (uint64_t first = 0; first < len; first += STREAM_CHUNK) {
          const uint64_t cnt = len - first < STREAM_CHUNK ? len - first : STREAM_CHUNK;
          int rc = jkq_get_snapshot_amplitudes_ordered(r.h, d.slot, d.qubit_of_bit.data(), first, cnt, amps.data());
          if (rc) r.die("jkq_get_snapshot_amplitudes_ordered", rc);
          for (uint64_t i = 0; i < cnt; i++) out << (first + i ? ", \"" : "\"") << format_amplitude(amps[2 * i], amps[2 * i + 1], r.tol) << "\"";
        }
        out << "]";
      }
      if (!d.statevector.empty()) {
        if (!d.probabilities.empty()) out << "," << std::endl;
        out << "      \"statevector\": [\"" << d.statevector[0];
        for (size_t i = 1; i < d.statevector.size(); i++) out << "\", \"" << d.statevector[i];
        out << "\"]";
      }
      out << std::endl << "    }" << (std::next(it, 1) != r.snapshots.end() ? "," : "") << std::endl;
    }
    out << "  }" << std::endl;
  }
  out << "}" << std::endl;

  if (opt.ps) {
    const std::chrono::duration<float> diff = t2 - t1;
    out << std::endl << "SIMULATION STATS: " << std::endl;
    out << "  Number of applied gates: " << r.gates << std::endl;
    out << "  Simulation time: " << diff.count() << " seconds" << std::endl;
    out << "  Maximal size of DD (number of nodes) during simulation: " << r.max_nodes << std::endl;
  }
  } catch (const Fatal &) {
    jkq_program_free(prog);
    return 1;
  }
  jkq_program_free(prog);
  return 0;
}

// ---------------------------------------------------------------------------------------------- resident server
// Wire format, both directions little-endian on one machine: a request is
//   u32 argc, argc x (u32 len, bytes)   the argv of the invocation (without argv[0])
//   u64 len, bytes                      the program text
//   u32 len, bytes                      the directory `include` statements are resolved against
// and the reply is  i32 exit code, u64 len + bytes of stdout, u64 len + bytes of stderr.
namespace {

std::string socket_path() {
  const char *e = getenv("JKU_SIMULATOR_SOCKET");
  if (e && *e) return e;
  char buf[96];
  snprintf(buf, sizeof(buf), "/tmp/jku_simulator.%u.sock", (unsigned)getuid());
  return buf;
}

bool write_all(int fd, const void *p, size_t n) {
  const char *c = (const char *)p;
  while (n) {
    ssize_t k = write(fd, c, n);
    if (k < 0) { if (errno == EINTR) continue; return false; }
    c += k; n -= (size_t)k;
  }
  return true;
}
bool read_all(int fd, void *p, size_t n) {
  char *c = (char *)p;
  while (n) {
    ssize_t k = read(fd, c, n);
    if (k < 0) { if (errno == EINTR) continue; return false; }
    if (k == 0) return false;
    c += k; n -= (size_t)k;
  }
  return true;
}
bool put_str(int fd, const std::string &s, bool wide) {
  if (wide) { uint64_t n = s.size(); if (!write_all(fd, &n, 8)) return false; }
  else { uint32_t n = (uint32_t)s.size(); if (!write_all(fd, &n, 4)) return false; }
  return write_all(fd, s.data(), s.size());
}
bool get_str(int fd, std::string &s, bool wide, uint64_t limit) {
  uint64_t n = 0;
  if (wide) { if (!read_all(fd, &n, 8)) return false; }
  else { uint32_t m = 0; if (!read_all(fd, &m, 4)) return false; n = m; }
  if (n > limit) return false;
  s.resize((size_t)n);
  return n == 0 || read_all(fd, &s[0], (size_t)n);
}

int connect_server() {
  const std::string path = socket_path();
  sockaddr_un a;
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  if (path.size() >= sizeof(a.sun_path)) return -1;
  strcpy(a.sun_path, path.c_str());
  int fd = socket(AF_UNIX, SOCK_STREAM, 0);
  if (fd < 0) return -1;
  if (connect(fd, (sockaddr *)&a, sizeof(a)) != 0) { close(fd); return -1; }
  return fd;
}

// forward this invocation; false = no server answered (the caller runs it itself)
bool try_forward(int argc, char **argv, const std::string &text, const std::string &dir, int *exit_code) {
  int fd = connect_server();
  if (fd < 0) return false;
  bool ok = true;
  uint32_t n = (uint32_t)(argc - 1);
  ok = ok && write_all(fd, &n, 4);
  for (int i = 1; ok && i < argc; i++) ok = put_str(fd, argv[i], false);
  ok = ok && put_str(fd, text, true) && put_str(fd, dir, false);
  int32_t code = 1;
  std::string so, se;
  ok = ok && read_all(fd, &code, 4) && get_str(fd, so, true, 1ull << 40) && get_str(fd, se, true, 1ull << 40);
  close(fd);
  if (!ok) return false;
  fwrite(so.data(), 1, so.size(), stdout);
  fwrite(se.data(), 1, se.size(), stderr);
  fflush(stdout);
  *exit_code = code;
  return true;
}

int serve(int idle_s) {
  const std::string path = socket_path();
  if (int fd = connect_server(); fd >= 0) {  // somebody is serving already
    close(fd);
    return 0;
  }
  unlink(path.c_str());
  sockaddr_un a;
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  if (path.size() >= sizeof(a.sun_path)) { std::cerr << "socket path too long" << std::endl; return 1; }
  strcpy(a.sun_path, path.c_str());
  int ls = socket(AF_UNIX, SOCK_STREAM, 0);
  if (ls < 0 || bind(ls, (sockaddr *)&a, sizeof(a)) != 0 || listen(ls, 16) != 0) {
    std::cerr << "cannot listen on " << path << ": " << strerror(errno) << std::endl;
    return 1;
  }
  chmod(path.c_str(), 0600);
  signal(SIGPIPE, SIG_IGN);
  HandleCache cache;
  {  // pay the CUDA start-up now, not on the first request
    jkq_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    if (jkq_create(&cache.h, &cfg) == 0) { cache.device = 0; cache.precision = 0.0; }
    else { if (cache.h) jkq_destroy(cache.h); cache.h = nullptr; }
  }
  for (;;) {
    pollfd pf;
    pf.fd = ls; pf.events = POLLIN; pf.revents = 0;
    int pr = poll(&pf, 1, idle_s > 0 ? idle_s * 1000 : -1);
    if (pr == 0) break;  // idle: leave
    if (pr < 0) { if (errno == EINTR) continue; break; }
    int fd = accept(ls, nullptr, nullptr);
    if (fd < 0) continue;
    uint32_t n = 0;
    std::vector<std::string> args;
    std::string text, dir;
    bool ok = read_all(fd, &n, 4) && n < 4096;
    for (uint32_t i = 0; ok && i < n; i++) { std::string s; ok = get_str(fd, s, false, 1 << 20); args.push_back(s); }
    ok = ok && get_str(fd, text, true, 1ull << 32) && get_str(fd, dir, false, 1 << 16);
    if (ok) {
      std::vector<char *> av;
      std::string a0 = "jku_simulator";
      av.push_back(&a0[0]);
      for (std::string &s : args) av.push_back(&s[0]);
      Options opt;
      std::string perr;
      std::ostringstream so, se;
      int32_t code = 1;
      if (!parse_args((int)av.size(), av.data(), opt, perr)) se << perr << std::endl;
      else if (opt.serve) code = 0;  // (a second --serve: nothing to do)
      else code = run_request(opt, text, dir, so, se, cache);
      const std::string o = so.str(), e = se.str();
      write_all(fd, &code, 4) && put_str(fd, o, true) && put_str(fd, e, true);
    }
    close(fd);
  }
  close(ls);
  unlink(path.c_str());
  if (cache.h) jkq_destroy(cache.h);
  return 0;
}

void autostart_server(const char *self, int idle_s) {
  pid_t p = fork();
  if (p != 0) return;  // parent (or failure): carry on
  setsid();
  if (fork() != 0) _exit(0);
  int nul = open("/dev/null", O_RDWR);
  if (nul >= 0) { dup2(nul, 0); dup2(nul, 1); dup2(nul, 2); }
  char idle[32];
  snprintf(idle, sizeof(idle), "--idle=%d", idle_s);
  execl(self, self, "--serve", idle, (char *)nullptr);
  _exit(1);
}

}  // namespace

int main(int argc, char **argv) {
  Options opt;
  std::string err;
  if (!parse_args(argc, argv, opt, err)) {
    std::cerr << err << std::endl;
    return 1;
  }
  if (opt.serve) return serve(opt.idle);
  if (opt.help || !opt.simulate) {
    std::cout << USAGE << "\n";
    return 1;
  }
  std::string text;
  {
    std::stringstream ss;
    if (opt.file.empty()) ss << std::cin.rdbuf();
    else {
      std::ifstream f(opt.file.c_str());
      if (!f) { std::cerr << "Failed to open file '" << opt.file << "'!" << std::endl; return 1; }
      ss << f.rdbuf();
    }
    text = ss.str();
  }
  std::string dir;
  if (!opt.file.empty()) {
    const size_t slash = opt.file.rfind('/');
    if (slash != std::string::npos) dir = opt.file.substr(0, slash);
    if (!dir.empty() && dir[0] != '/') {  // the server has another working directory
      char cwd[4096];
      if (getcwd(cwd, sizeof(cwd))) dir = std::string(cwd) + "/" + dir;
    }
  } else {
    char cwd[4096];
    if (getcwd(cwd, sizeof(cwd))) dir = cwd;  // `include "f";` next to the caller, as for an in-process run
  }
  if (!opt.no_server && !opt.dump_ops) {
    int code = 0;
    if (try_forward(argc, argv, text, dir, &code)) return code;
    const char *as = getenv("JKU_SIMULATOR_AUTOSERVE");
    if (as && *as && strcmp(as, "0") != 0) {
      char self[4096];
      ssize_t k = readlink("/proc/self/exe", self, sizeof(self) - 1);
      if (k > 0) { self[k] = 0; autostart_server(self, opt.idle); }
    }
  }
  HandleCache cache;
  const int rc = run_request(opt, text, opt.file.empty() ? std::string() : dir, std::cout, std::cerr, cache);
  if (cache.h) jkq_destroy(cache.h);
  return rc;
}
