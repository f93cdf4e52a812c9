"""GPU: the jku_simulator executable (csrc/jku_simulator.cpp) against what the UNMODIFIED reference executable printed
for the same QASM programs and flags (tests/golden/cli/*.json, see tests/golden/make_cli_golden.py) -- the exact
boundary the reference's Python backend crosses (qasm_simulator_jku.py:82-97) -- and the engine ops behind it
(IF, PROBABILITIES) through the C ABI."""
import json
import os
import subprocess

import numpy as np
import pytest

from test_qasm_frontend import CLI_DIR, HEADER, PROGRAMS, dense_run, load_case, parse_amplitude, reference_order

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli():
    from qiskit_jku_provider_b200 import _build
    assert os.path.exists(_build.CLI), "jku_simulator was not built (python -c 'import __graft_entry__ as g; g.build()')"
    return _build.CLI


def run_cli(cli, qasm, shots, seed, flags, cwd=None, env=None):
    out = subprocess.run([cli, "--simulate_qasm", "--shots=%d" % shots, "--seed=%d" % seed, "--ps"] + list(flags), input=qasm,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, universal_newlines=True, timeout=600, cwd=cwd,
                         env=dict(os.environ, **env) if env else None)
    assert out.returncode == 0, out.stderr
    text = out.stdout
    body = json.loads(text[text.index("{"):text.index("SIMULATION STATS")])
    stats = {}
    for line in text[text.index("SIMULATION STATS"):].splitlines():
        if ":" in line and not line.startswith("SIMULATION"):
            k, v = line.split(":", 1)
            stats[k.strip()] = v.strip()
    return body, stats, out


def ref_stats(rec):
    stats = {}
    text = rec["stdout"]
    for line in text[text.index("SIMULATION STATS"):].splitlines():
        if ":" in line and not line.startswith("SIMULATION"):
            k, v = line.split(":", 1)
            stats[k.strip()] = v.strip()
    return stats


def total_variation(a, b, shots_a, shots_b):
    keys = set(a) | set(b)
    return 0.5 * sum(abs(a.get(k, 0) / shots_a - b.get(k, 0) / shots_b) for k in keys)


@pytest.mark.parametrize("name", PROGRAMS)
def test_cli_matches_the_reference_executable(cli, name):
    qasm, rec, ref = load_case(name)
    got, stats, out = run_cli(cli, qasm, rec["shots"], rec["seed"], rec["flags"])
    # counts: same keys over all qubits (character i = qubit i); the RNG differs (Philox vs glibc rand), so the
    # histograms agree statistically.  Two samples of `shots` draws over k outcomes: TV distance ~ sqrt(k / shots).
    assert sum(got["counts"].values()) == rec["shots"]
    assert all(len(k) == len(next(iter(ref["counts"]))) for k in got["counts"])
    k = len(set(got["counts"]) | set(ref["counts"]))
    tv = total_variation(got["counts"], ref["counts"], rec["shots"], rec["shots"])
    assert tv < 2.5 * np.sqrt(k / rec["shots"]) + 1e-12, (tv, got["counts"], ref["counts"])
    if len(ref["counts"]) == 1:
        assert got["counts"] == ref["counts"]
    # an outcome the reference can never produce must not appear: supports agree up to rare outcomes
    rare = 5.0 / rec["shots"]
    for key, c in got["counts"].items():
        assert key in ref["counts"] or c / rec["shots"] < 10 * rare, key
    # snapshots
    assert set(got.get("snapshots", {})) == set(ref.get("snapshots", {}))
    for slot, want in ref.get("snapshots", {}).items():
        mine = got["snapshots"][slot]
        assert set(mine) == set(want), slot
        if "probabilities" in want:
            np.testing.assert_allclose(np.array(mine["probabilities"]), np.array(want["probabilities"]), rtol=2e-5, atol=2e-6)
            assert set(mine["probabilities_ket"]) == set(want["probabilities_ket"])
            for key, v in want["probabilities_ket"].items():
                assert abs(mine["probabilities_ket"][key] - v) <= 2e-5 * v + 2e-6
        if "statevector" in want:
            a = np.array([parse_amplitude(s) for s in mine["statevector"]])
            b = np.array([parse_amplitude(s) for s in want["statevector"]])
            np.testing.assert_allclose(a, b, rtol=0, atol=2e-6)
            same = sum(x == y for x, y in zip(mine["statevector"], want["statevector"]))
            assert same >= 0.9 * len(b), (mine["statevector"], want["statevector"])  # the printed strings themselves
    # --ps: ApplyGate calls of the (last) run and the largest DD
    rs = ref_stats(rec)
    if name != "teleport":  # (there the count depends on the last shot's measurement outcomes)
        assert stats["Number of applied gates"] == rs["Number of applied gates"]
        assert stats["Maximal size of DD (number of nodes) during simulation"] == rs["Maximal size of DD (number of nodes) during simulation"]
    else:
        assert 5 <= int(stats["Number of applied gates"]) <= 7
    # the reference complains on stderr about partial snapshots with --display_statevector; so do we
    assert ("Snapshot must contain all qubits" in out.stderr) == ("Snapshot must contain all qubits" in rec["stderr"])


def test_cli_teleport_conditionals_fire(cli):
    """teleportation of |1>: q[2] reads 1 in every shot only if both `if` corrections are applied on the right outcomes"""
    qasm, rec, ref = load_case("teleport")
    got, _, _ = run_cli(cli, qasm, 300, 21, [])
    assert set(got["counts"]) == {"001", "011", "101", "111"}
    assert all(40 <= c <= 110 for c in got["counts"].values())


def test_cli_is_reproducible_and_seed_dependent(cli):
    qasm, rec, _ = load_case("qelib_all")
    a, _, _ = run_cli(cli, qasm, 500, 5, [])
    b, _, _ = run_cli(cli, qasm, 500, 5, [])
    c, _, _ = run_cli(cli, qasm, 500, 6, [])
    assert a["counts"] == b["counts"] and a["counts"] != c["counts"]


def test_cli_counts_equal_the_in_process_engine(cli, eng):
    """same circuit as QASM text through the executable and as ops through the C ABI: identical histograms
    (same DD, same Philox stream)"""
    from qiskit_jku_provider_b200 import circuits as C
    for circ in (C.qft(10), C.random_circuit(9, 50, 4)):
        got, _, _ = run_cli(cli, circ.qasm(measure=True), 3000, 17, [])
        eng.load(circ)
        eng.run(3000, 17)
        assert got["counts"] == eng.counts_bitstrings()


def random_program(seed, n_qubits=6, n_statements=40):
    """random OpenQASM 2.0 text over the qelib1 vocabulary, two registers, a user gate, a full snapshot at the end"""
    import random
    rng = random.Random(seed)
    one = [("h", 0), ("x", 0), ("y", 0), ("z", 0), ("s", 0), ("sdg", 0), ("t", 0), ("tdg", 0), ("id", 0), ("rx", 1), ("ry", 1), ("rz", 1),
           ("u1", 1), ("u2", 2), ("u3", 3)]
    two = [("cx", 0), ("cz", 0), ("cy", 0), ("ch", 0), ("swap", 0), ("crz", 1), ("cu1", 1), ("cu3", 3), ("rzz", 1), ("pair", 2)]
    three = [("ccx", 0), ("cswap", 0)]
    na = n_qubits // 2
    names = ["a[%d]" % i for i in range(na)] + ["b[%d]" % i for i in range(n_qubits - na)]
    lines = ["OPENQASM 2.0;", 'include "qelib1.inc";',
             "gate pair(x, y) p, q { ry(x) p; cx p, q; rz(y - x/2) q; cu1(x*y) q, p; }",
             "qreg a[%d];" % na, "qreg b[%d];" % (n_qubits - na)]
    for _ in range(n_statements):
        r = rng.random()
        gates, k = (one, 1) if r < 0.5 else (two, 2) if r < 0.9 else (three, 3)
        name, n_par = rng.choice(gates)
        args = rng.sample(names, k)
        par = "(%s)" % ",".join(rng.choice(["%.6f" % rng.uniform(-3, 3), "pi/%d" % rng.randint(1, 8), "-pi*%d/%d" % (rng.randint(1, 5), rng.randint(2, 9))])
                                for _ in range(n_par)) if n_par else ""
        lines.append("%s%s %s;" % (name, par, ",".join(args)))
    if rng.random() < 0.5:
        lines.append("h a;")          # a whole-register statement
    lines.append("snapshot(0) %s;" % ",".join(names))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed", range(8))
def test_cli_statevector_of_random_programs_equals_a_dense_simulation(cli, seed):
    from qiskit_jku_provider_b200.qasm import parse_qasm
    text = random_program(seed)
    prog = parse_qasm(text, statevector=True)
    _, states = dense_run(prog)
    want = reference_order(states[0], prog.snapshots[0][1])
    got, _, _ = run_cli(cli, text, 64, seed, ["--display_statevector"])
    mine = np.array([parse_amplitude(a) for a in got["snapshots"]["0"]["statevector"]])
    np.testing.assert_allclose(mine, want, rtol=0, atol=2e-6)  # six printed digits
    assert abs(np.vdot(mine, mine).real - 1.0) < 1e-4
    p = np.abs(want) ** 2
    # counts keys: character i = qubit i; the snapshot lists the qubits in order, so int(key, 2) is its index
    assert all(p[int(k, 2)] > 1e-7 for k in got["counts"])


def test_cli_measured_circuit_too_large_for_one_sm_falls_back_to_the_full_engine(cli):
    """a measured circuit whose DD outgrows the shared-memory pools: the shots kernel reports every shot as unfinished
    and the executable redoes them on the full engine (grid path) -- same counts layout, every shot accounted for"""
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.batch import BatchRunner
    from qiskit_jku_provider_b200.qasm import parse_qasm
    body = C.random_circuit(14, 90, 2).qasm().replace("OPENQASM 2.0;\n", "").replace("qreg q[14];\n", "")
    text = HEADER + "qreg q[14];\ncreg c[1];\n" + body + "measure q[3] -> c[0];\nif (c == 1) x q[0];\nh q[5];\n"
    prog = parse_qasm(text)
    assert prog.gate_after_measure
    with BatchRunner(workers=8, device=0) as br:
        _, _, status, _, _ = br.run_shots(prog.n_qubits, prog.ops, 4, seed=2)
    assert status.all()                                   # none of them fits one SM
    got, stats, _ = run_cli(cli, text, 4, 2, [])
    assert sum(got["counts"].values()) == 4 and all(len(k) == 14 for k in got["counts"])
    assert int(stats["Maximal size of DD (number of nodes) during simulation"]) > 2048


def test_cli_show_probabilities_and_include_file(cli, tmp_path):
    (tmp_path / "qelib1.inc").write_text("gate h a { U(pi/2,0,pi) a; }\ngate cx c,t { CX c,t; }\n")
    out = subprocess.run([cli, "--simulate_qasm", "--shots=10", "--seed=1"], input=HEADER + "qreg q[2];\nh q[0];\ncx q[0],q[1];\nshow_probabilities;\n",
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, universal_newlines=True, timeout=600, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert lines[0].startswith("Probabilities of the states |q[1] q[0] >:")
    assert [l.strip() for l in lines[1:5]] == ["|00>: 0.5", "|01>: 0", "|10>: 0", "|11>: 0.5"]
    assert set(json.loads(out.stdout[out.stdout.index("{"):])["counts"]) <= {"00", "11"}


def test_cli_precision_flag_reaches_the_engine(cli):
    # with a tolerance of 0.2 the tiny rotation snaps to the identity: one outcome; with the default it does not
    prog = HEADER + "qreg q[1];\nry(0.2) q[0];\n"
    coarse, _, _ = run_cli(cli, prog, 2000, 3, ["--precision=0.2"])
    fine, _, _ = run_cli(cli, prog, 2000, 3, [])
    assert set(coarse["counts"]) == {"0"} and set(fine["counts"]) == {"0", "1"}


# ---- the engine ops behind the front end, through the C ABI
@pytest.fixture(scope="module")
def eng():
    from qiskit_jku_provider_b200.engine import Engine
    e = Engine(device=0, node_capacity=1 << 18)
    yield e
    e.close()


def test_probabilities_op_matches_a_dense_simulation(eng):
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.engine import OP_DTYPE, OP_PROBABILITIES, circuit_ops
    for circ, masks in ((C.random_circuit(7, 40, 3), (0b1111111, 0b0010110, 0b1000000, 0b0000001, 0b1100011)),
                        (C.qft(9), (0b111111111, 0b101010101)), (C.ghz(12), (0b100000000001, 0xFFF))):
        base = circuit_ops(circ)
        ops = np.zeros(len(base) + len(masks), dtype=OP_DTYPE)
        ops[:len(base)] = base
        for j, m in enumerate(masks):
            ops[len(base) + j]["kind"] = OP_PROBABILITIES
            ops[len(base) + j]["ctrl_pos"] = m
            ops[len(base) + j]["arg"] = 10 + j
        eng.load_ops(circ.n, ops)
        eng.apply()
        p = np.abs(eng.statevector()) ** 2
        for j, m in enumerate(masks):
            qubits = [q for q in range(circ.n) if (m >> q) & 1]
            want = np.zeros(1 << len(qubits))
            for basis, v in enumerate(p):
                i = sum(((basis >> q) & 1) << r for r, q in enumerate(qubits))
                want[i] += v
            got = eng.probabilities(10 + j)
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-9, err_msg="%s mask %x" % (circ.name, m))


def test_probabilities_op_on_a_large_dd_uses_the_grid_sampling_records(eng):
    """a DD of tens of thousands of nodes: subset probabilities still add up and match single-qubit marginals"""
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.engine import OP_DTYPE, OP_PROBABILITIES, circuit_ops
    circ = C.random_circuit(16, 110, 1)
    base = circuit_ops(circ)
    ops = np.zeros(len(base) + 1, dtype=OP_DTYPE)
    ops[:len(base)] = base
    ops[-1]["kind"] = OP_PROBABILITIES
    ops[-1]["ctrl_pos"] = 0b1000000000100101
    ops[-1]["arg"] = 1
    eng.load_ops(circ.n, ops)
    eng.apply()
    got = eng.probabilities(1)
    assert abs(got.sum() - 1.0) < 1e-8
    qubits = [0, 2, 5, 15]
    for r, q in enumerate(qubits):
        p0 = sum(v for i, v in enumerate(got) if not (i >> r) & 1)
        assert abs(p0 - eng.qubit_probability(q)) < 1e-8


def test_if_op_follows_the_measured_bits(eng):
    from qiskit_jku_provider_b200.engine import OP_IF, OP_MEASURE_ONE, gate_op, simple_op
    from qiskit_jku_provider_b200 import circuits as C
    h = C.Circuit(1).add("h", [0]).ops[0].matrix()
    x = [[0, 1], [1, 0]]
    fired = 0
    for seed in range(30):
        cond = simple_op(OP_IF, 0, 2)          # guards the next two ops
        cond.ctrl_pos, cond.ctrl_neg = 0b1000, 0b1000
        ops = [gate_op(0, h), simple_op(OP_MEASURE_ONE, 0, 3), cond, gate_op(1, x), gate_op(2, x), gate_op(3, x)]
        eng.set_option("measure_seed", 300 + seed)
        eng.load_ops(4, ops)
        eng.apply()
        bit = eng.mid_measurements()[3]
        fired += bit
        p = np.abs(eng.statevector()) ** 2
        want = 0b1000 | (0b0111 if bit else 0)   # q3 always flipped; q1, q2 only when the measured bit is 1; q0 = the outcome
        assert abs(p[want] - 1.0) < 1e-9, (seed, bit, np.argmax(p))
        assert eng.stats()["gates"] == (4 if bit else 2)
    assert 6 <= fired <= 24


def test_if_op_needs_something_to_guard(eng):
    from qiskit_jku_provider_b200.engine import OP_IF, simple_op
    from qiskit_jku_provider_b200.jkusimulatorerror import JKUSimulatorError
    with pytest.raises(JKUSimulatorError):
        eng.load_ops(2, [simple_op(OP_IF, 0, 1)])


def test_cli_guarded_first_measurement(cli):
    """`if (c == 0) measure ...` as the first measurement: the prefix "before the first measurement" ends in
    front of the guard instead of in a dangling IF op."""
    qasm = HEADER + "qreg q[2];\ncreg c[2];\nh q[0];\ncx q[0],q[1];\nif(c==0) measure q[0] -> c[0];\nmeasure q[1] -> c[1];\n"
    got, _, _ = run_cli(cli, qasm, 2000, 3, [])
    assert set(got["counts"]) == {"00", "11"} and sum(got["counts"].values()) == 2000
    assert abs(got["counts"]["00"] - 1000) < 150


def test_cli_resident_server_prints_what_a_fresh_process_prints(cli, tmp_path):
    """`jku_simulator --serve`: every ordinary invocation forwards to it (no CUDA start-up in the short-lived process)
    and prints the very same stdout -- counts, snapshots, statistics -- with the same exit code, also for errors."""
    import time
    env = dict(os.environ, JKU_SIMULATOR_SOCKET=str(tmp_path / "jku.sock"))
    server = subprocess.Popen([cli, "--serve", "--idle=120"], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    try:
        for _ in range(200):
            if os.path.exists(env["JKU_SIMULATOR_SOCKET"]):
                break
            time.sleep(0.05)
        assert os.path.exists(env["JKU_SIMULATOR_SOCKET"])
        for name in PROGRAMS[:4]:
            qasm, rec, _ = load_case(name)
            args = ["--simulate_qasm", "--shots=%d" % rec["shots"], "--seed=%d" % rec["seed"]] + list(rec["flags"])
            fresh = subprocess.run([cli, "--no_server"] + args, input=qasm, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                   universal_newlines=True, timeout=600)
            t0 = time.perf_counter()
            served = subprocess.run([cli] + args, input=qasm, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                    universal_newlines=True, timeout=600)
            dt = time.perf_counter() - t0
            assert served.returncode == fresh.returncode == 0
            assert served.stdout == fresh.stdout, name
            assert dt < 2.0
        bad = subprocess.run([cli, "--simulate_qasm", "--shots=1"], input="OPENQASM 2.0;\nqreg q[1];\nfoo q[0];\n", env=env,
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, universal_newlines=True, timeout=60)
        assert bad.returncode == 1 and "ERROR" in bad.stderr
        assert server.poll() is None  # a failed request does not end the server
    finally:
        server.terminate()
        server.wait(timeout=10)


@pytest.mark.parametrize("seed", range(3))
def test_cli_streamed_statevector_equals_the_one_printed_at_once(cli, seed):
    """A register of 21 qubits or more prints its state vector while reading it from the engine range by range
    (jkq_get_snapshot_amplitudes_ordered; QASMsimulator.cpp:1012-1030 holds 2^n strings).  JKQ_STREAM_MIN_QUBITS puts the
    small random programs of the test above on that path: same text, character for character."""
    text = random_program(seed)
    once, _, _ = run_cli(cli, text, 64, seed, ["--display_statevector"])
    streamed, _, _ = run_cli(cli, text, 64, seed, ["--display_statevector"], env={"JKQ_STREAM_MIN_QUBITS": "1"})
    assert streamed["snapshots"]["0"]["statevector"] == once["snapshots"]["0"]["statevector"]
    assert streamed["counts"] == once["counts"]


def test_cli_streams_the_statevector_of_a_22_qubit_register(cli):
    """2^22 amplitudes through the streaming read-out: GHZ with a phase, every entry but the two ends prints as 0."""
    n = 22
    text = "OPENQASM 2.0;\ninclude \"qelib1.inc\";\nqreg q[%d];\ncreg c[%d];\nh q[0];\n" % (n, n)
    text += "".join("cx q[%d],q[%d];\n" % (i, i + 1) for i in range(n - 1)) + "t q[%d];\nsnapshot(0) %s;\nmeasure q -> c;\n" % (n - 1, ",".join("q[%d]" % i for i in range(n)))
    got, _, _ = run_cli(cli, text, 16, 3, ["--display_statevector"])
    sv = got["snapshots"]["0"]["statevector"]
    assert len(sv) == 1 << n
    a0, a1 = parse_amplitude(sv[0]), parse_amplitude(sv[-1])   # (the U convention carries a global phase)
    assert abs(abs(a0) ** 2 - 0.5) < 1e-5 and abs(abs(a1) ** 2 - 0.5) < 1e-5 and abs(a1 / a0 - np.exp(0.25j * np.pi)) < 1e-5
    assert all(a == "0" for a in sv[1:-1])
