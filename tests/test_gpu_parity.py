"""GPU: parity of the CUDA engine, called through the C ABI (ctypes -> libjkq_b200.so), with the oracle
and with the golden outputs of the unmodified reference.  Run with `-m gpu` on the B200 box."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from test_oracle import level_parity
from qiskit_jku_provider_b200 import circuits as C

pytestmark = pytest.mark.gpu

# amplitude tolerance: the reference's complex-table tolerance Ctol = 1e-10 per snapped weight
# (QMDDcomplexD.cpp:336), an amplitude being a product of <= n+1 weights
AMP_TOL = 1e-9
TIE = ["--tie-tol"]  # the oracle restating the engine's tie rule (tests/test_emu_engine.py)


@pytest.fixture(scope="module")
def eng():
    from qiskit_jku_provider_b200.engine import Engine
    e = Engine(device=0, node_capacity=1 << 21)
    yield e
    e.close()


def text_to_circuit(g):
    import importlib.util
    import os
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return {c.name: c for c in mod.cases()}[g["name"]]


def run(eng, circ, shots, seed, mode, levels=True):
    eng.set_option("force_mode", mode)
    eng.set_option("level_log", 1 if levels else 0)
    eng.load(circ)
    eng.run(shots, seed)


@pytest.mark.parametrize("mode", [1, 2], ids=["cta", "grid"])
@pytest.mark.parametrize("name", golden_names())
def test_golden_reference_parity(eng, oracle_mod, name, mode):
    g = load_golden(name)
    circ = text_to_circuit(g)
    run(eng, circ, 2048, 5, mode)
    amps = eng.statevector()
    ref = np.array([complex(*a) for a in g["amps"]])
    assert np.abs(amps - ref).max() < AMP_TOL
    lg = eng.level_log(len(circ))
    o = oracle_mod.run_oracle(circ, shots=2048, seed=5, levels=True, extra=TIE)
    level_parity(g["level_counts_per_gate"], lg, name, o)  # vs the unmodified reference
    assert lg == o["level_counts_per_gate"]                    # vs the oracle, every gate, every level
    assert eng.level_counts() == o["level_counts"]
    assert eng.counts_bitstrings() == o["counts"]              # same Philox stream -> identical counts
    assert abs(eng.stats()["norm"] - 1.0) < 1e-9


CONFIGS = [C.ghz(24), C.ghz(24, star=True), C.qft(16), C.qft(24), C.grover(12, 20, marked=0x5A5),
           C.random_circuit(18, 40, 1), C.random_circuit(18, 40, 2), C.random_circuit(18, 40, 3),
           C.random_circuit(18, 80, 1), C.random_circuit(12, 80, 5)]


@pytest.mark.parametrize("circ", CONFIGS, ids=lambda c: c.name)
def test_oracle_parity_on_baseline_shapes(eng, oracle_mod, circ):
    run(eng, circ, 4096, 3, 0)
    k = min(1 << 12, 1 << circ.n)
    o = oracle_mod.run_oracle(circ, shots=4096, seed=3, levels=True, amps=True, extra=["--amp-limit", str(k)] + TIE)
    oa = np.array([complex(*a) for a in o["amps"]])
    assert np.abs(eng.amplitudes(0, k) - oa).max() < AMP_TOL
    assert eng.level_log(len(circ)) == o["level_counts_per_gate"]
    assert eng.counts_bitstrings() == o["counts"]


def test_qft32_full_size(eng, oracle_mod):
    """BASELINE config 2 at full size: 2528 applies, DD stays 32 nodes, |amp|^2 uniform = 2^-32."""
    circ = C.qft(32)
    run(eng, circ, 100000, 11, 1)
    o = oracle_mod.run_oracle(circ, shots=100000, seed=11, levels=True, amps=True, extra=["--amp-limit", "4096"] + TIE)
    assert eng.level_log(len(circ)) == o["level_counts_per_gate"]
    assert eng.level_counts() == [1] * 32
    a = eng.amplitudes(0, 4096)
    assert np.abs(a - np.array([complex(*x) for x in o["amps"]])).max() < AMP_TOL
    assert np.allclose(np.abs(a) ** 2 * 2.0 ** 32, 1.0, atol=1e-4)
    assert eng.counts_bitstrings() == o["counts"]
    assert abs(eng.stats()["norm"] - 1.0) < 1e-4


def test_grover20_801_iterations(eng):
    """BASELINE config 3 at full length: 801 iterations = 81 722 applies (the last count at which the reference
    itself stays at 39 nodes).  2 nodes per level throughout, success probability
    sin^2((2k+1) asin(2^-10)) = 0.99997."""
    k = 801
    circ = C.grover(20, k)
    run(eng, circ, 4096, 2, 1, levels=False)
    assert eng.level_counts() == [2] * 19 + [1]
    p = np.sin((2 * k + 1) * np.arcsin(2.0 ** -10)) ** 2
    assert abs(abs(eng.amplitudes(0x5A5A5, 1)[0]) ** 2 - p) < 1e-6 and p > 0.9999
    counts = eng.counts()
    assert counts.get(0x5A5A5, 0) >= 4090
    assert abs(eng.stats()["norm"] - 1.0) < 1e-6


def test_grover20_full_size(eng, oracle_mod):
    """BASELINE config 3: n = 20, native 19-control Z, 2 nodes per level; 60 iterations here (the
    801-iteration run is bench.py's job), checked against the oracle and the analytic success
    probability sin^2((2k+1) asin(2^-10))."""
    k = 60
    circ = C.grover(20, k)
    run(eng, circ, 20000, 2, 1, levels=False)
    o = oracle_mod.run_oracle(circ, shots=20000, seed=2, extra=TIE)
    assert eng.level_counts() == o["level_counts"] == [2] * 19 + [1]
    amp = eng.amplitudes(0x5A5A5, 1)[0]
    p = np.sin((2 * k + 1) * np.arcsin(2.0 ** -10)) ** 2
    assert abs(abs(amp) ** 2 - p) < 1e-6
    assert eng.counts_bitstrings() == o["counts"]


def test_random_blowup_uses_grid_path_and_collector(eng, oracle_mod):
    """BASELINE config 4b: n = 18, 100 ops -> ~1.5k live nodes, 25k created; auto mode must escalate."""
    circ = C.random_circuit(18, 100, 1)
    eng.set_option("cta_item_limit", 256)
    try:
        run(eng, circ, 4096, 3, 0, levels=False)
    finally:
        eng.set_option("cta_item_limit", 4096)
    st = eng.stats()
    assert st["grid_launches"] >= 1
    o = oracle_mod.run_oracle(circ, shots=4096, seed=3, amps=True, extra=["--amp-limit", "65536"] + TIE)
    assert np.abs(eng.amplitudes(0, 65536) - np.array([complex(*x) for x in o["amps"]])).max() < AMP_TOL
    assert abs(st["norm"] - 1.0) < 1e-9
    # sum over levels within 1 % of the oracle (blown-up random DDs: equality only while no two
    # distinct values fall within tol, SURVEY section 7 hard part 2)
    assert abs(sum(eng.level_counts()) - sum(o["level_counts"])) <= 0.01 * sum(o["level_counts"])


def test_collector_under_pressure():
    """Arena of 4096 nodes: many collections, identical result."""
    from qiskit_jku_provider_b200.engine import Engine
    circ = C.qft(12)
    # (the single-CTA path collects by its shared-memory pool sizes; gc_min drives the grid path)
    with Engine(device=0, node_capacity=1 << 12, gc_min=64, force_mode=2) as small, Engine(device=0, node_capacity=1 << 18) as big:
        for e in (small, big):
            e.load(circ)
            e.run(1000, 4)
        assert small.stats()["gc_runs"] >= 3
        assert np.abs(small.statevector() - big.statevector()).max() < AMP_TOL
        assert small.counts() == big.counts()


def test_arena_exhaustion_is_an_error_not_a_crash():
    from qiskit_jku_provider_b200 import JKUSimulatorError
    from qiskit_jku_provider_b200.engine import Engine
    with Engine(device=0, node_capacity=1 << 10, item_capacity=1 << 10) as e:
        e.load(C.random_circuit(14, 80, 2))
        with pytest.raises(JKUSimulatorError):
            e.run(10, 1)
        e.load(C.ghz(5))   # the handle stays usable
        e.run(100, 1)
        assert set(e.counts_bitstrings()) <= {"00000", "11111"}


def test_edge_cases(eng):
    # empty circuit on one qubit, zero shots, measured subset, snapshots mid-circuit
    from qiskit_jku_provider_b200.engine import OP_MEASURE, OP_SNAPSHOT, gate_op, simple_op
    eng.set_option("force_mode", 0)
    eng.set_option("level_log", 0)
    eng.load(C.Circuit(1))
    eng.run(10, 1)
    assert eng.counts() == {0: 10}
    assert eng.statevector().tolist() == [1 + 0j, 0j]
    eng.load(C.ghz(3))
    eng.run(0, 1)
    assert eng.counts() == {}
    h = C.expand("h", [0])[0]
    x = C.expand("cx", [0, 1])[0]
    ops = [gate_op(0, h.matrix()), simple_op(OP_SNAPSHOT, 0, 7), gate_op(1, x.matrix(), x.controls),
           simple_op(OP_MEASURE, 1, 0)]
    eng.load_ops(3, ops)
    eng.run(1000, 9)
    snap = eng.statevector(7)
    assert abs(abs(snap[0]) ** 2 - 0.5) < 1e-12 and abs(abs(snap[1]) ** 2 - 0.5) < 1e-12 and snap[3] == 0
    c = eng.counts()
    assert set(c) == {0, 1} and sum(c.values()) == 1000 and abs(c[0] - 500) < 80
    final = eng.statevector()
    assert final[1] == 0 and final[2] == 0 and abs(abs(final[3]) ** 2 - 0.5) < 1e-12


def test_shot_split_is_exact(eng):
    """Counts of [0,N) == counts of [0,a) + [a,N): the multi-GPU invariant, on one GPU."""
    circ = C.random_circuit(10, 50, 8)
    eng.set_option("level_log", 0)
    eng.load(circ)
    eng.apply()
    eng.sample(50000, seed=21)
    whole = eng.counts()
    eng.sample(12345, seed=21, shot_offset=0)
    eng.sample(50000 - 12345, seed=21, shot_offset=12345, accumulate=True)
    assert eng.counts() == whole


def test_dd_export_import_roundtrip(eng):
    from qiskit_jku_provider_b200.engine import Engine
    circ = C.qft(10)
    eng.set_option("level_log", 0)
    eng.load(circ, measure={q: q for q in range(10)})
    eng.apply()
    eng.sample(20000, seed=5)
    ref = eng.counts()
    nodes, root = eng.dd_export()
    assert len(nodes) == 11 and root == 1
    with Engine(device=0, node_capacity=1 << 12) as other:
        other.dd_import(10, nodes, root, eng.measure_map())
        other.sample(20000, seed=5)
        assert other.counts() == ref


def test_sampling_statistics_large(eng):
    """1e7 shots on GHZ-24: two outcomes, chi-squared p > 0.01 against 1/2, 1/2."""
    from scipy.stats import chisquare
    eng.set_option("level_log", 0)
    eng.load(C.ghz(24))
    eng.run(10 ** 7, 1234)
    c = eng.counts()
    assert set(c) == {0, (1 << 24) - 1} and sum(c.values()) == 10 ** 7
    assert chisquare(list(c.values())).pvalue > 0.01
    # total-variation distance to the exact distribution
    assert abs(c[0] / 1e7 - 0.5) < 1e-3


def test_cursor_handed_to_the_grid_path_when_the_cta_pools_run_out(oracle_mod):
    """A tiny shared-memory node pool: the single-CTA kernel cannot even rebuild the levels above its cursor,
    saves the cursor, and the grid path continues from it (jkq_apply.cuh run_gates / mode_enter)."""
    from qiskit_jku_provider_b200.engine import Engine
    for circ in (C.random_circuit(14, 80, 2), C.qft(12), C.grover(10, 4, marked=0x155)):
        o = oracle_mod.run_oracle(circ, amps=True, extra=["--amp-limit", "16384"] + TIE)
        oa = np.array([complex(*x) for x in o["amps"]])
        for limit in (96, 160):
            with Engine(device=0, node_capacity=1 << 18) as e:
                e.set_option("cta_node_limit", limit)
                e.load(circ)
                e.apply()
                if circ.name.startswith("random"):
                    assert e.stats()["grid_launches"] >= 1
                k = min(16384, 1 << circ.n)
                assert np.abs(e.amplitudes(0, k) - oa[:k]).max() < AMP_TOL, (circ.name, limit)
                lc = e.level_counts()
                if circ.name.startswith("random"):  # blown-up DD: equal up to values that straddle the tolerance
                    assert abs(sum(lc) - sum(o["level_counts"])) <= 0.01 * sum(o["level_counts"]), (circ.name, limit)
                else:
                    assert lc == o["level_counts"], (circ.name, limit)


def test_maximum_width_64_qubits(eng):
    """JKQ_MAXQ = 64 levels: 64-bit outcome keys, a 63-control gate, amplitudes addressed by a 64-bit index."""
    eng.set_option("force_mode", 0)
    eng.set_option("level_log", 0)
    n = 64
    eng.load(C.ghz(n))
    eng.run(2000, 5)
    c = eng.counts()
    assert set(c) <= {0, (1 << n) - 1} and sum(c.values()) == 2000 and abs(c[0] - 1000) < 120
    assert eng.level_counts() == [2] * (n - 1) + [1]
    assert abs(abs(eng.amplitudes((1 << n) - 1, 1)[0]) ** 2 - 0.5) < 1e-12
    circ = C.Circuit(n, name="mcz64")
    for q in range(n):
        circ.add("h", [q])
    circ.ops.append(C.MCG("z", n - 1, list(range(n - 1))))
    eng.load(circ)
    eng.apply()
    a0, a1 = eng.amplitudes(0, 1)[0], eng.amplitudes((1 << n) - 1, 1)[0]
    # only |1...1> changes sign; 2^-32 = 2.3e-10 sits at the table tolerance (Ctol = 1e-10 absolute,
    # QMDDcomplexD.cpp:336): (1/sqrt2)^63 and (1/sqrt2)^64 are one table entry, in the reference as well
    assert a1 == -a0 and abs(abs(a0) - 2.0 ** -32) < 1.01e-10
    assert eng.level_counts() == [2] * (n - 1) + [1]
    assert abs(eng.qubit_probability(17) - 0.5) < 1e-12


def test_failed_single_cta_attempt_leaves_the_arena_counters_alone():
    """Regression (found by scripts/parity_sweep.py): the state of random10_90_s11 has more reals than the shared
    value table by the time shots are asked for; the single-CTA attempt of build_sampling finds it too large to load,
    and used to clamp the arena's value count to the pool size on its way out -- every amplitude came out wrong."""
    import numpy as np
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.engine import Engine
    import oracle
    circ = C.random_circuit(10, 90, 11)
    ref = oracle.run_oracle(circ, amps=True)
    want = np.array([complex(*a) for a in ref["amps"]])
    with Engine(device=0, node_capacity=1 << 18) as eng:
        for mode in (0, 2):
            eng.set_option("force_mode", mode)
            eng.load(circ)
            eng.run(4000, 3)
            assert abs(eng.stats()["norm"] - 1.0) < 1e-9, mode
            assert np.abs(eng.statevector() - want).max() < 1e-9, mode


def test_gate_is_not_redone_when_the_collection_after_it_runs_out_of_the_shared_pools():
    """Regression (scripts/parity_sweep.py, seed 35): gate 183 of random6_90_s35 completes on the single-CTA path, the
    rebuild up to the root for the collection after it exhausts the shared value table, and the grid path used to
    start over AT that gate (next_gate was only advanced after the collection) -- the gate was applied twice."""
    import numpy as np
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.engine import Engine
    import oracle
    circ = C.random_circuit(6, 90, 35)
    ref = oracle.run_oracle(circ, amps=True)
    want = np.array([complex(*a) for a in ref["amps"]])
    with Engine(device=0, node_capacity=1 << 18) as eng:
        for mode in (0, 2):
            eng.set_option("force_mode", mode)
            eng.load(circ)
            eng.apply()
            st = eng.stats()
            assert np.abs(eng.statevector() - want).max() < 1e-9, mode
            if mode == 0:
                assert st["cta_launches"] >= 1 and st["grid_launches"] >= 1  # the hand-over did happen


def test_parity_sweep_subset():
    """scripts/parity_sweep.py on a slice of its seeds: random circuits through the single-CTA path, the grid path and
    the automatic mode against a dense simulation (1e-9), counts identical between two runs."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "parity_sweep.py"), "14", "6", "7", "28"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "parity sweep ok" in out.stdout


def test_snapshot_kept_as_a_diagram_and_read_range_by_range(eng):
    """Above `snapshot_dense_max` qubits a SNAPSHOT keeps a copy of the diagram instead of 2^n amplitudes, and
    jkq_get_snapshot_amplitudes walks it range by range (the streaming read-out of SURVEY 8f rank 2).  (1) forced on a
    10-qubit random circuit: bit for bit the dense snapshot, also in a permuted bit order; (2) a mid-circuit snapshot of a
    40-qubit register survives the gates and the collections that follow it."""
    import ctypes as Ct
    from qiskit_jku_provider_b200.engine import OP_SNAPSHOT, circuit_ops, gate_op, simple_op
    eng.set_option("force_mode", 0)
    eng.set_option("level_log", 0)
    circ = C.random_circuit(10, 60, 5)
    base = circuit_ops(circ)
    ops = np.zeros(len(base) + 1, dtype=base.dtype)
    half = len(base) // 2
    ops[:half] = base[:half]
    ops[half]["kind"] = OP_SNAPSHOT
    ops[half]["arg"] = 3
    ops[half + 1:] = base[half:]
    eng.load_ops(10, ops)
    eng.apply()
    dense = eng.statevector(3)
    eng.set_option("snapshot_dense_max", 4)
    try:
        eng.load_ops(10, ops)
        eng.apply()
        walked = np.concatenate([eng.snapshot_amplitudes(3, f, 256) for f in range(0, 1024, 256)])
        assert walked.tolist() == dense.tolist()
        assert eng.statevector(3).tolist() == dense.tolist()
        assert np.abs(dense - C.dense_statevector(C.Circuit(10, circ.ops[:half]))).max() < 1e-9
        perm = np.array([3, 0, 9, 1, 8, 2, 7, 4, 6, 5], dtype=np.uint32)
        out = np.zeros(2 * 1024)
        rc = eng._lib.jkq_get_snapshot_amplitudes_ordered(eng._h, 3, perm.ctypes.data_as(Ct.POINTER(Ct.c_uint32)), 0, 1024,
                                                          out.ctypes.data_as(Ct.POINTER(Ct.c_double)))
        assert rc == 0
        idx = np.arange(1024)
        e = np.zeros(1024, dtype=np.int64)
        for b in range(10):
            e |= ((idx >> b) & 1) << int(perm[b])
        assert out.view(np.complex128).tolist() == dense[e].tolist()
    finally:
        eng.set_option("snapshot_dense_max", 26)
    n = 40
    h, t = C.expand("h", [0])[0], C.expand("t", [n - 1])[0]
    ops = [gate_op(0, h.matrix()), simple_op(OP_SNAPSHOT, 0, 1)]
    for q in range(n - 1):
        cx = C.expand("cx", [q, q + 1])[0]
        ops.append(gate_op(cx.target, cx.matrix(), cx.controls))
    ops += [gate_op(n - 1, t.matrix()), simple_op(OP_SNAPSHOT, 0, 2)]
    eng.load_ops(n, ops)
    eng.apply()
    hm, tm = np.array(h.matrix(), dtype=complex), np.array(t.matrix(), dtype=complex)  # (the U convention carries a global phase)
    first = eng.snapshot_amplitudes(1, 0, 4)          # H|0> on qubit 0, nothing else yet
    assert np.abs(first - np.array([hm[0, 0], hm[1, 0], 0, 0])).max() < 1e-9
    assert eng.snapshot_amplitudes(1, (1 << n) - 2, 2).tolist() == [0j, 0j]
    last = eng.snapshot_amplitudes(2, (1 << n) - 2, 2)  # GHZ with the T phases on its two ends
    assert last[0] == 0 and abs(last[1] - hm[1, 0] * tm[1, 1]) < 1e-9
    assert abs(eng.snapshot_amplitudes(2, 0, 1)[0] - hm[0, 0] * tm[0, 0]) < 1e-9
    assert eng.amplitudes((1 << n) - 1, 1)[0] == last[1]


def test_sums_that_stay_a_chain_take_no_work_items(eng, oracle_mod):
    """H on the node that carries the marked path of a Grover state is a sum of two different sub-diagrams that shares
    everything but one path: thin_probe_sums walks it as a chain (jkq_thin.cuh lin_add) instead of through work items.
    The single-CTA kernel without that code asks for the one with it after a few such gates (JEng.sums_ask), so every
    later gate of the circuit is a thin gate; switched off (thin_off bit 6) the same circuit goes through the item phases
    and ends in the same diagram, gate by gate."""
    circ = C.grover(16, 12, marked=0xA5A5)
    eng.set_option("thin_off", 0)
    run(eng, circ, 4096, 4, 0, levels=False)
    st = eng.stats()
    assert st["thin_gates"] >= len(circ) - 16, (st["thin_gates"], len(circ))  # all but the gates before the relaunch
    a1, lc1, counts1 = eng.amplitudes(0, 1 << 12), eng.level_counts(), eng.counts()
    o = oracle_mod.run_oracle(circ, shots=4096, seed=4, levels=True, amps=True, extra=["--amp-limit", str(1 << 12)] + TIE)
    assert np.abs(a1 - np.array([complex(*x) for x in o["amps"]])).max() < AMP_TOL
    assert lc1 == o["level_counts"]
    try:
        eng.set_option("thin_off", 64)
        run(eng, circ, 4096, 4, 0, levels=False)
        assert eng.stats()["thin_gates"] < st["thin_gates"]
        assert np.abs(eng.amplitudes(0, 1 << 12) - a1).max() < 1e-12
        assert eng.level_counts() == lc1 and eng.counts() == counts1
        # per-gate level counts, both ways, against the oracle (the level log rebuilds to the root after every gate)
        for off in (0, 64):
            eng.set_option("thin_off", off)
            run(eng, circ, 16, 4, 0, levels=True)
            assert eng.level_log(len(circ)) == o["level_counts_per_gate"], off
    finally:
        eng.set_option("thin_off", 0)
