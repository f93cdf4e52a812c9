"""CPU: the oracle (oracle/qmdd_oracle.c, FP64 restatement) against golden outputs of the UNMODIFIED
reference engine (tests/golden/*.json, produced by tests/golden/make_golden.py from oracle/_ref) and
against the reference's own known-answer tests."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from qiskit_jku_provider_b200 import circuits as C

# Reference tolerance: two reals are one table entry iff they differ by <= Ctol = 1e-10
# (QMDDcomplexD.cpp:336); an amplitude is a product of <= n+1 snapped weights.
AMP_TOL = 1e-9

# Per-gate per-level node counts must equal the reference's exactly -- with TWO documented exceptions, both places
# where the reference's own result is decided by something other than the mathematics of the circuit:
#  (1) QMDDnormalize (QMDDpackage.cpp:541-570) picks the child of largest magnitude by exact compare, first index on
#      exact ties.  When two DIFFERENT canonical magnitudes are closer than the snapping error of their parts
#      (4 Ctol sqrt(m)) the true values are equal and the choice is made by 200-bit rounding noise: the two normal forms
#      of one vector then never merge (this is what blows the reference up at Grover iteration 802).
#  (2) Clookup (QMDDcomplexD.cpp:372-466) identifies a new value with the FIRST-inserted entry within the absolute
#      tolerance 1e-10.  Once two genuinely different values (further apart than any rounding of one quantity: > 1e-13)
#      have been identified, the diagram depends on which intermediate results the reference's recursion happened to
#      insert first (Grover n=20: amplitudes 2^-10 make second-order terms ~1e-10; first such snap at gate 102, first
#      differing level count at gate 105) -- SURVEY section 7 hard part 2.
# The oracle counts both per gate ("near_ties_per_gate", "lossy_snaps_per_gate").  The claim checked here:
#   * up to (not including) the first gate with a near tie or a lossy snap the counts are identical;
#   * a later difference is reported with its first gate and that cause (first_difference).
def first_cause(causes):
    """index of the first gate at which the oracle saw a near tie or a lossy snap (None: never, or no information)"""
    if not causes:
        return None
    both = [a or b for a, b in zip(causes.get("near_ties_per_gate", []), causes.get("lossy_snaps_per_gate", []))]
    return next((i for i, k in enumerate(both) if k), None)


def first_difference(ref, got):
    return next((i for i, (a, b) in enumerate(zip(ref, got)) if a != b), None)


def level_parity(ref, got, name, causes=None):
    """`causes`: an oracle result with levels=True (near_ties_per_gate / lossy_snaps_per_gate)"""
    assert len(ref) == len(got)
    d = first_difference(ref, got)
    if d is None:
        return None
    t = first_cause(causes)
    why = "no cause information" if not causes else (
        "first normalisation near tie / lossy value snap at gate %s (the reference's result depends on rounding noise or "
        "insertion order from there on)" % t)
    assert t is not None and t <= d, (
        "%s: level counts differ from the reference first at gate %d: ref %s got %s; %s" % (name, d, ref[d], got[d], why))
    print("%s: level counts identical to the reference for gates [0, %d); first difference at gate %d; %s" % (name, d, d, why))
    return d


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(oracle_mod, name):
    g = load_golden(name)
    r = oracle_mod.run_oracle(g["circuit"], levels=True, amps=True)
    ref = np.array([complex(*a) for a in g["amps"]])
    got = np.array([complex(*a) for a in r["amps"]])
    assert np.abs(ref - got).max() < AMP_TOL
    level_parity(g["level_counts_per_gate"], r["level_counts_per_gate"], name, r)


def _golden_circuits():
    import importlib.util
    import os
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return {c.name: c for c in mod.cases()}


@pytest.mark.parametrize("name", golden_names())
def test_golden_matches_dense_statevector(name):
    """Independent check of the fixtures themselves (stand-in for BasicAer in the reference's tests):
    the reference's amplitudes equal a plain NumPy state-vector simulation of the same circuit."""
    g = load_golden(name)
    circ = _golden_circuits()[name]
    assert circ.text() == g["circuit"]
    psi = C.dense_statevector(circ)
    ref = np.array([complex(*a) for a in g["amps"]])
    assert np.abs(ref - psi).max() < AMP_TOL


def test_kat_gate_x(oracle_mod):
    # test/test_qasm_simulator_jku.py:79-89 -- X, 100 shots -> {'1': 100} exactly
    c = C.Circuit(1).add("x", [0])
    assert oracle_mod.run_oracle(c, shots=100)["counts"] == {"1": 100}


def test_kat_entangle(oracle_mod):
    # test/test_qasm_simulator_jku.py:91-109 -- GHZ-5 from q0: only 00000 / 11111
    r = oracle_mod.run_oracle(C.ghz(5, star=True), shots=100)
    assert set(r["counts"]) <= {"00000", "11111"} and sum(r["counts"].values()) == 100


def test_kat_bell_amplitudes(oracle_mod):
    # test/test_jku_snapshot.py:29-39 -- |a0|^2 = |a3|^2 = 1/2 (5 places), a1 = a2 = 0 exactly
    c = C.Circuit(2).add("h", [0]).add("cx", [0, 1])
    a = [complex(*x) for x in oracle_mod.run_oracle(c, amps=True)["amps"]]
    assert abs(abs(a[0]) ** 2 - 0.5) < 1e-5 and abs(abs(a[3]) ** 2 - 0.5) < 1e-5
    assert a[1] == 0 and a[2] == 0


def test_kat_example_qasm(oracle_mod):
    # test/test_qasm_simulator.py:36-44 -- 8 outcomes 'xyz xyz', each 128 +- 0.04*1024
    g = load_golden("example_qasm")
    r = oracle_mod.run_oracle(g["circuit"], shots=1024, seed=11)
    assert len(r["counts"]) == 8
    for key, v in r["counts"].items():
        assert key[:3] == key[3:]
        assert abs(v - 128) <= 0.04 * 1024 + 1e-9


def test_philox_and_rand_streams_agree_statistically(oracle_mod):
    from scipy.stats import chi2_contingency
    c = C.random_circuit(5, 30, 3)
    a = oracle_mod.run_oracle(c, shots=4000, seed=5)["counts"]
    b = oracle_mod.run_oracle(c, shots=4000, seed=5, extra=["--rng", "rand"])["counts"]
    keys = sorted(k for k in set(a) | set(b) if a.get(k, 0) + b.get(k, 0) >= 10)
    table = np.array([[a.get(k, 0) for k in keys], [b.get(k, 0) for k in keys]])
    assert chi2_contingency(table)[1] > 0.01


def test_workload_sizes():
    # QFT-32 = 2528 applies, Grover-20 = 20 + 102 per iteration, GHZ-24 = 24
    assert len(C.qft(32)) == 2528
    assert len(C.grover(20, 3)) == 20 + 3 * 102
    assert len(C.ghz(24)) == 24
    assert len(C.random_circuit(18, 40, 1)) == 161
