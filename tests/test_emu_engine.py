"""CPU: the device code of the CUDA engine (csrc/*.cuh), compiled as ONE emulated host thread
(tests/emu), against the oracle and the reference's golden outputs.  This checks the algorithm
(frontier expansion, tables, collector, sampling walk); the GPU parity tests check the real kernels."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, run_emu
from qiskit_jku_provider_b200 import circuits as C

from test_oracle import level_parity

AMP_TOL = 1e-9

# The CUDA engine breaks normalisation ties with a tolerance (jkq_engine.cuh reduce_item, DESIGN.md
# "normalisation ties"); the oracle restates that rule behind --tie-tol.  With it, the two
# implementations -- recursive DFS over matrix DDs vs level-synchronous fused frontier -- must agree on
# every level count after every gate.
TIE = ["--tie-tol"]


@pytest.mark.parametrize("name", golden_names())
def test_emu_matches_reference_golden(emu_bin, oracle_mod, name):
    g = load_golden(name)
    r = run_emu(emu_bin, g["circuit"], levels=True, amps=True)
    near = oracle_mod.run_oracle(g["circuit"], levels=True, extra=TIE)
    ref = np.array([complex(*a) for a in g["amps"]])
    got = np.array([complex(*a) for a in r["amps"]])
    assert np.abs(ref - got).max() < AMP_TOL
    assert abs(r["norm"] - 1.0) < 1e-9
    level_parity(g["level_counts_per_gate"], r["level_counts_per_gate"], name, near)


@pytest.mark.parametrize("name", golden_names())
def test_emu_equals_tie_oracle_on_golden_circuits(emu_bin, oracle_mod, name):
    g = load_golden(name)
    r = run_emu(emu_bin, g["circuit"], levels=True)
    o = oracle_mod.run_oracle(g["circuit"], levels=True, extra=TIE)
    assert r["level_counts_per_gate"] == o["level_counts_per_gate"]


@pytest.mark.parametrize("circ", [C.qft(12), C.ghz(24), C.grover(10, 6, marked=0x155), C.random_circuit(10, 60, 4)],
                         ids=lambda c: c.name)
def test_emu_matches_oracle(emu_bin, oracle_mod, circ):
    o = oracle_mod.run_oracle(circ, shots=2000, seed=9, levels=True, amps=True, extra=["--amp-limit", "1024"] + TIE)
    e = run_emu(emu_bin, circ.text(), shots=2000, seed=9, levels=True, amps=True, extra=["--amp-limit", "1024"])
    oa = np.array([complex(*a) for a in o["amps"]])
    ea = np.array([complex(*a) for a in e["amps"]])
    assert np.abs(oa - ea).max() < AMP_TOL
    assert e["level_counts_per_gate"] == o["level_counts_per_gate"]
    assert e["counts"] == o["counts"]  # same Philox stream, same DD -> identical counts


def test_emu_collector_keeps_state(emu_bin, oracle_mod):
    """Tiny arena: the collector must run many times without changing the result."""
    circ = C.qft(10)
    e = run_emu(emu_bin, circ.text(), amps=True, extra=["--node-cap", "4096", "--gc-min", "64", "--grid"])
    o = oracle_mod.run_oracle(circ, amps=True)
    assert e["gc_runs"] >= 3
    assert np.abs(np.array([complex(*a) for a in o["amps"]]) - np.array([complex(*a) for a in e["amps"]])).max() < AMP_TOL


@pytest.mark.parametrize("circ", [C.random_circuit(10, 90, 11), C.random_circuit(12, 100, 5)], ids=lambda c: c.name)
def test_emu_state_too_large_for_the_single_cta_pools_survives_a_failed_attempt(emu_bin, oracle_mod, circ):
    """Regression: a single-CTA attempt that finds the state too large to load (more reals than the shared value table,
    or more nodes than the shared arena) must leave the arena's counters alone -- mode_leave used to clamp them to the
    pool sizes, which cut off live reals / nodes before the grid path took over (random10_90_s11: norm 9143)."""
    o = oracle_mod.run_oracle(circ, amps=True)
    e = run_emu(emu_bin, circ.text(), amps=True, extra=["--node-cap", str(1 << 20), "--gc-min", str(1 << 14)])
    oa = np.array([complex(*a) for a in o["amps"]])
    ea = np.array([complex(*a) for a in e["amps"]])
    assert e["grid_switches"] >= 1
    assert abs(e["norm"] - 1.0) < 1e-9
    assert np.abs(oa - ea).max() < AMP_TOL


@pytest.mark.parametrize("limit", [96, 160])
@pytest.mark.parametrize("circ", [C.grover(10, 4, marked=0x155), C.random_circuit(12, 70, 3)], ids=lambda c: c.name)
def test_emu_tiny_shared_node_pool_hands_over_without_losing_counts(emu_bin, oracle_mod, circ, limit):
    """A collection in the middle of a single-CTA run that finds the DD too large for a (here artificially tiny) shared
    node pool: the value count lives in shared memory on that path and must reach the arena before the grid path
    takes over."""
    o = oracle_mod.run_oracle(circ, amps=True)
    e = run_emu(emu_bin, circ.text(), amps=True, extra=["--node-cap", str(1 << 18), "--gc-min", str(1 << 15), "--cta-node-limit", str(limit)])
    assert abs(e["norm"] - 1.0) < 1e-9
    assert np.abs(np.array([complex(*a) for a in o["amps"]]) - np.array([complex(*a) for a in e["amps"]])).max() < AMP_TOL


@pytest.mark.parametrize("limit", [0, 512, 300, 200])
def test_emu_saved_cursor_hand_over(emu_bin, oracle_mod, limit):
    """The single-CTA path abandons a gate and (testing aid --cta-no-reserve 2) does not rebuild up to the root: the
    cursor itself is handed to the grid path, which re-hashes the arena as it is and continues from the cursor."""
    circ = C.random_circuit(6, 90, 35)
    o = oracle_mod.run_oracle(circ, amps=True)
    e = run_emu(emu_bin, circ.text(), amps=True, extra=["--node-cap", str(1 << 18), "--gc-min", str(1 << 14), "--cta-val-limit", str(limit), "--cta-no-reserve", "2"])
    assert e["grid_switches"] == 1
    assert abs(e["norm"] - 1.0) < 1e-9
    assert np.abs(np.array([complex(*a) for a in o["amps"]]) - np.array([complex(*a) for a in e["amps"]])).max() < AMP_TOL


@pytest.mark.parametrize("extra", [["--cta-val-limit", "60", "--cta-no-reserve", "2"], ["--cta-node-limit", "96"], ["--cta-val-limit", "48"]],
                         ids=["saved_cursor", "tiny_node_pool", "tiny_value_table"])
def test_emu_level_log_survives_the_hand_over(emu_bin, oracle_mod, extra):
    """Per-gate level counts when the single-CTA kernel hands over in the middle of the circuit: the grid kernel counts
    the levels of the last applied gate again on entry, every row equals the oracle's."""
    circ = C.random_circuit(8, 60, 9)
    o = oracle_mod.run_oracle(circ, levels=True, amps=True, extra=TIE)
    e = run_emu(emu_bin, circ.text(), levels=True, amps=True, extra=["--node-cap", str(1 << 18), "--gc-min", str(1 << 14)] + extra)
    assert e["grid_switches"] == 1
    assert e["level_counts_per_gate"] == o["level_counts_per_gate"]
    assert np.abs(np.array([complex(*a) for a in o["amps"]]) - np.array([complex(*a) for a in e["amps"]])).max() < AMP_TOL
