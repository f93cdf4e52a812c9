"""GPU: parity with the UNMODIFIED reference engine itself (oracle/_ref/jku_ref, built from /root/reference/src by
oracle/build_ref.sh; the binary travels to the GPU box, nothing here reads /root/reference) at BASELINE.json's
full sizes, through the C ABI.

Two kinds of check:
  * per gate: live nodes per level after EVERY gate (the reference's Active[v], QMDDpackage.h:199-201) with the level
    log, which collects after every gate;
  * prefix parity: the circuit cut after K gates and run the way bench.py runs it (no level log: the cursor is kept
    across gates, two or three collections per circuit), final counts against row K-1 of the reference -- the
    benchmarked control flow, checked at many K.
A difference is reported with the first differing gate and the first normalisation near tie (tests/test_oracle.py
level_parity); amplitudes within the reference's tolerance."""
import numpy as np
import pytest

from test_oracle import first_cause, level_parity
from qiskit_jku_provider_b200 import circuits as C

pytestmark = pytest.mark.gpu

AMP_TOL = 1e-9   # Ctol = 1e-10 per snapped weight (QMDDcomplexD.cpp:336) x <= n + 1 weights
TIE = ["--tie-tol"]


@pytest.fixture(scope="module")
def eng():
    from qiskit_jku_provider_b200.engine import Engine
    e = Engine(device=0, node_capacity=1 << 22)
    yield e
    e.close()


@pytest.fixture(scope="module")
def ref(oracle_mod):
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref/jku_ref was not built (needs /root/reference at build time)")
    return oracle_mod


def apply(eng, circ, levels, mode=0):
    eng.set_option("force_mode", mode)
    eng.set_option("level_log", 1 if levels else 0)
    eng.load(circ)
    eng.apply()


def per_gate(eng, ref, circ, amps=4096):
    r = ref.run_reference(circ, levels=True, amps=True, extra=["--amp-limit", str(min(amps, 1 << circ.n))])
    causes = ref.run_oracle(circ, levels=True, extra=TIE)
    apply(eng, circ, True)
    level_parity(r["level_counts_per_gate"], eng.level_log(len(circ)), circ.name, causes)
    k = min(amps, 1 << circ.n)
    ra = np.array([complex(*a) for a in r["amps"]])
    assert np.abs(eng.amplitudes(0, k) - ra).max() < AMP_TOL, circ.name
    return r


def test_qft32_per_gate_against_the_reference(eng, ref):
    r = per_gate(eng, ref, C.qft(32))
    assert r["level_counts"] == [1] * 32


def test_ghz24_per_gate_against_the_reference(eng, ref):
    per_gate(eng, ref, C.ghz(24))
    per_gate(eng, ref, C.ghz(24, star=True))


@pytest.mark.parametrize("seed", range(1, 21))
def test_random18_40_per_gate_against_the_reference(eng, ref, seed):
    per_gate(eng, ref, C.random_circuit(18, 40, seed))


def test_grover20_60_iterations_per_gate_against_the_reference(eng, ref):
    """Amplitudes of 2^-10 put second-order terms right at the tolerance: from gate 102 on the reference identifies
    genuinely different values (lossy snaps), and which ones depends on its own order of intermediate results.  Identical
    counts up to there; afterwards the transient counts inside an iteration may differ by a node per level, while every
    gate at which the reference is back to its steady 2 nodes per level must find the engine there as well."""
    circ = C.grover(20, 60)
    r = ref.run_reference(circ, levels=True)
    causes = ref.run_oracle(circ, levels=True, extra=TIE)
    apply(eng, circ, True)
    lg = eng.level_log(len(circ))
    d = level_parity(r["level_counts_per_gate"], lg, circ.name, causes)
    assert d is None or d >= 102
    steady = [2] * 19 + [1]
    n_steady = 0
    for g, row in enumerate(r["level_counts_per_gate"]):
        if g >= 20 and row == steady and (g - 20) % 102 == 101:  # the last gate of an iteration
            assert lg[g] == steady, "gate %d (end of iteration %d)" % (g, (g - 20) // 102)
            n_steady += 1
    assert n_steady == 60
    p = np.sin((2 * 60 + 1) * np.arcsin(2.0 ** -10)) ** 2
    assert abs(abs(eng.amplitudes(0x5A5A5, 1)[0]) ** 2 - p) < 1e-6
    differing = sum(1 for a, b in zip(r["level_counts_per_gate"], lg) if a != b)
    print("grover20_60: %d of %d gates with a different transient count" % (differing, len(lg)))


@pytest.mark.parametrize("circ,ks", [
    (C.qft(32), [1, 2, 3, 7, 40, 41, 42, 43, 44, 100, 333, 777, 1000, 1264, 1500, 1999, 2000, 2001, 2400, 2526, 2527, 2528]),
    (C.ghz(24), [1, 2, 5, 12, 23, 24]),
    (C.grover(20, 12), [20, 21, 30, 31, 41, 42, 61, 62, 82, 102, 122, 123, 500, 777, 1000, 1244]),
    (C.random_circuit(18, 40, 1), [5, 20, 50, 80, 100, 130, 161]),
    (C.random_circuit(18, 40, 7), [10, 33, 77, 120]),
], ids=lambda x: x.name if hasattr(x, "name") else "")
def test_prefix_parity_in_the_benchmarked_control_flow(eng, ref, circ, ks):
    """No level log: the kernel keeps its cursor across gates and collects only when the arena asks for it -- what
    bench.py times.  The live counts after K gates must still be the reference's."""
    rows = ref.run_reference(circ, levels=True)["level_counts_per_gate"]
    t = first_cause(ref.run_oracle(circ, levels=True, extra=TIE))
    for k in ks:
        sub = C.Circuit(circ.n, circ.ops[:k], "%s[:%d]" % (circ.name, k))
        apply(eng, sub, False)
        got = eng.level_counts()
        if got != rows[k - 1]:
            assert t is not None and t < k, "%s: after %d gates %s, reference %s, and no near tie / lossy snap before" % (circ.name, k, got, rows[k - 1])


@pytest.mark.parametrize("ops,amps", [(120, 1 << 16), (160, 1 << 16)])
def test_random18_blowup_against_the_reference_or_the_oracle(eng, oracle_mod, ops, amps):
    """BASELINE configs[3] stress ladder (seed 1): 120 ops -> 19.7 k live nodes, 160 ops -> 107 k.
    120 ops against the unmodified reference (3 s of MPFR); 160 ops would take it 65 s, so that one runs against the FP64
    restatement.  Amplitudes within tolerance, squared norm 1, live nodes in total within 1 % (on blown-up diagrams
    equality holds only while no two distinct values fall within the tolerance, SURVEY section 7 hard part 2)."""
    circ = C.random_circuit(18, ops, 1)
    use_ref = ops <= 120 and oracle_mod.have_reference()
    fn = oracle_mod.run_reference if use_ref else oracle_mod.run_oracle
    r = fn(circ, amps=True, extra=["--amp-limit", str(amps)] + ([] if use_ref else TIE), timeout=900)
    apply(eng, circ, False)
    st = eng.stats()
    assert st["grid_launches"] >= 1
    ra = np.array([complex(*a) for a in r["amps"]])
    assert np.abs(eng.amplitudes(0, amps) - ra).max() < AMP_TOL
    eng.sample(1000, seed=1)
    assert abs(eng.stats()["norm"] - 1.0) < 1e-9
    assert abs(sum(eng.level_counts()) - sum(r["level_counts"])) <= 0.01 * sum(r["level_counts"])


def test_batch_of_1024_equals_sequential_equals_oracle(oracle_mod):
    """BASELINE configs[4]: the 1 024-circuit batch (generator of config 4a, seeds 0..1023) through k_batch against the
    one-at-a-time engine (every circuit) and against the oracle's Philox counts (every 8th circuit)."""
    from qiskit_jku_provider_b200.batch import BatchRunner
    from qiskit_jku_provider_b200.engine import Engine
    circs = [C.random_circuit(18, 40, s) for s in range(1024)]
    with BatchRunner(workers=148) as br:
        got, gates, _ = br.run(circs, shots=1024, seed=1)
    assert gates == sum(len(c) for c in circs)
    with Engine(device=0, node_capacity=1 << 20) as e1:
        for i, c in enumerate(circs):
            e1.load(c)
            e1.run(1024, 1 + i)
            assert e1.counts() == got[i], c.name
            if i % 8 == 0:
                o = oracle_mod.run_oracle(c, shots=1024, seed=1 + i, extra=TIE)
                assert e1.counts_bitstrings() == o["counts"], c.name
