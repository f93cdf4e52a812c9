"""A/B of the batch fall-back: circuits that outgrew one SM redone on 1 / 2 / 4 / 8 full engines side by side."""
import sys, time
sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.batch import BatchRunner
from qiskit_jku_provider_b200.engine import circuit_ops

batch = [C.random_circuit(18, 40, s) for s in range(1024)]
for c in batch:
    circuit_ops(c)
ref = None
for k in (1, 2, 4, 1, 4):
    with BatchRunner(workers=148, fallback_engines=k) as br:
        br.run(batch, shots=1024, seed=1)
        t0 = time.perf_counter()
        out, gates, ms = br.run(batch, shots=1024, seed=1)
        wall = time.perf_counter() - t0
        print("fallback_engines %d: k_batch %.2f ms, fall-back %.2f ms (%d circuits), total %.2f ms -> %.2f M gate-applies/s; wall %.1f ms" %
              (k, br.last_kernel_ms, br.last_fallback_ms, br.fallbacks // 2, ms, gates / ms / 1e3, wall * 1e3), flush=True)
        print("    redone:", br.last_fallback_detail, flush=True)
        if ref is None: ref = out
        assert out == ref, "results differ with %d fall-back engines" % k
