"""Is the batch kernel bound by its slowest circuits?  Kernel time for 148 circuits (one per SM) at different seeds,
for 1024, and for 1024 with the heaviest circuits removed."""
import json, sys, time
sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.batch import BatchRunner
from qiskit_jku_provider_b200.engine import circuit_ops, Engine

circs = [C.random_circuit(18, 40, s) for s in range(1024)]
for c in circs:
    circuit_ops(c)
with BatchRunner(workers=148) as br:
    br.run(circs, shots=1024, seed=1)
    out, gates, ms = br.run(circs, shots=1024, seed=1)
    print("1024 circuits: kernel %.2f ms, %d gate applies, fallbacks %d" % (ms, gates, br.fallbacks))
    for lo in (0, 148, 296):
        sub = circs[lo:lo + 148]
        br.run(sub, shots=1024, seed=1)
        out, gates, ms = br.run(sub, shots=1024, seed=1)
        print("148 circuits [%d..): kernel %.2f ms (= slowest circuit), %d gate applies" % (lo, ms, gates))
    one = [circs[0]]
    br.run(one, shots=1024, seed=1)
    _, g1, ms1 = br.run(one, shots=1024, seed=1)
    print("1 circuit (seed 0, %d gates): kernel %.3f ms" % (g1, ms1))
    for shots in (1, 1024):
        _, g1, ms1 = br.run(one, shots=shots, seed=1)
        print("   shots=%d: %.3f ms" % (shots, ms1))
# per-circuit apply time on the full engine, to find the heavy ones
eng = Engine(device=0, node_capacity=1 << 18)
times = []
for i, c in enumerate(circs[:296]):
    eng.load(c); eng.apply()
    st = eng.stats()
    times.append((st["apply_ms"], i, len(c), st["items"], st["grid_launches"]))
times.sort(reverse=True)
print("heaviest of the first 296:", [(round(t, 2), i, n, it, g) for t, i, n, it, g in times[:8]])
print("median apply_ms %.3f  mean %.3f" % (sorted(t[0] for t in times)[len(times) // 2], sum(t[0] for t in times) / len(times)))
