import sys
sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.batch import BatchRunner
circs = [C.random_circuit(8, 20, s) for s in range(3)]
with BatchRunner(workers=2) as br:
    out, g, ms = br.run(circs, shots=64, seed=1)
    print(out, g, ms)
