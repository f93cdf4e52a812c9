"""One jkq_apply of a named circuit (for ncu).  usage: one_apply.py [qft32|ghz24|grover|r100|r120|r140|r160] [cta_threads] [repeat] [force_mode]"""
import sys

sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine

name = sys.argv[1] if len(sys.argv) > 1 else "qft32"
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 64
repeat = int(sys.argv[3]) if len(sys.argv) > 3 else 1
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
circ = {"qft32": lambda: C.qft(32), "ghz24": lambda: C.ghz(24), "grover": lambda: C.grover(20, 20),
        "r100": lambda: C.random_circuit(18, 100, 1), "r120": lambda: C.random_circuit(18, 120, 1),
        "r140": lambda: C.random_circuit(18, 140, 1), "r160": lambda: C.random_circuit(18, 160, 1)}[name]()
eng = Engine(device=0, node_capacity=1 << 23, value_capacity=1 << 22, cta_threads=threads)
eng.set_option("force_mode", mode)
eng.load(circ)
for _ in range(repeat):
    eng.apply()
st = eng.stats()
print("apply_ms", st["apply_ms"], "us_per_gate", 1e3 * st["apply_ms"] / len(circ), "items", st["items"], "nodes", st["nodes_created"],
      "grid_launches", st["grid_launches"], "cta_launches", st["cta_launches"], "gc", st["gc_runs"])
