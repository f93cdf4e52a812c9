"""Which circuits of the BASELINE batch outgrow one SM, where, and by which pool.  usage: batch_fallback_why.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.batch import BatchRunner
from qiskit_jku_provider_b200.engine import Engine

circs = [C.random_circuit(18, 40, s) for s in range(1024)]
with BatchRunner(workers=148) as br:
    br.fallback_engines = 1
    out, gates, ms = br.run(circs, shots=16, seed=1)
    print("fallbacks", br.fallbacks, "fast path said:", br.last_fast_path_error)
# the full engine on every circuit: who needs the grid path, at which gate, with how much state
eng = Engine(device=0, node_capacity=1 << 18)
heavy = []
grid_before = 0
for i, c in enumerate(circs):
    eng.load(c); eng.apply()
    st = eng.stats()
    took_grid, grid_before = st["grid_launches"] > grid_before, st["grid_launches"]  # (the launch counters run on over the handle's life)
    if took_grid:
        heavy.append(i)
        print("seed", i, "gates", len(c), "apply_ms %.2f" % st["apply_ms"], "items", st["items"], "max_frontier", st["max_frontier"], "gc", st["gc_runs"],
              "nodes_created", st["nodes_created"], "vals", st["value_inserts"], "live", st["live_nodes"], "live_vals", st.get("live_values"), flush=True)
eng.set_option("debug_stop_at_handover", 1)
for i in heavy:
    eng.load(circs[i])
    try:
        eng.apply()
    except Exception as e:
        print("seed", i, "stopped:", e)
eng.close()
