# cycle counters of the grid path (JKQ_PROFILE build in a scratch library): usage: bash scripts/grid_profile.sh
# (build it where nvcc is: JKQ_LIB=ab/libjkq_prof.so JKQ_DEFINES=JKQ_PROFILE python -c "from qiskit_jku_provider_b200 import _build; _build.build_native(force=True)")
set -e
export JKQ_LIB=${JKQ_LIB:-$PWD/ab/libjkq_prof.so}
[ -f "$JKQ_LIB" ] || JKQ_DEFINES="JKQ_PROFILE" python -c "from qiskit_jku_provider_b200 import _build; _build.build_native(force=True)" > /dev/null 2>&1 || true
python - <<'PY'
import sys
sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine
import os
big = os.environ.get("BIG")
cases = (("r22_200", C.random_circuit(22, 200, 1)),) if big else (("r120", C.random_circuit(18, 120, 1)), ("r160", C.random_circuit(18, 160, 1)))
for name, circ in cases:
    eng = Engine(device=0, node_capacity=1 << (25 if big else 20), value_capacity=1 << (26 if big else 20))
    eng.load(circ); eng.apply(); st = eng.stats()
    print(name, "apply_ms", st["apply_ms"], "items", st["items"], "gc", st["gc_runs"], "retry", st["retry_rounds"], "thin", st["thin_gates"], flush=True)
    eng.close()
PY
