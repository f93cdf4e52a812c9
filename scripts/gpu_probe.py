"""Ad-hoc GPU probe: engine vs oracle on the BASELINE configs, with timings."""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine
import oracle


def check(circ, eng, shots=4096, amp_limit=1 << 12, levels=True, mode=0):
    eng.set_option("force_mode", mode)
    eng.set_option("level_log", 1 if levels else 0)
    eng.load(circ)
    t = time.time()
    eng.run(shots=shots, seed=3)
    wall = time.time() - t
    st = eng.stats()
    k = min(amp_limit, 1 << circ.n)
    amps = eng.amplitudes(0, k)
    ref = oracle.run_oracle(circ, shots=shots, seed=3, levels=levels, amps=True, extra=["--amp-limit", str(k)])
    oamps = np.array([complex(*a) for a in ref["amps"]])
    err = float(np.abs(amps - oamps).max())
    res = {"name": circ.name, "mode": mode, "applies": len(circ), "amp_err": err,
           "final_levels_equal": eng.level_counts() == ref["level_counts"],
           "counts_equal": eng.counts_bitstrings() == ref["counts"],
           "apply_ms": st["apply_ms"], "us_per_gate": 1e3 * st["apply_ms"] / max(1, len(circ)),
           "sample_ms": st["sample_ms"], "items": st["items"], "nodes_created": st["nodes_created"],
           "retry": st["retry_rounds"], "gc": st["gc_runs"], "cta": st["cta_launches"], "grid": st["grid_launches"],
           "oracle_apply_ms": 1e3 * ref["t_apply_s"], "wall_s": wall, "norm": st["norm"]}
    if levels:
        lg = eng.level_log(len(circ))
        res["level_diff_gates"] = sum(1 for a, b in zip(lg, ref["level_counts_per_gate"]) if a != b)
    print(json.dumps(res), flush=True)
    return res


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    eng = Engine(device=0, node_capacity=1 << 22)
    if which in ("all", "small"):
        for mode in (1, 2):
            check(C.ghz(5, star=True), eng, mode=mode)
            check(C.qft(8), eng, mode=mode)
            check(C.grover(8, 12, marked=0xA5), eng, mode=mode)
            check(C.random_circuit(6, 40, 2), eng, mode=mode)
    if which in ("all", "big"):
        check(C.ghz(24), eng, mode=1)
        check(C.qft(32), eng, mode=1, levels=False)
        check(C.qft(32), eng, mode=1, levels=True)
        check(C.grover(20, 50), eng, mode=1, levels=False)
        check(C.random_circuit(18, 40, 1), eng, mode=0)
        check(C.random_circuit(18, 80, 1), eng, mode=0)
        check(C.random_circuit(18, 100, 1), eng, mode=0, levels=False)
        check(C.random_circuit(18, 100, 1), eng, mode=2, levels=False)
        check(C.random_circuit(18, 120, 1), eng, mode=0, levels=False)
    if which in ("all", "bw"):
        print(json.dumps(eng.random_sector_bench()), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
