"""Generates tests/golden/*.json by running the UNMODIFIED reference engine (oracle/_ref/jku_ref,
built from /root/reference/src by oracle/build_ref.sh) on small circuits.  Run in the build container
(where /root/reference exists); the JSON files are committed and are what travels to the GPU box.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from qiskit_jku_provider_b200 import circuits as C  # noqa: E402
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    out = []
    bell = C.Circuit(2, name="bell").add("h", [0]).add("cx", [0, 1])
    out.append(bell)
    out.append(C.Circuit(1, name="x1").add("x", [0]))
    out.append(C.ghz(5, star=True))
    out.append(C.ghz(8))
    out.append(C.qft(6))
    out.append(C.qft(10))
    out.append(C.grover(6, 4, marked=0x2B))
    out.append(C.grover(8, 12, marked=0xA5))
    multi = C.Circuit(4, name="multireg_x").add("x", [1]).add("x", [2])
    out.append(multi)
    ex = C.Circuit(6, name="example_qasm")
    for q in range(3):
        ex.add("h", [q])
    for q in range(3):
        ex.add("cx", [q, q + 3])
    out.append(ex)
    # every qelib1 gate once, controls above and below the target
    allg = C.Circuit(5, name="all_gates")
    for q in range(5):
        allg.add("h", [q])
    allg.add("u3", [0], [0.3, 0.7, 1.1]).add("u2", [1], [0.4, 0.9]).add("u1", [2], [0.25]).add("y", [3]).add("z", [4])
    allg.add("s", [0]).add("sdg", [1]).add("t", [2]).add("tdg", [3]).add("rx", [4], [0.8]).add("ry", [0], [0.6])
    allg.add("rz", [1], [0.5]).add("cz", [0, 4]).add("cy", [3, 1]).add("swap", [2, 4]).add("ch", [4, 0])
    allg.add("ccx", [0, 2, 4]).add("cswap", [3, 1, 2]).add("crz", [4, 1], [0.9]).add("cu1", [1, 3], [0.35])
    allg.add("cu3", [2, 0], [0.2, 0.4, 0.6]).add("rzz", [0, 3], [0.45]).add("iden", [2])
    out.append(allg)
    # native multi-controlled gates incl. negative controls on both sides of the target
    mc = C.Circuit(5, name="multi_control")
    for q in range(5):
        mc.add("ry", [q], [0.3 + 0.2 * q])
    mc.ops.append(C.MCG("z", 4, [0, 1, 2, 3]))
    mc.ops.append(C.MCG("x", 2, [(0, 1), (4, 0)]))
    mc.ops.append(C.MCU(1, 0.7, 0.2, 0.5, [(3, 1), (0, 0), (4, 1)]))
    mc.ops.append(C.MCG("h", 0, [(2, 0)]))
    mc.ops.append(C.MCG("s", 3, [(1, 1), (2, 1)]))
    out.append(mc)
    for seed in range(1, 9):
        out.append(C.random_circuit(5, 30, seed))
    for seed in (1, 2):
        out.append(C.random_circuit(8, 40, seed))
    return out


def main():
    oracle.build()
    if not oracle.have_reference():
        raise SystemExit("oracle/_ref/jku_ref missing: run oracle/build_ref.sh where /root/reference exists")
    index = []
    for c in cases():
        r = oracle.run_reference(c, shots=200, seed=1, levels=True, amps=True)
        doc = {"name": c.name, "n": c.n, "circuit": c.text(), "source": "oracle/_ref/jku_ref (unmodified reference, MPFR 200-bit)",
               "applies": r["applies"], "amps": r["amps"], "level_counts_per_gate": r["level_counts_per_gate"],
               "level_counts": r["level_counts"], "counts_rand_seed1_200": r["counts"], "nodes_created": r["nodes_created"],
               "peak_active": r["peak_active"]}
        with open(os.path.join(HERE, c.name + ".json"), "w") as f:
            json.dump(doc, f, separators=(",", ":"))
        index.append(c.name)
        print(c.name, r["applies"], r["peak_active"])
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f)


if __name__ == "__main__":
    main()
