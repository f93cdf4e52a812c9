"""Generates tests/golden/cli/*.json: outputs of the UNMODIFIED reference executable (oracle/_ref/jku_simulator, built by
oracle/build_ref.sh from /root/reference/src) on the QASM programs next to this script.  Run in the build container
(the reference does not travel to the GPU box; the fixtures do):

    python tests/golden/make_cli_golden.py

The reference needs qelib1.inc in its working directory; the text is the string its Python backend writes there
(qiskit_jku_provider/qasm_simulator_jku.py:32-60), read from the reference at generation time.
"""
import ast
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "jku_simulator")
REF_PY = "/root/reference/qiskit_jku_provider/qasm_simulator_jku.py"

# program -> (shots, seed, extra flags)
CASES = {
    "bell": (2000, 11, ["--display_statevector", "--display_probabilities"]),
    "broadcast": (4000, 5, ["--display_statevector", "--display_probabilities"]),
    "qelib_all": (4000, 3, ["--display_statevector", "--display_probabilities"]),
    "teleport": (400, 7, []),
    "reset_mid": (2000, 9, ["--display_statevector", "--display_probabilities"]),
    "partial_probs": (100, 1, ["--display_statevector", "--display_probabilities"]),
    "ghz8": (3000, 2, ["--display_statevector"]),
    "deterministic": (64, 4, []),
}


def qelib1_text():
    tree = ast.parse(open(REF_PY).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "qelib1":
            return ast.literal_eval(node.value)
    raise RuntimeError("qelib1 string not found in the reference")


def main():
    if not os.path.exists(REF_EXE):
        sys.exit("build the reference first: bash oracle/build_ref.sh")
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "oracle", "_ref")) as tmp:
        with open(os.path.join(tmp, "qelib1.inc"), "w") as f:
            f.write(qelib1_text())
        for name, (shots, seed, flags) in sorted(CASES.items()):
            if only and name not in only:
                continue
            qasm = open(os.path.join(HERE, "cli", name + ".qasm")).read()
            cmd = [REF_EXE, "--simulate_qasm", "--shots=%d" % shots, "--seed=%d" % seed, "--ps"] + flags
            out = subprocess.run(cmd, input=qasm, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, universal_newlines=True, timeout=300)
            if out.returncode != 0 or len(out.stdout) > (8 << 20):
                sys.exit("%s: reference failed (%d): %s" % (name, out.returncode, out.stderr[:400]))
            rec = {"name": name, "shots": shots, "seed": seed, "flags": flags, "stdout": out.stdout, "stderr": out.stderr}
            with open(os.path.join(HERE, "cli", name + ".json"), "w") as f:
                json.dump(rec, f, indent=1)
            print(name, "ok", len(out.stdout), "bytes", "stderr:", repr(out.stderr[:120]))


if __name__ == "__main__":
    main()
