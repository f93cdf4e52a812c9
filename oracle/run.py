"""ORACLE runner -- test infrastructure, not product code (see oracle/__init__.py)."""

import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "jku_ref")
REF_QASM_BIN = os.path.join(HERE, "_ref", "jku_simulator")
ORACLE_BIN = os.path.join(HERE, "_build", "qmdd_oracle")


def build(force=False):
    """Compile the C restatement (always) and, when /root/reference is present, the reference."""
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    src = os.path.join(HERE, "qmdd_oracle.c")
    if force or not os.path.exists(ORACLE_BIN) or os.path.getmtime(src) > os.path.getmtime(ORACLE_BIN):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-o", ORACLE_BIN, src, "-lm"])
    if os.path.isdir(os.environ.get("JKQ_REFERENCE_SRC", "/root/reference/src")):
        if force or not os.path.exists(REF_BIN):
            subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])
    return ORACLE_BIN


def have_reference():
    return os.path.exists(REF_BIN)


def _run(binary, circuit, shots, seed, levels, amps, extra, timeout):
    text = circuit if isinstance(circuit, str) else circuit.text()
    with tempfile.NamedTemporaryFile("w", suffix=".jkq", delete=False) as f:
        f.write(text)
        path = f.name
    try:
        cmd = [binary, "--circuit", path, "--seed", str(seed)]
        if shots:
            cmd += ["--shots", str(shots)]
        if levels:
            cmd.append("--levels")
        if amps:
            cmd.append("--amps")
        cmd += list(extra)
        out = subprocess.check_output(cmd, timeout=timeout)
    finally:
        os.unlink(path)
    return json.loads(out)


def run_oracle(circuit, shots=0, seed=1, levels=False, amps=False, extra=(), timeout=600):
    """FP64 C restatement.  counts use the product's Philox stream unless extra=('--rng','rand')."""
    build()
    return _run(ORACLE_BIN, circuit, shots, seed, levels, amps, extra, timeout)


def run_reference(circuit, shots=0, seed=1, levels=False, amps=False, extra=(), timeout=600):
    """The unmodified reference engine (MPFR).  Raises FileNotFoundError if it was never built."""
    if not have_reference():
        raise FileNotFoundError(REF_BIN)
    return _run(REF_BIN, circuit, shots, seed, levels, amps, extra, timeout)
