import json
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    with open(os.path.join(GOLDEN, "index.json")) as f:
        return json.load(f)


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def emu_bin():
    """Single-thread host emulation of the device headers (tests/emu) -- test infrastructure only."""
    src = os.path.join(ROOT, "tests", "emu", "jkq_emu.cpp")
    out = os.path.join(ROOT, "tests", "emu", "jkq_emu")
    deps = [src] + [os.path.join(ROOT, "qiskit_jku_provider_b200", "csrc", f)
                    for f in ("jkq_types.h", "jkq_engine.cuh", "jkq_thin.cuh", "jkq_grid.cuh", "jkq_apply.cuh", "jkq_sample.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-w", "-o", out, src])
    return out


def run_emu(emu_bin, circuit_text, shots=0, seed=1, levels=False, amps=False, extra=()):
    with tempfile.NamedTemporaryFile("w", suffix=".jkq", delete=False) as f:
        f.write(circuit_text)
        path = f.name
    try:
        cmd = [emu_bin, "--circuit", path, "--seed", str(seed), "--node-cap", str(1 << 16), "--gc-min", str(1 << 12)]
        if shots:
            cmd += ["--shots", str(shots)]
        if levels:
            cmd.append("--levels")
        if amps:
            cmd.append("--amps")
        cmd += list(extra)
        return json.loads(subprocess.check_output(cmd, timeout=300))
    finally:
        os.unlink(path)
