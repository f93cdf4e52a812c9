"""Builds libjkq_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc."""

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("JKQ_LIB") or os.path.join(HERE, "libjkq_b200.so")  # (JKQ_LIB: A/B runs of two builds on one box)
SOURCES = ["jkq_b200.cu", "jkq_qasm.cpp"]
CLI = os.path.join(HERE, "jku_simulator")  # the executable interface of the reference (src/main.cpp) on this engine
CLI_SOURCES = ["jku_simulator.cpp"]
HEADERS = ["jkq_types.h", "jkq_engine.cuh", "jkq_thin.cuh", "jkq_grid.cuh", "jkq_apply.cuh", "jkq_sample.cuh", os.path.join("..", "..", "include", "jkq.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine is CUDA only and cannot be built without the CUDA toolkit")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def cli_stale():
    if not os.path.exists(CLI):
        return True
    t = os.path.getmtime(CLI)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in CLI_SOURCES + [HEADERS[-1]]) or os.path.getmtime(LIB) > t


def build_cli(force=False):
    """g++ csrc/jku_simulator.cpp -> qiskit_jku_provider_b200/jku_simulator, linked against libjkq_b200.so next to it."""
    if not force and not cli_stale():
        return CLI
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-o", CLI] + [os.path.join(CSRC, s) for s in CLI_SOURCES] +
                          ["-L" + HERE, "-l:libjkq_b200.so", "-Wl,-rpath,$ORIGIN"])
    return CLI


def build_native(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> qiskit_jku_provider_b200/libjkq_b200.so"""
    if not force and not stale():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    for d in os.environ.get("JKQ_DEFINES", "").split():  # e.g. JKQ_DEFINES="JKQ_PROFILE" for the cycle counters
        cmd.insert(1, "-D" + d)
    subprocess.check_call(cmd)
    build_cli(force=True)
    return LIB
