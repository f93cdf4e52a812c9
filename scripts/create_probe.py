"""Where the start-up time of a fresh process goes: CUDA context, jkq_create (allocations + table clears), first run."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
t0 = time.perf_counter()
import ctypes as C
from qiskit_jku_provider_b200 import _native
lib = _native.load()
t1 = time.perf_counter()
rt = C.CDLL("libcudart.so.12")
rt.cudaFree(0)
t2 = time.perf_counter()
from qiskit_jku_provider_b200.engine import Engine
from qiskit_jku_provider_b200 import circuits as Cc
e = Engine(device=0)
t3 = time.perf_counter()
e2 = Engine(device=0)
t4 = time.perf_counter()
e.load(Cc.qft(32)); e.run(1024, 1)
t5 = time.perf_counter()
e.load(Cc.qft(32)); e.run(1024, 1)
t6 = time.perf_counter()
small = Engine(device=0, node_capacity=1 << 16, value_capacity=1 << 14, item_capacity=1 << 16)
t7 = time.perf_counter()
print("import+dlopen %.3f  cuda context %.3f  jkq_create #1 %.3f  #2 %.3f  first run %.3f  second run %.3f  small create %.3f" % (
    t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6))
