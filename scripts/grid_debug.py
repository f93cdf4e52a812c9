import sys, time
sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine
eng = Engine(device=0, node_capacity=1 << 21)
eng.set_option("force_mode", 2)
for rep in range(6):
    for lv in (1, 0):
        eng.set_option("level_log", lv)
        for c in (C.qft(10), C.grover(8, 12, marked=0xA5), C.qft(12)):
            eng.load(c)
            t = time.time()
            try:
                eng.run(2048, 5)
                print(rep, lv, c.name, "ok %.2fs" % (time.time() - t), flush=True)
            except Exception as e:
                print(rep, lv, c.name, "FAIL %.2fs" % (time.time() - t), str(e)[:600], flush=True)
