"""Batches of independent circuits (BASELINE.json configs[4]).

A canonical DD does not shard, but independent circuits do: the single-CTA path of the engine keeps one
circuit's whole working set in ONE SM's shared memory, so a B200 runs up to 148 circuits side by side.
``BatchRunner`` owns a pool of small engines (one ``jkq_handle`` = one CUDA stream each) and a thread per
engine; the ctypes calls release the GIL, so the persistent circuit kernels of different engines overlap on
different SMs.  Across GPUs the batch is split by circuit index (``i % world == rank``), one process per
GPU, no data-path collective; per-circuit counts are gathered on the host.

The reference runs one circuit per ``jku_simulator`` process, sequentially
(``qiskit_jku_provider/qasm_simulator_jku.py:340-359``).
"""

import threading
from concurrent.futures import ThreadPoolExecutor

from .engine import Engine, circuit_ops
from .jkusimulatorerror import JKUSimulatorError

# small pools: a circuit that outgrows them is redone on the shared big engine
SMALL = dict(node_capacity=1 << 15, value_capacity=1 << 13, item_capacity=1 << 15, gc_min=1 << 12)


class BatchRunner:
    def __init__(self, workers=64, device=0, small=None, big=None):
        self.device = int(device)
        self.workers = int(workers)
        self.small = dict(SMALL, **(small or {}))
        self.big_opts = dict(big or {})
        self._local = threading.local()
        self._engines = []
        self._lock = threading.Lock()
        self._big = None
        self._big_lock = threading.Lock()
        self._pool = ThreadPoolExecutor(max_workers=self.workers)
        self.fallbacks = 0

    def _engine(self):
        e = getattr(self._local, "engine", None)
        if e is None:
            e = Engine(device=self.device, **self.small)
            self._local.engine = e
            with self._lock:
                self._engines.append(e)
        return e

    def _run_one(self, job):
        idx, circ, shots, seed = job
        ops = circuit_ops(circ)
        eng = self._engine()
        try:
            eng.load_ops(circ.n, ops)
            eng.run(shots, seed)
            keys, vals = eng.counts_arrays()
            st = eng.stats()
        except JKUSimulatorError:
            with self._big_lock:  # the DD outgrew the small pools: one big engine, serialised
                if self._big is None:
                    self._big = Engine(device=self.device, **self.big_opts)
                self.fallbacks += 1
                self._big.load_ops(circ.n, ops)
                self._big.run(shots, seed)
                keys, vals = self._big.counts_arrays()
                st = self._big.stats()
        return idx, dict(zip(keys.tolist(), vals.tolist())), st["gates"], st["apply_ms"]

    def run(self, circuits, shots=1024, seed=1, rank=0, world=1):
        """Runs ``circuits[i]`` for ``i % world == rank``; returns {i: {outcome: count}}, total gate applies,
        sum of the per-circuit device times in ms.  Circuit i samples with seed ``seed + i``."""
        jobs = [(i, c, shots, seed + i) for i, c in enumerate(circuits) if i % world == rank]
        out, gates, dev_ms = {}, 0, 0.0
        for idx, counts, g, ms in self._pool.map(self._run_one, jobs):
            out[idx] = counts
            gates += g
            dev_ms += ms
        return out, gates, dev_ms

    def close(self):
        self._pool.shutdown(wait=True)
        for e in self._engines:
            e.close()
        self._engines = []
        if self._big is not None:
            self._big.close()
            self._big = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
