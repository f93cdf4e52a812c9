"""Batches of independent circuits (BASELINE.json configs[4]).

A canonical DD does not shard, but independent circuits do: the single-CTA path of the engine keeps one
circuit's whole working set in ONE SM's shared memory, so a B200 runs up to 148 circuits side by side.
``BatchRunner`` wraps ``jkq_batch_*`` (include/jkq.h): a pool of small engines, one CUDA stream each; every
circuit is one asynchronous kernel chain (init, gates, sampling records, shots, histogram compaction, results
into pinned memory) and the host only synchronises with an engine when it needs it for the next circuit.
A circuit that outgrows one SM's shared memory is redone on a full engine (grid path).
Across GPUs the batch is split by circuit index (``i % world == rank``), one process per GPU, no data-path
collective; per-circuit counts are gathered on the host.

The reference runs one circuit per ``jku_simulator`` process, sequentially
(``qiskit_jku_provider/qasm_simulator_jku.py:340-359``).
"""

import ctypes as C

import numpy as np

from . import _native
from .engine import OP_DTYPE, Engine, circuit_ops
from .jkusimulatorerror import JKUSimulatorError


class BatchRunner:
    def __init__(self, workers=64, device=0, small=None, big=None, fallback_engines=4):
        """``workers``: engines in the pool (circuits in flight); ``small``: jkq_config fields of the pool engines;
        ``big``: options of the fall-back Engine."""
        self._lib = _native.load()
        self.device = int(device)
        self.workers = int(workers)
        cfg = _native.JkqConfig()
        cfg.device = self.device
        for k, v in (small or {}).items():
            if not hasattr(cfg, k):
                raise TypeError("unknown engine option %r" % k)
            setattr(cfg, k, v)
        self._b = C.c_void_p()
        rc = self._lib.jkq_batch_create(C.byref(self._b), self.device, self.workers, C.byref(cfg))
        if rc != 0:
            msg = self._lib.jkq_batch_last_error(self._b).decode() if self._b else "no usable CUDA device"
            if self._b:
                self._lib.jkq_batch_destroy(self._b)
                self._b = None
            raise JKUSimulatorError("jkq_batch_create failed (%d): %s" % (rc, msg))
        self.big_opts = dict(big or {})
        self._bigs = []
        self.fallback_engines = max(1, int(fallback_engines))
        self.fallbacks = 0

    def run(self, circuits, shots=1024, seed=1, rank=0, world=1):
        """Runs ``circuits[i]`` for ``i % world == rank``; returns {i: {outcome: count}}, total gate applies,
        sum of the per-circuit device times in ms.  Circuit i samples with seed ``seed + i``."""
        mine = [i for i in range(len(circuits)) if i % world == rank]
        if not mine:
            return {}, 0, 0.0
        return self.run_ops([(circuits[i].n, circuit_ops(circuits[i])) for i in mine], shots, [seed + i for i in mine], index=mine)

    def run_ops(self, programs, shots, seeds, index=None):
        """``programs``: sequence of (n_qubits, numpy array of OP_DTYPE) -- GATE / MEASURE / BARRIER ops, e.g. what
        ``JKUSimulatorWrapper.convert_qobj_circuit_to_ops`` makes of a Qobj experiment; program j samples with
        ``seeds[j]`` and is reported under key ``index[j]`` (default j).  Same return value as ``run``."""
        mine = list(index) if index is not None else list(range(len(programs)))
        if not mine:
            return {}, 0, 0.0
        arrs = [np.ascontiguousarray(a, dtype=OP_DTYPE) for _, a in programs]
        offs = np.zeros(len(mine) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(a) for a in arrs])
        ops = np.ascontiguousarray(np.concatenate(arrs) if len(arrs) > 1 else arrs[0], dtype=OP_DTYPE)
        if len(ops) == 0:
            ops = np.zeros(1, dtype=OP_DTYPE)  # (a valid pointer for a batch of empty programs)
        nq = np.array([n for n, _ in programs], dtype=np.uint32)
        seeds = np.array([int(x) for x in seeds], dtype=np.uint64)
        n_of = {i: int(programs[j][0]) for j, i in enumerate(mine)}
        cap = int(max(1, min(shots, 1 << 16)))
        keys = np.zeros((len(mine), cap), dtype=np.uint64)
        cnts = np.zeros((len(mine), cap), dtype=np.uint64)
        nd = np.zeros(len(mine), dtype=np.uint32)
        status = np.zeros(len(mine), dtype=np.int32)
        gates, ms = C.c_uint64(0), C.c_double(0)
        u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        import time
        t0 = time.perf_counter()
        rc = self._lib.jkq_batch_run(self._b, len(mine), nq.ctypes.data_as(u32p), ops.ctypes.data_as(C.POINTER(_native.JkqOp)),
                                     offs.ctypes.data_as(u64p), int(shots), seeds.ctypes.data_as(u64p), cap,
                                     keys.ctypes.data_as(u64p), cnts.ctypes.data_as(u64p), nd.ctypes.data_as(u32p),
                                     status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(gates), C.byref(ms))
        self.native_s = time.perf_counter() - t0  # time inside the C call (launches + waits), without the dict building
        if rc != 0:
            raise JKUSimulatorError("jkq_batch_run failed (%d): %s" % (rc, self._lib.jkq_batch_last_error(self._b).decode()))
        self.last_fast_path_error = self._lib.jkq_batch_last_error(self._b).decode() if status.any() else ""
        out, total_gates, dev_ms = {}, int(gates.value), float(ms.value)
        self.last_kernel_ms, self.last_fallback_ms = dev_ms, 0.0  # k_batch alone / circuits redone on the full engine
        redo = []
        for j, i in enumerate(mine):
            if status[j] == 0:
                k = int(nd[j])
                out[i] = dict(zip(keys[j, :k].tolist(), cnts[j, :k].tolist()))
            else:
                redo.append((j, i))
        if redo:
            # outgrew one SM's shared memory (or the small pools): the full engine, grid path included.  Up to
            # ``fallback_engines`` of them side by side, one host thread and one stream each: their single-CTA prefixes
            # (the circuit is redone from gate 0) overlap on different SMs, the grid kernels take turns.
            want = min(self.fallback_engines, len(redo))
            try:
                while len(self._bigs) < want:
                    self._bigs.append(Engine(device=self.device, **self.big_opts))
            except JKUSimulatorError as e:
                raise JKUSimulatorError("%s (batch fast path said: %s)" % (e, self.last_fast_path_error))
            self.fallbacks += len(redo)

            def work(w):
                eng, res = self._bigs[w], []
                for j, i in redo[w::want]:
                    eng.load_ops(n_of[i], arrs[j])
                    eng.run(shots, int(seeds[j]))
                    st = eng.stats()
                    res.append((i, eng.counts(), st["gates"], st["apply_ms"]))
                return res

            t1 = time.perf_counter()
            if want == 1:
                parts = [work(0)]
            else:
                from concurrent.futures import ThreadPoolExecutor
                with ThreadPoolExecutor(max_workers=want) as pool:
                    parts = list(pool.map(work, range(want)))
            wall_ms = (time.perf_counter() - t1) * 1e3
            for res in parts:
                for i, counts, g, _ in res:
                    out[i] = counts
                    total_gates += g
            # one engine: the sum of its device times; several side by side: the host clock around all of them (an upper
            # bound of the device time, sampling and the read-back of the counts included)
            self.last_fallback_ms = sum(r[3] for r in parts[0]) if want == 1 else wall_ms
            self.last_fallback_detail = [(i, round(ms_, 3)) for res in parts for i, _, _, ms_ in res]  # (circuit, its apply ms)
            dev_ms += self.last_fallback_ms
        return out, total_gates, dev_ms

    def run_shots(self, n_qubits, ops, shots, seed=1, shot_offset=0, measure_seed0=None, physical_reset=False):
        """Per-shot re-simulation of ONE circuit with mid-circuit measure / reset / if ops (``jkq_batch_run_shots``):
        ``shots`` instances side by side, instance s drawing its measurement outcomes from the stream
        ``measure_seed0 + s``.  Returns (keys[shots], mid_bits[shots], status[shots], gate applies, kernel ms);
        status != 0 marks an instance that outgrew the single-CTA path."""
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        shots = int(shots)
        keys = np.zeros(max(1, shots), dtype=np.uint64)
        bits = np.zeros(max(1, shots), dtype=np.uint64)
        status = np.zeros(max(1, shots), dtype=np.int32)
        gates, ms = C.c_uint64(0), C.c_double(0)
        u64p = C.POINTER(C.c_uint64)
        if measure_seed0 is None:
            measure_seed0 = (int(seed) * 0x9E3779B97F4A7C15) & 0x3FFFFFFFFFFFFFFF  # (same default as multi_gpu.run_shots_sharded)
        rc = self._lib.jkq_batch_run_shots(self._b, int(n_qubits), ops.ctypes.data_as(C.POINTER(_native.JkqOp)) if len(ops) else None, len(ops),
                                           shots, int(seed), int(shot_offset), int(measure_seed0) & 0xFFFFFFFFFFFFFFFF, int(bool(physical_reset)),
                                           keys.ctypes.data_as(u64p), bits.ctypes.data_as(u64p), status.ctypes.data_as(C.POINTER(C.c_int32)),
                                           C.byref(gates), C.byref(ms))
        if rc != 0:
            raise JKUSimulatorError("jkq_batch_run_shots failed (%d): %s" % (rc, self._lib.jkq_batch_last_error(self._b).decode()))
        return keys[:shots], bits[:shots], status[:shots], int(gates.value), float(ms.value)

    def close(self):
        if getattr(self, "_b", None):
            self._lib.jkq_batch_destroy(self._b)
            self._b = None
        for e in self._bigs:
            e.close()
        self._bigs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
