"""Backend for the JKU decision-diagram simulator, B200-native engine.

Mirror of the reference's ``qiskit_jku_provider/qasm_simulator_jku.py``: same classes
(``QasmSimulator``, ``JKUSimulatorWrapper``), same Qobj fields read, same result dictionary.  What the
reference does by writing OpenQASM text and ``qelib1.inc`` to disk and piping them to the
``jku_simulator`` executable (``:82-97,149-171``) happens here through the C ABI of
``libjkq_b200.so`` (``include/jkq.h``) via ctypes: the Qobj instructions are macro-expanded to the
engine's primitives (``circuits.expand`` restates ``qelib1``, ``:32-60``) and handed over as an array of
``jkq_op``; counts and state vectors come back in caller-owned buffers instead of JSON text.
"""

import logging
import random
import time
import uuid

import numpy as np

from . import circuits
from .compat import BackendConfiguration, BackendStatus, BaseBackend, Result
from .engine import OP_BARRIER, OP_DTYPE, OP_GATE, OP_MEASURE, OP_RESET, OP_SNAPSHOT, Engine
from .jkujob import JKUJob
from .jkusimulatorerror import JKUSimulatorError

RUN_MESSAGE = """DD-based simulator by JKU Linz, Austria
Developer: Alwin Zulehner, Robert Wille
For more information, please visit http://iic.jku.at/eda/research/quantum_simulation"""

VERSION = '1.0.0'  # VERSION.txt of the reference

logger = logging.getLogger(__name__)

# The reference always appends a final snapshot and therefore always materialises 2^n amplitudes
# (``:116-120,138``).  Above this many qubits the final state vector is only produced on request
# (experiment.config.data contains 'statevector').
STATEVECTOR_AUTO_MAX_QUBITS = 20
BATCH_MIN_EXPERIMENTS = 4      # from here on the experiments of a Qobj that only ask for counts go through k_batch
BATCH_MAX_SHOTS = 1 << 16      # (capacity of the per-circuit result block of jkq_batch_run)
MULTI_GPU_MIN_SHOTS = 1 << 22  # from here on the shots of one experiment are split over the backend's devices


class _Templates:
    """Registry of macro-expanded gates: (name, number of qubits, params) -> a run of primitives in flat tables
    (target position, control positions / values, matrix index into a table of distinct 2x2 matrices), positions into
    the instruction's qubit list.  ``circuits.expand`` restates qelib1 (reference ``:32-60``).  The tables are kept as
    numpy arrays so that a whole experiment is stamped out with a handful of vector operations instead of a Python
    loop over its primitives."""

    def __init__(self):
        self.ids = {}
        self.start, self.count = [], []
        self.p_kind, self.p_t, self.p_m, self.p_ctl = [], [], [], []   # per primitive; p_ctl: ((position, value), ...)
        self.mats, self.mat_ids = [], {}
        self._np = None
        for kind in (OP_MEASURE, OP_SNAPSHOT, OP_BARRIER, OP_RESET):  # the non-gate instructions are one-slot templates
            self._add(("#", kind), [(kind, 0 if kind in (OP_MEASURE, OP_RESET) else -1, (), (0.0,) * 8)])

    def _add(self, key, prims):
        tid = len(self.start)
        self.start.append(len(self.p_kind)); self.count.append(len(prims))
        for kind, t_pos, ctl, flat in prims:
            mi = self.mat_ids.get(flat)
            if mi is None:
                mi = self.mat_ids[flat] = len(self.mats)
                self.mats.append(flat)
            self.p_kind.append(kind); self.p_t.append(t_pos); self.p_ctl.append(ctl); self.p_m.append(mi)
        self.ids[key] = tid
        self._np = None
        return tid

    def special(self, kind):
        return self.ids[("#", kind)]

    def gate(self, key):
        """key = (name, n_qubits, params).  Raises JKUSimulatorError for a gate qelib1 does not define."""
        name, n_qubits, params = key
        try:
            prims = circuits.expand(name, list(range(n_qubits)), params)
        except (KeyError, IndexError, ValueError):
            raise JKUSimulatorError("Undefined gate: {}".format(name))
        out = []
        for p in prims:
            m = p.matrix()
            flat = tuple(x for row in m for z in row for x in (complex(z).real, complex(z).imag))
            out.append((OP_GATE, p.target, tuple(p.controls), flat))
        return self._add(key, out)

    def tables(self):
        if self._np is None:
            width = max(1, max(len(c) for c in self.p_ctl))
            cpos = np.full((len(self.p_ctl), width), -1, dtype=np.int64)
            cval = np.zeros((len(self.p_ctl), width), dtype=bool)
            for i, ctl in enumerate(self.p_ctl):
                for k, (pos, val) in enumerate(ctl):
                    cpos[i, k] = pos
                    cval[i, k] = bool(val)
            self._np = (np.array(self.start, dtype=np.int64), np.array(self.count, dtype=np.int64),
                        np.array(self.p_kind, dtype=np.uint32), np.array(self.p_t, dtype=np.int64),
                        np.array(self.p_m, dtype=np.int64), cpos, cval, np.array(self.mats, dtype=np.float64).reshape(-1, 8))
        return self._np


_TEMPLATES = None


def _templates():
    global _TEMPLATES
    if _TEMPLATES is None:
        _TEMPLATES = _Templates()
    return _TEMPLATES


class JKUSimulatorWrapper:
    """Converter between Qobj experiments and the engine."""

    def __init__(self, engine=None, silent=False, physical_reset=False, shots_runner=None, devices=None, statevector='auto'):
        self.physical_reset = bool(physical_reset)
        self.statevector = statevector
        self.devices = list(devices) if devices else None  # CUDA ordinals the shots of one experiment may be split over
        self.shots_runner = shots_runner  # batch.BatchRunner for per-shot re-simulation, or None: sequential
        self.seed = 0
        self.shots = 1
        self.engine = engine
        self.additional_output_data = []
        self.silent = silent
        self.max_snapshot_index = 0

    def set_config(self, global_config, experiment_config):
        # reference :74-80 -- note that `seed_simulator` is ignored there as well
        self.additional_output_data = []
        if hasattr(experiment_config, 'data'):
            self.additional_output_data = experiment_config.data
        if hasattr(global_config, 'seed'):
            self.seed = global_config.seed
        else:
            self.seed = random.getrandbits(32)

    def compute_measurement_data(self, experiment):
        """qubit -> classical bit map of the end-of-circuit measurements (reference :238-256)."""
        header = experiment.header
        measurement_data = {'mapping': {},
                            'clbits': header.creg_sizes,
                            'clbits_num': len(header.clbit_labels),
                            'qubits_num': len(header.qubit_labels)}
        for op in experiment.instructions:
            if op.name == "measure":
                measurement_data['mapping'][op.qubits[0]] = op.memory[0]
            else:
                for qubit in op.qubits:
                    if qubit in measurement_data['mapping'].keys() and not op.name == 'snapshot':
                        raise JKUSimulatorError(
                            "Error: qubit {} was used after being measured. This is currently "
                            "not supported by JKU".format(qubit))
        return measurement_data

    def convert_qobj_circuit_to_ops(self, experiment, want_final_statevector):
        """Qobj instructions -> packed array of jkq_op (replaces convert_qobj_circuit_to_jku_qasm, :123-140).

        Each named gate is macro-expanded once per (name, params) into a template of primitives
        (``circuits.expand`` restates qelib1) and stamped onto the instruction's qubits; the result is one
        numpy structured array with the memory layout of ``jkq_op`` (include/jkq.h)."""
        T = _templates()
        if len(T.start) >= 65536:  # (circuits of all-different angles: start over rather than grow without bound)
            T.__init__()
        ids = T.ids
        t_measure, t_snapshot, t_barrier, t_reset = (T.special(k) for k in (OP_MEASURE, OP_SNAPSHOT, OP_BARRIER, OP_RESET))
        tids, args, qlens, flatq = [], [], [], []   # per instruction: template, argument, its qubit list (flattened)
        self.per_shot = False
        self.max_snapshot_index = 0
        snapshot_slots = []
        for op in experiment.instructions:
            name = op.name
            if name == 'measure':
                tids.append(t_measure); args.append(op.memory[0]); qlens.append(1); flatq.append(op.qubits[0])
            elif name == 'snapshot':
                params = list(getattr(op, 'params', []) or [])
                slot = int(params[0]) if params else int(getattr(op, 'label', 0))
                self.max_snapshot_index = max(self.max_snapshot_index, slot)
                snapshot_slots.append(slot)
                tids.append(t_snapshot); args.append(slot); qlens.append(0)
            elif name == 'barrier':
                tids.append(t_barrier); args.append(0); qlens.append(0)
            elif name == 'reset':
                # ResetQubit (Simulator.cpp:291-339) PROJECTS on |0> and rescales -- deterministic, so the shots are
                # drawn from the one final state.  With ``physical_reset`` the qubit is measured and flipped back
                # instead; that outcome is random, and the circuit is re-simulated per shot like the reference does
                # when a state is not reusable (QASMsimulator.cpp:501-522).
                if self.physical_reset:
                    self.per_shot = True
                tids.append(t_reset); args.append(0); qlens.append(1); flatq.append(op.qubits[0])
            else:
                params = getattr(op, 'params', None)
                qubits = op.qubits
                key = (name, len(qubits), tuple(params) if params else ())
                tid = ids.get(key)
                if tid is None:
                    tid = T.gate(key)
                    ids = T.ids
                tids.append(tid); args.append(0); qlens.append(len(qubits)); flatq.extend(qubits)
        self.max_snapshot_index += 1  # add_final_snapshot, :116-120
        if want_final_statevector:
            tids.append(t_snapshot); args.append(self.max_snapshot_index); qlens.append(0)
            snapshot_slots.append(self.max_snapshot_index)
        if not tids:
            return np.zeros(0, dtype=OP_DTYPE), snapshot_slots
        # stamp the templates onto the instructions' qubits, all primitives of the experiment at once
        t_start, t_count, p_kind, p_t, p_m, p_cpos, p_cval, mats = T.tables()
        tids = np.array(tids, dtype=np.int64)
        cnt = t_count[tids]
        first = np.cumsum(cnt) - cnt                       # index of the instruction's first primitive in the output
        inst = np.repeat(np.arange(len(tids)), cnt)        # primitive -> instruction
        prim = t_start[tids][inst] + (np.arange(len(inst)) - first[inst])   # primitive -> row of the template tables
        qlens = np.array(qlens, dtype=np.int64)
        qoff = (np.cumsum(qlens) - qlens)[inst]            # where the instruction's qubits start in flatq
        flatq = np.array(flatq + [0], dtype=np.int64)
        ops = np.zeros(len(inst), dtype=OP_DTYPE)
        ops['kind'] = p_kind[prim]
        tp = p_t[prim]
        ops['target'] = np.where(tp >= 0, flatq[np.minimum(qoff + np.maximum(tp, 0), len(flatq) - 1)], 0)
        cp = p_cpos[prim]                                  # (primitives, widest control list), -1 = no control
        bit = np.left_shift(np.uint64(1), flatq[np.minimum(qoff[:, None] + np.maximum(cp, 0), len(flatq) - 1)].astype(np.uint64))
        has = cp >= 0
        val = p_cval[prim]
        ops['ctrl_pos'] = np.bitwise_or.reduce(np.where(has & val, bit, np.uint64(0)), axis=1)
        ops['ctrl_neg'] = np.bitwise_or.reduce(np.where(has & ~val, bit, np.uint64(0)), axis=1)
        ops['m'] = mats[p_m[prim]]
        ops['arg'] = np.array(args, dtype=np.int64)[inst]
        return ops, snapshot_slots

    def wants_statevector(self, n_qubits):
        if 'statevector' in self.additional_output_data:
            return True
        return self.statevector == 'auto' and n_qubits <= STATEVECTOR_AUTO_MAX_QUBITS

    def run_experiment(self, config, experiment):
        self.set_config(config, experiment.config)
        measurement_data = self.compute_measurement_data(experiment)
        n = measurement_data['qubits_num']
        want_sv = self.wants_statevector(n)
        self.start_time = time.time()
        ops, slots = self.convert_qobj_circuit_to_ops(experiment, want_sv)
        if not self.silent:
            print(RUN_MESSAGE)
        eng = self.engine
        eng.set_option("physical_reset", int(self.physical_reset))
        eng.load_ops(n, ops)
        if self.per_shot:
            # Every shot is its own run with fresh measurement outcomes (QASMsimulator.cpp:501-522).  All shots go
            # through one launch of the shots kernel (jkq_batch_run_shots: one instance per SM at a time); a shot the
            # single-CTA path cannot finish is redone on the full engine with the same Philox streams.
            shots = int(self.shots)
            mseed0 = (int(self.seed) * 0x9E3779B97F4A7C15) & 0x3FFFFFFFFFFFFFFF
            merged = {}
            status = np.ones(shots, dtype=np.int32)
            if self.shots_runner is not None:
                keys, _, status, _, _ = self.shots_runner.run_shots(n, ops, shots, seed=int(self.seed), measure_seed0=mseed0,
                                                                    physical_reset=self.physical_reset)
                for k in keys[status == 0].tolist():
                    merged[k] = merged.get(k, 0) + 1
            for shot in np.nonzero(status)[0].tolist():
                eng.set_option("measure_seed", mseed0 + shot)
                eng.apply()
                eng.sample(1, seed=int(self.seed), shot_offset=shot)
                for k, v in eng.counts().items():
                    merged[k] = merged.get(k, 0) + v
            if not np.any(status):  # the final state of the last shot, for the snapshots
                eng.set_option("measure_seed", mseed0 + shots - 1)
                eng.apply()
            output_data = self.collect_output(eng, measurement_data, slots, want_sv, counts_override=merged)
        else:
            if self.devices and len(self.devices) > 1 and int(self.shots) >= MULTI_GPU_MIN_SHOTS:
                # the shots of one state over several GPUs (jkq_sample_multi_gpu): same counts as on one
                eng.apply()
                self.multi_gpu_stats = eng.sample_multi(self.devices, self.shots, self.seed)
            else:
                eng.run(self.shots, self.seed)
            output_data = self.collect_output(eng, measurement_data, slots, want_sv)
        self.end_time = time.time()
        stats = eng.stats()
        result_dict = {'header': {'name': experiment.header.name,
                                  'memory_slots': experiment.config.memory_slots,
                                  'creg_sizes': experiment.header.creg_sizes
                                  },
                       'status': 'DONE', 'time_taken': self.end_time - self.start_time,
                       'seed': self.seed, 'shots': self.shots,
                       'data': output_data,
                       'success': True,
                       'metadata': {'engine': stats},
                       }
        return result_dict

    def collect_output(self, eng, measurement_data, slots, want_sv, counts_override=None):
        """Counts as hex keys over the classical bits (reference convert_counts/qubits_to_clbits,
        :219-235) and snapshots in Qiskit order (:190-211).  The engine already keys outcomes by
        classical bit and indexes amplitudes with bit q = qubit q, so no bit-reversal table is needed."""
        result = {}
        clbits_num = measurement_data['clbits_num']
        if clbits_num > 0:
            if measurement_data['mapping']:
                if counts_override is not None:
                    result['counts'] = {hex(int(k)): int(v) for k, v in counts_override.items()}
                else:
                    keys, vals = eng.counts_arrays()
                    result['counts'] = dict(zip(map(hex, keys.tolist()), vals.tolist()))
            else:  # nothing measured: every shot reads all-zero classical bits
                result['counts'] = {hex(0): int(self.shots)}
        else:
            result['counts'] = {}
        snapshots = {}
        for slot in slots:
            sv = eng.statevector(slot)
            snap = {'statevector': np.stack([sv.real, sv.imag], axis=1).tolist()}
            probs = (sv.real ** 2 + sv.imag ** 2)
            if 'probabilities' in self.additional_output_data:
                snap['probabilities'] = probs.tolist()
            if 'probabilities_ket' in self.additional_output_data:
                nq = measurement_data['qubits_num']
                snap['probabilities_ket'] = {format(i, '0%db' % nq): float(p) for i, p in enumerate(probs) if p > 0.0}
            snapshots[str(slot)] = snap
        if snapshots:
            result['snapshots'] = snapshots
        if want_sv:
            result['statevector'] = snapshots[str(self.max_snapshot_index)]['statevector']
        return result


class QasmSimulator(BaseBackend):
    """Python interface to the JKU decision-diagram simulator (B200 engine)."""

    DEFAULT_CONFIGURATION = {
        'backend_name': 'qasm_simulator',
        'backend_version': VERSION,
        'url': 'https://github.com/Qiskit/qiskit-jku-provider',
        'simulator': True,
        'local': True,
        'description': 'JKU C++ simulator',
        'basis_gates': ['u0', 'u1', 'u2', 'u3', 'cx', 'x', 'y', 'z', 'h', 's', 't', 'snapshot'],
        'memory': False,
        'n_qubits': 30,
        'conditional': False,
        'max_shots': 100000,
        'coupling_map': None,
        'open_pulse': False,
        'gates': [
            {
                'name': 'TODO',
                'parameters': [],
                'qasm_def': 'TODO'
            }
        ]
    }

    def __init__(self, configuration=None, provider=None, silent=False, device=0, physical_reset=False, devices=None,
                 statevector='auto', **engine_options):
        """
        Args:
            configuration (dict): backend configuration
            device (int): CUDA ordinal the engine runs on
            devices (list of int): CUDA ordinals of ONE process to spread work over (default: ``[device]``).  The
                                   experiments of a Qobj go round-robin to one engine per device and run side by side
                                   (the reference runs them one after the other, ``:340-359``); the shots of a single
                                   experiment are split over all of them once there are at least 2^22
                                   (``jkq_sample_multi_gpu``: peer copy of the sampling DD, device-side merge).
                                   Counts do not depend on the number of devices.
            statevector: 'auto' (default) -- every experiment of at most 20 qubits also returns its final state vector,
                                   like the reference, which always appends a final snapshot (``:116-120``); False -- only
                                   when the experiment asks for it (``config.data`` contains 'statevector').  A Qobj of
                                   at least 4 experiments that return counts only runs as ONE batch: one persistent
                                   kernel per device, one circuit per SM at a time (``jkq_batch_run``), instead of
                                   one experiment after the other (``:340-359``) -- same counts either way.
            physical_reset (bool): ``reset`` = measure and flip back (per-shot re-simulation) instead of the
                                   reference's projection on |0> (Simulator.cpp:291-339)
            engine_options: forwarded to ``jkq_config`` (node_capacity, tolerance, ...)
        Raises:
            ImportError: if libjkq_b200.so was not built (the reference raises FileNotFoundError when
                         its executable is missing, ``:326-332``).
        """
        super().__init__(configuration=(configuration or
                                        BackendConfiguration.from_dict(self.DEFAULT_CONFIGURATION)),
                         provider=provider)
        from . import _native
        _native.load()  # fail early and loudly, like the reference's executable check
        self.silent = silent
        self.physical_reset = physical_reset
        self.statevector = statevector
        self._batch_runners = {}  # device ordinal -> batch.BatchRunner (one engine per SM) for count-only Qobjs
        self._devices = [int(d) for d in devices] if devices else [int(device)]
        self._device = self._devices[0]
        self._engine_options = engine_options
        self._engine = None
        self._more_engines = {}   # device ordinal -> Engine, for the experiments of a Qobj that run side by side
        self._shots_runner = None

    def _get_engine(self, device=None):
        if device is None or device == self._device:
            if self._engine is None:  # lazy: creating the CUDA context is deferred to the first job
                self._engine = Engine(device=self._device, **self._engine_options)
            return self._engine
        if device not in self._more_engines:
            self._more_engines[device] = Engine(device=device, **self._engine_options)
        return self._more_engines[device]

    def _get_shots_runner(self):
        if self._shots_runner is None:  # pool of small engines, one per SM, for per-shot re-simulation
            from .batch import BatchRunner
            small = {"tolerance": self._engine_options["tolerance"]} if "tolerance" in self._engine_options else None
            self._shots_runner = BatchRunner(workers=1 << 16, device=self._device, small=small)
        return self._shots_runner

    def run(self, qobj):
        job_id = str(uuid.uuid4())
        local_job = JKUJob(self, job_id, self._run_job, qobj)
        local_job.submit()
        return local_job

    def _run_job(self, job_id, qobj):
        """Run circuits in q_job"""
        result_list = []
        self._validate(qobj)
        start = time.time()
        n_dev = len(self._devices)
        batch = self._batch_plan(qobj)
        if batch is not None:
            result_list = self._run_batch(qobj, batch)
        elif n_dev > 1 and len(qobj.experiments) > 1:
            # one engine per device, experiment i on device i % n_dev, every device in its own host thread (ctypes
            # releases the GIL inside the library)
            from concurrent.futures import ThreadPoolExecutor

            def lane(k):
                w = JKUSimulatorWrapper(self._get_engine(self._devices[k]), silent=self.silent, physical_reset=self.physical_reset,
                                        shots_runner=None, statevector=self.statevector)
                w.shots = qobj.config.shots
                return [(i, w.run_experiment(qobj.config, qobj.experiments[i])) for i in range(k, len(qobj.experiments), n_dev)]

            for k in range(n_dev):
                self._get_engine(self._devices[k])  # (created here: one CUDA context after the other)
            with ThreadPoolExecutor(n_dev) as ex:
                done = dict(p for part in ex.map(lane, range(n_dev)) for p in part)
            result_list = [done[i] for i in range(len(qobj.experiments))]
        else:
            s = JKUSimulatorWrapper(self._get_engine(), silent=self.silent, physical_reset=self.physical_reset,
                                    shots_runner=self._get_shots_runner() if self.physical_reset else None,
                                    devices=self._devices, statevector=self.statevector)
            s.shots = qobj.config.shots
            for experiment in qobj.experiments:
                result_list.append(s.run_experiment(qobj.config, experiment))
        end = time.time()
        result = {'backend_name': self._configuration.backend_name,
                  'backend_version': VERSION,
                  'qobj_id': qobj.qobj_id,
                  'job_id': job_id,
                  'results': result_list,
                  'status': 'COMPLETED',
                  'success': True,
                  'time_taken': (end - start)}
        return Result.from_dict(result)

    def _batch_plan(self, qobj):
        """The converted experiments [(wrapper, n_qubits, ops, measurement data)] if the whole Qobj can go through the
        batch kernel (counts only: no state vector, no snapshot, no per-shot re-simulation), else None."""
        exps = qobj.experiments
        if len(exps) < BATCH_MIN_EXPERIMENTS or int(qobj.config.shots) > BATCH_MAX_SHOTS:
            return None
        plan = []
        for experiment in exps:
            w = JKUSimulatorWrapper(None, silent=True, physical_reset=self.physical_reset, statevector=self.statevector)
            w.shots = qobj.config.shots
            w.set_config(qobj.config, experiment.config)
            md = w.compute_measurement_data(experiment)
            n = md['qubits_num']
            if w.wants_statevector(n) or not md['mapping'] or md['clbits_num'] == 0 or n > 64:
                return None
            ops, slots = w.convert_qobj_circuit_to_ops(experiment, False)
            if slots or w.per_shot or bool(np.any(ops['kind'] == OP_RESET)):
                return None
            plan.append((w, n, ops, md))
        return plan

    def _get_batch_runner(self, device):
        if device not in self._batch_runners:
            from .batch import BatchRunner
            small = {"tolerance": self._engine_options["tolerance"]} if "tolerance" in self._engine_options else None
            self._batch_runners[device] = BatchRunner(workers=1 << 16, device=device, small=small, big=dict(self._engine_options))
        return self._batch_runners[device]

    def _run_batch(self, qobj, plan):
        """All experiments of the Qobj as one batch per device (experiment i on device i % n_dev, one host thread each);
        experiment i samples with the seed ``run_experiment`` would use, so the counts are those of the sequential path."""
        if not self.silent:
            print(RUN_MESSAGE)
        start = time.time()
        n_dev = len(self._devices)
        shots = int(qobj.config.shots)

        def lane(k):
            idx = list(range(k, len(plan), n_dev))
            if not idx:
                return {}, {}
            br = self._get_batch_runner(self._devices[k])
            out, _, _ = br.run_ops([(plan[i][1], plan[i][2]) for i in idx], shots, [plan[i][0].seed for i in idx], index=idx)
            return out, {"fallbacks": br.fallbacks, "kernel_ms": br.last_kernel_ms, "fallback_ms": br.last_fallback_ms}

        if n_dev == 1:
            parts = [lane(0)]
        else:
            from concurrent.futures import ThreadPoolExecutor
            for k in range(n_dev):
                self._get_batch_runner(self._devices[k])  # (created here: one CUDA context after the other)
            with ThreadPoolExecutor(n_dev) as ex:
                parts = list(ex.map(lane, range(n_dev)))
        elapsed = time.time() - start
        result_list = []
        for i, (w, n, ops, md) in enumerate(plan):
            k = i % n_dev
            counts = parts[k][0][i]
            experiment = qobj.experiments[i]
            result_list.append({'header': {'name': experiment.header.name,
                                           'memory_slots': experiment.config.memory_slots,
                                           'creg_sizes': experiment.header.creg_sizes},
                                'status': 'DONE', 'time_taken': elapsed / len(plan),
                                'seed': w.seed, 'shots': shots,
                                'data': {'counts': {hex(int(key)): int(v) for key, v in counts.items()}},
                                'success': True,
                                'metadata': {'batch': dict(parts[k][1], device=self._devices[k], experiments=len(plan))}})
        return result_list

    def _validate(self, qobj):
        # no-op in the reference as well (:361-364)
        return

    def status(self):
        return BackendStatus(backend_name=self.name(),
                             backend_version=self.configuration().backend_version,
                             operational=True,
                             pending_jobs=0,
                             status_msg='')
