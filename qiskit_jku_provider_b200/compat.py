"""Qiskit-Terra-0.8-era base classes, or minimal stand-ins when qiskit is not installed.

The reference is written against qiskit-terra 0.8 (``setup.py:115``): ``BaseProvider``, ``BaseBackend``,
``BaseJob``, ``JobStatus``, ``JobError``, ``BackendConfiguration``, ``BackendStatus``, ``Result`` and the
``Qobj`` model.  Those classes no longer exist in current qiskit and qiskit is absent from this image,
so the provider API is kept on duck-typed stand-ins with the same attribute names the reference reads
(``qasm_simulator_jku.py:75-80,124-126,162-164,241-249,346-353``).  A small circuit builder
(``QuantumRegister`` / ``ClassicalRegister`` / ``QuantumCircuit`` / ``assemble`` / ``execute``) produces
such Qobj objects so that the reference's tests and examples can be restated without qiskit.
"""

import enum
import uuid
from types import SimpleNamespace


class QiskitError(Exception):
    def __init__(self, *message):
        super().__init__(' '.join(str(m) for m in message))
        self.message = ' '.join(str(m) for m in message)

    def __str__(self):
        return repr(self.message)


class JobError(QiskitError):
    """Base class for errors raised by jobs."""


class JobStatus(enum.Enum):
    INITIALIZING = 'job is being initialized'
    QUEUED = 'job is queued'
    VALIDATING = 'job is being validated'
    RUNNING = 'job is actively running'
    CANCELLED = 'job has been cancelled'
    DONE = 'job has successfully run'
    ERROR = 'job incurred error'


class _Model(SimpleNamespace):
    def get(self, key, default=None):
        return getattr(self, key, default)

    def to_dict(self):
        return dict(self.__dict__)

    @classmethod
    def from_dict(cls, d):
        return cls(**d)


class BackendConfiguration(_Model):
    pass


class BackendStatus(_Model):
    pass


class BaseProvider:
    def __init__(self, *args, **kwargs):
        pass

    def get_backend(self, name=None, **kwargs):
        backends = self.backends(name, **kwargs)
        if not backends:
            raise QiskitError('No backend matches the criteria')
        return backends[0]


class BaseBackend:
    def __init__(self, configuration, provider=None):
        self._configuration = configuration
        self._provider = provider

    def configuration(self):
        return self._configuration

    def properties(self):
        return None

    def provider(self):
        return self._provider

    def name(self):
        return self._configuration.backend_name

    def __str__(self):
        return self.name()


class BaseJob:
    def __init__(self, backend, job_id):
        self._job_id = job_id
        self._backend = backend

    def job_id(self):
        return self._job_id

    def backend(self):
        return self._backend


# ------------------------------------------------------------------------------------- Result
def _format_counts(hex_counts, creg_sizes, memory_slots):
    """hex keys -> bit strings with a space between classical registers (what qiskit's
    Result.get_counts does; gives '01 10' / '100 100' in the reference's tests)."""
    out = {}
    if not (creg_sizes and len(creg_sizes) > 1):  # one register: no spaces to put in
        fmt = '0%db' % memory_slots
        for key, val in hex_counts.items():
            bits = format(int(key, 16), fmt)
            out[bits] = out.get(bits, 0) + val
        return out
    for key, val in hex_counts.items():
        bits = bin(int(key, 16))[2:].rjust(memory_slots, '0')
        if creg_sizes and len(creg_sizes) > 1:
            parts, pos = [], len(bits)
            for _, size in creg_sizes:
                parts.insert(0, bits[pos - size:pos])
                pos -= size
            bits = ' '.join(parts)
        out[bits] = out.get(bits, 0) + val
    return out


class ExperimentResult(_Model):
    pass


class Result:
    """Subset of qiskit.result.Result the reference's tests and examples use."""

    def __init__(self, d):
        self._d = d
        self.backend_name = d.get('backend_name')
        self.backend_version = d.get('backend_version')
        self.qobj_id = d.get('qobj_id')
        self.job_id = d.get('job_id')
        self.success = d.get('success')
        self.status = d.get('status')
        self.time_taken = d.get('time_taken')
        self.results = [ExperimentResult(**r) for r in d.get('results', [])]

    @classmethod
    def from_dict(cls, d):
        return cls(d)

    def to_dict(self):
        return self._d

    def _get_experiment(self, key=None):
        if key is None:
            if len(self.results) != 1:
                raise QiskitError('You have to select a circuit when there is more than one available')
            return self.results[0]
        if isinstance(key, int):
            return self.results[key]
        name = key if isinstance(key, str) else getattr(key, 'name', None)
        for r in self.results:
            if r.header.get('name') == name:
                return r
        raise QiskitError('Data for experiment "%s" could not be found.' % name)

    def data(self, experiment=None):
        return self._get_experiment(experiment).data

    def get_counts(self, experiment=None):
        exp = self._get_experiment(experiment)
        if 'counts' not in exp.data:
            raise QiskitError('No counts for experiment "%s"' % exp.header.get('name'))
        return _format_counts(exp.data['counts'], exp.header.get('creg_sizes'), exp.header.get('memory_slots', 0))

    def get_statevector(self, experiment=None):
        import numpy as np
        exp = self._get_experiment(experiment)
        if 'statevector' not in exp.data:
            raise QiskitError('No statevector for experiment "%s"' % exp.header.get('name'))
        sv = np.asarray(exp.data['statevector'], dtype=float)
        return sv[:, 0] + 1j * sv[:, 1]


# ------------------------------------------------------------------------------------- Qobj model
class QobjInstruction(_Model):
    pass


class QobjExperiment(_Model):
    pass


class Qobj(_Model):
    pass


def validate_qobj_against_schema(qobj):
    """The reference validates against the JSON schema (jkujob.py:70); here: the fields it reads."""
    for attr in ('qobj_id', 'config', 'experiments'):
        if not hasattr(qobj, attr):
            raise QiskitError('Qobj misses "%s"' % attr)
    if not hasattr(qobj.config, 'shots'):
        raise QiskitError('Qobj config misses "shots"')


# ------------------------------------------------------------------------------------- circuit builder
class _Register:
    def __init__(self, size, name=None):
        self.size = int(size)
        self.name = name or ('%s%d' % (self.prefix, _Register._counter))
        _Register._counter += 1

    _counter = 0

    def __getitem__(self, i):
        if not 0 <= i < self.size:
            raise IndexError(i)
        return (self, i)

    def __len__(self):
        return self.size

    def __iter__(self):
        return iter((self, i) for i in range(self.size))


class QuantumRegister(_Register):
    prefix = 'q'


class ClassicalRegister(_Register):
    prefix = 'c'


_GATES = {  # name -> (number of qubits, number of parameters)
    'u0': (1, 1), 'u1': (1, 1), 'u2': (1, 2), 'u3': (1, 3), 'x': (1, 0), 'y': (1, 0), 'z': (1, 0), 'h': (1, 0),
    's': (1, 0), 'sdg': (1, 0), 't': (1, 0), 'tdg': (1, 0), 'iden': (1, 0), 'id': (1, 0), 'rx': (1, 1), 'ry': (1, 1),
    'rz': (1, 1), 'cx': (2, 0), 'cy': (2, 0), 'cz': (2, 0), 'ch': (2, 0), 'swap': (2, 0), 'crz': (2, 1),
    'cu1': (2, 1), 'cu3': (2, 3), 'rzz': (2, 1), 'ccx': (3, 0), 'cswap': (3, 0),
}


class QuantumCircuit:
    """Just enough of qiskit.QuantumCircuit for the reference's tests: registers, the qelib1 gates,
    measure, barrier, snapshot, ``+``."""

    _n = 0

    def __init__(self, *regs, name=None):
        self.qregs, self.cregs, self.data = [], [], []
        self.name = name or ('circuit%d' % QuantumCircuit._n)
        QuantumCircuit._n += 1
        self.add_register(*regs)

    def add_register(self, *regs):
        for r in regs:
            if isinstance(r, QuantumRegister) and r not in self.qregs:
                self.qregs.append(r)
            elif isinstance(r, ClassicalRegister) and r not in self.cregs:
                self.cregs.append(r)

    def _qindex(self, bit):
        reg, i = bit
        off = 0
        for r in self.qregs:
            if r is reg:
                return off + i
            off += r.size
        raise QiskitError('qubit register not in circuit')

    def _cindex(self, bit):
        reg, i = bit
        off = 0
        for r in self.cregs:
            if r is reg:
                return off + i
            off += r.size
        raise QiskitError('classical register not in circuit')

    def __getattr__(self, name):
        if name in _GATES:
            nq, np_ = _GATES[name]

            def gate(*args):
                params, qargs = list(args[:np_]), list(args[np_:])
                if len(qargs) != nq:
                    raise QiskitError('%s takes %d qubits' % (name, nq))
                # register-wide broadcast, like `h q;` (QASMsimulator.cpp:294,324-350)
                width = max([a.size for a in qargs if isinstance(a, _Register)] or [1])
                for k in range(width):
                    bits = [a[k] if isinstance(a, _Register) else a for a in qargs]
                    self.data.append((name, [float(p) for p in params], [self._qindex(b) for b in bits], []))
                return self
            return gate
        raise AttributeError(name)

    def reset(self, q):
        if isinstance(q, _Register):
            for k in range(q.size):
                self.reset(q[k])
            return self
        self.data.append(('reset', [], [self._qindex(q)], []))
        return self

    def measure(self, q, c):
        if isinstance(q, _Register):
            for k in range(q.size):
                self.measure(q[k], c[k])
            return self
        self.data.append(('measure', [], [self._qindex(q)], [self._cindex(c)]))
        return self

    def barrier(self, *qargs):
        bits = []
        for a in (qargs or self.qregs):
            bits += [self._qindex(b) for b in (a if isinstance(a, _Register) else [a])]
        self.data.append(('barrier', [], bits, []))
        return self

    def snapshot(self, slot):
        self.data.append(('snapshot', [int(slot)], list(range(sum(r.size for r in self.qregs))), []))
        return self

    def __add__(self, other):
        out = QuantumCircuit(*(self.qregs + self.cregs), name=self.name)
        out.add_register(*(other.qregs + other.cregs))
        for src in (self, other):
            for name, params, qbits, cbits in src.data:
                if name == 'measure':
                    out.data.append((name, params, [out._qindex_by_reg(src, qbits[0])], [out._cindex_by_reg(src, cbits[0])]))
                else:
                    out.data.append((name, params, [out._qindex_by_reg(src, q) for q in qbits], []))
        return out

    def _bit_of(self, regs, flat):
        for r in regs:
            if flat < r.size:
                return (r, flat)
            flat -= r.size
        raise IndexError(flat)

    def _qindex_by_reg(self, src, flat):
        return self._qindex(src._bit_of(src.qregs, flat))

    def _cindex_by_reg(self, src, flat):
        return self._cindex(src._bit_of(src.cregs, flat))


def assemble(circuits, shots=1024, seed=None, qobj_id=None, data=None):
    """QuantumCircuit(s) -> Qobj with the fields the backend reads."""
    if isinstance(circuits, QuantumCircuit):
        circuits = [circuits]
    exps = []
    for qc in circuits:
        qubit_labels = [[r.name, i] for r in qc.qregs for i in range(r.size)]
        clbit_labels = [[r.name, i] for r in qc.cregs for i in range(r.size)]
        instrs = []
        for name, params, qbits, cbits in qc.data:
            ins = QobjInstruction(name=name, qubits=list(qbits))
            if params:
                ins.params = list(params)
            if cbits:
                ins.memory = list(cbits)
            instrs.append(ins)
        cfg = _Model(memory_slots=len(clbit_labels), n_qubits=len(qubit_labels))
        if data is not None:
            cfg.data = list(data)
        header = _Model(name=qc.name, qubit_labels=qubit_labels, clbit_labels=clbit_labels,
                        creg_sizes=[[r.name, r.size] for r in qc.cregs],
                        qreg_sizes=[[r.name, r.size] for r in qc.qregs],
                        memory_slots=len(clbit_labels), n_qubits=len(qubit_labels))
        exps.append(QobjExperiment(instructions=instrs, header=header, config=cfg))
    config = _Model(shots=int(shots), memory_slots=max([e.config.memory_slots for e in exps] or [0]))
    if seed is not None:
        config.seed = int(seed)
    return Qobj(qobj_id=qobj_id or str(uuid.uuid4()), config=config, experiments=exps, header=_Model(), type='QASM',
                schema_version='1.1.0')


def execute(circuits, backend, shots=1024, seed=None, **kwargs):
    """qiskit.execute stand-in: assemble and run.  (The reference relies on qiskit's transpiler to
    unroll to basis_gates; this engine expands every qelib1 gate itself, so nothing to unroll.)"""
    seed = kwargs.pop('seed_simulator', seed)
    return backend.run(assemble(circuits, shots=shots, seed=seed, data=kwargs.pop('data', None)))
