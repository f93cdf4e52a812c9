"""``JKUProvider``: the entry point a user of the reference already has in their code.

Same surface as ``qiskit_jku_provider/jkuprovider.py:18-45`` -- ``get_backend(name)``, ``backends()``,
``available_backends(filters)``, ``str(provider) == 'JKUProvider'`` -- over one backend, ``qasm_simulator``, whose
engine is the CUDA library (no executable is looked up; a missing ``libjkq_b200.so`` raises on construction).
"""

from .compat import BaseProvider
from .qasm_simulator_jku import QasmSimulator


def _matches(backend, filters):
    """every (key, value) of ``filters`` equals the backend's configuration entry"""
    config = backend.configuration()
    return all(config.get(key) == wanted for key, wanted in (filters or {}).items())


class JKUProvider(BaseProvider):
    """Provider for the local JKU backend."""

    BACKEND_NAMES = ('qasm_simulator',)

    def __init__(self, *args, **kwargs):
        super().__init__(args, kwargs)
        simulator = QasmSimulator(provider=self)
        assert simulator.name() in self.BACKEND_NAMES
        self.backends_list = {simulator.name(): simulator}  # (attribute name kept: callers of the reference read it)

    def __str__(self):
        return 'JKUProvider'

    def backends(self, name=None, **kwargs):
        # like the reference (:41-42), the name and filter arguments are accepted and ignored
        return [self.backends_list[key] for key in self.BACKEND_NAMES]

    def available_backends(self, filters=None):
        return [backend for backend in self.backends() if _matches(backend, filters)]

    def get_backend(self, name, **kwargs):
        """KeyError for anything but 'qasm_simulator', as in the reference (:28-29)"""
        return self.backends_list[name]
