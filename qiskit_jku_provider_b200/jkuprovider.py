"""Provider for the local JKU backend (mirror of ``qiskit_jku_provider/jkuprovider.py:18-45``)."""

from .compat import BaseProvider
from .qasm_simulator_jku import QasmSimulator


class JKUProvider(BaseProvider):
    """Provider for the local JKU backend."""

    def __init__(self, *args, **kwargs):
        super().__init__(args, kwargs)
        self.backends_list = {'qasm_simulator': QasmSimulator(provider=self)}

    def get_backend(self, name, **kwargs):
        return self.backends_list[name]

    def available_backends(self, filters=None):
        backends = self.backends_list
        filters = filters or {}
        for key, value in filters.items():
            backends = {name: instance for name, instance in backends.items()
                        if instance.configuration().get(key) == value}
        return list(backends.values())

    def backends(self, name=None, **kwargs):
        return list(self.backends_list.values())

    def __str__(self):
        return 'JKUProvider'
