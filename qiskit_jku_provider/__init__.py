"""Drop-in import name of the reference package (``from qiskit_jku_provider import JKUProvider,
QasmSimulator``), served by the B200-native engine in ``qiskit_jku_provider_b200``."""

from qiskit_jku_provider_b200.jkusimulatorerror import JKUSimulatorError  # noqa: F401


def __getattr__(name):
    import qiskit_jku_provider_b200 as _impl
    return getattr(_impl, name)
