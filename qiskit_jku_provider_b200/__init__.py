"""B200-native QMDD engine behind the JKUProvider / QasmSimulator API."""

from .jkusimulatorerror import JKUSimulatorError  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not need the GPU library
    if name in ("JKUProvider",):
        from .jkuprovider import JKUProvider
        return JKUProvider
    if name in ("QasmSimulator",):
        from .qasm_simulator_jku import QasmSimulator
        return QasmSimulator
    if name in ("JKUJob",):
        from .jkujob import JKUJob
        return JKUJob
    if name in ("Engine",):
        from .engine import Engine
        return Engine
    raise AttributeError(name)
