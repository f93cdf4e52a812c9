"""``JKUSimulatorError``: what the backend raises for anything the engine or the front end rejects.

Surface of ``qiskit_jku_provider/jkusimulatorerror.py:16-26``: built from message parts, ``.message`` is their
space-joined text and ``str()`` is its ``repr``.  A subclass of ``QiskitError`` when qiskit is importable, so
existing ``except QiskitError`` clauses keep catching it.
"""

try:  # pragma: no cover - qiskit is optional in this image
    from qiskit import QiskitError as _Parent
except Exception:
    _Parent = Exception


class JKUSimulatorError(_Parent):
    def __init__(self, *parts):
        super().__init__(*parts)
        self.message = ' '.join(map(str, parts))

    def __str__(self):
        return repr(self.message)
