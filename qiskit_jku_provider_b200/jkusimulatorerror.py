"""Exception for errors raised by the JKU simulator (mirror of the reference's
``qiskit_jku_provider/jkusimulatorerror.py:16-26``; derives from QiskitError when qiskit is there)."""

try:  # pragma: no cover - qiskit is optional
    from qiskit import QiskitError as _Base
except Exception:  # qiskit absent: plain Exception keeps the same surface
    _Base = Exception


class JKUSimulatorError(_Base):
    """Class for errors raised by the JKU simulator."""

    def __init__(self, *message):
        """Set the error message."""
        super().__init__(*message)
        self.message = ' '.join(str(m) for m in message)

    def __str__(self):
        """Return the message."""
        return repr(self.message)
