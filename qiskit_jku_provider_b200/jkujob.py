"""Job class of the backend (mirror of the reference's ``qiskit_jku_provider/jkujob.py:21-131``).

The one deliberate difference: the executor is a ThreadPoolExecutor on every platform.  The reference
uses a ProcessPoolExecutor on Linux (``jkujob.py:47-50``), i.e. fork -- which is unsafe once a CUDA
context exists in the parent; ctypes releases the GIL during the engine call, so a thread suffices.
"""

import functools
from concurrent import futures

from .compat import BaseJob, JobError, JobStatus, validate_qobj_against_schema


def requires_submit(func):
    """Decorator to ensure that a submit has been performed before calling the method."""
    @functools.wraps(func)
    def _wrapper(self, *args, **kwargs):
        if self._future is None:
            raise JobError("Job not submitted yet!. You have to .submit() first!")
        return func(self, *args, **kwargs)
    return _wrapper


class JKUJob(BaseJob):
    """JKU Job class."""

    _executor = futures.ThreadPoolExecutor(max_workers=1)

    def __init__(self, backend, job_id, fn, qobj):
        super().__init__(backend, job_id)
        self._fn = fn
        self._qobj = qobj
        self._future = None

    def submit(self):
        """Submit the job to the backend for execution.  Raises JobError on a second submit."""
        if self._future is not None:
            raise JobError("We have already submitted the job!")
        validate_qobj_against_schema(self._qobj)
        self._future = self._executor.submit(self._fn, self._job_id, self._qobj)

    @requires_submit
    def result(self, timeout=None):
        return self._future.result(timeout=timeout)

    @requires_submit
    def cancel(self):
        return self._future.cancel()

    @requires_submit
    def status(self):
        # The order is important here
        if self._future.running():
            _status = JobStatus.RUNNING
        elif self._future.cancelled():
            _status = JobStatus.CANCELLED
        elif self._future.done():
            _status = JobStatus.DONE if self._future.exception() is None else JobStatus.ERROR
        elif self._future._state == 'PENDING':
            _status = JobStatus.QUEUED
        else:
            raise JobError('Unexpected behavior of {0}'.format(self.__class__.__name__))
        return _status

    def backend(self):
        return self._backend

    def qobj(self):
        return self._qobj
