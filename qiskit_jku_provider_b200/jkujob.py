"""``JKUJob``: the future-like handle ``QasmSimulator.run`` returns (surface of ``qiskit_jku_provider/jkujob.py:21-131``:
``submit / result(timeout) / cancel / status / backend / qobj``, ``JobError`` on a second submit or on use before
submit).

One deliberate difference: the work runs on a single worker THREAD on every platform.  The reference forks a
process pool on Linux (``jkujob.py:47-50``), which is not an option once the parent holds a CUDA context; ctypes
releases the GIL for the duration of the engine call, so a thread gives the same concurrency.
"""

from concurrent.futures import ThreadPoolExecutor

from .compat import BaseJob, JobError, JobStatus, validate_qobj_against_schema

_WORKER = ThreadPoolExecutor(max_workers=1, thread_name_prefix="jkq-job")


class JKUJob(BaseJob):
    """One submitted Qobj."""

    _executor = _WORKER  # (class attribute kept for code that swaps the executor, as with the reference)

    def __init__(self, backend, job_id, fn, qobj):
        super().__init__(backend, job_id)
        self._fn, self._qobj, self._future = fn, qobj, None

    # -- life cycle
    def submit(self):
        if self._future is not None:
            raise JobError("We have already submitted the job!")
        validate_qobj_against_schema(self._qobj)
        self._future = self._executor.submit(self._fn, self._job_id, self._qobj)

    def _submitted(self):
        if self._future is None:
            raise JobError("Job not submitted yet!. You have to .submit() first!")
        return self._future

    def result(self, timeout=None):
        return self._submitted().result(timeout=timeout)

    def cancel(self):
        return self._submitted().cancel()

    def status(self):
        future = self._submitted()
        # running / cancelled / done are tested in this order: a finished future also answers False to running()
        if future.running():
            return JobStatus.RUNNING
        if future.cancelled():
            return JobStatus.CANCELLED
        if future.done():
            return JobStatus.ERROR if future.exception() is not None else JobStatus.DONE
        return JobStatus.QUEUED  # accepted by the executor, not started yet

    # -- accessors
    def backend(self):
        return self._backend

    def qobj(self):
        return self._qobj
