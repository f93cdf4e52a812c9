"""ORACLE package -- test infrastructure, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  It wraps two CPU checkers:

* ``oracle/_ref/jku_ref``      the UNMODIFIED reference engine built from /root/reference/src
                               (``oracle/build_ref.sh``; MPFR 200-bit, single-threaded);
* ``oracle/_build/qmdd_oracle`` the plain-C FP64 restatement ``oracle/qmdd_oracle.c``.

Both read the same circuit text and print the same JSON (see ``oracle/ref_harness.cpp``).
"""

from .run import build, have_reference, run_oracle, run_reference  # noqa: F401
