"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU (numpy/scipy) restatement of the exact-GP hot path of mwhoffman/pygp,
used as the checker for the CUDA path.  Nothing under ``pygp_b200/`` imports
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may.

Parity pin: ``oracle/make_golden.py`` runs the UNMODIFIED reference (imported
from /root/reference through ``oracle/ref_loader.py``) and this restatement on
the same inputs, asserts they agree, and writes the reference's outputs to
``tests/golden/``; ``tests/test_oracle.py`` re-checks the restatement against
those committed fixtures and against the known answers of SURVEY.md section 8c.
"""

from .pygp_oracle import *  # noqa: F401,F403
