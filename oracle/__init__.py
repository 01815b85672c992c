"""CPU oracle for the matchgate-circuit contraction hot path of MatchCake.

TEST INFRASTRUCTURE ONLY.  This package is a float64 numpy restatement of the
reference's algorithm (``/root/reference/src/matchcake``); every function cites
the reference ``file:line`` it follows.  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` -- never by the product package ``matchcake_b200``.

Parity status: PINNED.  The reference itself cannot be imported in this image
(PennyLane, torch_pfaffian, autoray are absent), so the oracle is
pinned against every golden value the reference publishes for this path
(``tests/golden/reference_golden.json``: JOSS paper circuit, the probabilities
and expectation-values tutorials) and against the known-answer tests of the
reference's own test-suite (Pfaffian KATs, lookup-table KATs, Jordan-Wigner
table, GramMatrix index rules).  See ``tests/test_oracle_golden.py``.

Third-party arithmetic that is not under /root/reference:
``torchpfaffian==0.0.3`` (``pyproject.toml:36``).  Its two entry points used by
the path (``torch_pfaffian.pfaffian(sign=True)`` -- Parlett-Reid, and
``PfaffianDet`` -- ``sqrt(|det|)``) are restated from their published
algorithms in ``oracle/pfaffian.py``.
"""

from . import covariance, gates, kernels, lookup_table, pfaffian, readout  # noqa: F401
