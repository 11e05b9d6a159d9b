"""CPU oracle for the STCN/MiVOS space-time memory read (TEST INFRASTRUCTURE ONLY).

Everything under ``oracle/`` is a checker: a CPU restatement of the reference's
algorithm for the hot path (numpy fp64 for the tie-aware checker, torch-CPU fp32
for the op-for-op port that is timed as the CPU baseline).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product package ``evavos_b200`` never does; it fails
loudly when its CUDA library is missing.

Parity pinning: the reference ships NO tests, golden vectors or known-answer
fixtures for this path.  The oracle is therefore pinned
against outputs of the unmodified reference itself, imported from
``/root/reference`` in the build container by ``oracle/make_golden.py``; the
resulting vectors are committed under ``tests/golden/`` and
``tests/test_oracle_golden.py`` checks the oracle against them.
"""
