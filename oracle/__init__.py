"""CPU oracle for the delivery_drone hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``reinforcement-learning-101_b200/``) imports
this directory.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and
only as the checker / the CPU baseline -- never as the thing shipped.

Parity pinning: the reference has no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself: ``tests/golden/make_golden.py`` imports the unmodified reference
``DroneGame`` from /root/reference (pygame stubbed), records KAT1..KAT7 and a
randomised corpus, and commits them under ``tests/golden/``.
"""
