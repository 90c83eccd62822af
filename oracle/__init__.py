"""CPU oracle for the EgoNet per-crop inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs (``cpu_baseline`` and ``--impl reference``) may import it, and there only
as the checker / the timed CPU baseline -- never as a fallback of the CUDA path.

Every function restates one piece of the reference's algorithm (file:line
cited in its docstring, paths relative to the upstream repository root) with
plain numpy / torch-CPU arithmetic.

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against the reference's own
code *executed in the build container* -- ``tests/golden/make_golden.py``
imports the upstream modules from /root/reference, runs them on seeded inputs
and stores their outputs under ``tests/golden/*.npz``; ``tests/test_oracle_*``
replays those vectors against this package on any machine.
"""
