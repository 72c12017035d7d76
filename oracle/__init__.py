"""ORACLE — test infrastructure, not product code.

CPU restatements (numpy) of the reference's MANO / forward-kinematics / MPJPE
path.  May be imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` — as the checker or
the CPU baseline, never as the thing shipped.  Parity pin: outputs of the
unmodified reference captured in ``tests/golden/`` (the reference has no golden
vectors of its own).
"""
