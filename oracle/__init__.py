"""CPU oracle for the DRBA hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or
the CPU baseline.  Nothing under ``drba_b200/`` imports it.

* ``oracle.cport``  -- ctypes binding of ``drba_oracle.c`` (splat, DRM, backwarp,
  resize; plain C restatement of models/softsplat, models/drm.py, warplayer.py).
* ``oracle.ifnet``  -- torch-fp32 functional restatement of RIFE 4.26-heavy
  IFNet (models/rife_426_heavy/IFNet_HDv3.py) and the RIFE wrapper
  (models/rife.py) on top of the C splat.

Parity status: pinned against the reference's own outputs, see
tests/golden/make_golden.py and tests/test_oracle_golden.py.
"""
