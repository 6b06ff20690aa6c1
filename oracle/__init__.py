"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DNS-SLAM render-and-optimise hot path.

Nothing in the product package (``dns_slam_b200``) may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` use it, and only as the checker / the CPU baseline.

Parity status
-------------
* Everything the reference itself owns on this path (pixel / ray sampling, far plane,
  depth-guided z sampling, point build, feature matching, occupancy compositing, all
  losses, TV smoothness, quaternion maths) is PINNED: ``oracle/make_golden.py`` imports
  the reference's own Python from ``/root/reference`` and stores its outputs as golden
  vectors under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
  restatement against them.
* The arithmetic inside ``tinycudann`` (HashGrid, OneBlob, CutlassMLP) is a third-party,
  un-vendored, un-pinned dependency (reference ``requirements.txt:35``; README fallback
  commit 91ee479d275d322a65726435040fc20b56b9c991) that is absent from this machine.
  ``oracle/tcnn_standin.py`` restates its published algorithm in fp32 PyTorch.  The
  reference holds no test, fixture or golden vector at that boundary, so for those three
  operators the status is **parity unpinned**.
"""
