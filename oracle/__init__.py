"""ORACLE package -- CPU checker for the B200 visual-odometry hot path.  Test infrastructure only:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it;
nothing under ``droplet_visual_odometry_b200/`` does.

Parity status: PINNED to outputs of the reference run in the build container.  The reference repository holds no tests, fixtures
or golden vectors for this path, so the pins are:
  (a) the reference's own module, imported unmodified from /root/reference/scripts/visual_odometry_v3.py by ``ref_loader`` (stub
      ROS / plotting modules only) and executed -- fixture ``tests/golden/golden_reference_module.npz`` written by
      ``tests/golden/make_reference_golden.py``, re-derived by tests/test_reference_module.py whenever the checkout is present;
  (b) the reference's third-party arithmetic engine, cv2 4.13.0, called through the reference's literal call sites
      (``cv2_chain``) -- live wherever cv2 is importable, and through ``tests/golden/golden_480x360.npz``
      (``tests/golden/make_golden.py``);
  (c) numpy / C++ restatements of what cv2 does (``orb_np``, ``pose_np``, ``ingest_np``, ``retain_best.cpp``), each checked
      against (b) by tests/test_oracle_golden.py.
"""
