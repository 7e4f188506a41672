"""CPU oracle for the SAC minibatch gradient step (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker.  The product path (``deep-rl-grasping_b200``) never imports this package.

PARITY UNPINNED: the reference's learner is stable-baselines==2.10.1 on
tensorflow==1.14 (``/root/reference/setup.py:7-8``); neither is in the reference tree nor
installable here, and the reference holds no golden vectors for this path
(``tests_gripper/test_sim.py`` never builds a model).  The oracle restates the published SB2 SAC
algorithm and is pinned only by artefacts: variable names/shapes and
hyper-parameters in ``trained_models/**.zip``, the normalisation statistics in
``vecnormalize.pkl`` and early-training scalars in ``logs.csv`` (tests/test_oracle_pins.py).
"""
