"""CPU oracle for the replay-sample -> learn_from_batch hot path of IntelLabs/coach.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import,
call, link or execute it, and there only as the checker (or as the CPU baseline being *reported*), never as the
thing measured or shipped.  ``coach_b200`` never imports this package.

Contents
--------
* ``ref_loader``      -- imports the *unmodified* reference (``/root/reference/rl_coach``) with ``tensorflow`` and
                         ``redis`` stubbed in ``sys.modules``.  Works only in the build container (the reference tree
                         does not exist on the GPU box); used to pin the restatement and to generate
                         ``tests/golden/*.npz`` (``oracle/make_golden.py``).
* ``segment_tree.c``  -- plain-C restatement of ``SegmentTree`` / ``PrioritizedExperienceReplay`` arithmetic (fp64,
                         glibc ``pow``) so that 2^20-leaf cases finish in seconds.
* ``memory.py``       -- Python restatement of ExperienceReplay / PrioritizedExperienceReplay / Batch gather,
                         written to follow the reference's per-sample control flow (it is also the timed "port"
                         CPU baseline, because that is how the reference runs: one Python thread).
* ``nets.py``         -- torch-CPU fp32 restatement of the TF-1.x graph semantics (embedders, heads, losses,
                         global-norm clipping, TF Adam, polyak) for DQN/DDQN/dueling, ClippedPPO, DDPG/TD3, SAC.
* ``rl_math.py``      -- GAE, n-step returns, running-stats normalisation (numpy fp64).

Parity status: memory / filters / GAE / n-step are PINNED against the reference's own code (imported here) and its
SegmentTree known-answer tests.  The neural-network arithmetic (``nets.py``) is **parity unpinned**: it restates
TensorFlow-1.14 semantics, TensorFlow is not installable here and the reference holds no unit test of any loss,
gradient or optimizer step.
"""
