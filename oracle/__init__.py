"""CPU oracle for the LittleGAN hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (`/root/reference`, TensorFlow 1.15.4 / Keras)
ships no tests, golden vectors or fixtures, and TensorFlow cannot be imported
in this container (no cp312 wheel, no network).  Everything under `oracle/` is
a restatement of the reference's arithmetic in PyTorch-CPU / NumPy, following
`model.py`, `instance.py`, `eager_trainer.py:85-169,265-298`, `utils.py:47-56`
and `fid.py:112-188`; the TensorFlow semantics that are not visible in the
reference source (SAME padding, kernel layouts, Keras BCE, TF-1.x Adam) are
restated from TF 1.15's published behaviour and cross-checked here by
definition-level loops and adjoint identities (tests/test_oracle.py).  The FID
statistics / Frechet distance rows are the exception: NumPy / SciPy are
installed, so `oracle.fid_oracle` *is* the reference's own arithmetic.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package, and only as the checker.
Nothing under `littlegan_b200/` imports it.
"""
