"""CPU oracle for the MMaDA masked-diffusion denoising path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product package ``mmada_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it, and there only
as the checker (or as the timed CPU baseline), never as the thing shipped.

The oracle is a from-scratch PyTorch-on-CPU restatement of the reference's
algorithm for the hot path.  The reference is
pure Python/PyTorch with no tests and no golden vectors of its own, so the
oracle is pinned the only way available: ``oracle/make_goldens.py`` imports the
REAL reference from ``/root/reference`` in the build container, runs it on
synthetic weights produced by ``oracle/weights.py`` and

  * asserts the restatement reproduces the reference bit-for-bit, and
  * writes the reference's own outputs to ``tests/golden/*.npz``.

On the GPU box ``/root/reference`` is absent: tests compare the CUDA path with
this restatement and with the committed goldens.
"""
