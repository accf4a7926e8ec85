"""CPU oracle for the DPS hot path -- TEST INFRASTRUCTURE, not product code.

Everything under ``oracle/`` restates, on CPU in plain PyTorch fp32, the
algorithm of the reference's per-timestep posterior-sampling update so that the
CUDA path in ``samplers_b200`` can be checked against it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it; the product package never does (it raises when
its CUDA extension is missing instead of falling back to this code).

Pinning status:
  * identity / inpainting operators, Gaussian / Poisson noise, Tweedie, bridge
    statistics, the DPS loop: pinned against the *unmodified reference itself*,
    imported from /root/reference in the build container by
    ``oracle/ref_shim.py`` and recorded into ``tests/golden/*.npz`` by
    ``oracle/make_golden.py``.
  * Gaussian blur, motion blur and box super-resolution operators do not exist
    in the reference: for those rows parity is UNPINNED -- the oracle operator
    defined in ``oracle/operators.py`` is run *through the reference's own
    DPSSampler* (so every other line of the step is pinned) but the operator
    arithmetic itself has no reference counterpart.
"""
