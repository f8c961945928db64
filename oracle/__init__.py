"""CPU oracle for the per-voxel T2 fit hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / reference
legs of ``bench.py`` may import it, and there only as the checker or the timed
CPU arm.  ``fetal_t2mapping_b200`` never imports this package and has no CPU
fallback for the fit.

Contents
--------
``fit_oracle.py``   numpy/scipy restatement of the reference hot path
                    (``fit_voxel``, ``set_fit_params``, the hot block of
                    ``process_t2maps`` and ``compute_residuals``), each function
                    citing the reference file:line it follows.  The optimiser
                    itself (L-BFGS-B + 2-point finite differences) lives in
                    the third-party dependency scipy (reference pin 1.11.3,
                    ``requirements_frozen.txt:144``; this image ships 1.18.1)
                    and is *called*, exactly as the reference calls it.
``ref_loader.py``   imports the UNMODIFIED reference from ``/root/reference``
                    (only exists in the build container) to pin the oracle and
                    to generate ``tests/golden/*.npz``.

Pinning status: the reference ships no tests or golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against outputs of the reference itself
run in the build container (``tests/golden/make_golden.py`` → committed
fixtures; ``tests/test_oracle_pinned.py``), including a second run of the
unmodified reference under a one-ulp jitter of ``np.exp`` that measures how well the
reference reproduces itself (``tests/golden/make_jitter.py``).
"""
