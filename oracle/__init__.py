"""CPU oracle for the GP-C-OCBA decision + PCS hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (pure PyTorch / numpy) of the reference algorithm for the
one hot path this repository accelerates.  It is *the checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  Nothing under ``contextual_rs_b200/`` imports it and the product path has no CPU
fallback: it fails loudly when the CUDA library is missing.

Parity status.  The reference's arithmetic for ``model.posterior`` / ``posterior.rsample`` lives in
BoTorch (>=0.4) / GPyTorch (>=1.4) (``/root/reference/setup.py:3-11``, not version-pinned, no lock
file), neither of which is vendored in the reference tree nor importable in this image, and no
reference TEST pins a numeric posterior value.  ``gp_oracle.py`` therefore restates the *published*
exact-GP algorithm of those libraries (SingleTaskGP = Normalize -> ScaleKernel(Matern-5/2 ARD | RBF
ARD) + ConstantMean + homoskedastic noise -> Standardize), anchored on the reference's own call
sites — and the whole chain (transforms, fit, posterior, zeta table, OCBA step 2(b)) is pinned on
OUTPUTS OF THE REFERENCE ITSELF: the recorded run ``experiments/wsc_experiments/config_b_3_mean/
0000_ML_Gao.pt`` is replayed by ``trace_replay.py`` and 90 of its first 100 recorded decisions are
reproduced from re-fitted hyper-parameters (fixture ``tests/golden/reference_trace_ml_gao.json``).
Still unpinned: bit-level agreement of posterior VALUES with GPyTorch (not observable here).

What else is pinned (see ``tests/test_oracle_golden.py``):
  * the decision logic (sort / zeta / argmin / OCBA ratio) against the mock posterior tables of
    ``test/test_contextual_rs_strategies.py:169-176`` and the ``(arm 2, context 1)`` known answer
    of ``test/test_contextual_rs_strategies.py:22-62``;
  * the Philox4x32-10 stream against the Random123 known-answer vectors;
  * the PCS estimator against exact Gauss-Hermite quadrature and the [0,1] / monotonicity
    properties of ``test/test_generalized_pcs.py:242-329``.
"""
