"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy/scipy) of the per-pixel statistical pipeline behind
``HiC3DeFDR.run_to_qvalues()`` of thomasgilgenast/hic3defdr v0.2.1.  It is the
checker the CUDA path is compared with; it is never the product.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  ``hic3defdr_b200`` never does.

Pinning status (see DESIGN.md "Oracle"):
  * fit_mu_hat, conditional_mor, sparse_union: pinned by the reference's own
    doctest vectors (tests/golden/reference_kats.json) AND by outputs of the
    real reference run in the build container (tests/golden/*.npz, made by
    tests/golden/make_golden.py through oracle/refrun.py).
  * qcml/cml/q2qnbinom/lrt/disp_idx/load_bias: pinned by outputs of the real
    reference only (the reference has no tests for them).
  * BH (lib5c.adjust_pvalues -> statsmodels fdr_bh) and lowess (lib5c port of
    statsmodels _lowess.pyx): lib5c 0.6.0 / statsmodels 0.10.2 are not in the
    image and not under /root/reference -> restated from the published
    algorithms: PARITY UNPINNED for those two functions.
"""
