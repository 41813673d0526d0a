"""CPU oracle for the WISKI online-update hot path of wjmaddox/online_gp.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and
there only as the checker / the CPU arm, never as the thing measured or shipped.  The product package
``online_gp_b200`` never imports this package and fails loudly when its CUDA library is missing.

Parity status: the reference is pure Python on top of GPyTorch/BoTorch, neither of which is installed (or
installable: no network, no wheels) in the build container, so the reference itself cannot be run here and it
ships no golden output vectors.  At the GPyTorch-version level this oracle is therefore
**parity unpinned**.  What it *is* pinned against: the dense exact-GP identity that the reference's single live
numerical test asserts (``tests/mlls/test_batched_woodbury_marginal_log_likelihood.py:55-73``: WISKI MLL and
hyper-parameter gradients == exact GP on the SKI kernel), evaluated independently in fp64 by
``oracle.exact_gp`` on the fixtures of the reference's own tests; the frozen outputs live
in ``tests/golden/`` together with the generating script ``oracle/make_golden.py``.

Modules
  interp      - GPyTorch ``create_grid`` / ``Interpolation.interpolate`` / ``left_interp`` restated
  interp_np   - the same stencils a second time, numpy scalar loops, no torch (independent check of the bit-exact indices)
  gridkernel  - per-dimension Toeplitz columns of K_uu and the Kronecker-Toeplitz MVM
  exact_gp    - dense exact GP on the SKI kernel  W K W^T + sigma^2 D   (the independent check)
  wiski_ref   - reference-literal restatement (dense W^T D^-1 W, Cholesky dispatch, SVD root update)
  wiski_matfree - matrix-free restatement (m x r panels) used as the CPU baseline at north-star grid sizes
"""
