"""CPU oracle for the ParaDiag block-circulant preconditioner hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product path
(``optimal_control_paradiag_b200``) never imports it and has no CPU fallback.

What it restates (all citations are into the read-only upstream checkout,
``Code/Control_Wave_PC.py`` unless another file is named):

* ``fem1d``        1-D uniform P1 mass / stiffness matrices (what Firedrake
                   assembles for ``UnitIntervalMesh`` + ``CG1``, :17, :33, :42).
* ``eigs``         circulant eigenvalues :387-388, the 2x2 blocks :418-419,
                   ``np.linalg.eig`` / ``inv`` :421-425 and the closed forms.
* ``pc_ref_route`` line-by-line restatement of ``DiagFFTPC.initialize`` /
                   ``apply`` :380-553 (eig per k, S^-1, shifted solves, S,
                   1/lambda_2, fft).
* ``pc_explicit``  the explicit sparse block-circulant matrix P, factorised by
                   SuperLU -- semantic ground truth at small sizes.
* ``pc_fast``      the division-free decoupled form (same operator): the large-size
                   checker (``stage_columns``: sampled frequencies at any size,
                   also in 80-bit) and, as ``apply_threaded`` (scipy.fft workers +
                   the fused pthread stage of ``csrc/pc_solve.c``), the timed CPU
                   baseline.
* ``pc_longdouble`` the same in 80-bit extended precision (conditioning).
* ``operator``     the all-at-once matrix of ``Build_L`` :86-179 and the
                   manufactured right-hand side of ``Build_f/g/IC`` :48-83.
* ``gmres``        PETSc-KSPGMRES semantics selected by the options :347-359
                   (``gmres_lean``: same iteration, growing 2-D basis, for the
                   BASELINE-size counts of ``tests/golden/gmres_counts.json``).
* ``pc_alpha``     the alpha EXTENSION (no upstream counterpart): explicit P_alpha,
                   per-frequency block LU, decoupled closed form.

PARITY UNPINNED, except for everything of the reference that runs without
Firedrake.  The upstream repository holds no golden vectors, fixtures or recorded
logs for PC-apply outputs or GMRES iteration counts, and none of Firedrake /
petsc4py / MUMPS is installable in this image, so ``DiagFFTPC.apply`` and the
Krylov solve themselves cannot be run.  Pinned against EXECUTED upstream code
(the generators read the upstream files at generation time, run the lines
unmodified and store numerical outputs only):

* the eigen-set-up stage :387-436 (Lambda_1, Lambda_2, the per-frequency
  ``eig`` / ``inv`` loop): ``tests/golden/make_reference_setup_golden.py``;
  ``eigs`` reproduces it bit for bit and every route agrees with the
  line-by-line route driven by those arrays
  (``tests/test_reference_setup_golden.py``);
* the whole known-answer notebook ``Code/mat_test.ipynb`` (cells executed from
  its JSON) and the closed forms of ``Code/pre_cond.py:32-38``:
  ``tests/golden/make_reference_notebook_golden.py``;
  ``tests/test_reference_notebook_golden.py`` pins ``eigs.lambdas``,
  ``eigs.closed_form``, the FFT / circulant conventions and the numpy eig route
  to them (``tests/test_oracle_notebook.py`` re-derives the same identities).

Unpinned (restated from the source, checked only against itself -- three
independent routes agreeing to ~1e-13 at small sizes): the P1 mass / stiffness
assembly, the shifted solves with Dirichlet rows, the operator / right-hand
side of ``Build_L`` / ``Build_f/g/IC`` and PETSc's GMRES semantics.
"""
