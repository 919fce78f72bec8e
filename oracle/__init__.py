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

PARITY PINNED AGAINST EXECUTED UPSTREAM CODE, with one stated substitution.  The upstream repository holds no golden
vectors, fixtures or recorded logs, and none of Firedrake / petsc4py / MUMPS is installable in this image, so the
script cannot run as it is.  What runs (the generators read the upstream files at generation time, execute the lines
unmodified and store numerical outputs only):

* the two CLASSES of ``Code/Control_Wave_PC.py`` -- ``Optimal_Control_Wave_Equation`` (:13-179: ``__init__``,
  ``Build_f``, ``Build_g``, ``Build_Initial_Condition``, ``Build_L``) and ``DiagFFTPC`` (:376-558: ``initialize``,
  ``update``, ``apply``, ``applyTranspose``) -- executed with ``fd`` bound to ``tests/golden/firedrake_standin.py``,
  which supplies only the textbook part (P1 mass / stiffness on the uniform interval, UFL forms affine in their
  coefficient functions, homogeneous Dirichlet rows, sparse LU instead of MUMPS, the mixed-vector layout) and nothing
  that knows about the preconditioner: ``tests/golden/make_reference_executed_golden.py``.
  ``tests/test_reference_executed_golden.py`` pins every route of ``DiagFFTPC.apply`` here to the executed apply
  (2e-14 ... 1.5e-12), ``operator.AllAtOnce`` to the executed forms (right-hand side and matvec to 2e-16, the
  1/2-weight rows and the :138 quirk included), the direct baseline, and GMRES (same counts -- 5 at the upstream
  constants -- and residual histories to 1e-12 when ``gmres`` runs on the executed operator with the executed apply);
  ``tests/test_gpu_reference_executed.py`` does the same for the CUDA path;
* the eigen-set-up stage :387-436 on its own (Lambda_1, Lambda_2, the per-frequency ``eig`` / ``inv`` loop):
  ``tests/golden/make_reference_setup_golden.py``; ``eigs`` reproduces it bit for bit
  (``tests/test_reference_setup_golden.py``);
* the whole known-answer notebook ``Code/mat_test.ipynb`` (cells executed from its JSON) and the closed forms of
  ``Code/pre_cond.py:32-38``: ``tests/golden/make_reference_notebook_golden.py``,
  ``tests/test_reference_notebook_golden.py``.

Still UNPINNED (restated from documentation, checked only against itself): what the stand-in replaces -- Firedrake's
assembly of P1 forms and its boundary-condition handling, MUMPS -- and PETSc's KSPGMRES (``gmres`` restates the
semantics the options :347-359 select; the fixtures above run that restatement on executed operators).
"""
