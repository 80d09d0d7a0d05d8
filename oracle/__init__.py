"""CPU oracle for the artery.fe per-timestep network solve.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker.  The product
path (``bloodflow_b200``) never imports this package and has no CPU fallback.

Parity status (see DESIGN.md "Oracle"):

* The boundary-condition half of the path (flux, source, compute_U_half,
  windkessel, problem_function, jacobian, newton, adjust_bifurcation_step,
  CFL_term, adjust_dex, host set-up) is PINNED: ``tests/golden/make_golden.py``
  runs the reference's own Python code for those functions (imported from
  /root/reference under a minimal ``dolfin`` stand-in that only provides
  point-evaluated Expressions and P1 Functions) and the fixtures it wrote are
  checked against this oracle in ``tests/test_oracle_golden.py``.
* The per-artery variational solve lives in FEniCS 2017.2.0 (un-vendored third
  party, ``setup.py:56`` / ``Dockerfile:1`` of the reference), which cannot be
  installed here and for which the reference ships no golden values:
  **parity unpinned** for that half.  It is restated from the reference's UFL
  form (``arteryfe/artery.py:191-244``) and the published DOLFIN/FFC semantics
   and cross-checked by symbolic differentiation.
"""
