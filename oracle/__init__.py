"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the AttentionShift hot path.

Nothing in ``attentionshift_b200/`` (the product) may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker or the timed CPU arm.

Two layers:

* ``oracle.ref_loader``  -- executes the *unmodified* reference functions straight
  out of ``/root/reference`` (AST extraction, last-definition-wins).  It only works
  where ``/root/reference`` exists (the build container); it is used to pin the
  restatement and to generate ``tests/golden/*.pt``.
* ``oracle.vit`` / ``oracle.attnshift`` -- a self-contained torch-CPU fp32
  restatement of the same algorithms (travels to the GPU box).  Every function
  cites the reference file:line it follows.

Parity status: the reference ships NO golden vectors or tests for this path
(SURVEY.md section 4), so the restatement is pinned against outputs of the reference
itself run in the build container (tests/golden/make_golden.py, tests/test_oracle_*).
The single exception is connected-component labelling (cc_torch): its source is
absent from the reference tree, so that one stage is "parity unpinned"
(8-connectivity assumed, see oracle/attnshift.py::ccl_label).
"""
