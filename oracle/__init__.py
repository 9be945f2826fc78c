"""CPU oracle for the super-resolution forward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline -- never as the thing shipped.  The product package
(``mobilesuperresolution_b200``) must not import this package.

Two restatements live here:

* ``oracle.port``      -- the reference's op sequence restated with
  ``torch.nn.functional`` on the CPU (the reference *is* PyTorch, so this is
  the faithful "port" that is timed as the CPU baseline).  Pinned bit-exactly
  against the imported reference modules by ``oracle/make_golden.py`` and by
  ``tests/test_oracle_vs_reference.py`` (which runs wherever ``/root/reference``
  exists).
* ``oracle.c_oracle``  -- an independent plain-C restatement (direct loops,
  float64 accumulation, no torch) in ``oracle/csrc/oracle_sr.c``.  Pinned
  against the committed golden vectors in ``tests/golden/``.

Parity status: PINNED -- against golden vectors generated from the imported
reference (``tests/golden/*.npz`` + ``kat.json``; generator committed as
``oracle/make_golden.py``) and against SURVEY.md Appendix D known answers.
The mmedit boundary (``mmedit.flow_warp`` / ``mmedit.SPyNet``) is NOT in
``/root/reference`` (un-vendored, unpinned dependency): parity there is
anchored on the in-repo twin ``models/spynet_arch.py`` only.
"""
