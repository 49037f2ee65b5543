"""oracle/ — CPU checkers for the emission + Viterbi hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs (``cpu_baseline`` and
``--impl reference``) may import this package.  Nothing under ``exomedepth_b200/`` does: the product
path is CUDA-only and fails loudly without its extension.

* ``oracle.ref``   — the reference's own C/C++ compiled unmodified (``oracle/_ref/libexomedepth_ref.so``,
                     built in place from /root/reference/src by ``oracle/Makefile``), driven through
                     fake SEXPs exactly as R's ``.Call`` would.
* ``oracle.port``  — this repo's plain-C restatement (``oracle/oracle.c``), pinned against ``oracle.ref``
                     and against the known-answer vectors of SURVEY.md §8c; it also defines the
                     S>3 / forward / MLE extensions, for which no reference exists (parity unpinned).
* ``oracle.framing`` — numpy restatement of the R-side CallCNVs framing (R/class_definition.R:311-419).
"""
