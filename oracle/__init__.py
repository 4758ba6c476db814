"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the CLIP-Lite JSD estimator.

Nothing under ``oracle/`` is part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and there only as the checker or as the timed CPU
baseline -- never as a fallback for the CUDA extension.

Parity status: the reference ships no tests or golden vectors for this path
(SURVEY.md section 8c).  The oracle is therefore pinned against outputs of the
reference's own ``loss.py`` executed in the build container; the vectors and
the script that made them live in ``tests/golden/``.
"""
