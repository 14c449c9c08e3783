"""CPU oracle for the CenterNet heatmap hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``centernet_b200/`` may import this
package: it exists so that ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` can check (and
time) the hand-written CUDA path against an independent restatement of the
reference algorithm.

Parity status: PINNED.  The reference ships no golden vectors for this path
(SURVEY.md section 8c), so the oracle is pinned against outputs of the
unmodified reference functions imported from ``/root/reference/src/lib`` in the
build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``,
checked by ``tests/test_oracle_golden.py``) plus the DCNv2 known-answer test of
``DCNv2/test.py:32-65``.
"""
