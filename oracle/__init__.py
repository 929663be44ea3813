"""CPU oracle for the dnpy hot path (TEST INFRASTRUCTURE ONLY).

This package is a NumPy float64 restatement of the reference's layer
forward/backward/optimizer arithmetic (``/root/reference/dnpy``).  It exists so
that parity tests, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` have something to check the CUDA path
against on a GPU box where ``/root/reference`` does not exist.

Nothing under ``dnpy_b200/`` may import it: the product path is CUDA-only and
fails loudly when the extension is missing.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference in the build container, records the reference's own unit-test vectors
(``tests/coverage/test_layers.py`` / ``test_losses.py``) plus seeded
differential cases for every row of SURVEY.md §8a, and ``tests/test_oracle.py``
checks every oracle function against those fixtures.
"""
