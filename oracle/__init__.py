"""oracle/ — TEST INFRASTRUCTURE ONLY.

A CPU restatement (numpy / torch-CPU) of the reference's rollout-and-update hot
path, used as the checker for the CUDA path.  Nothing in the product package
``ia2c_b200`` may import from here; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs do.

Parity pin: the reference (thinclab/IA2C) ships no tests or golden vectors
(SURVEY.md §4).  The pins are therefore *outputs of the unmodified reference
itself*, executed in the build container under a record/replay tape by
``oracle/gen_golden.py`` and committed under ``tests/golden/``.  Every function
in this package is checked against those fixtures by ``tests/test_oracle_*.py``.
"""
