"""CPU oracle for the heatmap-codec hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  Nothing under
``infantposeestimation_gaussianbias_b200/`` imports it: the product path is the
CUDA library and fails loudly when that library is missing.
"""
