#!/usr/bin/env python
"""The dense Hamiltonians the reference's end-to-end eigensolver test reads (reference examples/*.hamiltonian: first the
dimension, then the n x n elements row by row; test/itsolv/test_LinearEigensystem.cpp:52-59, file_eigen :346-351) as one
small fixture, tests/golden/hamiltonians.npz, so that tests can run where /root/reference does not exist.

    python tests/golden/make_hamiltonian_fixtures.py        (in the build container, where /root/reference is mounted)
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("ITSOLV_REFERENCE", "/root/reference")

out = {}
for name in ("he", "hf", "bh"):  # phenol.hamiltonian is not part of the checkout
    tokens = open(os.path.join(REFERENCE, "examples", name + ".hamiltonian")).read().split()
    n = int(tokens[0])
    out[name] = np.array([float(t) for t in tokens[1:1 + n * n]]).reshape(n, n)
    assert len(tokens) >= 1 + n * n
np.savez_compressed(os.path.join(HERE, "hamiltonians.npz"), **out)
print({k: v.shape for k, v in out.items()})
