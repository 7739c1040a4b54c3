import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import iterative_solver_b200 as pkg
from iterative_solver_b200 import _native as N, harness as H
import itsolv_oracle_lib
kw = dict(n=30000, nroots=16, max_size_qspace=8, nbuffers=8, hermitian=1)
mode = sys.argv[1]
if mode == "ref":
    r, _ = itsolv_oracle_lib.load().ref.solve(H.make_spec(kind=N.KIND_DAVIDSON, **kw))
else:
    ctx = pkg.Context(0); ctx.init_comm(0, 1, b"\0" * 128)
    r, _ = H.solve(ctx, H.make_spec(kind=N.KIND_DAVIDSON, fused=int(mode), **kw))
print(mode, os.environ.get("ITSOLV_REMEASURE_FROM"), r.iterations, r.r_creations, r.q_creations, r.d_creations, " ".join(f"{r.errors[i]:.2e}" for i in range(16)))
