cd $GRAFT_REPO_ROOT
export PYTHONFAULTHANDLER=1
export ITSOLV_BACKTRACE=1
python -X faulthandler - <<'PY' > gpurun_out/diag1.log 2>&1
import sys
sys.path[:0]=['.', 'tests']
print("import", flush=True)
import torch, numpy as np
import iterative_solver_b200 as pkg
from iterative_solver_b200 import _native as N, harness as H
print("context", flush=True)
ctx = pkg.Context(0)
ctx.init_comm(0, 1, b"\0"*128)
print("gemm_inner", flush=True)
x = [torch.randn(10000, dtype=torch.float64, device="cuda") for _ in range(4)]
print(ctx.gemm_inner(x, x)[0], flush=True)
print("dot", ctx.dot(x[0], x[1]), flush=True)
for fused in (0, 1):
    for kind in (N.KIND_DAVIDSON, N.KIND_LINEQ, N.KIND_DIIS):
        print("solve kind", kind, "fused", fused, flush=True)
        res, _ = H.solve(ctx, H.make_spec(20000, kind=kind, nroots=4, hermitian=1, fused=fused))
        print("  ->", res.converged, res.iterations, flush=True)
print("done", flush=True)
PY
tail -40 gpurun_out/diag1.log
python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/diag_pytest.log 2>&1; tail -30 gpurun_out/diag_pytest.log
