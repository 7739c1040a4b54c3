cd $GRAFT_REPO_ROOT
export ITSOLV_BACKTRACE=1
python - <<'PY' 2>&1 | tail -60
import sys, os
sys.path[:0]=['.','tools','tests']
import numpy as np, torch
import iterative_solver_b200 as pkg, config_runs as CR
from iterative_solver_b200 import _native as N, harness as H
ctx = pkg.Context(0); ctx.init_comm(0,1,b"\0"*128)
n=200000
def run(tag, fused, env):
    for k,v in env.items(): os.environ[k]=v
    spec=H.make_spec(n, kind=N.KIND_DAVIDSON, hermitian=1, nroots=16, max_size_qspace=8, nbuffers=8, fused=fused)
    res,sol=H.solve(ctx,spec,want_solutions=True)
    for k in env: os.environ.pop(k)
    out=[]
    for k in range(16):
        x=torch.from_numpy(sol[k]).cuda(); y=CR.torch_banded_apply(x,n,0,0,1); r=y-res.eigenvalues[k]*x
        out.append(float(r.norm()/x.norm()))
    print(tag,'it',res.iterations,'err %.1e'%max(res.errors[i] for i in range(16)),'true max %.1e'%max(out),'creations',res.r_creations,res.q_creations,res.d_creations, flush=True)
run('unfused',0,{})
run('fused',1,{})
run('fused2',2,{})
PY
