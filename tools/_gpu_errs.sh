cd "${GRAFT_REPO_ROOT:-/root/repo}"
python tools/_errs.py ref
python tools/_errs.py 0
ITSOLV_REMEASURE_FROM=1 python tools/_errs.py 1
ITSOLV_REMEASURE_FROM=2 python tools/_errs.py 1
ITSOLV_REMEASURE_FROM=99 python tools/_errs.py 1
