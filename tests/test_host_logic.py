"""Host-side logic that needs no GPU: sharding, the candidate merge of the distributed select, the packing of P-space
maps, the host CSR generator, and the N>1 plumbing over gloo with world_size 2."""
import os
import subprocess
import sys
import textwrap

import numpy as np

import iterative_solver_b200 as pkg
from iterative_solver_b200 import harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_distribution_matches_reference_rule(oracle):
    """reference array/util/Distribution.h:99-110: block = n / P, first n % P chunks one longer"""
    for n, p in ((10, 3), (0, 4), (7, 8), (2_000_000_000, 8), (1_000_000_007, 5)):
        got = pkg.distribution(n, p)
        assert np.array_equal(got, oracle.c.distribution(n, p))
        sizes = np.diff(got)
        assert got[0] == 0 and got[-1] == n and sizes.max() - sizes.min() <= 1 and np.all(np.diff(sizes) <= 0)


def test_select_merge_is_the_reference_order(oracle):
    """merging per-shard candidates gives what the reference's heap gives on the whole vector"""
    rng = np.random.default_rng(4)
    x = np.round(rng.standard_normal(1000), 1)
    for kw in ({}, {"max": True}, {"max": True, "ignore_sign": True}, {"ignore_sign": True}):
        nsel = 17
        want_idx, want_val = oracle.c.select(x, nsel, **kw)
        cidx, cval = [], []
        for lo, hi in ((0, 334), (334, 667), (667, 1000)):  # three ranks
            i, v = oracle.c.select(x[lo:hi].copy(), nsel, **kw)
            cidx += (i + lo).tolist()
            cval += v.tolist()
        cidx += [-1, -1]  # empty candidate slots of a short shard
        cval += [0.0, 0.0]
        gi, gv = pkg.select_merge(np.array(cidx), np.array(cval), nsel, **kw)
        assert np.array_equal(gi, want_idx) and np.array_equal(gv, want_val)


def test_pack_maps_and_host_csr():
    ptr, idx, val = H.pack_maps([{5: 1.0, 2: -1.0}, {}, {9: 3.0}])
    assert ptr.tolist() == [0, 2, 2, 3] and idx.tolist() == [2, 5, 9] and val.tolist() == [-1.0, 1.0, 3.0]
    row_ptr, col, v, diag = H.banded_csr_host(10, 2, 1e-3)
    assert row_ptr[-1] == col.size == v.size == 10 * 5 - 2 * 3
    assert diag.tolist() == list(range(1, 11))
    dense = np.zeros((10, 10))
    for r in range(10):
        dense[r, col[row_ptr[r]:row_ptr[r + 1]]] = v[row_ptr[r]:row_ptr[r + 1]]
    assert np.array_equal(dense, dense.T) and dense[3, 4] == 1e-3 * (1 + (7 % 7)) and dense[3, 5] == 1e-3 * (1 + 8 % 7)
    # a shard of rows equals the corresponding rows of the whole operator
    rp2, c2, v2, d2 = H.banded_csr_host(10, 2, 1e-3, 4, 8)
    assert np.array_equal(c2, col[row_ptr[4]:row_ptr[8]]) and np.array_equal(v2, v[row_ptr[4]:row_ptr[8]])


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    import iterative_solver_b200 as pkg
    from iterative_solver_b200 import distributed as D
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank = dist.get_rank()
    # 1. the communicator id travels from rank 0 to every rank unchanged
    payload = bytes(range(128)) if rank == 0 else None
    got = D.broadcast_bytes(payload, 128, 0)
    assert got == bytes(range(128))
    # 2. every rank derives the same sharding and owns a disjoint, covering range
    n = 1001
    b = pkg.distribution(n, 2)
    lo, hi = int(b[rank]), int(b[rank + 1])
    owned = torch.zeros(n, dtype=torch.int64); owned[lo:hi] = 1
    dist.all_reduce(owned)
    assert bool((owned == 1).all())
    # 3. sharded select: local candidates gathered and merged give the global answer on both ranks
    x = np.round(np.random.default_rng(0).standard_normal(n), 1)
    order = sorted(range(lo, hi), key=lambda i: (-x[i], i), reverse=True)[:5]
    cand = torch.tensor([[float(i), x[i]] for i in order], dtype=torch.float64)
    allc = [torch.zeros_like(cand) for _ in range(2)]
    dist.all_gather(allc, cand)
    allc = torch.cat(allc).numpy()
    gi, gv = pkg.select_merge(allc[:, 0].astype(np.int64), allc[:, 1].copy(), 5)
    want = sorted(sorted(range(n), key=lambda i: (-x[i], i), reverse=True)[:5])
    assert gi.tolist() == want, (gi.tolist(), want)
    # 4. bench aggregation: whole-job value = units of all ranks / max over ranks of the time
    t = torch.tensor([0.5 + 0.25 * rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == 0.75
    dist.barrier()
    print("rank", rank, "ok")
""")


def test_two_rank_plumbing_over_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_the_configurations_a_bench_run_covers_fit_their_gpus():
    """bench.config_plan: which of BASELINE.json's configurations run at N GPUs, against the measured high-water marks
    (profiles/memory_table_r02.jsonl, vectors of 8 n / N bytes, 180 GB per GPU with 6 GB kept free as run_configs does)"""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import config_runs as CR
    peaks = {}
    for line in open(os.path.join(ROOT, "profiles", "memory_table_r02.jsonl")):
        r = json.loads(line)
        peaks[(r["nroots"], r["nbuffers"], r["max_size_qspace"])] = r["peak_vectors"]
    for world in (1, 2, 4, 8):
        plan = dict(bench.config_plan(world))
        assert ("c4" in plan) == (world > 1) and ("c3" in plan) == (world <= 2)
        if "c4" in plan:
            base = CR.CONFIGS["c4"]
            shape = {**base, **plan["c4"]}
            key = (shape["nroots"], shape.get("nbuffers", shape["nroots"]), shape["max_size_qspace"])
            need = peaks[key] * 8.0 * base["n"] / world
            assert need + 6e9 <= 180e9, (world, key, need)
            assert world < 8 or plan["c4"] == {}  # the stated shape on 8 GPUs
