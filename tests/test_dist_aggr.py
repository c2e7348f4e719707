"""N > 1 host logic on CPU: world_size-2 gloo run of the sum/avg/min/max/count-by protocol:
rank-local partial {values, counts}[G x P] (here folded by the oracle, on the GPU by vmb_rollup_aggr_partial),
identity fill of empty cells, one all-reduce of values (operator per aggregate) and one of counts, finalize.
The result must equal the single-process fold over all series."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from rollup_names import AGGR


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fold(name, rolled, groups, G, series):
    P = rolled.shape[1]
    v = np.zeros((G, P))
    c = np.zeros((G, P))
    for s in series:
        row = np.ascontiguousarray(rolled[s])
        g = int(groups[s])
        O.lib().vmo_aggr_update(AGGR[name], v[g].ctypes.data_as(O.f64p), c[g].ctypes.data_as(O.f64p), row.ctypes.data_as(O.f64p), P)
    return v, c


def _finalize(name, v, c):
    for g in range(v.shape[0]):
        O.lib().vmo_aggr_finalize(AGGR[name], v[g].ctypes.data_as(O.f64p), c[g].ctypes.data_as(O.f64p), v.shape[1])
    return v


def _worker(rank, world, port, rolled, keys, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from victoriametrics_b200 import promql
    groups, G = promql.dense_group_ids(keys)  # same global key list on every rank => identical ids
    mine = promql.shard_series(len(keys), rank, world)
    res = {}
    for name in ("sum", "avg", "min", "max", "count", "sum2", "geomean"):
        v, c = _fold(name, rolled, groups, G, mine)
        if name not in ("count",):
            v[c == 0] = promql.ALLREDUCE_IDENTITY[name]  # == vmb_aggr_prepare_allreduce
        tv, tc = torch.from_numpy(v), torch.from_numpy(c)
        promql.torch_all_reduce(tv, tc, promql.ALLREDUCE_OP[name])
        res[name] = _finalize(name, tv.numpy(), tc.numpy())
    if rank == 0:
        out_q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_sum_by_protocol_world2_gloo():
    rng = np.random.default_rng(5)
    S, P = 41, 17
    rolled = np.round(rng.normal(10, 5, (S, P)), 3)
    rolled[rng.random((S, P)) < 0.3] = np.nan
    rolled[:, 3] = np.nan  # a point nobody has
    keys = ["mode=%d" % (i % 5) for i in range(S)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, rolled, keys, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from victoriametrics_b200 import promql
    groups, G = promql.dense_group_ids(keys)
    for name, got in res.items():
        v, c = _fold(name, rolled, groups, G, range(S))
        exp = _finalize(name, v, c)
        assert np.array_equal(np.isnan(got), np.isnan(exp)), name
        assert np.allclose(got, exp, rtol=1e-12, atol=0, equal_nan=True), name


def test_shard_and_group_ids():
    from victoriametrics_b200 import promql
    parts = [promql.shard_series(10, r, 4) for r in range(4)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(10))
    ids, g = promql.dense_group_ids(["a", "b", "a", "c", "b"])
    assert ids.tolist() == [0, 1, 0, 2, 1] and g == 3
