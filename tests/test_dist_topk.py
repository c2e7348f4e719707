"""N > 1 host logic on CPU for topk/bottomk ("per-GPU per-point top-k candidates -> all-gather -> final select"):
a world_size-2 gloo run of the protocol promql.topk drives on the GPU (vmb_topk_candidates -> all-gather -> vmb_topk_merge ->
vmb_topk_apply), with numpy stand-ins for the three device steps.  The masked shards put together must equal the
single-process newAggrFuncTopK (aggr.go:646) over all series."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _better(reverse):
    return (lambda a: np.sort(a)) if reverse else (lambda a: np.sort(a)[::-1])


def _candidates(vals, groups, G, kmax, reverse):
    """== vmb_topk_candidates: per (group, point) the kmax best non-NaN values of the local series, NaN padded"""
    S, P = vals.shape
    cand = np.full((G, P, kmax), np.nan)
    best = _better(reverse)
    for g in range(G):
        rows = np.nonzero(groups == g)[0]
        for p in range(P):
            col = vals[rows, p]
            col = best(col[~np.isnan(col)])[:kmax]
            cand[g, p, :len(col)] = col
    return cand


def _merge(parts, kmax, reverse):
    """== vmb_topk_merge over [nparts x G x P x kmax]"""
    nparts, G, P, _ = parts.shape
    out = np.full((G, P, kmax), np.nan)
    best = _better(reverse)
    for g in range(G):
        for p in range(P):
            col = parts[:, g, p, :].reshape(-1)
            col = best(col[~np.isnan(col)])[:kmax]
            out[g, p, :len(col)] = col
    return out


def _apply(vals, groups, group_sizes, cand, ks, reverse):
    """== vmb_topk_apply: blank everything that is not among the k best of its (group, point); k per point, clamped to the
    group's size over ALL processes (aggr.go:670)"""
    out = vals.copy()
    S, P = vals.shape
    for s in range(S):
        g = int(groups[s])
        for p in range(P):
            k = ks[p]
            kn = 0 if (np.isnan(k) or k < 0) else int(min(k, group_sizes[g]))
            v = vals[s, p]
            if kn == 0 or np.isnan(v):
                out[s, p] = np.nan
                continue
            have = cand[g, p, :kn]
            have = have[~np.isnan(have)]
            if len(have) < kn:
                continue  # fewer values than k in the group: all of them are kept
            thr = have[-1]
            if (v > thr) if reverse else (v < thr):
                out[s, p] = np.nan
    return out


def _topk_single(vals, ks, groups, reverse):
    out = vals.copy()
    S, P = vals.shape
    for g in np.unique(groups):
        rows = np.nonzero(groups == g)[0]
        for p in range(P):
            k = ks[p]
            kn = 0 if (np.isnan(k) or k < 0) else int(min(k, len(rows)))
            col = vals[rows, p]
            key = np.where(np.isnan(col), np.inf if reverse else -np.inf, col)
            order = np.argsort(-key if reverse else key, kind="stable")
            out[rows[order[:len(rows) - kn]], p] = np.nan
    return out


def _worker(rank, world, port, vals, groups, G, ks, kmax, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from victoriametrics_b200 import promql
    mine = promql.shard_series(vals.shape[0], rank, world)
    res = {}
    for reverse in (False, True):
        local = vals[mine]
        cand = _candidates(local, groups[mine], G, kmax, reverse)
        t = torch.from_numpy(cand)
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        merged = _merge(np.stack([x.numpy() for x in gathered]), kmax, reverse)
        sizes = np.bincount(groups, minlength=G)  # series per group over all processes
        res[reverse] = (mine, _apply(local, groups[mine], sizes, merged, ks, reverse))
    out_q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_topk_protocol_world2_gloo():
    rng = np.random.default_rng(9)
    S, P, G = 37, 11, 3
    vals = rng.normal(0, 10, (S, P))
    vals[rng.random((S, P)) < 0.25] = np.nan
    vals[:, 4] = np.nan                       # a point nobody has
    groups = rng.integers(0, G, S).astype(np.uint32)
    ks = np.array([3, 1, 2, 0, 5, 100, -1, np.nan, 4, 2, 7], dtype=np.float64)
    kmax = int(min(np.nanmax(ks), np.bincount(groups, minlength=G).max()))  # like promql.topk: largest k, clamped to the largest group
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, vals, groups, G, ks, kmax, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for reverse in (False, True):
        got = np.full_like(vals, np.nan)
        for rank in (0, 1):
            mine, out = results[rank][reverse]
            got[mine] = out
        exp = _topk_single(vals, ks, groups, reverse)
        assert np.array_equal(np.isnan(got), np.isnan(exp)), reverse
        assert np.array_equal(got[~np.isnan(got)], exp[~np.isnan(exp)]), reverse
