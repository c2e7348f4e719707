"""experiment: the same samples as bench.py's 20 000 x 8192, but every series stored as 4 blocks of 2048 rows (time-disjoint,
arriving in reverse order) -- exercises the host plan's re-layout and the multi-block paths at scale"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
import victoriametrics_b200 as vm
from victoriametrics_b200 import encoding, promql, storage

NS, NB, ROWS = 20000, 4, 2048
rng = np.random.default_rng(7)
T0 = bench.T0
pieces, pos = [], 0
cols = {k: [] for k in ("first_value", "val_off", "val_size", "val_mt", "ts_off", "ts_size", "min_ts", "max_ts", "series_idx")}
tsp = []
for b in range(NB):  # one shared delta-const timestamp payload per block position
    ts = T0 + 15000 * (b * ROWS + np.arange(ROWS, dtype=np.int64))
    td, tmt, tf = encoding.marshal_timestamps(ts)
    tsp.append((pos, td.size, tf, int(ts[-1]), tmt))
    pieces.append(td); pos += td.size
inc = rng.integers(0, 1501, (NS, NB * ROWS), dtype=np.int64)
v = np.cumsum(inc, axis=1)
for b in reversed(range(NB)):  # arrival order: newest block first
    payload, offs, mts, firsts = encoding.marshal_columns(np.ascontiguousarray(v[:, b * ROWS:(b + 1) * ROWS]))
    pieces.append(payload)
    cols["first_value"].append(firsts); cols["val_off"].append(offs[:-1] + pos); cols["val_size"].append(np.diff(offs).astype(np.uint32))
    cols["val_mt"].append(mts); pos += payload.size
    cols["ts_off"].append(np.full(NS, tsp[b][0])); cols["ts_size"].append(np.full(NS, tsp[b][1], dtype=np.uint32))
    cols["min_ts"].append(np.full(NS, tsp[b][2])); cols["max_ts"].append(np.full(NS, tsp[b][3]))
    cols["series_idx"].append(np.arange(NS, dtype=np.uint32))
cat = {k: np.concatenate(vv) for k, vv in cols.items()}
order = np.argsort(cat["series_idx"], kind="stable")  # blocks of a series consecutive, still newest first inside a series
descs = storage.descs_from_arrays(rows=np.full(NS * NB, ROWS, dtype=np.uint32), scale=-2, ts_mt=tsp[0][4], precision_bits=64,
                                  **{k: vv[order] for k, vv in cat.items()})
payload = np.concatenate(pieces)
ctx = vm.default_context()
B = storage.Blocks(descs, payload)
start, end, step = bench.query_range(NB * ROWS, 300000, 15000)
points = 1 + (end - start) // step
out = torch.empty((NS, points), dtype=torch.float64, device="cuda")
for _ in range(3):
    promql.eval_rollup_func("rate", B, start, end, step, 300000, out_dev_ptr=out.data_ptr())
ctx.enable_stage_timing(True)
promql.eval_rollup_func("rate", B, start, end, step, 300000, out_dev_ptr=out.data_ptr())
st = ctx.stage_ms()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    promql.eval_rollup_func("rate", B, start, end, step, 300000, out_dev_ptr=out.data_ptr())
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print("4 blocks x 2048 rows per series: %.3f ms/step (%.1f G samples/s); stages zstd %.3f decode %.3f preamble %.3f rollup %.3f" %
      (dt * 1e3, NS * NB * ROWS / dt / 1e9, st[0], st[1], st[2], st[3]))
# same result as the single-block layout?
d1, p1, _ = bench.gen_blocks(64, NB * ROWS, 99)
print("nan rows:", int(torch.isnan(out).all(dim=1).sum().item()), "of", NS)
