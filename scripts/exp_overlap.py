"""experiment: does running K independent sub-batches on K streams (K host threads, K contexts) overlap the stages?"""
import sys, time, threading
sys.path.insert(0, ".")
import numpy as np, torch
import bench
import victoriametrics_b200 as vm
from victoriametrics_b200 import promql, storage

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
start, end, step = bench.query_range(8192, 300000, 15000)
points = 1 + (end - start) // step
descs, payload, _ = bench.gen_blocks(NB, 8192, 1234)
for K in (1, 2, 4):
    per = NB // K
    parts = []
    for k in range(K):
        ctx = vm.Context(0)
        st = torch.cuda.Stream()
        ctx.set_stream(st.cuda_stream)
        d = descs[k * per:(k + 1) * per].copy()
        d["series_idx"] -= d["series_idx"][0]
        B = storage.Blocks(d, payload, ctx)
        out = torch.empty((per, points), dtype=torch.float64, device="cuda")
        parts.append((ctx, st, B, out))
    def work(i, reps):
        ctx, st, B, out = parts[i]
        for _ in range(reps):
            promql.eval_rollup_func("rate", B, start, end, step, 300000, out_dev_ptr=out.data_ptr())
    def run(reps):
        th = [threading.Thread(target=work, args=(i, reps)) for i in range(K)]
        [t.start() for t in th]; [t.join() for t in th]
    run(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(10)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print("K=%d: %.3f ms per %d blocks -> %.2f G samples/s" % (K, dt * 1e3, NB, NB * 8192 / dt / 1e9), flush=True)
    del parts
