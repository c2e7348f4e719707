"""debug driver of the fused kernel: a few blocks, fused vs un-fused vs oracle"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import victoriametrics_b200 as vm
import blockgen
T0 = 1_700_000_000_000
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
kind = sys.argv[2] if len(sys.argv) > 2 else "counter"
func = sys.argv[3] if len(sys.argv) > 3 else "rate"
rows = int(sys.argv[4]) if len(sys.argv) > 4 else 8192
rng = np.random.default_rng(1)
blocks = [blockgen.OBlock(blockgen.gen_timestamps(rng, "regular", rows, T0), blockgen.gen_values(rng, kind, rows), -2, 64, i) for i in range(nb)]
print("mt", {b.vmt for b in blocks}, flush=True)
descs, payload = blockgen.to_blockset(blocks)
ctx = vm.default_context()
B = vm.storage.Blocks(descs, payload, ctx)
start, end, step, window = T0 + 300000, T0 + 15000 * (rows - 1), 15000, 300000
P = 1 + (end - start) // step
res = {}
for fused in (False, True):
    out = torch.full((nb, P), -7.0, dtype=torch.float64, device="cuda")
    ctx.set_fused(fused)
    t = time.time()
    print("launch fused=%s" % fused, flush=True)
    _, sc = vm.promql.eval_rollup_func(func, B, start, end, step, window, out_dev_ptr=out.data_ptr())
    torch.cuda.synchronize()
    print("done fused=%s in %.3f s scanned=%d" % (fused, time.time() - t, sc), flush=True)
    res[fused] = out.cpu().numpy()
a, b = res[True], res[False]
eq = a.view(np.uint64) == b.view(np.uint64)
print("bit-equal:", eq.all(), "mismatches:", (~eq).sum())
if not eq.all():
    idx = np.argwhere(~eq)[:10]
    for i, j in idx:
        print(i, j, a[i, j], b[i, j])
