"""experiment (not a bench leg): decode time of blocks marshaled by the REFERENCE encoder (oracle + libzstd 1.5.7),
per value kind -- shows what the sequences path costs next to the Huffman-only fast path."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import blockgen
import victoriametrics_b200 as vm
from victoriametrics_b200 import storage, promql

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
ctx = vm.default_context()
T0 = 1_700_000_000_000
TKIND = sys.argv[2] if len(sys.argv) > 2 else "regular"
KINDS = sys.argv[3].split(",") if len(sys.argv) > 3 else ("counter", "counter_smooth", "gauge", "gauge_small")
for kind in KINDS:
    rng = np.random.default_rng(1)
    uniq = [blockgen.OBlock(blockgen.gen_timestamps(rng, TKIND, 8192, T0), blockgen.gen_values(rng, kind, 8192), -2, 64, 0)
            for _ in range(64)]
    blocks = []
    for i in range(NB):
        b = uniq[i % 64]
        c = blockgen.OBlock.__new__(blockgen.OBlock)
        c.__dict__.update(b.__dict__)
        c.series_idx = i
        blocks.append(c)
    descs, payload = blockgen.to_blockset(blocks)
    B = storage.Blocks(descs, payload)
    mts = sorted(set((int(b.tmt), int(b.vmt)) for b in uniq))
    ratio = sum(b.vdata.size for b in uniq) / (64 * 8192)
    start, end, step = T0 + 300000, int(uniq[0].ts[-1]), 15000
    points = 1 + (end - start) // step
    out = torch.empty((NB, points), dtype=torch.float64, device="cuda")
    for _ in range(2):
        promql.eval_rollup_func("rate", B, start, end, step, 300000, out_dev_ptr=out.data_ptr())
    ctx.enable_stage_timing(True)
    promql.eval_rollup_func("rate", B, start, end, step, 300000, out_dev_ptr=out.data_ptr())
    st = ctx.stage_ms()
    ctx.enable_stage_timing(False)
    print("%-15s val_mt %s  %.2f B/sample  stages ms zstd %.3f decode %.3f preamble %.3f rollup %.3f fused %.3f  (%d blocks)" %
          (kind, mts, ratio, st[0], st[1], st[2], st[3], st[5], NB), flush=True)
