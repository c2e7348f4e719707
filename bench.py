#!/usr/bin/env python3
"""bench.py -- the hot path of BASELINE.json on N B200s of one node.

Workload (BASELINE.json configs[1], the configuration `metric` is quoted on for one GPU):
  decode 100 000 blocks (8192 samples each; timestamps delta-const, values ZSTD nearest-delta2 counters at scale -2,
  rare resets) + rate(m[5m]) at step 15 s  ->  [100 000 x 8172] float64.
One "step" = one pass of the hot path over that batch.  N > 1: every rank owns its own 100 000 blocks (series shard by
TSID, no data-path collective) -> weak scaling.

  value : samples/s with the compressed blocks already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e   : the same through the public host-buffer call (vmb_eval_rollup_host): H2D of descriptors+payload from pinned
          host memory, decode, rollup, D2H of the result, all inside the timed region
  roofline : the dominant kernel stage, algorithmic bytes / measured stage time vs MEASURED_PEAKS.json
  cpu_baseline : the oracle (C++ restatement of the Go path, zstd through the reference's own libzstd when
          oracle/_ref is present) on all host cores over a bounded sample of the same blocks  (rank 0, N = 1)

  --impl reference : times the reference's CPU implementation of the path (see cpu_baseline) on the same config.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T0 = 1_700_000_000_000
SCRAPE_MS = 15000
SCALE = -2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--blocks", type=int, default=100_000, help="blocks (= series) per GPU")
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--func", default="rate")
    ap.add_argument("--ts", default="regular", choices=["regular", "jitter"],
                    help="timestamps: regular = t0 + 15 s * i (one shared MarshalTypeDeltaConst payload, configs[1]); jitter = "
                         "every series has its own +-50 ms scrape jitter (zstd nearest-delta2 timestamp columns)")
    ap.add_argument("--kind", default="counter", choices=["counter", "gauge", "mixed"],
                    help="synthetic values: counter = configs[1] (default), gauge = configs[2]-style round(N(5000,300)) at "
                         "scale -2, mixed = configs[4]-style 40%% counters / 30%% gauges / 20%% const / 10%% delta-const")
    ap.add_argument("--window-ms", type=int, default=300_000)
    ap.add_argument("--step-ms", type=int, default=15_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--aggr", default="", help="configs[4] variant: aggr(func(m[d])) by (label) with an NCCL all-reduce of the "
                                               "per-GPU partial states, e.g. --aggr sum (the host-buffer arm folds on the GPU too and returns [groups x points])")
    ap.add_argument("--groups", type=int, default=1000, help="label groups for --aggr")
    ap.add_argument("--cpu-seconds", type=float, default=24.0, help="target CPU work of the cpu_baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ synthetic input
GEN_STATS = {"series": 0, "series_with_drop": 0, "rows": 0, "rows_from_first_drop": 0}
RCR_FUNCS = ("rate", "increase", "irate", "increase_pure", "increase_prometheus", "rate_prometheus", "rollup_rate", "rollup_increase")


def gen_blocks(nblocks, rows, seed, chunk=4000, kind="counter", ts_kind="regular"):
    """node_cpu_seconds_total-like counters, marshaled by the product's own encoder
    (vmb_marshal_columns).  -> (descs structured array, payload np.uint8)"""
    from victoriametrics_b200 import encoding, storage
    rng = np.random.default_rng(seed)
    ts = T0 + SCRAPE_MS * np.arange(rows, dtype=np.int64)
    tdata, tmt, tfirst = encoding.marshal_timestamps(ts)
    assert tmt == encoding.MarshalTypeDeltaConst
    pieces = [tdata]
    pos = tdata.size
    cols = {k: [] for k in ("first_value", "val_off", "val_size", "val_mt", "ts_off", "ts_size", "ts_mt", "min_ts", "max_ts")}
    for c0 in range(0, nblocks, chunk):
        n = min(chunk, nblocks - c0)
        inc = rng.integers(0, 1501, (n, rows), dtype=np.int64)
        inc[:, 0] += rng.integers(0, 10 ** 9, n)
        v = np.cumsum(inc, axis=1)
        resets = rng.random((n, rows)) < 1e-4
        resets[:, 0] = False
        base = np.maximum.accumulate(np.where(resets, v, 0), axis=1)
        v -= base
        if kind != "counter":
            g = np.rint(rng.normal(5000.0, 300.0, (n, rows))).astype(np.int64)
            if kind == "gauge":
                v = g
            else:  # mixed: series i takes its kind from i mod 10
                k = (np.arange(c0, c0 + n) % 10)[:, None]
                const = np.broadcast_to(rng.integers(0, 10 ** 6, (n, 1)), (n, rows))
                dconst = rng.integers(0, 10 ** 6, (n, 1)) + rng.integers(1, 100, (n, 1)) * np.arange(rows, dtype=np.int64)[None, :]
                v = np.where(k < 4, v, np.where(k < 7, g, np.where(k < 9, const, dconst)))
        GEN_STATS["series"] += n
        dropped = np.diff(v, axis=1) < 0
        has = dropped.any(axis=1)
        GEN_STATS["series_with_drop"] += int(np.count_nonzero(has))
        GEN_STATS["rows"] += n * rows
        # removeCounterResets touches a series from the 128-row group of its first value drop on
        GEN_STATS["rows_from_first_drop"] += int((rows - ((np.argmax(dropped, axis=1)[has] + 1) & ~127)).sum())
        if ts_kind == "jitter":  # every series has its own scrape jitter of +-50 ms (SURVEY.md 8d config 2 variant)
            tj = ts[None, :] + rng.integers(-50, 51, (n, rows))
            tp_, toffs, tmts, tfirsts = encoding.marshal_columns(tj)
            pieces.append(tp_)
            cols["ts_off"].append(toffs[:-1] + pos)
            cols["ts_size"].append(np.diff(toffs).astype(np.uint32))
            cols["ts_mt"].append(tmts)
            cols["min_ts"].append(tfirsts)
            cols["max_ts"].append(tj[:, -1].copy())
            pos += tp_.size
        payload, offs, mts, firsts = encoding.marshal_columns(v)
        pieces.append(payload)
        cols["first_value"].append(firsts)
        cols["val_off"].append(offs[:-1] + pos)
        cols["val_size"].append(np.diff(offs).astype(np.uint32))
        cols["val_mt"].append(mts)
        pos += payload.size
    tcols = dict(min_ts=tfirst, max_ts=int(ts[-1]), ts_off=0, ts_size=tdata.size, ts_mt=tmt)
    if ts_kind == "jitter":
        tcols = {k: np.concatenate(cols[k]) for k in ("min_ts", "max_ts", "ts_off", "ts_size", "ts_mt")}
    descs = storage.descs_from_arrays(
        first_value=np.concatenate(cols["first_value"]), val_off=np.concatenate(cols["val_off"]),
        val_size=np.concatenate(cols["val_size"]), rows=np.full(nblocks, rows, dtype=np.uint32),
        series_idx=np.arange(nblocks, dtype=np.uint32), scale=SCALE, val_mt=np.concatenate(cols["val_mt"]), precision_bits=64,
        **tcols)
    return descs, np.concatenate(pieces)


def query_range(rows, window_ms, step_ms):
    start = T0 + window_ms
    end = T0 + SCRAPE_MS * (rows - 1)
    return start, end, step_ms


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """SM clock / throttle reasons / power sampled DURING the timed region: NVML in-process every ~2 ms (the timed region of
    the device-resident arm is < 100 ms), `nvidia-smi -lms` as the fallback"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index, pci_bus_id=None):
        self.gpu = gpu_index
        self.pci = pci_bus_id
        self.rows = []   # (time, sm_mhz, max_mhz, power_w, [reasons])
        self.proc = None
        self.nvml = None
        self.stop_flag = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.pci:
                try:
                    h = pynvml.nvmlDeviceGetHandleByPciBusId(self.pci.encode() if isinstance(self.pci, str) else self.pci)
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nvml = (pynvml, h)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        pynvml, h = self.nvml
        R = (("hw_slowdown", getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8)),
             ("hw_thermal_slowdown", getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
             ("sw_thermal_slowdown", getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
             ("sw_power_cap", getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)))
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = int(get_reasons(h))
                try:
                    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = 0.0
                self.rows.append((time.time(), sm, self.max_mhz, pw, [n for n, bit in R if mask & bit]))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) >= 9:
                try:
                    reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9])
                               if v.lower().startswith("active")]
                    self.rows.append((time.time(), float(r[1]), float(r[2]), float(r[3]), reasons))
                except ValueError:
                    pass

    def window(self, t_begin, t_end):
        rows = [r for r in self.rows if t_begin <= r[0] <= t_end]
        src = "inside the timed region"
        if not rows:
            rows = [r for r in self.rows if t_begin - 0.2 <= r[0] <= t_end + 0.2]
            src = "within 0.2 s of the timed region"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(r[1] for r in rows)
        reasons = sorted({n for r in rows for n in r[4]})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": rows[0][2], "reasons": reasons, "samples": len(rows),
                "power_w_max": round(max(r[3] for r in rows), 1), "sampled": src,
                "via": "nvml" if self.nvml else "nvidia-smi"}

    def stop(self, t_begin=None, t_end=None):
        self.stop_flag = True
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
        if t_begin is None:
            return None
        return self.window(t_begin, t_end)


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(descs, payload, func, start, end, step, window, target_seconds, repeats=1):
    """the oracle's threaded per-series loop (oracle/cpu_pipeline.cpp) on a bounded sample -> dict"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from rollup_names import RF
    from victoriametrics_b200 import promql
    L = O.lib()
    fn = L.vmo_cpu_eval_rollup
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_size_t, O.u8p, C.c_int64, C.c_int64, C.POINTER(O.RollupCfg), C.c_int, C.c_int, O.f64p,
                   C.POINTER(C.c_uint64), C.c_int, C.c_int]
    rc = promql.get_rollup_configs(func, start, end, step, window)
    phis = np.full(rc.points, 0.99) if func == "quantile_over_time" else None  # kept alive by this frame
    cfg = O.RollupCfg(RF[func], start, end, step, window, 0, 0, int(rc.MayAdjustWindow), int(rc.isDefaultRollup),
                      rc.samplesScannedPerCall, phis.ctypes.data_as(O.f64p) if phis is not None else None, None)
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    P = rc.points
    rows = int(descs["rows"][0])
    kind = "reference" if L.vmo_zstd_ref_available() else "port"

    def run(nb, out, nthreads=None):
        scanned = C.c_uint64(0)
        t = time.perf_counter()
        r = fn(descs.ctypes.data, nb, payload.ctypes.data_as(O.u8p), -(1 << 63), (1 << 63) - 1, C.byref(cfg),
               int(rc.removeCounterResets), int(rc.dropStaleNaNs), out.ctypes.data_as(O.f64p), C.byref(scanned),
               nthreads or cores, 1)
        dt = time.perf_counter() - t
        assert r == 0, r
        return dt

    nb0 = min(len(descs), max(cores * 8, 256))
    out0 = np.empty((nb0, P), dtype=np.float64)
    out0.fill(0.0)  # touch the pages outside the timed region (the Go code reuses pooled buffers)
    run(nb0, out0)  # warm-up
    dt0 = run(nb0, out0)
    rate0 = nb0 * rows / dt0
    wall_target = max(0.2, target_seconds / cores)  # target_seconds of CPU work spread over all cores, >= 0.2 s of wall
    nb = int(wall_target * rate0 / rows)
    nb = max(nb0, min(nb, len(descs)))
    out = np.empty((nb, P), dtype=np.float64)
    out.fill(0.0)
    # one software thread per hardware thread is not always the fastest on a hyper-threaded host: also try one per two and
    # report whichever is faster (the baseline gets the benefit of the doubt)
    tried = {}
    for nt in ([cores, cores // 2] if cores >= 16 else [cores]):
        for _ in range(max(repeats, 2)):
            dt = run(nb, out, nt)
            tried[nt] = min(tried.get(nt, dt), dt)
    used = min(tried, key=tried.get)
    best = tried[used]
    return {"value": nb * rows / best, "unit": "samples/s", "cores": used, "kind": kind,
            "sample": "%d of %d blocks x %d rows, %.2f s wall on %d threads (C++ restatement of the Go path%s)%s" %
                      (nb, len(descs), rows, best, used, "; zstd via the reference's libzstd 1.5.7" if kind == "reference" else "",
                       "; thread counts tried: " + ", ".join("%d -> %.2f s" % (k, v) for k, v in sorted(tried.items())) if len(tried) > 1 else ""),
            "seconds": best, "blocks": nb, "host_threads_available": cores}, out


# ------------------------------------------------------------------------------------------------ main
def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    start, end, step = query_range(a.rows, a.window_ms, a.step_ms)
    points = 1 + (end - start) // step
    what = {"counter": "configs[1]: decode %d blocks x %d samples (ts delta-const, values zstd nearest-delta2 counters, scale %d)",
            "gauge": "configs[2]-style: decode %d blocks x %d samples (ts delta-const, values zstd nearest-delta gauges, scale %d)",
            "mixed": "configs[4]-style: decode %d blocks x %d samples (40%% counters, 30%% gauges, 20%% const, 10%% delta-const, scale %d)"}[a.kind]
    if a.ts == "jitter":
        what = what.replace("ts delta-const", "ts zstd nearest-delta2 with +-50 ms jitter")
    workload = (what + " + %s()[%ds] step=%ds per GPU") % (a.blocks, a.rows, SCALE, a.func, a.window_ms // 1000, a.step_ms // 1000)
    base = {"metric": "rollup samples/sec (block decode + %s, raw samples decoded and scanned per second)" % a.func,
            "unit": "samples/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64 codec -> f64 rollup", "data": "synthetic",
            "config": {"workload": workload, "blocks_per_gpu": a.blocks, "rows_per_block": a.rows, "points_per_series": int(points),
                       "parallelism": "series sharded by TSID across %d GPU(s), no data-path collective" % a.gpus,
                       "l2": "inputs (>=1.3 GB compressed, 13 GB decoded) exceed the 126 MB L2; no flush needed"}}

    if a.impl == "reference":
        if rank != 0:
            return 0
        descs, payload = gen_blocks(min(a.blocks, 20000), a.rows, seed=1234, kind=a.kind, ts_kind=a.ts)
        vals = []
        for _ in range(a.warmup):
            cpu_reference(descs, payload, a.func, start, end, step, a.window_ms, 2.0)
        t_all = time.perf_counter()
        last = None
        for _ in range(a.steps):
            last, _ = cpu_reference(descs, payload, a.func, start, end, step, a.window_ms, a.cpu_seconds / max(a.steps, 1) + 1.0)
            vals.append(last["value"])
        v = float(np.median(vals))
        out = dict(base)
        out.update({"impl": "reference", "value": v, "ms_per_step": 1e3 * a.blocks * a.rows / v,
                    "cpu_baseline": dict(last, value=v), "gpu_launches": 0,
                    "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "wall_s": time.perf_counter() - t_all})
        out["config"]["note"] = "CPU reference arm: each step is a bounded sample of the workload; value = samples/s on all host cores"
        print(json.dumps(out))
        return 0

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import victoriametrics_b200 as vm
    from victoriametrics_b200 import _lib, promql, storage
    ctx = vm.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    t_gen = time.perf_counter()
    descs, payload = gen_blocks(a.blocks, a.rows, seed=1234 + rank, kind=a.kind, ts_kind=a.ts)
    gen_s = time.perf_counter() - t_gen
    rows_total = int(a.blocks) * int(a.rows)
    compressed = int(descs["val_size"].sum()) + (int(descs["ts_size"].sum()) if a.ts == "jitter" else int(descs["ts_size"][0]))

    blocks = storage.Blocks(descs, payload, ctx)
    out_dev = torch.empty((a.blocks, points), dtype=torch.float64, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    func_args = np.full(points, 0.99) if a.func == "quantile_over_time" else None  # quantile_over_time(0.99, m[d])

    def dev_step():
        return promql.eval_rollup_func(a.func, blocks, start, end, step, a.window_ms, args=func_args, out_dev_ptr=out_dev.data_ptr())

    if a.aggr:
        # sum(rate(m[5m])) by (label): every rank folds its own series into [groups x points] partial states, one NCCL
        # all-reduce of values and one of counts merges them, every rank finalizes
        rc_aggr = promql.get_rollup_configs(a.func, start, end, step, a.window_ms)
        group_ids = ((np.arange(a.blocks, dtype=np.int64) * world + rank) % a.groups).astype(np.uint32)

        class Buf:
            def __init__(self, nbytes):
                self.t = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
                self.ptr = self.t.data_ptr()
        ia = promql.IncrementalAggr(a.aggr, a.groups, points, Buf)
        reduce_cb = (lambda v, c, op: promql.torch_all_reduce(v.t, c.t, op)) if world > 1 else None
        h_aggr = _lib.lib().vmb_host_alloc(a.groups * points * 8)
        assert h_aggr, "pinned host allocation failed"
        aggr_host = np.ctypeslib.as_array(C.cast(h_aggr, C.POINTER(C.c_double)), shape=(a.groups, points))

        def dev_step():  # noqa: F811
            scanned_ = ia.update_blocks(blocks, rc_aggr, group_ids)
            ia.finalize(ctx, all_reduce=reduce_cb, out=aggr_host)  # D2H of the [groups x points] query result included
            return None, scanned_
        base["metric"] = "rollup samples/sec (block decode + %s(%s) by label, raw samples decoded and scanned per second)" % (a.aggr, a.func)
        base["config"]["workload"] = workload.replace("configs[1]", "configs[4]-style") + "; %s by %d groups" % (a.aggr, a.groups)
        base["config"]["parallelism"] = ("series sharded by TSID across %d GPU(s); per-GPU partial [groups x points] states merged by "
                                         "NCCL all-reduce (values: %s, counts: sum)" % (a.gpus, promql.ALLREDUCE_OP[a.aggr]))

    # ---- kernel-only: compressed blocks resident in HBM
    for _ in range(a.warmup):
        dev_step()
    pci = None
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        pci = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except Exception:
        pass
    # NUMA placement: run this rank (and first-touch its pinned buffers) on the CPUs next to its GPU's PCIe root, like any
    # multi-GPU host process would be deployed; restored before the CPU baseline, which uses every core
    affinity0 = None
    numa_cpus = None
    try:
        affinity0 = os.sched_getaffinity(0)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % pci[4:]) as f:  # sysfs uses a 4-digit PCI domain
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= affinity0
        if cpus:
            os.sched_setaffinity(0, cpus)
            numa_cpus = len(cpus)
    except Exception:
        pass
    sampler = ClockSampler(local_rank, pci)
    sampler.start()
    barrier()
    l0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tb = time.time()
    ev0.record(stream)
    scanned = 0
    for _ in range(a.steps):
        _, scanned = dev_step()
    ev1.record(stream)
    barrier()
    te = time.time()
    launches = ctx.launch_count - l0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.window(tb, te)
    # per-stage device times of one extra step (CUDA events inside the library, same stream)
    ctx.enable_stage_timing(True)
    dev_step()
    stage_ms = ctx.stage_ms()
    ctx.enable_stage_timing(False)

    # ---- e2e: host buffers in, host result out
    e2e = None
    if not a.no_e2e:
        nbytes_out = (a.groups if a.aggr else a.blocks) * points * 8
        hp = _lib.lib().vmb_host_alloc(payload.size + 64)
        ho = _lib.lib().vmb_host_alloc(nbytes_out)
        hd = _lib.lib().vmb_host_alloc(descs.nbytes)
        assert hp and ho and hd, "pinned host allocation failed"
        h_payload = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(payload.size,))
        h_payload[:] = payload
        h_descs = np.ctypeslib.as_array(C.cast(hd, C.POINTER(C.c_uint8)), shape=(descs.nbytes,)).view(descs.dtype)
        h_descs[:] = descs
        h_out = np.ctypeslib.as_array(C.cast(ho, C.POINTER(C.c_double)), shape=((a.groups if a.aggr else a.blocks), points))

        def host_step():
            if a.aggr and world > 1:  # per-rank partial from host buffers, NCCL all-reduce, finalize + D2H of the result
                sc_ = ia.update_host(h_descs, h_payload, rc_aggr, group_ids, ctx)
                ia.finalize(ctx, all_reduce=reduce_cb, out=h_out)
                return h_out, sc_
            if a.aggr:
                return promql.eval_rollup_aggr_host(a.aggr, a.func, h_descs, h_payload, group_ids, a.groups, start, end, step,
                                                    a.window_ms, args=func_args, out=h_out, ctx=ctx)
            return promql.eval_rollup_func_host(a.func, h_descs, h_payload, start, end, step, a.window_ms, args=func_args,
                                                out=h_out, nseries=a.blocks, ctx=ctx)
        for _ in range(max(1, min(a.warmup, 2))):
            host_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw = time.perf_counter()
        tb2 = time.time()
        e0.record(stream)
        for _ in range(a.steps):
            host_step()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - tw
        e2e_clocks = sampler.window(tb2, time.time())
        e2e_ms = max(e0.elapsed_time(e1), 0.0)
        e2e = {"ms": e2e_ms, "wall_ms": wall * 1e3, "h2d": int(descs.nbytes + payload.size), "d2h": int(nbytes_out)}
        if a.aggr:  # same query result as the device-resident arm (chunk-wise folding only reorders float additions)
            assert np.allclose(h_out, aggr_host, rtol=1e-9, atol=0, equal_nan=True)
        else:
            check = float(np.nansum(h_out[: min(a.blocks, 64)]))
            dcheck = float(torch.nansum(out_dev[: min(a.blocks, 64)]).item())
            assert abs(check - dcheck) <= 1e-9 * max(1.0, abs(dcheck)), (check, dcheck)

    sampler.stop()
    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e["ms"] if e2e else 0.0, e2e["wall_ms"] if e2e else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_wall_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        ms_per_step = dev_ms / a.steps
        value = world * rows_total / (ms_per_step / 1e3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
        varint_bytes = int(2 * rows_total) * (2 if a.ts == "jitter" else 1)  # ~2 B/sample zig-zag varints per zstd column
        drop_frac = (GEN_STATS["rows_from_first_drop"] / max(GEN_STATS["rows"], 1)) if a.func in RCR_FUNCS else 0.0
        stage_names = ["zstd", "column_decode", "series_preamble", "rollup", "aggregate"]
        stage_bytes = [compressed + varint_bytes,                          # zstd: read frames, write varint bytes
                       varint_bytes + 64 * a.blocks + rows_total * 16,     # decode: read varints + descs, write ts+val
                       int(rows_total * 16 * drop_frac),                   # preamble: read+write values from the first value drop
                                                                           # of a series on (removeCounterResets); the rest is skipped
                       rows_total * 16 + a.blocks * points * 8, 0]         # rollup: read ts+val, write result
        stages = {}
        for n_, ms_, b_ in zip(stage_names, stage_ms, stage_bytes):
            if ms_ > 0:
                stages[n_] = {"ms": round(ms_, 4), "algorithmic_GB": round(b_ / 1e9, 3), "GBps": round(b_ / 1e9 / (ms_ / 1e3), 1)}
        dom = max(stages, key=lambda k: stages[k]["ms"]) if stages else None

        def traffic_of(stage):
            """dram__bytes_read.sum + dram__bytes_write.sum of the stage's kernel from the committed `ncu --set full`
            capture (profiles/r01/traffic.json: bytes per launch at 20 000 blocks, rate workload), scaled to this launch"""
            try:
                t = json.load(open(os.path.join(ROOT, "profiles", "r01", "traffic.json")))
                if a.func != t["func"] or a.rows != t["rows"]:
                    return None
                return int(t["bytes_per_launch"][stage] * (a.blocks / t["blocks"]))
            except Exception:
                return None
        fused_bytes = compressed + a.blocks * points * 8
        out = dict(base)
        out.update({"value": value, "ms_per_step": ms_per_step, "gpu_launches": int(launches), "clocks": clocks,
                    "samples_scanned_per_step": int(scanned)})
        if dom:
            ach = stages[dom]["GBps"]
            out["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                               "traffic": traffic_of(dom), "peak_source": peak_src, "stages": stages,
                               "whole_step": {"fused_algorithmic_GB": round(fused_bytes / 1e9, 3),
                                              "GBps": round(fused_bytes / 1e9 / (ms_per_step / 1e3), 1),
                                              "frac": round(fused_bytes / 1e9 / (ms_per_step / 1e3) / peak, 4)}}
        out["decode_GBps_decoded_basis"] = round(rows_total * 16 / 1e9 / ((stage_ms[0] + stage_ms[1]) / 1e3), 1) if stage_ms[1] > 0 else None
        if e2e:
            per = e2e_ms / a.steps
            out["e2e"] = {"value": world * rows_total / (per / 1e3), "unit": "samples/s", "ms_per_step": per,
                          "wall_ms_per_step": e2e_wall_ms / a.steps, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                          "api": (("vmb_eval_rollup_aggr_host_partial + NCCL all-reduce + vmb_aggr_finalize" if world > 1 else
                                   "vmb_eval_rollup_aggr_host") + " (pinned host descriptors+payload in, [groups x points] result out)"
                                  if a.aggr else "vmb_eval_rollup_host (pinned host descriptors+payload in, pinned host result out)"),
                          "clocks": e2e_clocks}
        out["config"]["compressed_bytes_per_gpu"] = compressed
        out["config"]["bytes_per_sample_compressed"] = round(compressed / rows_total, 3)
        out["config"]["input_generation_s"] = round(gen_s, 1)
        out["config"]["series_with_a_counter_reset"] = round(GEN_STATS["series_with_drop"] / max(GEN_STATS["series"], 1), 3)
        out["config"]["host_affinity"] = ("GPU-local NUMA node, %d CPUs" % numa_cpus) if numa_cpus else "unchanged"
        if affinity0:
            try:
                os.sched_setaffinity(0, affinity0)
            except Exception:
                pass
        if world == 1 and a.cpu_seconds > 0:
            try:
                cb, _ = cpu_reference(descs, payload, a.func, start, end, step, a.window_ms, a.cpu_seconds)
                out["cpu_baseline"] = cb
            except Exception as e:  # the oracle is optional for the product; report instead of failing the bench
                out["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
