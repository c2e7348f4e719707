#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors from its Go test files into tests/golden/go_kats.json.

Run in the build container only (needs /root/reference):   python tests/golden/extract_go_kats.py
The Go sources are parsed as text (there is no Go toolchain here); only literal vectors are taken, no code.
Numeric literals are evaluated with exact rational arithmetic and rounded once, like Go untyped constants.

Float encoding in the JSON: strings -- float.hex() or one of "nan", "inf", "-inf", "stale" (Prometheus StaleNaN).
"""
import json
import os
import re
import sys
from fractions import Fraction

REF = os.environ.get("VM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "go_kats.json")

V_INF_POS = (1 << 63) - 1
V_INF_NEG = -(1 << 63)
V_STALE = (1 << 63) - 2
V_MAX = (1 << 63) - 3
V_MIN = -(1 << 63) + 1
CONSTS = {"vInfPos": V_INF_POS, "vInfNeg": V_INF_NEG, "vStaleNaN": V_STALE, "vMax": V_MAX, "vMin": V_MIN,
          "MarshalTypeZSTDNearestDelta2": 1, "MarshalTypeDeltaConst": 2, "MarshalTypeConst": 3,
          "MarshalTypeZSTDNearestDelta": 4, "MarshalTypeNearestDelta2": 5, "MarshalTypeNearestDelta": 6}
SPECIAL = {"nan": "nan", "math.NaN()": "nan", "inf": "inf", "-inf": "-inf", "infPos": "inf", "infNeg": "-inf",
           "math.Inf(1)": "inf", "math.Inf(+1)": "inf", "math.Inf(-1)": "-inf", "StaleNaN": "stale",
           "decimal.StaleNaN": "stale"}


def func_body(src, name):
    m = re.search(r"^func %s\(.*?\{" % re.escape(name), src, re.M)
    if not m:
        raise KeyError(name)
    i = m.end()
    depth = 1
    while depth:
        c = src[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
        elif c == '"':
            i = src.index('"', i + 1)
        elif c == "/" and src[i + 1] == "/":
            i = src.index("\n", i)
        i += 1
    return src[m.end():i - 1]


def split_top(s):
    out, depth, cur, i = [], 0, "", 0
    while i < len(s):
        c = s[i]
        if c == '"':
            j = s.index('"', i + 1)
            cur += s[i:j + 1]
            i = j + 1
            continue
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += c
        i += 1
    if cur.strip():
        out.append(cur.strip())
    return out


NUM_RE = re.compile(r"(?<![\w.])(\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)")


def eval_num(expr):
    """Evaluate a Go constant expression exactly -> Fraction (or a special-name string)."""
    e = expr.strip()
    if e in SPECIAL:
        return SPECIAL[e]
    if e.startswith("-") and e[1:].strip() in SPECIAL:
        v = SPECIAL[e[1:].strip()]
        return {"inf": "-inf", "-inf": "inf"}.get(v, v)
    for k, v in CONSTS.items():
        e = re.sub(r"\b%s\b" % k, "(%d)" % v, e)
    # Go: << has multiplication precedence (Python: lower than +/-) -> fold literal shifts first
    e = re.sub(r"(\d+)\s*<<\s*(\d+)", lambda m: "(%d)" % (int(m.group(1)) << int(m.group(2))), e)
    e = NUM_RE.sub(lambda m: 'F("%s")' % m.group(1), e)
    val = eval(e, {"F": _F, "__builtins__": {}})
    return val


class _F(Fraction):
    """Fraction that supports << with integer operands (Go constant shifts)."""

    def __new__(cls, s):
        return super().__new__(cls, Fraction(s))

    def __lshift__(self, other):
        return Fraction(int(self) << int(other))


def as_int(expr):
    v = eval_num(expr)
    assert not isinstance(v, str), expr
    assert Fraction(v).denominator == 1, expr
    return int(v)


def as_float(expr):
    v = eval_num(expr)
    if isinstance(v, str):
        return v
    f = float(Fraction(v))  # correctly rounded, like Go constant conversion
    return f.hex()


def parse_slice(expr, conv):
    e = expr.strip()
    if e == "nil":
        return []
    m = re.match(r"^\[\](?:int64|float64)\{(.*)\}$", e, re.S)
    assert m, expr
    return [conv(x) for x in split_top(m.group(1))]


def calls(body, fname):
    """yield the argument lists of every `fname(...)` call statement in body"""
    for m in re.finditer(r"^\s*%s\(" % re.escape(fname), body, re.M):
        i = m.end()
        depth = 1
        j = i
        while depth:
            c = body[j]
            if c == '"':
                j = body.index('"', j + 1)
            elif c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
            j += 1
        yield split_top(body[i:j - 1])


def unq(s):
    assert s[0] == '"' and s[-1] == '"', s
    return s[1:-1]


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def main():
    K = {}
    # ------------------------------------------------------------------ lib/encoding
    nd2 = read("lib/encoding/nearest_delta2_test.go")
    K["nearest_delta"] = [[as_int(a[1]), as_int(a[2]), as_int(a[3]), as_int(a[4]), as_int(a[5])]
                          for a in calls(func_body(nd2, "TestNearestDelta"), "testNearestDelta")]
    K["marshal_nearest_delta2"] = [[parse_slice(a[1], as_int), as_int(a[2]), as_int(a[3]), unq(a[4])]
                                   for a in calls(func_body(nd2, "TestMarshalInt64NearestDelta2"),
                                                  "testMarshalInt64NearestDelta2")]
    nd = read("lib/encoding/nearest_delta_test.go")
    K["marshal_nearest_delta"] = [[parse_slice(a[1], as_int), as_int(a[2]), as_int(a[3]), unq(a[4])]
                                  for a in calls(func_body(nd, "TestMarshalInt64NearestDelta"),
                                                 "testMarshalInt64NearestDelta")]
    enc = read("lib/encoding/encoding_test.go")
    for key, fn in (("is_const", "TestIsConst"), ("is_delta_const", "TestIsDeltaConst"), ("is_gauge", "TestIsGauge")):
        K[key] = [[parse_slice(a[0], as_int), a[1].split("//")[0].strip() == "true"] for a in calls(func_body(enc, fn), "f")]
    K["ensure_non_decreasing"] = [[parse_slice(a[1], as_int), as_int(a[2]), as_int(a[3]), parse_slice(a[4], as_int)]
                                  for a in calls(func_body(enc, "TestEnsureNonDecreasingSequence"),
                                                 "testEnsureNonDecreasingSequence")]
    K["marshal_array_generic"] = [[parse_slice(a[1], as_int), as_int(a[2]), as_int(a[3])]
                                  for a in calls(func_body(enc, "TestMarshalUnmarshalInt64ArrayGeneric"),
                                                 "testMarshalUnmarshalInt64Array")]
    # ------------------------------------------------------------------ lib/decimal
    dec = read("lib/decimal/decimal_test.go")
    K["positive_float_to_decimal"] = [[as_float(a[0]), as_int(a[1]), as_int(a[2].split("//")[0])]
                                      for a in calls(func_body(dec, "TestPositiveFloatToDecimal"), "f")]
    K["append_decimal_to_float"] = [[parse_slice(a[1], as_int), as_int(a[2]), parse_slice(a[3], as_float)]
                                    for a in calls(func_body(dec, "TestAppendDecimalToFloat"), "testAppendDecimalToFloat")]
    K["calibrate_scale"] = [[parse_slice(a[1], as_int), parse_slice(a[2], as_int), as_int(a[3]), as_int(a[4]),
                             parse_slice(a[5], as_int), parse_slice(a[6], as_int), as_int(a[7])]
                            for a in calls(func_body(dec, "TestCalibrateScale"), "testCalibrateScale")]
    K["append_float_to_decimal"] = [[parse_slice(a[1], as_float), parse_slice(a[2], as_int), as_int(a[3])]
                                    for a in calls(func_body(dec, "TestAppendFloatToDecimal"), "testAppendFloatToDecimal")]
    K["from_float"] = [[as_float(a[0]), as_int(a[1]), as_int(a[2])] for a in calls(func_body(dec, "TestFloatToDecimal"), "f")]
    # ------------------------------------------------------------------ app/vmselect/promql rollup
    rt = read("app/vmselect/promql/rollup_test.go")
    m = re.search(r"testValues\s*=\s*(\[\]float64\{.*?\})\s*\n\s*testTimestamps\s*=\s*(\[\]int64\{.*?\})", rt, re.S)
    K["test_values"] = parse_slice(m.group(1), as_float)
    K["test_timestamps"] = parse_slice(m.group(2), as_int)
    K["rollup_func_success"] = [[unq(a[0]), as_float(a[1])]
                                for a in calls(func_body(rt, "TestRollupNewRollupFuncSuccess"), "f")]
    one_arg = {"TestRollupDurationOverTime": "duration_over_time", "TestRollupShareLEOverTime": "share_le_over_time",
               "TestRollupShareGTOverTime": "share_gt_over_time", "TestRollupShareEQOverTime": "share_eq_over_time",
               "TestRollupCountLEOverTime": "count_le_over_time", "TestRollupCountGTOverTime": "count_gt_over_time",
               "TestRollupCountEQOverTime": "count_eq_over_time", "TestRollupCountNEOverTime": "count_ne_over_time",
               "TestRollupSumLEOverTime": "sum_le_over_time", "TestRollupSumGTOverTime": "sum_gt_over_time",
               "TestRollupSumEQOverTime": "sum_eq_over_time", "TestRollupQuantileOverTime": "quantile_over_time",
               "TestRollupPredictLinear": "predict_linear", "TestRollupHoeffdingBoundLower": "hoeffding_bound_lower",
               "TestRollupHoeffdingBoundUpper": "hoeffding_bound_upper"}
    K["rollup_func_one_arg"] = []
    for fn, name in one_arg.items():
        for a in calls(func_body(rt, fn), "f"):
            K["rollup_func_one_arg"].append([name, as_float(a[0]), as_float(a[1])])
    K["rollup_holt_winters"] = [[as_float(a[0]), as_float(a[1]), as_float(a[2])]
                                for a in calls(func_body(rt, "TestRollupHoltWinters"), "f")]
    K["rollup_outlier_iqr"] = [[parse_slice(a[0], as_float), as_float(a[1])]
                               for a in calls(func_body(rt, "TestRollupOutlierIQR"), "f")]
    K["rollup_delta"] = [[as_float(a[0]), as_float(a[1]), as_float(a[2]), parse_slice(a[3], as_float), as_float(a[4])]
                         for a in calls(func_body(rt, "TestRollupDelta"), "f")]
    K["rollup_deriv_fast_prometheus"] = [[parse_slice(a[0], as_float), as_int(a[1]), as_float(a[2])]
                                         for a in calls(func_body(rt, "TestRollupDerivFastPrometheus"), "f")]
    K["linear_regression"] = [[parse_slice(a[0], as_float), parse_slice(a[1], as_int), as_float(a[2]), as_float(a[3])]
                              for a in calls(func_body(rt, "TestLinearRegression"), "f")]

    # rollupConfig.Do sub-tests: every t.Run block holding an `rc := rollupConfig{...}` literal
    do_tests = []
    for tm in re.finditer(r"^func (Test\w+)\(t \*testing\.T\) \{", rt, re.M):
        tname = tm.group(1)
        body = func_body(rt, tname)
        if "rollupConfig{" not in body or tname == "TestRollupBigNumberOfValues":
            continue
        # variables defined at the top of the Test function (before the first t.Run)
        scope = {}

        def grab_vars(text, into):
            for vm in re.finditer(r"^\s*(\w+)\s*:?=\s*(\[\](?:int64|float64)\{[^}]*\})", text, re.M):
                conv = as_int if "int64" in vm.group(2)[:8] else as_float
                into[vm.group(1)] = parse_slice(vm.group(2), conv)

        first_run = body.find("t.Run(")
        grab_vars(body[:first_run if first_run >= 0 else len(body)], scope)
        for rm in re.finditer(r't\.Run\("([^"]*)", func\(t \*testing\.T\) \{', body):
            # brace-match the sub-test body
            i = rm.end()
            depth = 1
            j = i
            while depth:
                c = body[j]
                if c == '"':
                    j = body.index('"', j + 1)
                elif c == "/" and body[j + 1] == "/":
                    j = body.index("\n", j)
                elif c == "{":
                    depth += 1
                elif c == "}":
                    depth -= 1
                j += 1
            sub = body[i:j - 1]
            cm = re.search(r"rc := rollupConfig\{(.*?)\n\t\t\}", sub, re.S)
            if not cm:
                continue
            local = {}
            grab_vars(body[:rm.start()], local)  # latest (re)assignment before this sub-test wins
            grab_vars(sub, local)
            cfg = {}
            for line in cm.group(1).strip().splitlines():
                line = line.split("//")[0].strip().rstrip(",")
                if not line:
                    continue
                k, v = [x.strip() for x in line.split(":", 1)]
                cfg[k] = v
            dm = re.search(r"rc\.Do\(nil, (\w+), (\w+)\)", sub)
            vname, tname2 = dm.group(1), dm.group(2)
            values = K["test_values"] if vname == "testValues" else local[vname]
            timestamps = K["test_timestamps"] if tname2 == "testTimestamps" else local[tname2]
            sm = re.search(r"samplesScanned != (\d+)", sub)
            em = re.search(r"valuesExpected :?= (\[\]float64\{[^}]*\})", sub)
            if not em:
                continue
            do_tests.append({
                "test": tname + "/" + rm.group(1),
                "func": cfg["Func"],
                "start": as_int(cfg.get("Start", "0")), "end": as_int(cfg.get("End", "0")),
                "step": as_int(cfg.get("Step", "0")), "window": as_int(cfg.get("Window", "0")),
                "lookback_delta": as_int(cfg.get("LookbackDelta", "0")),
                "may_adjust_window": cfg.get("MayAdjustWindow", "false") == "true",
                "values": values, "timestamps": timestamps,
                "samples_scanned": int(sm.group(1)) if sm else None,
                "expected": parse_slice(em.group(1), as_float),
            })
    K["rollup_do"] = do_tests

    # ------------------------------------------------------------------ lib/storage/dedup.go, netstorage.mergeSortBlocks
    def dur_ms(expr):
        e = expr.replace("time.Millisecond", "1").replace("time.Second", "1000")
        return as_int(e)
    src = read("lib/storage/dedup_test.go")
    K["needs_dedup"] = [{"interval": as_int(a[0]), "timestamps": parse_slice(a[1], as_int), "expected": a[2] == "true"}
                        for a in calls(func_body(src, "TestNeedsDedup"), "f")]
    dd = []
    for a in calls(func_body(src, "TestDeduplicateSamples"), "f"):
        ts = parse_slice(a[1], as_int)
        dd.append({"interval": dur_ms(a[0]), "timestamps": ts, "values": [float(i).hex() for i in range(len(ts))],
                   "timestamps_expected": parse_slice(a[2], as_int), "values_expected": parse_slice(a[3], as_float)})
    for tname in ("TestDeduplicateSamplesWithIdenticalTimestamps", "TestDeduplicateSamples_KeepsFirstAndLast"):
        for a in calls(func_body(src, tname), "f"):
            dd.append({"interval": dur_ms(a[0]), "timestamps": parse_slice(a[1], as_int), "values": parse_slice(a[2], as_float),
                       "timestamps_expected": parse_slice(a[3], as_int), "values_expected": parse_slice(a[4], as_float)})
    K["dedup_samples"] = dd

    def parse_struct(expr):  # {Timestamps: []int64{..}, Values: []float64{..}} (either may be missing)
        body = expr.strip()
        body = body[body.index("{") + 1:body.rindex("}")]
        d = {"Timestamps": [], "Values": []}
        for fld in split_top(body):
            k, v = fld.split(":", 1)
            d[k.strip()] = parse_slice(v, as_int if k.strip() == "Timestamps" else as_float)
        return d
    src = read("app/vmselect/netstorage/netstorage_test.go")
    mm = []
    for a in calls(func_body(src, "TestMergeSortBlocks"), "f"):
        blocks = []
        if a[0].strip() != "nil":
            inner = a[0].strip()
            inner = inner[inner.index("{") + 1:inner.rindex("}")]
            blocks = [parse_struct(b) for b in split_top(inner)]
        exp = parse_struct(a[2])
        mm.append({"blocks": [{"timestamps": b["Timestamps"], "values": b["Values"]} for b in blocks], "dedup_interval": as_int(a[1]),
                   "timestamps_expected": exp["Timestamps"], "values_expected": exp["Values"]})
    K["merge_sort_blocks"] = mm

    with open(OUT, "w") as f:
        json.dump(K, f, indent=0, separators=(",", ":"))
    print("wrote", OUT, {k: (len(v) if isinstance(v, list) else 1) for k, v in K.items()})


if __name__ == "__main__":
    sys.exit(main())
