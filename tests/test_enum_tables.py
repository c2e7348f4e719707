"""The name -> id tables of the Python host mirror must follow the enums of include/vmb200.h (and, for the kernels, the internal
enums of csrc/*.inc that mirror them): a reordered enum would silently run another operator."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "vmb200.h")).read()


def _enum(text, name):
    """enum constants in order with their values (explicit `= n` restarts the count)"""
    body = text[text.index("enum %s" % name):]
    body = body[body.index("{") + 1:body.index("};")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    body = re.sub(r"//[^\n]*", "", body)
    out, nxt = [], 0
    for item in body.split(","):
        item = item.strip()
        if not item:
            continue
        m = re.match(r"([A-Za-z0-9_]+)\s*(?:=\s*(\d+))?$", item)
        assert m, item
        if m.group(2) is not None:
            nxt = int(m.group(2))
        out.append((m.group(1), nxt))
        nxt += 1
    return out


def test_transform_ids_follow_the_header_and_the_kernel_enum():
    from victoriametrics_b200 import promql
    pub = _enum(HDR, "vmb_transform_func")
    assert {n[len("VMB_TF_"):].lower(): v for n, v in pub} == promql.TRANSFORM_FUNCS
    inc = open(os.path.join(ROOT, "victoriametrics_b200", "csrc", "transform.inc")).read()
    internal = {n[len("TF_"):].lower(): v for n, v in _enum(inc, "vmb_transform_internal") if not n.startswith("TF__")}
    assert internal == promql.TRANSFORM_FUNCS


def test_binary_operator_ids_follow_the_header_and_the_kernel_enum():
    from victoriametrics_b200 import promql
    names = {"PLUS": "+", "MINUS": "-", "MUL": "*", "DIV": "/", "MOD": "%", "POW": "^", "ATAN2": "atan2", "EQ": "==", "NEQ": "!=", "GT": ">",
             "LT": "<", "GTE": ">=", "LTE": "<=", "DEFAULT": "default", "IF": "if", "IFNOT": "ifnot"}
    pub = _enum(HDR, "vmb_binop")
    assert {names[n[len("VMB_BO_"):]]: v for n, v in pub} == promql.BINARY_OPS
    inc = open(os.path.join(ROOT, "victoriametrics_b200", "csrc", "matrix_ops.inc")).read()
    internal = {names[n[len("BO_"):]]: v for n, v in _enum(inc, "vmb_binop_internal") if not n.startswith("BO__")}
    assert internal == promql.BINARY_OPS


def test_aggregate_ids_follow_the_header():
    from victoriametrics_b200 import promql
    pub = _enum(HDR, "vmb_aggr_func")
    assert {n[len("VMB_AGGR_"):].lower(): v for n, v in pub} == promql.AGGR_FUNCS


def test_rollup_function_ids_follow_the_header():
    from victoriametrics_b200 import promql
    pub = [(n, v) for n, v in _enum(HDR, "vmb_rollup_func") if n != "VMB_RF__COUNT"]
    ids = sorted(set(promql.ROLLUP_FUNCS.values()))
    assert ids == [v for _, v in pub] and len(pub) == len(ids)  # every enum value is reachable by at least one MetricsQL name
