"""The Go side's function-id table (integration/go/app/vmselect/promql/vmb200_func_ids.go) is generated from include/vmb200.h:
it must be up to date, cover every rollup function name of the host mirror, and agree with the enum order."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "integration", "go", "gen_func_ids.py")
OUT = os.path.join(ROOT, "integration", "go", "app", "vmselect", "promql", "vmb200_func_ids.go")


def test_go_func_id_table_is_current_and_complete():
    assert subprocess.run([sys.executable, GEN, "--check"]).returncode == 0, "run integration/go/gen_func_ids.py"
    from victoriametrics_b200 import promql
    src = open(OUT).read()
    table = dict((m.group(1), int(m.group(2))) for m in re.finditer(r'"([a-z0-9_]+)":\s+(\d+), // VMB_RF_', src))
    assert table == promql.ROLLUP_FUNCS
    hdr = open(os.path.join(ROOT, "include", "vmb200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr[hdr.index("enum vmb_rollup_func {"):hdr.index("VMB_RF__COUNT")], flags=re.S)
    enum = re.findall(r"\b(VMB_RF_[A-Z0-9_]+)\b", body)
    for name, idx in table.items():
        assert "// %s\n" % enum[idx] in src.split('"%s":' % name)[1].split("\n")[0] + "\n"
    go = open(os.path.join(ROOT, "integration", "go", "app", "vmselect", "promql", "eval_vmb200.go")).read()
    assert "vmb200FuncIDs[funcName]" in go and "ok" in go.split("vmb200FuncIDs[funcName]")[0].splitlines()[-1]  # checked lookup
