"""host-side mirror of getRollupConfigs (rollup.go:374-504): which configs, tags, preFuncs and flags each function expands to
(no GPU needed: only the Python mirror is exercised)"""
import numpy as np
import pytest

from victoriametrics_b200 import promql


def _cfgs(name, **kw):
    return promql.get_rollup_configs_multi(name, 0, 1000, 100, 300, **kw)


def test_rollup_family_tags_and_prefuncs():
    for name, pre, rcr in (("rollup", None, False), ("rollup_rate", "deriv", True), ("rollup_deriv", "deriv", False),
                           ("rollup_increase", "delta", True), ("rollup_delta", "delta", False),
                           ("rollup_scrape_interval", "scrape_interval", False)):
        rcs = _cfgs(name)
        assert [(r.Func, r.TagValue) for r in rcs] == [("min_over_time", "min"), ("max_over_time", "max"), ("avg_over_time", "avg")]
        assert all(r.preFunc == pre and r.removeCounterResets == rcr and r.dropStaleNaNs for r in rcs), name
        one = _cfgs(name, tag="max")
        assert [(r.Func, r.TagValue) for r in one] == [("max_over_time", "")]      # rollup.go:421-426: tag given => no label
    with pytest.raises(ValueError):
        _cfgs("rollup", tag="median")
    assert _cfgs("rollup_rate")[0].MayAdjustWindow and not _cfgs("rollup_delta")[0].MayAdjustWindow  # rollup.go:199


def test_candlestick_aggr_over_time_quantiles():
    assert [r.TagValue for r in _cfgs("rollup_candlestick")] == ["open", "close", "low", "high"]
    assert [(r.Func, r.TagValue) for r in _cfgs("rollup_candlestick", tag="low")] == [("rollup_low", "low")]
    rcs = _cfgs("aggr_over_time", aggr_funcs=["min_over_time", "rate", "count_over_time"])
    assert [r.TagValue for r in rcs] == ["min_over_time", "rate", "count_over_time"]
    assert all(r.removeCounterResets for r in rcs)           # one rate() makes the shared preFunc remove resets (rollup.go:480)
    assert all(r.samplesScannedPerCall == 0 for r in rcs)    # looked up under "aggr_over_time" (rollup.go:394)
    with pytest.raises(ValueError):
        _cfgs("aggr_over_time", aggr_funcs=["quantile_over_time"])
    q = _cfgs("quantiles_over_time", phis=[0.5, 0.99, 1])
    assert [(r.Func, r.TagValue, float(np.asarray(r.args))) for r in q] == [
        ("quantile_over_time", "0.5", 0.5), ("quantile_over_time", "0.99", 0.99), ("quantile_over_time", "1", 1.0)]
    with pytest.raises(KeyError):
        _cfgs("rate")


def test_flag_word_carries_the_prefunc():
    rc = _cfgs("rollup_increase")[0]
    flags = rc._cfg().flags
    assert flags & promql.RC_PRE["delta"] and flags & promql.RC_REMOVE_COUNTER_RESETS and flags & promql.RC_DROP_STALE_NANS
    assert not flags & (promql.RC_PRE["deriv"] | promql.RC_PRE["scrape_interval"])
