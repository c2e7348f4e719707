"""Name tables shared by the tests: reference rollup function names -> ids, and their per-function flags
(rollup.go:24-108 rollupFuncs, :199 rollupFuncsCanAdjustWindow, :223 rollupFuncsRemoveCounterResets,
:238 rollupFuncsSamplesScannedPerCall).  The numeric ids are the enum order of include/vmb200.h (and of the oracle)."""

RF_IDS = ["default_rollup", "rate", "delta", "avg_over_time", "min_over_time", "max_over_time", "sum_over_time",
          "count_over_time", "quantile_over_time", "first_over_time", "last_over_time", "range_over_time",
          "sum2_over_time", "stddev_over_time", "stdvar_over_time", "ideriv", "idelta", "deriv", "increase_pure",
          "changes", "changes_prometheus", "resets", "increases_over_time", "integrate", "lag", "lifetime",
          "scrape_interval", "tmin_over_time", "tmax_over_time", "tfirst_over_time", "tlast_over_time",
          "tlast_change_over_time", "mode_over_time", "mad_over_time", "outlier_iqr_over_time", "zscore_over_time",
          "ascent_over_time", "descent_over_time", "distinct_over_time", "geomean_over_time", "predict_linear",
          "holt_winters", "hoeffding_bound_lower", "hoeffding_bound_upper", "duration_over_time", "count_le_over_time",
          "count_gt_over_time", "count_eq_over_time", "count_ne_over_time", "share_le_over_time", "share_gt_over_time",
          "share_eq_over_time", "sum_le_over_time", "sum_gt_over_time", "sum_eq_over_time", "present_over_time",
          "absent_over_time", "stale_samples_over_time", "median_over_time", "rate_over_sum", "delta_prometheus",
          "rate_prometheus", "rollup_open", "rollup_close", "rollup_high", "rollup_low"]
RF = {n: i for i, n in enumerate(RF_IDS)}
# aliases: several reference names share one implementation (rollup.go:24-108)
ALIASES = {"deriv_fast": "rate", "increase": "delta", "irate": "ideriv", "decreases_over_time": "resets",
           "timestamp": "tlast_over_time", "timestamp_with_name": "tlast_over_time",
           "increase_prometheus": "delta_prometheus"}
for a, b in ALIASES.items():
    RF[a] = RF[b]

# Go identifiers used as rollupConfig.Func in rollup_test.go -> reference function name
GO_FUNC = {"rollupFirst": "first_over_time", "rollupLast": "last_over_time", "rollupDefault": "default_rollup",
           "rollupDelta": "delta", "rollupCount": "count_over_time", "rollupMin": "min_over_time",
           "rollupMax": "max_over_time", "rollupSum": "sum_over_time", "rollupDeltaPrometheus": "delta_prometheus",
           "rollupIdelta": "idelta", "rollupLag": "lag", "rollupLifetime": "lifetime",
           "rollupScrapeInterval": "scrape_interval", "rollupChanges": "changes",
           "rollupChangesPrometheus": "changes_prometheus", "rollupResets": "resets", "rollupAvg": "avg_over_time",
           "rollupDerivSlow": "deriv", "rollupDerivFast": "rate", "rollupIderiv": "ideriv",
           "rollupStddev": "stddev_over_time", "rollupIntegrate": "integrate", "rollupDistinct": "distinct_over_time",
           "rollupModeOverTime": "mode_over_time", "rollupRateOverSum": "rate_over_sum",
           "rollupZScoreOverTime": "zscore_over_time", "rollupIncreasePure": "increase_pure"}

REMOVE_COUNTER_RESETS = {"increase", "increase_prometheus", "increase_pure", "irate", "rate", "rate_prometheus",
                         "rollup_increase", "rollup_rate"}
CAN_ADJUST_WINDOW = {"default_rollup", "deriv", "deriv_fast", "ideriv", "irate", "rate", "rate_over_sum", "rollup",
                     "rollup_candlestick", "rollup_deriv", "rollup_rate", "rollup_scrape_interval", "scrape_interval",
                     "timestamp"}
SAMPLES_SCANNED_PER_CALL = {"absent_over_time": 1, "count_over_time": 1, "default_rollup": 1, "delta": 2,
                            "delta_prometheus": 2, "deriv_fast": 2, "first_over_time": 1, "idelta": 2, "ideriv": 2,
                            "increase": 2, "increase_prometheus": 2, "increase_pure": 2, "irate": 2, "lag": 1,
                            "last_over_time": 1, "lifetime": 2, "present_over_time": 1, "rate": 2,
                            "rate_prometheus": 2, "scrape_interval": 2, "tfirst_over_time": 1, "timestamp": 1,
                            "timestamp_with_name": 1, "tlast_over_time": 1}
AGGR = {"sum": 0, "min": 1, "max": 2, "avg": 3, "count": 4, "sum2": 5, "geomean": 6, "any": 7, "group": 8}
