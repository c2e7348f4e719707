"""victoriametrics_b200 -- B200-native block codec + rollup executor for VictoriaMetrics' query hot path.

Host-side mirror of the reference's Go API for this path (names follow the reference):
  encoding.marshal_values / unmarshal_values / marshal_timestamps / unmarshal_timestamps   (lib/encoding/encoding.go)
  decimal.append_decimal_to_float / append_float_to_decimal                                 (lib/decimal/decimal.go)
  storage.Block / Blocks / decode_blocks                                                    (lib/storage/block.go)
  promql.RollupConfig.do / get_rollup_configs / eval_rollup_func / IncrementalAggr          (app/vmselect/promql)
All compute goes through libvmb200.so (hand-written sm_100a CUDA behind the C ABI of include/vmb200.h).
"""
from . import _lib  # noqa: F401
from ._lib import Context, VmbError, default_context  # noqa: F401
from . import decimal, encoding, promql, storage  # noqa: F401
