"""Print the field / wide-multiply micro-benchmark rates (zkp_bench_field)."""
import json
import sys

sys.path.insert(0, ".")
from zkp_b200 import Engine  # noqa: E402

eng = Engine(0)
names = ["fe_mul 8x32", "fe_sq 8x32", "fe51_mul 5x51", "fe25_mul 10x25.5", "32 plain IMAD.WIDE / iter", "32 carry-chained IMAD.WIDE / iter"]
out = {}
for k, n in enumerate(names):
    r = eng.bench_field(k, 2048)
    out[n] = r
    extra = "  => %.3e wide mults/s" % (r * 32) if k >= 4 else ""
    print("%-36s %.4e /s%s" % (n, r, extra))
json.dump(out, open("gpurun_out/field_rates.json", "w"), indent=1)
