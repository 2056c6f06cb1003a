"""Print the field / wide-multiply micro-benchmark rates (zkp_bench_field)."""
import json
import sys

sys.path.insert(0, ".")
from zkp_b200 import Engine  # noqa: E402

eng = Engine(0)
names = ["fe_mul 8x32", "fe_sq 8x32", "fe51_mul 5x51", "fe25_mul 10x25.5", "32 plain IMAD.WIDE / iter", "32 carry-chained IMAD.WIDE / iter",
         "fe_mul 8x32, shift-add reduce", "fe_sq 8x32, shift-add reduce", "fe_mul 8x32, vartime tail", "fe_sq 8x32, vartime tail",
         "mixed point addition (7M), ct tails", "mixed point addition (7M), vartime tails",
         "squaring chain, 8 int warps / block", "squaring chain, 6 int + 2 FP64 warps", "squaring chain, 5 int + 3 FP64 warps",
         "squaring chain, 4 int + 4 FP64 warps", "squaring chain, 8 FP64 warps"]
out = {}
for k, n in enumerate(names):
    r = eng.bench_field(k, 2048)
    out[n] = r
    extra = "  => %.3e wide mults/s" % (r * 32) if k in (4, 5) else ""
    print("%-36s %.4e /s%s" % (n, r, extra))
import ctypes
for mode, name in [(1, "integer warps only (4 of 8 warps/block)"), (2, "FP64 warps only"), (0, "both side by side")]:
    ms = ctypes.c_double(0)
    assert eng._lib.zkp_bench_dual(eng._ctx, mode, 2048, ctypes.byref(ms)) == 0
    out["dual_" + name] = ms.value
    print("dual: %-42s %.3f ms" % (name, ms.value))
json.dump(out, open("gpurun_out/field_rates.json", "w"), indent=1)
