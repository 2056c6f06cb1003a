"""Quick on-GPU probe: field-multiplier rates, MSM timings by size / window / lanes.  Not the bench."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tests import util_data as U  # noqa: E402
from zkp_b200 import Engine  # noqa: E402

eng = Engine(0)
out = {}
print("device", torch.cuda.get_device_name(0))
for kind, name in enumerate(["fe_mul 8x32", "fe_sq 8x32", "fe51_mul 5x51", "fe25_mul 10x25.5"]):
    r = eng.bench_field(kind, 4096)
    out[name] = r
    print("%-18s %.3e ops/s" % (name, r), flush=True)

K = 4096
base = np.frombuffer(b"".join(U.base_points(K)), dtype=np.uint8).reshape(K, 32)
rng = np.random.default_rng(1)


def make(n, bits=252):
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    nb = bits // 8
    sc[:, nb:] = 0
    if bits % 8:
        sc[:, nb] &= (1 << (bits % 8)) - 1
    idx = rng.integers(0, K, size=n)
    return sc, base[idx]


stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
eng.set_option("profile", 1)
res = torch.zeros(64, dtype=torch.uint8, device="cuda")


def time_dev(n, reps=3, **opts):
    sc, pt = make(n)
    dsc = torch.from_numpy(sc).cuda()
    dpt = torch.from_numpy(pt).cuda()
    for k, v in opts.items():
        eng.set_option(k, v)
    eng.msm_vartime_dev(dsc.data_ptr(), dpt.data_ptr(), n, res.data_ptr())
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.msm_vartime_dev(dsc.data_ptr(), dpt.data_ptr(), n, res.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = res.cpu().numpy()
    status = int(np.frombuffer(st[32:36].tobytes(), dtype=np.int32)[0])
    for k in opts:
        eng.set_option(k, 0)
    sm = eng.stage_ms()
    print("    stages:", " ".join("%s=%.3f" % (k, v) if isinstance(v, float) else "%s=%d" % (k, v) for k, v in sm.items()), flush=True)
    return best, status


for lg in [8, 12, 16, 18, 20, 22]:
    ms, st = time_dev(1 << lg)
    print("n=2^%d  %.3f ms  %.3e terms/s status=%d" % (lg, ms, (1 << lg) / ms * 1e3, st), flush=True)
    out["msm_2^%d_ms" % lg] = ms
eng.set_option("window_cap", 24)
for c in [13, 14, 15, 16, 17, 18, 19, 20]:
    for g in ([0] if c != 16 else [0, 32, 128, 1024]):
        try:
            ms, st = time_dev(1 << 22, reps=2, window=c, chunk=g)
            print("n=2^22 c=%d chunk=%d  %.3f ms status=%d" % (c, g, ms, st), flush=True)
            out["msm_2^22_c%d_s%d_ms" % (c, g)] = ms
        except Exception as ex:
            print("c=%d chunk=%d failed: %s" % (c, g, ex))
# CMZ-like mix at 2^24 terms: 13/24 full-size, 11/24 128-bit-negated
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
