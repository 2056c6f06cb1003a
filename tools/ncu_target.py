"""Small deterministic workload for ncu captures: field micro-kernels + one 2^20-term MSM (+ optional CMZ mix)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from zkp_b200 import Engine  # noqa: E402

eng = Engine(0)
if "--field" in sys.argv:
    for kind in range(4):
        eng.bench_field(kind, 512)
n = 1 << int(sys.argv[sys.argv.index("--log2") + 1]) if "--log2" in sys.argv else 1 << 20
rng = np.random.default_rng(5)
B = np.frombuffer(bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76"), dtype=np.uint8)
K = 1 << 14
r = rng.integers(0, 256, size=(K, 32), dtype=np.uint8)
r[:, 31] &= 0x0F
pts = eng.msm_ct_batched(r, np.broadcast_to(B, (K, 32)).copy(), np.arange(K + 1, dtype=np.uint64))
sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
sc[:, 31] &= 0x0F
P = pts[rng.integers(0, K, size=n)]
for _ in range(2):
    enc, ident, _ = eng.msm_vartime(sc, P)
print("done", enc.hex())
