#!/usr/bin/env python
"""bench.py -- proofs verified/sec (batch) for the CMZ'13 10-attribute credential on N B200s.

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1) prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[3] scaled per GPU, weak scaling): every GPU batch-verifies 2^21 CMZ'13
`cred_show_10` proofs = ONE combined multi-scalar multiplication of 12 + 24 * 2^21 = 50,331,660 terms
(/root/reference/src/toolbox/batch_verifier.rs:219-230), so N = 8 is exactly configs[3] (2^24 proofs, one accept bit
per GPU gathered over NCCL).  A "step" is one such verification.  Inputs are synthetic but valid (data:
"synthetic"): coefficient scalars have the distribution batch_verifier.rs:173-206 produces (13 instance rows
full-size, 11 commitment rows = -rho with 128-bit rho), points are 2^22 distinct valid ristretto255 encodings
per GPU, and the scalars attached to each point cancel so the sum is the identity and every step must ACCEPT
(checked every step; a flipped scalar must REJECT, checked once).  Host transcript hashing and the coefficient
fold stay on the host by north_star and are outside this MSM-boundary metric.

  value     device-resident: scalars+points already in HBM (3.2 GB per GPU > 126 MB L2), CUDA-event timed
  e2e       same verification through the C ABI zkp_batch_verify() from pinned HOST buffers, H2D inside
  roofline  the dominant kernel (k_ingest2: decompression + digit sort, two launches per step, timed live with CUDA
            events inside the timed steps) against the HBM copy peak of MEASURED_PEAKS.json, plus the integer-pipe
            figures that actually bound this path
  cpu_baseline  the C port of the reference's serial u64 CPU algorithms (oracle/_ref) on all host threads
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L = 2**252 + 27742317777372353535851937790883648493
ROWS, NUM_S = 24, 12           # CMZ: 13 instance + 11 commitment rows, 12 static points
FULL_ROWS = 13                  # rows 0..12 carry full-size coefficients, rows 13..23 are -rho (128-bit rho)
BYTES_PER_PROOF = 1536          # 24 terms x (32 B scalar + 32 B point)   


def bucket_adds_per_proof(c):
    """Non-zero signed c-bit digits per proof: 13 full-size rows (253 bits after the sign fold) and 11 rows of -rho with
    128-bit rho (129 bits with the carry)."""
    return FULL_ROWS * -(-253 // c) + (ROWS - FULL_ROWS) * -(-129 // c)


def _l_words():
    return np.frombuffer(L.to_bytes(32, "little"), dtype=np.uint64).copy()


def _sub_from_l(x):
    """l - x for x (k,4) uint64 little-endian words with x < l; vectorised borrow chain."""
    lw = _l_words()
    out = np.empty_like(x)
    borrow = np.zeros(x.shape[0], dtype=np.uint64)
    for w in range(4):
        a = np.full(x.shape[0], lw[w], dtype=np.uint64)
        b = x[:, w]
        d = a - b
        b1 = (a < b).astype(np.uint64)
        d2 = d - borrow
        b2 = (d < borrow).astype(np.uint64)
        out[:, w] = d2
        borrow = b1 | b2
    return out


def _colsum_mod_l(x):
    """sum over axis 0 of (t, k, 4)-word scalars, mod l, returned as python ints per k (t is small)."""
    lo = (x & np.uint64(0xFFFFFFFF)).sum(axis=0, dtype=np.uint64)
    hi = (x >> np.uint64(32)).sum(axis=0, dtype=np.uint64)
    k = x.shape[1]
    out = []
    for i in range(k):
        v = 0
        for w in range(4):
            v += (int(lo[i, w]) + (int(hi[i, w]) << 32)) << (64 * w)
        out.append(v % L)
    return out


def make_cmz_batch(n_proofs, seed, n_points=None):
    """Synthetic valid batch-verification MSM instance (see module docstring).
    Returns (static_coeffs[12,32], instance_coeffs[24*N,32] row-major, point index arrays)."""
    rng = np.random.default_rng(seed)
    N = n_proofs
    K = n_points or 2 * N
    half = K // 2
    sc = rng.integers(0, 2**64, size=(ROWS, N, 4), dtype=np.uint64)
    sc[:FULL_ROWS, :, 3] &= np.uint64((1 << 60) - 1)          # < 2^252 < l
    # commitment rows: -rho mod l with 128-bit rho  (batch_verifier.rs:183)
    rho = sc[FULL_ROWS:].reshape(-1, 4).copy()
    rho[:, 2:] = 0
    sc[FULL_ROWS:] = _sub_from_l(rho).reshape(ROWS - FULL_ROWS, N, 4)
    # term (row j, proof p) uses point (p mod half) + (j & 1) * half; make each point's scalars cancel by fixing
    # the term in row 0 (even rows) / row 1 (odd rows) of proof p < half
    return sc, K, half


def finish_cancellation(sc, N, half):
    """Overwrite rows 0 and 1 for proofs < half so that the scalars of every point sum to 0 mod l."""
    for parity in (0, 1):
        rows = sc[parity::2]                                   # (12, N, 4)
        t = N // half
        grp = rows.reshape(rows.shape[0], t, half, 4).reshape(rows.shape[0] * t, half, 4).copy()
        grp[0] = 0                                             # the compensator slot (row `parity`, proofs < half)
        sums = _colsum_mod_l(grp)
        comp = np.zeros((half, 4), dtype=np.uint64)
        for i, s in enumerate(sums):
            v = (L - s) % L
            comp[i] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint64)
        sc[parity, :half] = comp
    return sc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5),
                              ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 500] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank (CPU affinity, hence first-touch placement of its pinned host buffers) to the NUMA node its GPU hangs
    off: with 8 ranks pulling 3.2 GB per step each from one host, cross-socket reads are what limits the end-to-end leg.
    Plumbing only; silently does nothing where sysfs does not tell."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return None


def cpu_backend():
    """Which restatement of the reference's CPU path this host can run: the 4-way vector one (the reference's
    `simd_backend`, oracle/c/ref_ifma.h: needs AVX-512 IFMA) or the serial u64 one (its default backend)."""
    from oracle import cref
    if cref.simd_available():
        return True, "simd_backend restatement (4-way AVX-512 IFMA point arithmetic, serial decompression)"
    return False, "u64 serial backend restatement (this CPU lacks avx512ifma: no vector backend)"


def cpu_baseline(sc_rows, pt_enc, sample_proofs, threads, steps=1, simd=None):
    """The C port of the reference's CPU path (oracle/_ref, kind "port") on a bounded sample of the SAME
    workload: the first `sample_proofs` proofs' columns of every row, i.e. one MSM of 24*sample terms.
    simd=None picks the vector backend when the CPU has it (the stronger baseline)."""
    from oracle import cref
    if simd is None:
        simd = cpu_backend()[0]
    s = np.ascontiguousarray(sc_rows[:, :sample_proofs]).reshape(-1, 32)
    p = np.ascontiguousarray(pt_enc[:, :sample_proofs]).reshape(-1, 32)
    best = 1e30
    for _ in range(steps):
        t = time.perf_counter()
        out = cref.msm_vartime(s, p, threads=threads, simd=simd)
        best = min(best, time.perf_counter() - t)
        assert out is not None
    return sample_proofs / best, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-proofs", type=int, default=21, help="proofs per GPU (default 2^21)")
    ap.add_argument("--cpu-sample-log2", type=int, default=16, help="proofs in the CPU baseline sample")
    ap.add_argument("--ref-sample-log2", type=int, default=17, help="proofs per step of the --impl reference arm")
    ap.add_argument("--window", type=int, default=0, help="tuning: force the Pippenger window width")
    ap.add_argument("--chunk", type=int, default=0, help="tuning: force the accumulate work-item length")
    ap.add_argument("--no-proofs-leg", action="store_true", help="skip the e2e_from_proofs leg (real proofs, device front end)")
    ap.add_argument("--chunk-terms-log2", type=int, default=0, help="tuning: H2D pipeline chunk (terms) for the e2e leg")
    ap.add_argument("--ingest-variant", type=int, default=-1, help="tuning: occupancy point of k_ingest2 (0..3)")
    ap.add_argument("--phase1-percent", type=int, default=0, help="tuning: share of point chunks in the first ingestion phase (e2e)")
    ap.add_argument("--no-numa-bind", action="store_true", help="tuning: do not pin ranks to their GPU's NUMA node")
    ap.add_argument("--bv-chunk-terms-log2", type=int, default=0, help="tuning: slab size (terms) of the from-proofs leg")
    ap.add_argument("--dual-stream", type=int, default=-1, help="tuning: chunk kernels of the e2e leg on two alternating streams (0/1)")
    ap.add_argument("--accumulate-variant", type=int, default=0, help="tuning: occupancy point of k_accumulate (4, 5, 6 resident blocks)")
    ap.add_argument("--scatter-batch", type=int, default=-1, help="tuning: scatter phase with four cursor atomics in flight (0/1)")
    ap.add_argument("--bv-phase1-rows", type=int, default=0, help="tuning: rows decompressed in phase 1 of the from-proofs leg")
    ap.add_argument("--bv-prep-stream", type=int, default=-1, help="tuning: front-end kernel of the from-proofs leg on its own stream (0/1)")
    ap.add_argument("--bv-prep-smem-kb", type=int, default=-1, help="tuning: residency cap of that kernel (unused dynamic shared memory, KB)")
    ap.add_argument("--no-fused-sort", action="store_true", help="tuning: separate scatter pass instead of the two-phase ingestion")
    ap.add_argument("--sweep", default="", help="tuning: comma list of windows; prints stage times per window and exits")
    ap.add_argument("--no-configs", action="store_true", help="skip the legs of the other BASELINE configs (`configs` block)")
    ap.add_argument("--dleq-log2", type=int, default=20, help="DLEQ proofs per GPU of the dleq_batch_verify leg")
    ap.add_argument("--prove-log2", type=int, default=16, help="CMZ proofs per GPU of the cmz_prove leg")
    ap.add_argument("--sweep-max-log2", type=int, default=22, help="largest size of the raw MSM sweep")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N = 1 << args.log2_proofs
    n_terms = NUM_S + ROWS * N
    host_threads = os.cpu_count() or 1
    config = {"workload": "CMZ13 cred_show_10 BatchVerifier::verify_batchable, 2^%d proofs per GPU "
                          "(one %d-term ristretto255 MSM per GPU; BASELINE configs[3] at 8 GPUs)" % (args.log2_proofs, n_terms),
              "proofs_per_gpu": N, "msm_terms_per_gpu": n_terms, "distinct_points_per_gpu": 2 * N,
              "l2_policy": "inputs (3.2 GB/GPU) and workspace exceed the 126 MB L2; no explicit flush",
              "parallelism": "proof-sharded x%d, 1 accept bit per GPU" % args.gpus}

    if args.impl == "reference":
        if rank != 0:
            return
        # the reference itself is Rust (no toolchain here): its CPU algorithms are timed through the C port
        # a bounded sample of the workload per step (the same 24-row coefficient mix, DISTINCT points, one MSM of
        # 24 * sample terms): 2^21 proofs per step would take ~20 s on these cores, i.e. minutes per run
        sample = 1 << args.ref_sample_log2
        sc, K, half = make_cmz_batch(sample, seed=1234, n_points=2 * sample)
        sc = finish_cancellation(sc, sample, half)
        pts = _host_points(2 * sample, distinct=2 * sample, threads=host_threads)   # valid encodings made on the CPU
        pidx = (np.arange(sample) % half)[None, :] + (np.arange(ROWS) & 1)[:, None] * half
        pt_rows = pts[pidx]
        sc_rows = sc.view(np.uint8).reshape(ROWS, sample, 32)
        for _ in range(args.warmup):
            cpu_baseline(sc_rows, pt_rows, min(sample, 1 << 12), host_threads)
        t0 = time.perf_counter()
        rates = [cpu_baseline(sc_rows, pt_rows, sample, host_threads)[0] for _ in range(args.steps)]
        dt = (time.perf_counter() - t0) / args.steps
        v = float(np.median(rates))
        simd, backend = cpu_backend()
        u64_v = cpu_baseline(sc_rows, pt_rows, sample, host_threads, simd=False)[0] if simd else v
        line = {"impl": "reference", "metric": "proofs verified/sec (batch) CMZ13 10-attr credential", "value": v,
                "unit": "proofs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": config,
                "same_workload_size": False, "sample_proofs_per_step": sample,
                "what": "the reference is Rust (no toolchain in this image): this arm times a C port of its CPU algorithms "
                        "(kind: port), on a %d-proof sample per step of the 2^%d-proof workload the GPU arm runs in full; "
                        "the rate is per proof, so the two are comparable, the sizes are not equal"
                        % (sample, args.log2_proofs),
                "cpu_baseline": {"value": v, "unit": "proofs/s", "cores": host_threads, "kind": "port",
                                 "backend": backend, "u64_serial_backend_value": u64_v,
                                 "sample": "%d proofs (%d-term MSM over %d distinct points) per step, C port of dalek's CPU "
                                           "algorithms (%s), %d threads"
                                           % (sample, ROWS * sample, 2 * sample,
                                              "vector backend" if simd else "u64 serial backend", host_threads)},
                "e2e": {"value": v, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 from C (NCCL prints its version
    # banner there) are sent to stderr for the duration of the run, and the line is written to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from zkp_b200 import Engine

    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa_bind else None
    host_threads = len(os.sched_getaffinity(0))     # what this rank may actually use
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = Engine(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)

    # ---- synthetic valid instance (per rank, different seed) ----------------------------------------------
    t_setup = time.perf_counter()
    sc, K, half = make_cmz_batch(N, seed=1000 + rank)
    sc = finish_cancellation(sc, N, half)
    pts = gpu_points(eng, K, seed=77 + rank)                       # (K,32) distinct valid encodings
    pidx = (np.arange(N) % half)[None, :] + (np.arange(ROWS) & 1)[:, None] * half
    inst_points = torch.from_numpy(pts)[torch.from_numpy(pidx.reshape(-1))].contiguous()   # (24N,32)
    inst_coeffs = torch.from_numpy(sc.view(np.uint8).reshape(ROWS * N, 32))
    static_points = torch.from_numpy(pts[:NUM_S].copy())
    static_coeffs = torch.zeros(NUM_S, 32, dtype=torch.uint8)       # static rows: zero net coefficient
    h_scal = torch.cat([static_coeffs, inst_coeffs]).pin_memory()
    h_pts = torch.cat([static_points, inst_points]).pin_memory()
    d_scal, d_pts = h_scal.cuda(non_blocking=True), h_pts.cuda(non_blocking=True)
    d_res = torch.zeros(64, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    def step_dev():
        eng.msm_vartime_dev(d_scal.data_ptr(), d_pts.data_ptr(), n_terms, d_res.data_ptr())

    def read_result():
        r = d_res.cpu().numpy()
        status, ident = np.frombuffer(r[32:40].tobytes(), dtype=np.int32)
        return int(status), int(ident)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.no_fused_sort:
        eng.set_option("fused_sort", 0)
    if args.bv_chunk_terms_log2:
        eng.set_option("bv_chunk_terms", 1 << args.bv_chunk_terms_log2)
    if args.dual_stream >= 0:
        eng.set_option("dual_stream", args.dual_stream)
    if args.phase1_percent:
        eng.set_option("phase1_percent", args.phase1_percent)
    if args.accumulate_variant:
        eng.set_option("accumulate_variant", args.accumulate_variant)
    if args.scatter_batch >= 0:
        eng.set_option("scatter_batch", args.scatter_batch)
    if args.bv_phase1_rows:
        eng.set_option("bv_phase1_rows", args.bv_phase1_rows)
    if args.bv_prep_stream >= 0:
        eng.set_option("bv_prep_stream", args.bv_prep_stream)
    if args.bv_prep_smem_kb >= 0:
        eng.set_option("bv_prep_smem_kb", args.bv_prep_smem_kb)
    if args.ingest_variant >= 0:
        eng.set_option("ingest_variant", args.ingest_variant)
    if args.chunk_terms_log2:
        eng.set_option("chunk_terms", 1 << args.chunk_terms_log2)
    if args.window:
        eng.set_option("window", args.window)
    if args.chunk:
        eng.set_option("chunk", args.chunk)
    if args.sweep:
        eng.set_option("profile", 1)
        for c in [int(x) for x in args.sweep.split(",")]:
            eng.set_option("window", c)
            step_dev()
            step_dev()
            st = eng.stage_ms()
            tot = sum(v for k, v in st.items() if k not in ("window", "lanes"))
            print("c=%d total=%.2f ms  " % (c, tot) + " ".join("%s=%.2f" % (k, v) for k, v in st.items()), flush=True)
            assert read_result() == (0, 1)
        return
    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    assert read_result() == (0, 1), "valid batch was not accepted: %r" % (read_result(),)
    # negative control: one flipped coefficient byte must reject
    d_bad = d_scal.clone()
    d_bad[NUM_S + 5, 0] ^= 1
    eng.msm_vartime_dev(d_bad.data_ptr(), d_pts.data_ptr(), n_terms, d_res.data_ptr())
    torch.cuda.synchronize()
    assert read_result() == (0, 0), "tampered batch was accepted"
    del d_bad

    # per-stage device times (CUDA events on the launching stream) for the roofline block
    eng.set_option("profile", 1)
    stage_acc = {}
    for _ in range(2):
        step_dev()
        for k, v in eng.stage_ms().items():
            stage_acc.setdefault(k, []).append(v)
    eng.set_option("profile", 0)
    stages = {k: float(np.mean(v)) for k, v in stage_acc.items()}

    # ---- timed region: device-resident ---------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(2):                 # keep the GPU busy while nvidia-smi starts sampling
        step_dev()
    launches0 = eng.launch_count
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    accepts = 0
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    launches = eng.launch_count - launches0
    ms_total = e0.elapsed_time(e1)
    live = eng.live_ms()            # events around the dominant kernels of the last TIMED step (always recorded)
    assert read_result() == (0, 1)

    # ---- e2e through the C ABI with host buffers -------------------------------------------------------------
    h_sc_np, h_pt_np = h_scal.numpy(), h_pts.numpy()
    for _ in range(1):
        ok, rc = eng.batch_verify(h_sc_np[:NUM_S], h_pt_np[:NUM_S], h_sc_np[NUM_S:], h_pt_np[NUM_S:], ROWS, N)
        assert ok and rc == 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ok, rc = eng.batch_verify(h_sc_np[:NUM_S], h_pt_np[:NUM_S], h_sc_np[NUM_S:], h_pt_np[NUM_S:], ROWS, N)
        accepts += int(ok)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    assert accepts == args.steps

    # register-resident field-op rates of this GPU (the calibration every integer-pipe fraction refers to)
    fe_sq_rate = eng.bench_field(9, 2048)                     # variable-time tails: what the verifier's kernels run
    fe_mul_rate = eng.bench_field(8, 2048)
    madd_rate = eng.bench_field(11, 512)
    fe_sq_ct, fe_mul_ct = eng.bench_field(1, 2048), eng.bench_field(0, 2048)   # branch-free tails: the prover's kernels
    configs = {}
    from tools import bench_legs as BL
    if not args.no_configs:
        # the H2D copies of one e2e step alone, all ranks at once: the ceiling of the e2e leg at N GPUs
        configs["h2d_ceiling"] = BL.h2d_ceiling([(d_scal, h_scal), (d_pts, h_pts)], args.steps, barrier, world)
        # ONE batch, ONE verdict over all ranks (partial sums over NCCL)
        configs["single_verdict"] = BL.single_verdict(eng, stream, d_scal, d_pts, n_terms, NUM_S, rank, world,
                                                      max(2, min(args.steps, 5)))

    # ---- e2e from PROOFS: 2^16 real CMZ proofs (made here by zkp_prove_batch = the cmz_prove leg, BASELINE configs[1])
    # tiled to 2^21 batch entries; transcripts, challenges, weights, coefficient fold and MSM all on the GPU
    # (zkp_batch_verify_proofs, SURVEY 8f f1+f2) ---------------------------------------------------------------------
    proofs_leg = None
    if not args.no_proofs_leg:
        from tools.workloads import cmz_instances
        from zkp_b200 import toolbox as PT
        eng.set_stream(None)
        n_real = min(N, 1 << args.prove_log2)
        st_cmz, sec, limbs, enc_pts = cmz_instances(eng, n_real, np.random.default_rng(500 + rank))
        entropy = np.random.default_rng(600 + rank).integers(0, 256, size=(n_real, 32), dtype=np.uint8)
        configs["cmz_prove"], (enc_p, com, resp) = BL.cmz_prove(eng, st_cmz, sec, limbs, entropy, (fe_sq_ct, fe_mul_ct), 3,
                                                               barrier, world)
        assert (enc_p == enc_pts).all()
        del limbs
        reps = N // n_real
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        inst_h = pin(np.tile(np.ascontiguousarray(enc_p[:, :13].transpose(1, 0, 2)), (1, reps, 1)))
        com_h = pin(np.tile(com, (reps, 1, 1)))
        resp_h = pin(np.tile(resp, (reps, 1, 1)))
        common_h = np.ascontiguousarray(enc_p[0, 13:])
        seed = bytes(range(32))
        st_cmz.batch_verify_device(eng, com_h, resp_h, b"CMZ", inst_h, common_h, seed)          # warm-up + must accept
        bad = resp_h[:n_real].copy()
        bad[7, 3, 0] ^= 1
        try:
            st_cmz.batch_verify_device(eng, com_h[:n_real], bad, b"CMZ", np.ascontiguousarray(inst_h[:, :n_real]),
                                       common_h, seed)
            raise SystemExit("tampered proof accepted by the device front end")
        except PT.VerificationFailure:
            pass
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st_cmz.batch_verify_device(eng, com_h, resp_h, b"CMZ", inst_h, common_h, seed)
        barrier()
        proofs_leg = {"seconds": (time.perf_counter() - t0) / args.steps,
                      "h2d_bytes_per_step": int(inst_h.nbytes + com_h.nbytes + resp_h.nbytes + common_h.nbytes),
                      "distinct_real_proofs": int(n_real)}
        del inst_h, com_h, resp_h
    # the CPU baseline's sample (rank 0 at N = 1): the first proofs' columns of every row of THIS workload
    cpu_sample = min(1 << args.cpu_sample_log2, N)
    cpu_sc_rows = cpu_pt_rows = None
    if rank == 0 and world == 1:
        cpu_sc_rows = np.ascontiguousarray(sc.view(np.uint8).reshape(ROWS, N, 32)[:, :cpu_sample])
        cpu_pt_rows = np.ascontiguousarray(inst_points.numpy().reshape(ROWS, N, 32)[:, :cpu_sample])
    if not args.no_configs:
        h_sc_np = h_pt_np = None
        del d_scal, d_pts, h_scal, h_pts, inst_points, inst_coeffs, sc
        torch.cuda.empty_cache()
        eng.set_stream(None)
        configs["dleq_batch_verify"] = BL.dleq_batch_verify(eng, stream, args.dleq_log2, max(2, min(args.steps, 5)), rank,
                                                            barrier, world)
        configs["raw_msm_sweep"] = BL.raw_msm_sweep(eng, stream, list(range(8, args.sweep_max_log2 + 1, 2)), rank, world, 18,
                                                    host_threads)

    # ---- gather: max time over ranks, accept bits over NCCL ---------------------------------------------------
    from zkp_b200 import parallel
    tm = torch.tensor([ms_total, e2e_s * 1e3, (proofs_leg["seconds"] if proofs_leg else 0.0) * 1e3],
                      dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    # the single NCCL gather of accept bits (one per GPU shard)
    accept_bitmap, verdict = parallel.gather_accept_bits(accepts == args.steps, device="cuda")
    assert verdict
    ms_total, e2e_ms, proofs_ms = float(tm[0].item()), float(tm[1].item()), float(tm[2].item())

    if rank == 0:
        ms_step = ms_total / args.steps
        value = world * N / (ms_step * 1e-3)
        e2e_value = world * N / (e2e_ms * 1e-3 / args.steps)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        dom = max((k for k in stages if k not in ("window", "lanes")), key=lambda k: stages[k])
        # The dominant kernel of a step is k_ingest2: two launches, each decompresses half of the points and reads all the
        # scalars (digit histogram in phase 1, counting-sort scatter in phase 2).  ALGORITHMIC bytes of one launch by
        # SURVEY 8(d) (64 B read per MSM term = 32 B scalar + 32 B encoding): n/2 encodings + n scalars.
        ing = [live.get("ingest_phase1", -1.0), live.get("ingest_phase2", -1.0)]
        ing_ms = float(np.mean(ing)) if min(ing) > 0 else 0.0
        ing_bytes = n_terms * (0.5 * 32.0 + 32.0)                # algorithmic bytes of ONE launch
        achieved = ing_bytes / (ing_ms * 1e-3) / 1e9 if ing_ms else 0.0
        # integer-pipe view: field mults per point (decode 3 S + 2 M, inverse square root 252 S + 13 M, coordinates 6 M, Niels
        # form 1 M = 255 S + 22 M; ge.cuh ristretto_decode) against the calibrated register-resident rates
        dec_time_at_peak = 0.5 * n_terms * (255.0 / fe_sq_rate + 22.0 / fe_mul_rate)
        acc_ms = live.get("accumulate", -1.0)
        traffic, traffic_each, traffic_src, pipe = None, None, None, None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
            prof = json.load(open(os.path.join(ROOT, "profiles", "r02_ingest_ncu.json")))
            scale = n_terms / prof["n_terms"]
            ing_l = [l for l in prof["launches"] if l["kernel"].startswith("k_ingest2")]
            traffic_each = [l["traffic_bytes"] * scale for l in ing_l]
            traffic = traffic_each[0]
            traffic_src = ("profiles/r02_ingest_ncu.json (ncu --set full at n=%d; `traffic` = the phase-1 launch, "
                           "`traffic_each` = both launches: the phase-2 launch carries the counting-sort scatter)" % prof["n_terms"])
            pipe = {l["kernel"]: {"fmaheavy_busy_pct": l["fmaheavy_pct"], "issue_active_pct": l["issue_active_pct"]}
                    for l in prof["launches"]}
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "k_ingest2 (2 launches per step: decompress half of the points + digit "
                                              "histogram / scatter of all scalars)",
                    "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": traffic, "traffic_each": traffic_each,
                    "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ing_bytes, "kernel_ms": ing_ms, "kernel_ms_each": ing,
                    "share_of_step": (sum(ing) / ms_step) if ing_ms else None,
                    "whole_path_hbm": {"bytes_per_proof": BYTES_PER_PROOF,
                                       "achieved_GBs": N * BYTES_PER_PROOF / (ms_step * 1e-3) / 1e9,
                                       "frac": N * BYTES_PER_PROOF / (ms_step * 1e-3) / 1e9 / hbm_peak},
                    "integer_pipe": {"what": "the path is bound by the FMA-heavy pipe (32x32->64 IMAD.WIDE: 85-86 % busy in "
                                             "k_ingest2 and k_accumulate, profiles/r02_ingest_ncu.json), not by HBM: kernel "
                                             "time against the register-resident field-op rates of this run",
                                     "fe_sq_per_s_calibrated": fe_sq_rate, "fe_mul_per_s_calibrated": fe_mul_rate,
                                     "madd_per_s_calibrated": madd_rate,
                                     "k_ingest2_frac_of_calibrated": (dec_time_at_peak / (ing_ms * 1e-3)) if ing_ms else None,
                                     "k_accumulate_ms": acc_ms if acc_ms > 0 else None,
                                     "k_accumulate_frac_of_calibrated":
                                         (bucket_adds_per_proof(int(stages.get("window", 19))) * N / madd_rate / (acc_ms * 1e-3))
                                         if acc_ms > 0 else None,
                                     "ncu_pipe_counters": pipe},
                    "stage_ms_unfused_profile_mode": stages, "dominant_stage_unfused": dom}
        sample, sc_rows, pt_rows = cpu_sample, cpu_sc_rows, cpu_pt_rows
        # the CPU baseline is taken at N = 1 only (the other ranks would compete for the same host cores)
        cpu_v, cpu_t = cpu_baseline(sc_rows, pt_rows, sample, host_threads) if world == 1 else (None, 0.0)
        cpu_1 = cpu_baseline(sc_rows, pt_rows, max(256, sample // 8), 1)[0] if world == 1 else None   # what `cargo bench` would see
        cpu_simd, cpu_backend_name = cpu_backend() if world == 1 else (False, "")
        cpu_u64 = (cpu_baseline(sc_rows, pt_rows, sample, host_threads, simd=False)[0] if cpu_simd else cpu_v) if world == 1 else None
        line = {"metric": "proofs verified/sec (batch) CMZ13 10-attr credential", "value": value, "unit": "proofs/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (8x32-bit saturated limbs, 64-bit products)",
                "data": "synthetic", "config": config,
                "e2e": {"value": e2e_value, "unit": "proofs/s", "h2d_bytes_per_step": int(n_terms * 64),
                        "d2h_bytes_per_step": 48, "ms_per_step": e2e_ms / args.steps,
                        "api": "zkp_batch_verify (C ABI), pinned host buffers"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "cpu_baseline": ({"value": cpu_v, "unit": "proofs/s", "cores": host_threads, "kind": "port",
                                  "single_thread_value": cpu_1, "backend": cpu_backend_name,
                                  "u64_serial_backend_value": cpu_u64,
                                  "sample": "first %d proofs of this workload (one %d-term MSM), C port of dalek's "
                                            "CPU algorithms (%s) sharded over %d threads, %.2f s"
                                            % (sample, ROWS * sample, "vector backend" if cpu_simd else "u64 serial backend",
                                               host_threads, cpu_t)} if world == 1 else
                                 {"value": None, "unit": "proofs/s", "cores": 0, "kind": "port",
                                  "sample": "taken at N = 1 only (bench.py --gpus 1)"}),
                "accept_bits": accept_bitmap, "setup_s": setup_s, "numa_binding_rank0": numa, "configs": configs}
        if proofs_leg:
            line["e2e_from_proofs"] = {
                "value": world * N / (proofs_ms * 1e-3), "unit": "proofs/s", "ms_per_step": proofs_ms,
                "h2d_bytes_per_step": proofs_leg["h2d_bytes_per_step"], "d2h_bytes_per_step": 48,
                "api": "zkp_batch_verify_proofs via zkph_batch_verify_device (C ABI), pinned host buffers",
                "what": "Merlin transcripts, challenges, per-proof weights, coefficient fold AND the MSM on the GPU "
                        "(SURVEY 8f rows f1+f2); %d distinct real CMZ proofs tiled to 2^%d batch entries with distinct "
                        "weights; accept checked, tampered response rejected" % (proofs_leg["distinct_real_proofs"], args.log2_proofs)}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def gpu_points(eng, K, seed):
    """K distinct valid ristretto255 encodings r_k * B computed by the engine's constant-time batched path."""
    rng = np.random.default_rng(seed)
    r = rng.integers(0, 256, size=(K, 32), dtype=np.uint8)
    r[:, 31] &= 0x0F
    B = np.frombuffer(bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76"), dtype=np.uint8)
    out = np.empty((K, 32), dtype=np.uint8)
    chunk = 1 << 20
    for lo in range(0, K, chunk):
        m = min(chunk, K - lo)
        pts = np.broadcast_to(B, (m, 32)).copy()
        off = np.arange(m + 1, dtype=np.uint64)
        out[lo:lo + m] = eng.msm_ct_batched(r[lo:lo + m], pts, off)
    return out


def _host_points(K, distinct=4096, threads=1):
    """K valid encodings without a GPU (reference arm): r*B for `distinct` random r via the C port, tiled."""
    from oracle import cref
    rng = np.random.default_rng(4242)
    d = min(K, distinct)
    r = rng.integers(0, 256, size=(d, 32), dtype=np.uint8)
    r[:, 31] &= 0x0F
    B = np.frombuffer(bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76"), dtype=np.uint8)
    out = np.empty((d, 32), dtype=np.uint8)
    threads = max(1, min(int(threads), 64))
    bounds = [d * i // threads for i in range(threads + 1)]

    def work(i):
        lo, hi = bounds[i], bounds[i + 1]
        if hi > lo:
            o, valid = cref.msm_vartime_batched(r[lo:hi], np.broadcast_to(B, (hi - lo, 32)).copy(),
                                                np.arange(hi - lo + 1, dtype=np.uint64))
            assert valid.all()
            out[lo:hi] = o
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]   # ctypes releases the GIL in the C call
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return out[np.arange(K) % d]


if __name__ == "__main__":
    main()
