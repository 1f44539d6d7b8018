#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 EqF-VIO filter hot path.

Metric (BASELINE.json): filter steps/sec with IMU @ 200 Hz + vision @ 20 Hz (a filter step = one
processIMUData or processVisionData call; `fastRiccati: false`, so every step carries a Riccati
propagate) at N tracked features, and the fp64 rate of the Sigma contractions against the measured
DGEMM peak.  A bench "step" is one vision period = 10 IMU ticks + 1 vision frame = 11 filter steps.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--features 512] [--impl reference]

The headline line is N = 512 (the size the north-star target is quoted on).  The same JSON line carries one
sub-record per other BASELINE.json configuration under "configs" — N = 64 (config 2), N = 256 (config 3),
N = 1024 (config 4, fewer periods, stated) at one GPU, "config5_N256" (one N = 256 session per GPU) under
torchrun — plus the reference's other shipped operating points at N = 512: "fastRiccati" (EQVIO_config.yaml)
and "churn" (template outlierThreshold, 5 % of the features replaced in every frame).

N > 1 is launched by torchrun, one rank per GPU: the recursion does not shard, so every rank runs its
own session (seed 1000 + features + rank) and the ranks all-gather an 8-double pose record after each
vision update (NCCL, on the handle's gather stream — never in front of the next IMU tick).  `--impl reference`
times the reference-equivalent CPU path (numpy restatement of the reference with OpenBLAS for the dense
products, oracle/eqvio_numpy.py) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

STEPS_PER_PERIOD = 11  # 10 IMU ticks + 1 vision frame
METRIC = "filter steps/sec (IMU@200Hz+vision@20Hz, fastRiccati=false)"


def flop_model(N):
    """Dense flop counts of what the reference executes."""
    n, m, p, l = 11 + 3 * N, 2 * N, 5 + 3 * N, 3 * N
    P = 4 * n**3 + 18 * n**2 + 72 * n
    G = 2 * m * n**2 + 2 * m**2 * n + 2 * m**3 + 2 * n**2 * m + 2 * n * m**2 + 2 * n * m
    J = 2 * n**2 * m + 2 * n**3 + n**2
    L = 2 * p**3 + 2 * l * p**2 + 2 * l**2 * p + 16 * l**2
    return {"P": P, "G": G, "J": J, "L": L, "period": 11 * P + G + J + L}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--features", type=int, default=512)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-periods", type=int, default=3)
    ap.add_argument("--no-sub-configs", action="store_true", help="headline line only (no configs / fastRiccati / churn sub-records)")
    return ap.parse_args()


class ClockSampler:
    """SM clock, power and throttle reasons sampled every 20 ms through NVML while the timed region runs
    (the B200_PROFILING.md clocks line; falls back to polling nvidia-smi if pynvml is unavailable)."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.samples, self._stop, self.t, self.proc = index, [], threading.Event(), None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.nv is not None:
            self._stop.set()
            self.t.join(timeout=1)
            sm = [s[0] for s in self.samples]
            busy = [c for c in sm if c > 0.5 * max(sm)] if sm else []
            reasons = set()
            for _, _, rs in self.samples:
                for name, bit in self.REASONS.items():
                    if rs & bit:
                        reasons.add(name)
            return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": self.sm_max,
                    "power_w_max": max((s[1] for s in self.samples), default=None), "samples": len(sm), "reasons": sorted(reasons),
                    "how": "NVML, 20 ms period, over warm-up + timed region of the device-timed pass"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        self.proc.terminate()
        out = self.proc.communicate(timeout=2)[0]
        sm, mx, rs = [], [], set()
        for line in out.splitlines():
            c = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
                v = int(c[3], 16)
                for name, bit in self.REASONS.items():
                    if v & bit:
                        rs.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(rs), "how": "nvidia-smi -lms 100"}


def set_host_threads(n: int | None) -> int:
    """Thread count of the BLAS behind numpy for what follows (None = every host core).  The CPU arm sets it itself
    whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1 for its workers)."""
    cores = os.cpu_count() or 1
    want = cores if n is None else n
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=want)
        got = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(got) if got else want
    except Exception:
        return want


def bench_settings(**overrides):
    from eqf_vio_b200.settings import conditioned_settings

    return conditioned_settings(**overrides)


def config_dict(N: int, variant: str = "") -> dict:
    """The workload description, shared word for word by the B200 arm and the reference arm."""
    d = {
        "workload": f"N={N} features (n={11+3*N}), IMU 200 Hz / vision 20 Hz synthetic sequence, EQVIO_config_template.yaml settings with "
                    "outlierThreshold=1e9 (N stays fixed), initialSceneDepth=8, initialPointVariance=100 (start-up matched to the 3-15 m synthetic scene; "
                    "BASELINE.md section 7)",
        "step": "one vision period = 10 IMU ticks + 1 vision frame = 11 filter steps, 11 Riccati propagates + 1 update",
        "l2": "flushed between bench steps (256 MiB write, untimed)",
        "association": "reference order for (F Sigma) F^T, (C Sigma) C^T, (Sigma C^T) S^-1; the covariance update K C Sigma is evaluated as (K C) Sigma by the "
                       "reference arm and as K (C Sigma), with the C Sigma of the S formation, by the B200 arm (same matrix by associativity, 5e-15 apart)",
    }
    if variant:
        d["variant"] = variant
    return d


def cpu_periods_timed(N, periods, warm, threads, settings_overrides=None):
    """Times `periods` vision periods of the numpy restatement of the reference (after `warm` untimed ones) with
    `threads` BLAS threads (None = all cores)."""
    from eqf_vio_b200.synthetic import period_sequence
    from oracle import eqvio_numpy as onp

    cores = set_host_threads(threads)
    s = bench_settings(**(settings_overrides or {}))
    seq = period_sequence(N, periods + warm, camera_offset=tuple(s.cameraOffset))
    f = onp.VIOFilter(onp.Settings(**s.as_dict()))
    t_start, per = None, []
    for kind, i in seq.events():
        if kind == "imu":
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
            if t_start is not None and i > warm:
                per.append(time.perf_counter() - t_start)
            t_start = time.perf_counter()
    total = sum(per)
    return STEPS_PER_PERIOD * len(per) / total, cores, len(per), 1e3 * total / len(per)


def cpu_baseline(N, periods, warm=1, one_thread_periods=1, settings_overrides=None):
    """The reference-equivalent CPU path on the host cores: all cores (a courtesy to the CPU: the reference's own build
    is single-threaded Eigen) and, when asked, one thread (the reference as shipped)."""
    v, cores, k, ms = cpu_periods_timed(N, periods, warm, None, settings_overrides)
    out = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
           "sample": f"{k} vision period(s) ({STEPS_PER_PERIOD * k} filter steps) of the same N={N} workload after {warm} warm-up period(s); numpy restatement of the reference "
                     f"(oracle/eqvio_numpy.py), dense products on OpenBLAS with {cores} threads (Eigen3 is absent here, the reference binary cannot be built)",
           "ms_per_period": ms}
    if one_thread_periods > 0:
        v1, c1, k1, ms1 = cpu_periods_timed(N, one_thread_periods, min(warm, 1), 1, settings_overrides)
        out["one_thread"] = {"value": v1, "unit": "steps/s", "cores": c1, "ms_per_period": ms1,
                             "sample": f"{k1} vision period(s) with OpenBLAS limited to 1 thread: the reference as shipped is a 1-thread Eigen build (eqf_vio/CMakeLists.txt:26,34-43)"}
    set_host_threads(None)
    return out


def run_reference(args, rank, world):
    """Reference arm: the reference-equivalent CPU path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    N, K, W = args.features, args.steps, args.warmup
    value, cores, k, ms = cpu_periods_timed(N, K, W + 1, None)
    one = None
    if not args.no_sub_configs:
        v1, c1, k1, ms1 = cpu_periods_timed(N, 1, 1, 1)
        one = {"value": v1, "unit": "steps/s", "cores": c1, "ms_per_period": ms1,
               "sample": "1 vision period with OpenBLAS limited to 1 thread: the reference as shipped is a 1-thread Eigen build (eqf_vio/CMakeLists.txt:26,34-43)"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": k, "warmup": W,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_dict(N),
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": f"{k} vision periods after {W} warm-up; numpy restatement of the reference (oracle/eqvio_numpy.py), dense products on OpenBLAS with {cores} threads "
                                   "(Eigen3 is absent here, the reference binary cannot be built; oracle/_ref — the reference's own sources against an Eigen stand-in with plain-loop products — is slower and is not used)",
                         "one_thread": one},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops_dense_equiv": flop_model(N)["period"] * k / (ms * k * 1e-3) / 1e9,
    }
    emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    """The contract is ONE JSON line on stdout: native libraries (NCCL prints its version banner to fd 1 when the
    box sets NCCL_DEBUG) are kept off it by pointing fd 1 at stderr for the run and writing the line to the saved fd."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def riccati_kernel_name(N):
    n_sigma = 11 + 3 * N
    t1 = ((n_sigma + 31) // 32) ** 2
    if 324 <= t1 < 888:
        return True, ("eqvio::dgemm_pair_kernel<TileCfg<32,32,16,16,...>> with every tile split over 3-4 k-ranges (fp64 DMMA.8x8x4, TMA-staged; W = F Sigma and Sigma' = [W|T B R][F|B]^T + T P in one launch, "
                      "partials summed in a fixed order by the chunk that arrives last, second product gated per row block of W)")
    if (2 * t1 <= 148) or (t1 >= 888):
        return True, ("eqvio::dgemm_pair_kernel<TileCfg<32,32,16,16,...>> (fp64 DMMA.8x8x4, TMA-staged; W = F Sigma and Sigma' = [W|T B R][F|B]^T + T P in one launch, "
                      "second product gated per row block of W)")
    return False, "eqvio::dgemm_dmma_tma_kernel<TileCfg<32,32,16,16,4,4>> (fp64 DMMA.8x8x4, TMA-staged), two launches per Riccati step: W = F Sigma, Sigma' = [W|T B R][F|B]^T + T P"


class Session:
    """One filter session on this rank's GPU and the three timed passes over it."""

    def __init__(self, torch, dist, N, K, W, rank, local_rank, world, settings, churn=0.0, n_pass=3):
        from eqf_vio_b200.filter import VIOFilter
        from eqf_vio_b200.sessions import session_seed
        from eqf_vio_b200.synthetic import period_sequence

        self.torch, self.dist, self.N, self.K, self.W, self.world, self.rank = torch, dist, N, K, W, world, rank
        self.dev = torch.device("cuda", local_rank)
        self.seq = period_sequence(N, n_pass * (W + K) + 1, seed=session_seed(N, rank), camera_offset=tuple(settings.cameraOffset))
        self.f = VIOFilter(settings, device=local_rank)
        self.ext = torch.cuda.ExternalStream(self.f.stream_ptr(), device=self.dev)
        self.ydev = torch.tensor(self.seq.bearings, dtype=torch.float64, device=self.dev).contiguous()
        self.gather_stream = None
        if world > 1:
            gs, pub = self.f.posePublish()
            self.gather_stream = torch.cuda.ExternalStream(gs, device=self.dev)

            class PoseRec:  # the published pose record, viewed by torch without a copy
                __cuda_array_interface__ = {"shape": (8,), "typestr": "<f8", "data": (pub, False), "version": 3}

            self.pose_dev = torch.as_tensor(PoseRec(), device=self.dev)
        torch.cuda.synchronize()
        # per-frame id sets: all N ids, or (churn) a fraction of them replaced by fresh ids in every frame
        M = len(self.seq.vision_stamps)
        self.frame_sel = None
        if churn > 0:
            rng = np.random.default_rng(7 + N)
            k = max(1, int(round(churn * N)))
            ids = np.arange(N, dtype=np.int64)      # id carried by bearing slot j
            nxt = N
            self.frame_ids, self.frame_perm = [], []
            for j in range(M):
                if j > 0:
                    slots = rng.choice(N, size=k, replace=False)
                    ids[slots] = np.arange(nxt, nxt + k)
                    nxt += k
                order = np.argsort(ids, kind="stable")
                self.frame_ids.append(ids[order].astype(np.int32))
                self.frame_perm.append(order)
            self.y_host = [np.ascontiguousarray(self.seq.bearings[j][self.frame_perm[j]]) for j in range(M)]
            self.ydev = torch.tensor(np.stack(self.y_host), dtype=torch.float64, device=self.dev).contiguous()
        else:
            self.frame_ids = [self.seq.ids] * M
            self.y_host = [self.seq.bearings[j] for j in range(M)]
        periods, cur = [], []
        for kind, i in self.seq.events():
            cur.append((kind, i))
            if kind == "vision":
                periods.append(cur)
                cur = []
        init, self.periods = periods[0], periods[1:]
        for kind, i in init:
            self.imu(i) if kind == "imu" else self.f.processVisionData(self.seq.vision_stamps[i], self.frame_ids[i], self.y_host[i])
        self.f.synchronize()
        self.it = iter(self.periods)

    def imu(self, i):
        s = self.seq
        self.f.processIMUData(s.imu[i, 0], s.imu[i, 1:4], s.imu[i, 4:7])

    def gather(self):
        if self.world > 1:
            from eqf_vio_b200.sessions import gather_pose_records

            with self.torch.cuda.stream(self.gather_stream):
                gather_pose_records(self.pose_dev)

    def run_period_resident(self, evs):
        for kind, i in evs:
            if kind == "imu":
                self.imu(i)
            else:
                self.f.processVisionDataDevice(self.seq.vision_stamps[i], self.frame_ids[i], self.ydev[i].data_ptr())
                self.gather()

    def run_period_e2e(self, evs):
        out = None
        for kind, i in evs:
            if kind == "imu":
                self.imu(i)
            else:
                self.f.processVisionData(self.seq.vision_stamps[i], self.frame_ids[i], self.y_host[i])  # host buffers in
                self.gather()
                out = self.f.stateEstimate()  # D2H read of the result, as the reference's callers do (main.cpp:134)
        return out

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def device_timed(self, flush):
        """Pass 1: inputs resident in HBM, per-period CUDA events on the handle's stream, max over ranks."""
        torch, f, K, W = self.torch, self.f, self.K, self.W
        for _ in range(W):
            self.run_period_resident(next(self.it))
        self.barrier()
        f.launch_count(reset=True)
        pairs = []
        for _ in range(K):
            with torch.cuda.stream(self.ext):
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.ext)
            self.run_period_resident(next(self.it))
            e1.record(self.ext)
            pairs.append((e0, e1))
        self.barrier()
        launches = f.launch_count(reset=True)
        dev_ms = self.max_over_ranks(sum(a.elapsed_time(b) for a, b in pairs))
        return dev_ms, launches

    def end_to_end(self, flush):
        """Pass 2: host numpy buffers through the C ABI, stateEstimate() read back after every frame, wall clock."""
        torch, f, K, W = self.torch, self.f, self.K, self.W
        for _ in range(W):
            self.run_period_e2e(next(self.it))
        self.barrier()
        e2e_s = 0.0
        for _ in range(K):
            with torch.cuda.stream(self.ext):
                flush.zero_()
            f.synchronize()
            t0 = time.perf_counter()
            self.run_period_e2e(next(self.it))
            f.synchronize()
            e2e_s += time.perf_counter() - t0
        self.barrier()
        n_state = f.numLandmarks
        h2d = 10 * 7 * 8 + 3 * self.N * 8            # 10 IMU samples (kernel arguments) + one frame of bearings
        d2h = 248 + 8 * 8 * n_state                  # stateEstimate(): base state + 8 landmark fields x N
        return self.max_over_ranks(e2e_s), h2d, d2h

    def profiled(self):
        """Pass 3: every GEMM / chain launch bracketed by CUDA events inside the library (direct launches)."""
        f, K, W = self.f, self.K, self.W
        for _ in range(W):
            self.run_period_resident(next(self.it))
        f.synchronize()
        f.profile_enable(True)
        f.profile_read(reset=True)
        for _ in range(K):
            self.run_period_resident(next(self.it))
        classes = f.profile_read_classes(reset=False)
        g = f.profile_read(reset=True)
        f.profile_enable(False)
        return classes, g

    def close(self):
        self.f.close()


def int8_roofline_record(N, K, classes, g, dev_ms_local, peak_tf, peak_i8, slices):
    """Roofline of the dominant kernel when the Riccati products run on the int8 tensor cores: executed int8 operations
    (2 M N K per slice pair, S (S + 1) / 2 pairs per fp64 product) against the int8 dense peak."""
    g_launches, g_ms, g_flops = g
    dom = classes["riccati_i8_gemm"]
    d_ms, d_flops, d_launches = dom["ms"], dom["flops"], dom["launches"]
    pairs = slices * (slices + 1) // 2
    tops = d_flops * pairs / (d_ms * 1e-3) / 1e12 if d_ms > 0 else 0.0
    eq_tf = d_flops / (d_ms * 1e-3) / 1e12 if d_ms > 0 else 0.0
    n_sigma = 11 + 3 * N
    mc = (n_sigma - 11) // 128 * 128
    kpad = (n_sigma + 31) // 32 * 32
    whole = {k: classes[k] for k in ("riccati_gemm", "riccati_i8_gemm")}
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_oz_gemm_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"N{N}")
            traffic_src = "NOT measured by this run: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this kernel (profiles/r02_oz_gemm_traffic.json; cold L2)"
        except Exception:
            traffic = None
    return {
        "bound": "tensor",
        "kernel": f"eqvio::k_oz_riccati<{slices}> (tcgen05.mma.cta_group::1.kind::i8 128x128x32, int32 accumulators in TMEM, bulk-copy staged pre-swizzled int8 slice tiles): "
                  f"two launches per Riccati step, W = F Sigma and Sigma' = W F^T + T B R B^T + T P; the {mc} x {mc} landmark block of each fp64 product is an exact sum of {pairs} int8 products "
                  f"({slices} slices of 7 bits per operand), the rows / columns in front of it are fp64 dots in the same launch, and the epilogue emits the result as the next product's int8 operand",
        "achieved": tops, "peak": peak_i8["value"], "unit": "TOP/s (int8 dense)", "frac": tops / peak_i8["value"] if peak_i8["value"] else None,
        "peak_source": peak_i8["source"],
        "traffic": traffic, "traffic_source": traffic_src,
        "launches": d_launches, "avg_launch_ms": d_ms / d_launches if d_launches else None,
        "int8_ops_per_launch": 2.0 * mc * mc * n_sigma * pairs,
        "algorithmic_bytes_per_launch": 3 * slices * mc * kpad + 8 * mc * mc,
        "algorithmic_bytes_note": "both operands' int8 slice arrays read once + the fp64 result block and its int8 slices (the next product's operand) written once",
        "fp64_equivalent": {"achieved_tflops": eq_tf, "dgemm_peak_tflops": peak_tf, "ratio": eq_tf / peak_tf if peak_tf else None,
                            "note": "2 M N K of the fp64 product the int8 launches stand for, over their time; the DMMA pipe tops out at 37.1 TFLOP/s (profiles/r01_dmma_microbench.md)"},
        "riccati_step_fp64_equivalent_tflops": (sum(v["flops"] for v in whole.values()) / (K * STEPS_PER_PERIOD)) / 1e12,
        "how": "every launch of the int8 products bracketed by CUDA events on its stream inside the library (eqvio_profile_enable, direct launches instead of graph replay) over K periods; "
               "achieved = executed int8 operations / summed launch time",
        "kernel_share_of_step": d_ms / dev_ms_local if dev_ms_local else None,
        "all_gemm_launches": {"launches": g_launches, "ms": g_ms, "tflops": g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0,
                              "note": "fp64-equivalent; includes the 64-deep Schur panel / trailing GEMMs, which overlap each other on five streams"},
        "by_class": {k: {"launches": v["launches"], "ms_per_period": v["ms"] / K, "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0)}
                     for k, v in classes.items()},
        "by_class_note": "in-stream time between two events around each launch, summed per class: classes overlap each other (five update streams + the state stream), and a launch's time includes its wait for SM slots; riccati_i8_gemm tflops are fp64-equivalent",
    }


def roofline_record(N, K, classes, g, dev_ms_local, peak_tf, peak_i8=None, slices=0):
    if slices > 0 and classes.get("riccati_i8_gemm", {}).get("launches", 0) > 0:
        return int8_roofline_record(N, K, classes, g, dev_ms_local, peak_tf, peak_i8, slices)
    g_launches, g_ms, g_flops = g
    paired, kernel_name = riccati_kernel_name(N)
    dom = classes["riccati_gemm"]
    d_ms, d_flops, d_launches = dom["ms"], dom["flops"], dom["launches"]
    achieved_tf = d_flops / (d_ms * 1e-3) / 1e12 if d_ms > 0 else 0.0
    n_sigma = 11 + 3 * N
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01c_gemm_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"N{N}")
            traffic_src = "NOT measured by this run: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this kernel (profiles/r01c_gemm_traffic.json; ncu flushes L2 before each replay, so it is the cold-cache figure)"
        except Exception:
            traffic = None
    return {
        "bound": "tensor", "kernel": kernel_name,
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None,
        "traffic": traffic, "traffic_source": traffic_src,
        "launches": d_launches, "avg_launch_ms": d_ms / d_launches if d_launches else None,
        "flops_per_launch_avg": d_flops / d_launches if d_launches else None,
        "how": "every launch of the Riccati contractions bracketed by CUDA events on its stream inside the library (eqvio_profile_enable, direct launches instead of graph replay) over K periods; achieved = executed 2MNK flops (4n^3 + 2n^2(n16-n+6) per step) / summed launch time",
        "algorithmic_bytes_per_launch": (4 if paired else 3) * 8 * n_sigma * n_sigma,
        "peak_source": "measured live: torch.matmul fp64 8192^3 (cuBLAS DGEMM), best of 5, same GPU; MEASURED_PEAKS.json holds no fp64 figure. DMMA pipe ceiling measured at 37.1 TFLOP/s (profiles/r01_dmma_microbench.md)",
        # this rank's Riccati launch time over this rank's device-timed region (both per K periods)
        "kernel_share_of_step": d_ms / dev_ms_local if dev_ms_local else None,
        "all_gemm_launches": {"launches": g_launches, "ms": g_ms, "tflops": g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0,
                              "note": "includes the 64-deep Schur panel / trailing GEMMs, which overlap each other on five streams"},
        "by_class": {k: {"launches": v["launches"], "ms_per_period": v["ms"] / K, "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0)}
                     for k, v in classes.items()},
        "by_class_note": "in-stream time between two events around each launch, summed per class: classes overlap each other (five update streams + the state stream), and a launch's time includes its wait for SM slots — the small state kernels of tick t+1 sit under the Riccati launch of tick t, which costs no wall time",
    }


def measure_peak_tf(torch, dev):
    """fp64 GEMM peak measured live on this GPU (MEASURED_PEAKS.json has no fp64 entry)."""
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * 8192**3 / (best * 1e-3) / 1e12


def measure_peak_int8(torch, dev):
    """int8 tensor-core peak for the roofline of the int8 Riccati products: cuBLASLt int8 GEMM (torch._int_mm) 8192^3 measured
    live; else twice the measured dense bf16 figure of MEASURED_PEAKS.json (the architecture's int8 : bf16 ratio); else nominal."""
    try:
        a = torch.randint(-64, 64, (8192, 8192), dtype=torch.int8, device=dev)
        b = torch.randint(-64, 64, (8192, 8192), dtype=torch.int8, device=dev)
        torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch._int_mm(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        live = 2 * 8192**3 / (best * 1e-3) / 1e12
    except Exception:
        live = None
    mp = None
    try:
        mp = 2.0 * float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    if live is not None and (mp is None or live >= 0.5 * mp):
        return {"value": max(live, mp or 0.0), "source": f"max of: torch._int_mm 8192^3 (cuBLASLt int8 GEMM) measured live, best of 8: {live:.0f} TOP/s; 2 x MEASURED_PEAKS.json bf16_tflops (burst): {mp if mp else float('nan'):.0f} TOP/s"}
    if mp is not None:
        return {"value": mp, "source": "2 x bf16_tflops (burst) of MEASURED_PEAKS.json: no int8 entry there, int8 dense rate is twice bf16 on sm_100"}
    return {"value": 4500.0, "source": "nominal B200 dense int8 (no measured figure available)"}


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # The only collective is a 64-byte all-gather per vision frame: one NCCL channel (one CTA) is plenty, and every SM NCCL does not
        # occupy matters — the int8 Riccati kernel needs 144 whole SMs of 148 (4 GPUs: 12362 -> 12528 steps/s with one channel).
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "1")
        os.environ.setdefault("NCCL_MIN_NCHANNELS", "1")
        dist.init_process_group("nccl", device_id=dev)

    N, K, W = args.features, args.steps, args.warmup
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2
    peak_tf = measure_peak_tf(torch, dev) if rank == 0 else 0.0
    peak_i8 = measure_peak_int8(torch, dev) if rank == 0 else {"value": 0.0, "source": ""}

    # ---------------- headline: N features, fastRiccati = false, fixed N ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ses = Session(torch, dist, N, K, W, rank, local_rank, world, bench_settings())
    dev_ms, launches = ses.device_timed(flush)
    graph_replays, graphs_held = ses.f.graph_stats()
    e2e_s, h2d, d2h = ses.end_to_end(flush)
    clocks = sampler.stop()
    classes, g = ses.profiled()
    slices = ses.f.riccati_int8_slices()
    ses.close()
    total_steps = STEPS_PER_PERIOD * K * world
    value = total_steps / (dev_ms * 1e-3)
    e2e_value = total_steps / e2e_s
    fm = flop_model(N)
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "arithmetic": ("state, Sigma and every result in IEEE fp64; the Riccati step's two n x n x n contractions are assembled exactly from int8 tensor-core products "
                           f"({slices} slices of 7 bits per operand = 55 mantissa bits, int32 accumulation, fp64 recombination: error 7e-16 of |row| |column|, as an fp64 GEMM)"
                           if slices > 0 else "IEEE fp64 throughout (DMMA tensor instructions for the Sigma contractions)"),
            "data": "synthetic",
            "config": config_dict(N),
            "sessions": f"{world} independent filter session(s), one per GPU" + (", NCCL all-gather of the 8-double pose record per vision frame on the handle's gather stream" if world > 1 else ""),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "how": "host numpy buffers through eqvio_process_imu / eqvio_process_vision, stateEstimate() read back after every vision frame; wall clock between stream synchronisations"},
            "gpu_launches": launches,
            "cuda_graphs": {"replays_so_far": graph_replays, "instantiated": graphs_held,
                            "note": "gpu_launches counts this library's kernels, those inside replayed graphs included"},
            "roofline": roofline_record(N, K, classes, g, dev_ms, peak_tf, peak_i8, slices),
            "gflops_dense_equiv": fm["period"] * K * world / (dev_ms * 1e-3) / 1e9,
            "flop_model_period_gflop": fm["period"] / 1e9,
        }

    def sub_record(Ns, Ks, Ws, settings, variant, churn=0.0, with_roofline=True, cpu=None):
        """value / e2e (/ roofline / cpu_baseline) of another configuration, same passes as the headline."""
        sp = ClockSampler(local_rank)
        sp.start()
        s2 = Session(torch, dist, Ns, Ks, Ws, rank, local_rank, world, settings, churn=churn)
        ms, nl = s2.device_timed(flush)
        replays, held = s2.f.graph_stats()
        es, hi, ho = s2.end_to_end(flush)
        ck = sp.stop()
        rec = {"config": config_dict(Ns, variant), "n_gpus": world, "steps": Ks, "warmup": Ws,
               "value": STEPS_PER_PERIOD * Ks * world / (ms * 1e-3), "unit": "steps/s", "ms_per_step": ms / Ks,
               "e2e": {"value": STEPS_PER_PERIOD * Ks * world / es, "unit": "steps/s", "h2d_bytes_per_step": hi, "d2h_bytes_per_step": ho},
               "gpu_launches": nl, "cuda_graph_replays": replays, "landmarks_at_end": s2.f.numLandmarks,
               "clocks": {k: ck.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples")}}
        if with_roofline:
            cl, gg = s2.profiled()
            if rank == 0:
                r = roofline_record(Ns, Ks, cl, gg, ms, peak_tf, peak_i8, s2.f.riccati_int8_slices())
                rec["roofline"] = {k: r[k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "launches", "avg_launch_ms", "kernel_share_of_step", "algorithmic_bytes_per_launch", "fp64_equivalent") if k in r}
                rec["roofline"]["by_class"] = r["by_class"]
        s2.close()
        if cpu is not None and rank == 0:
            rec["cpu_baseline"] = cpu_baseline(Ns, **cpu)
            rec["speedup_vs_cpu_baseline_e2e"] = rec["e2e"]["value"] / rec["cpu_baseline"]["value"]
        return rec

    if not args.no_sub_configs:
        configs = {}
        if world == 1:
            for Ns in (64, 256, 1024):
                if Ns == N:
                    continue
                big = Ns >= 1024
                Ks, Ws = (min(K, 10), 3) if big else (K, W)
                cpu = {"periods": 1, "warm": 0, "one_thread_periods": 0} if big else {"periods": args.cpu_periods, "warm": 1, "one_thread_periods": 1}
                rec = sub_record(Ns, Ks, Ws, bench_settings(), "", cpu=None if args.no_cpu_baseline else cpu)
                if big:
                    rec["truncated"] = f"{Ks} timed vision periods ({Ks * 0.05:.2f} s of the 60 s sequence of BASELINE config 4) after {Ws} warm-up; CPU baseline: 1 period, no warm-up, all cores only"
                configs[f"N{Ns}"] = rec
        else:
            # BASELINE config 5: one N = 256 session per GPU + NCCL pose gather
            configs["config5_N256"] = sub_record(256, K, W, bench_settings(), f"{world} sessions x N=256, one per GPU, NCCL pose all-gather per frame")
        # the reference's other shipped operating points, at the headline size
        fast = sub_record(N, K, W, bench_settings(fastRiccati=True), "fastRiccati: true (EQVIO_config.yaml:18): one Riccati propagate per vision period")
        fast["flop_model_period_gflop"] = (fm["P"] + fm["G"] + fm["J"] + fm["L"]) / 1e9
        churn = sub_record(N, K, W, bench_settings(outlierThreshold=0.01), "outlierThreshold 0.01 (template value) and 5 % of the features replaced by new ids in every frame (landmark churn: removeOldLandmarks / removeOutliers / addNewLandmarks run every frame)",
                           churn=0.05, with_roofline=False)
        if rank == 0:
            line["configs"] = configs
            line["fastRiccati"] = fast
            line["churn"] = churn
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_baseline(N, args.cpu_periods)
            line["cpu_baseline"] = cb
            line["speedup_vs_cpu_baseline_e2e"] = e2e_value / cb["value"]
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
