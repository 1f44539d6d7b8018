#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 EqF-VIO filter hot path.

Metric (BASELINE.json): filter steps/sec with IMU @ 200 Hz + vision @ 20 Hz (a filter step = one
processIMUData or processVisionData call; `fastRiccati: false`, so every step carries a Riccati
propagate) at N tracked features, and the fp64 rate of the Sigma contractions against the measured
DGEMM peak.  A bench "step" is one vision period = 10 IMU ticks + 1 vision frame = 11 filter steps.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--features 512] [--impl reference]

N > 1 is launched by torchrun, one rank per GPU: the recursion does not shard, so every rank runs its
own session (seed 1000 + features + rank) and the ranks all-gather an 8-double pose record after each
vision update (NCCL).  `--impl reference` times the reference-equivalent CPU path (numpy restatement of
the reference with OpenBLAS for the dense products, oracle/eqvio_numpy.py) on the host cores.
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


def use_all_host_threads() -> int:
    """The CPU arm runs its dense products on every host core whatever the launcher exported (torchrun sets
    OMP_NUM_THREADS=1 for its workers, which would make the baseline ~10x slower than the box can do)."""
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=cores)
        got = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(got) if got else cores
    except Exception:
        return cores


def run_reference(args, rank, world):
    """Reference arm: the reference-equivalent CPU path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from eqf_vio_b200.settings import conditioned_settings
    from eqf_vio_b200.synthetic import period_sequence
    from oracle import eqvio_numpy as onp

    N, K, W = args.features, args.steps, args.warmup
    cores = use_all_host_threads()
    s = conditioned_settings()
    seq = period_sequence(N, W + K, camera_offset=tuple(s.cameraOffset))
    f = onp.VIOFilter(onp.Settings(**s.as_dict()))
    ev = list(seq.events())
    per, t_start, done = [], None, 0
    for kind, i in ev:
        if kind == "imu":
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
            if t_start is not None and i > W:
                per.append(time.perf_counter() - t_start)
            t_start = time.perf_counter()
            done = i
    total = sum(per)
    value = STEPS_PER_PERIOD * len(per) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": len(per), "warmup": W,
        "ms_per_step": 1e3 * total / len(per), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"N={N} features, n={11+3*N}, IMU 200 Hz / vision 20 Hz, template settings with outlierThreshold=1e9, initialSceneDepth=8, initialPointVariance=100",
                   "step": "one vision period = 10 IMU ticks + 1 vision frame = 11 filter steps"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": f"{len(per)} vision periods after {W} warm-up; numpy restatement of the reference, dense products on OpenBLAS with {cores} threads (Eigen3 is absent here, the reference binary cannot be built)"},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops_dense_equiv": flop_model(N)["period"] * len(per) / total / 1e9,
    }
    emit(line)


def cpu_baseline(N, periods):
    from eqf_vio_b200.settings import conditioned_settings
    from eqf_vio_b200.synthetic import period_sequence
    from oracle import eqvio_numpy as onp

    cores = use_all_host_threads()
    s = conditioned_settings()
    seq = period_sequence(N, periods + 1, camera_offset=tuple(s.cameraOffset))
    f = onp.VIOFilter(onp.Settings(**s.as_dict()))
    t_start, per = None, []
    for kind, i in seq.events():
        if kind == "imu":
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
            if t_start is not None and i > 1:
                per.append(time.perf_counter() - t_start)
            t_start = time.perf_counter()
    total = sum(per)
    return {"value": STEPS_PER_PERIOD * len(per) / total, "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"{len(per)} vision periods ({STEPS_PER_PERIOD * len(per)} filter steps) of the same N={N} workload after 1 warm-up period; numpy restatement of the reference, OpenBLAS {cores} threads",
            "ms_per_period": 1e3 * total / len(per)}


_REAL_STDOUT = None


def emit(line: dict):
    """The contract is ONE JSON line on stdout: native libraries (NCCL prints its version banner to fd 1 when the
    box sets NCCL_DEBUG) are kept off it by pointing fd 1 at stderr for the run and writing the line to the saved fd."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


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

    from eqf_vio_b200.filter import VIOFilter
    from eqf_vio_b200.sessions import gather_pose_records, session_seed
    from eqf_vio_b200.settings import conditioned_settings
    from eqf_vio_b200.synthetic import period_sequence

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    N, K, W = args.features, args.steps, args.warmup
    s = conditioned_settings()
    n_pass = 3
    seq = period_sequence(N, n_pass * (W + K) + 1, seed=session_seed(N, rank), camera_offset=tuple(s.cameraOffset))
    f = VIOFilter(s, device=local_rank)
    ext = torch.cuda.ExternalStream(f.stream_ptr(), device=dev)
    ydev = torch.tensor(seq.bearings, dtype=torch.float64, device=dev).contiguous()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2

    class PoseRec:  # the handle's device-resident pose record, viewed by torch without a copy
        __cuda_array_interface__ = {"shape": (8,), "typestr": "<f8", "data": (f.poseRecordDevicePtr(), False), "version": 3}

    pose_dev = torch.as_tensor(PoseRec(), device=dev)
    torch.cuda.synchronize()

    events = list(seq.events())
    # split into: init (up to and including vision 0), then one list of events per vision period
    periods, cur = [], []
    for kind, i in events:
        cur.append((kind, i))
        if kind == "vision":
            periods.append(cur)
            cur = []
    init, periods = periods[0], periods[1:]

    def imu(i):
        f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])

    for kind, i in init:
        imu(i) if kind == "imu" else f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
    f.synchronize()

    def run_period_resident(evs):
        for kind, i in evs:
            if kind == "imu":
                imu(i)
            else:
                f.processVisionDataDevice(seq.vision_stamps[i], seq.ids, ydev[i].data_ptr())
                if world > 1:
                    with torch.cuda.stream(ext):
                        gather_pose_records(pose_dev)

    def run_period_e2e(evs):
        out = None
        for kind, i in evs:
            if kind == "imu":
                imu(i)
            else:
                f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])  # host buffers in
                if world > 1:
                    with torch.cuda.stream(ext):
                        gather_pose_records(pose_dev)
                out = f.stateEstimate()  # D2H read of the result, as the reference's callers do (main.cpp:134)
        return out

    def flush_l2():
        with torch.cuda.stream(ext):
            flush.zero_()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    it = iter(periods)
    # ---------------- pass 1: inputs resident in HBM, device-timed (value) ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(W):
        run_period_resident(next(it))
    barrier()
    f.launch_count(reset=True)
    pairs = []
    for _ in range(K):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        run_period_resident(next(it))
        e1.record(ext)
        pairs.append((e0, e1))
    barrier()
    clocks = sampler.stop()
    launches = f.launch_count(reset=True)
    graph_replays, graphs_held = f.graph_stats()
    dev_ms = sum(a.elapsed_time(b) for a, b in pairs)
    dev_ms = max_over_ranks(dev_ms)

    # ---------------- pass 2: end to end through the host-buffer API, wall clock ----------------
    for _ in range(W):
        run_period_e2e(next(it))
    barrier()
    e2e_s = 0.0
    for _ in range(K):
        flush_l2()
        f.synchronize()
        t0 = time.perf_counter()
        run_period_e2e(next(it))
        f.synchronize()
        e2e_s += time.perf_counter() - t0
    barrier()
    e2e_s = max_over_ranks(e2e_s)
    n_state = f.numLandmarks
    h2d = 10 * 7 * 8 + 3 * N * 8                     # 10 IMU samples (kernel arguments) + one frame of bearings
    d2h = 248 + 8 * 8 * n_state                      # stateEstimate(): base state + 8 landmark fields x N

    # ---------------- pass 3: the GEMM kernel's own launches bracketed by CUDA events ----------------
    for _ in range(W):
        run_period_resident(next(it))
    f.synchronize()
    f.profile_enable(True)
    f.profile_read(reset=True)
    for _ in range(K):
        run_period_resident(next(it))
    classes = f.profile_read_classes(reset=False)
    g_launches, g_ms, g_flops = f.profile_read(reset=True)
    f.profile_enable(False)

    total_steps = STEPS_PER_PERIOD * K * world
    value = total_steps / (dev_ms * 1e-3)
    e2e_value = total_steps / e2e_s

    if rank == 0:
        # fp64 GEMM peak measured live on this GPU (MEASURED_PEAKS.json has no fp64 entry)
        a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        peak_tf = 2 * 8192**3 / (best * 1e-3) / 1e12
        del a, b
        fm = flop_model(N)
        # dominant kernel: the Riccati step's two Sigma contractions (F Sigma) F^T — one launch of the pair kernel when the
        # first product runs for more than a wave of CTA slots (or both fit one CTA per SM), else two launches of the
        # single-product kernel; same 32x32 DMMA tile code either way.  The update's GEMMs, the 64-deep panel / trailing
        # GEMMs of the Schur eliminations and their chain kernels are reported per class below.
        n_sigma = 11 + 3 * N
        t1 = ((n_sigma + 31) // 32) ** 2
        paired = (2 * t1 <= 148) or (t1 >= 1110)
        dom = [classes["riccati_gemm"]]
        d_ms, d_flops, d_launches = sum(c["ms"] for c in dom), sum(c["flops"] for c in dom), sum(c["launches"] for c in dom)
        achieved_tf = d_flops / (d_ms * 1e-3) / 1e12 if d_ms > 0 else 0.0
        kernel_name = ("eqvio::dgemm_pair_kernel<TileCfg<32,32,16,16,...>> (fp64 DMMA.8x8x4, TMA-staged; W = F Sigma and Sigma' = [W|T B R][F|B]^T + T P in one launch, "
                       "second product gated per row block of W)" if paired else
                       "eqvio::dgemm_dmma_tma_kernel<TileCfg<32,32,16,16,4,4>> (fp64 DMMA.8x8x4, TMA-staged), two launches per Riccati step: W = F Sigma, Sigma' = [W|T B R][F|B]^T + T P")
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01c_gemm_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(f"N{N}")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": f"N={N} features (n={11+3*N}), IMU 200 Hz / vision 20 Hz synthetic sequence, EQVIO_config_template.yaml settings with outlierThreshold=1e9 (N stays fixed), initialSceneDepth=8, initialPointVariance=100 (start-up matched to the 3-15 m synthetic scene)",
                "step": "one vision period = 10 IMU ticks + 1 vision frame = 11 filter steps, 11 Riccati propagates + 1 update",
                "sessions": f"{world} independent filter session(s), one per GPU" + (", NCCL all-gather of the 8-double pose record per vision frame" if world > 1 else ""),
                "l2": "flushed between bench steps (256 MiB write, untimed); per-period working set is ~170 MB at N=512",
                "association": "reference order: (F Sigma) F^T, (C Sigma) C^T, (Sigma C^T) S^-1, (K C) Sigma",
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "how": "host numpy buffers through eqvio_process_imu / eqvio_process_vision, stateEstimate() read back after every vision frame; wall clock between stream synchronisations"},
            "gpu_launches": launches,
            "cuda_graphs": {"replays_so_far": graph_replays, "instantiated": graphs_held,
                            "note": "gpu_launches counts this library's kernels, those inside replayed graphs included"},
            "roofline": {
                "bound": "tensor", "kernel": kernel_name,
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None,
                "traffic": traffic,
                "launches": d_launches, "avg_launch_ms": d_ms / d_launches if d_launches else None,
                "flops_per_launch_avg": d_flops / d_launches if d_launches else None,
                "how": "every launch of the Riccati contractions bracketed by CUDA events on its stream inside the library (eqvio_profile_enable, direct launches instead of graph replay) over K periods; achieved = executed 2MNK flops (4n^3 + 2n^2(n16-n+6) per step) / summed launch time",
                "algorithmic_bytes_per_launch": (4 if paired else 3) * 8 * n_sigma * n_sigma,
                "peak_source": "measured live: torch.matmul fp64 8192^3 (cuBLAS DGEMM), best of 5, same GPU; MEASURED_PEAKS.json holds no fp64 figure. DMMA pipe ceiling measured at 37.1 TFLOP/s (profiles/r01_dmma_microbench.md)",
                "kernel_share_of_step": d_ms / (dev_ms / world) if dev_ms else None,
                "all_gemm_launches": {"launches": g_launches, "ms": g_ms, "tflops": g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0,
                                      "note": "includes the 64-deep Schur panel / trailing GEMMs, which overlap each other on five streams"},
                "by_class": {k: {"launches": v["launches"], "ms_per_period": v["ms"] / K, "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0)}
                             for k, v in classes.items()},
                "by_class_note": "in-stream time between two events around each launch, summed per class: classes overlap each other (five update streams + the state stream), and a launch's time includes its wait for SM slots — the small state kernels of tick t+1 sit under the Riccati launch of tick t, which costs no wall time",
            },
            "gflops_dense_equiv": fm["period"] * K * world / (dev_ms * 1e-3) / 1e9,
            "flop_model_period_gflop": fm["period"] / 1e9,
        }
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
