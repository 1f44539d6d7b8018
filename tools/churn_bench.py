"""The churn sub-record of bench.py on its own (template outlier threshold, 5 % of the ids replaced per frame).   python tools/churn_bench.py [periods=40]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
K = int(sys.argv[1]) if len(sys.argv) > 1 else 40
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
ses = bench.Session(torch, None, 512, K, 5, 0, 0, 1, bench.bench_settings(outlierThreshold=0.01), churn=0.05)
ms, nl = ses.device_timed(flush)
print(f"churn: {11 * K / (ms * 1e-3):.1f} steps/s, {ms / K:.3f} ms/period, graph stats {ses.f.graph_stats()}, landmarks {ses.f.numLandmarks}")
