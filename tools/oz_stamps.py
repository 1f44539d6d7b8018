"""Where the time of the int8 Riccati kernel goes: per-CTA clock64 stamps (EQVIO_OZ_STAMPS=1) of the last two launches of a short run.
    python tools/oz_stamps.py [features]"""
import ctypes as C, os, sys
os.environ["EQVIO_OZ_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eqf_vio_b200 import abi
from eqf_vio_b200.filter import VIOFilter
from eqf_vio_b200.settings import conditioned_settings
from eqf_vio_b200.synthetic import period_sequence

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
s = conditioned_settings(outlierThreshold=1e9)
seq = period_sequence(N, 3, camera_offset=tuple(s.cameraOffset))
f = VIOFilter(s, device=0)
ev = list(seq.events())
for kind, i in ev[:-3]:   # stop in the middle of a period: the last launches are steady-state ticks
    if kind == "imu":
        f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
    else:
        f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
f.synchronize()
cnt = C.c_size_t()
L = abi.lib()
abi.check(L.eqvio_oz_stamps(f._h, None, 0, C.byref(cnt)), "stamps")
buf = np.zeros(cnt.value, dtype=np.int64)
abi.check(L.eqvio_oz_stamps(f._h, buf.ctypes.data_as(C.POINTER(C.c_longlong)), cnt.value, C.byref(cnt)), "stamps")
n = 11 + 3 * N
T = ((n - 11) // 128) ** 2
os.environ.setdefault("EQVIO_GRAPHS", "1")
names = ["start", "jobs0 done", "acc0 full", "acc0 read", "acc1 full", "acc1 read", "fp64 stored", "row barrier passed", "emitted"]
blk = lambda i: buf[i * 1024 * 16:(i * 1024 + T) * 16].reshape(T, 16)
spans = []
for par in range(2):
    for ph in range(2):
        st = blk(2 * par + ph)
        spans.append((st[:, 11].min(), st[:, 15].max(), par, ph))
order = sorted(spans)
last = [x for x in order][-4:]
print("launch timeline of the last two steps (%globaltimer, us relative to the first launch's first CTA): first CTA started / last CTA ended")
t0 = last[0][0]
prev_end = None
for a, b, par, ph in last:
    gap = "" if prev_end is None else f"   gap to the previous launch's end {(a - prev_end) / 1e3:6.1f} us"
    print(f"  parity {par} phase {ph + 1}: {(a - t0) / 1e3:8.1f} -> {(b - t0) / 1e3:8.1f}  ({(b - a) / 1e3:6.1f} us){gap}")
    prev_end = b
par_last = last[-1][2]
for ph in range(2):
    st = blk(2 * par_last + ph)
    t0c = st[:, 0:1]
    rel = (st[:, :9] - t0c) / 1.965e3   # us at 1.965 GHz
    print(f"--- phase {ph + 1}: {T} CTAs, microseconds since the CTA's start (median / min / max over CTAs)")
    for k, nm in enumerate(names):
        print(f"  {nm:20s} {np.median(rel[:, k]):8.1f} {rel[:, k].min():8.1f} {rel[:, k].max():8.1f}")
    w = st[:, 9:11] / 1.965e3
    print(f"  issuer waited for operands: batch 0 {np.median(w[:, 0]):.1f} us, batch 1 {np.median(w[:, 1]):.1f} us (median); waits for acc_empty {np.median((st[:, 13] - st[:, 12]) / 1.965e3):.1f} us; all MMAs issued at {np.median((st[:, 14] - st[:, 0]) / 1.965e3):.1f} us")
