"""Timing of the set passes alone: the fused cooperative kernel (so_sets_fused) against the three chained kernels, on Q / S of a
given size, with CUDA events (device time per call) and the host clock (call + wait).  Usage: python tools/time_sets.py"""
import os, sys, time
os.environ["SO_FUSED_DEBUG_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from safeopt_b200.engine import DeviceEngine

eng = DeviceEngine(max_gps=1)
rs = np.random.RandomState(0)
for M in [40_000, 250_000, 781_250, 6_250_000]:
    Qh = np.sort(rs.randn(M, 2), axis=1)
    Sh = (rs.rand(M) < 0.02).astype(np.uint8)
    Q, S, Mm = eng.to_device(Qh), eng.to_device(Sh), eng.zeros((M,), "u8")
    key, row = eng.empty((M,)), eng.empty((M,), "i64")
    rec_s, rec_m, cnt = eng.zeros((1, 64), "u8"), eng.zeros((1, 64), "u8"), eng.zeros((1,), "i64")
    sc, th = np.array([1.4]), np.array([0.4])

    def fused():
        return eng.sets_fused(Q, 1, 0, S, sc, th, True, Mm, key, row)

    def fused_async():
        eng.sets_fused(Q, 1, 0, S, sc, th, True, Mm, key, row, fetch=False)

    def chain():
        eng.reduce_safe(Q, 1, 0, S, rec_s)
        eng.maximizers_chain(Q, 1, 0, S, rec_s, 1, sc, Mm, rec_m)
        eng.candidates_chain(Q, 1, 0, S, Mm, rec_m, 1, sc, th, None, key, row, cnt)

    for name, fn, sync in [("fused+fetch", fused, False), ("fused kernel only", fused_async, True), ("3 chained kernels", chain, True)]:
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 200
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / n * 1e6
        print("M=%8d  %-18s device %.1f us/call   host %.1f us/call" % (M, name, e0.elapsed_time(e1) / n * 1e3, wall))
    # host cost of the launch call alone, and the in-kernel phase times of the last launch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fused_async()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    st = np.zeros(16, dtype=np.int64)
    eng.lib.so_debug_fused_times(eng.handle, st.ctypes.data)
    d = np.diff(st[:13].astype(float))
    print("   launch call %.1f us, then wait %.1f us; in-kernel ns (globaltimer; stamps of whichever block wrote last): scanA %d - %d pubA %d waitA %d scanB %d - %d pubB %d waitB %d scanC %d arrive %d pubwaitC %d copy %d  total %d" % (
        (t1 - t0) * 1e6, (t2 - t1) * 1e6, *d, float(st[12] - st[0])))
