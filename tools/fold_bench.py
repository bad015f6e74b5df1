#!/usr/bin/env python
"""Host-side microbench of jr_fsm_fold / jr_fsm_fold_mt on a synthetic batch shaped like config #3's steady state
(per group: the leader's APPLY + NOTIFY + PATTERN records and one masked follower APPLY record).  No GPU needed.

usage: fold_bench.py [groups] [replicas]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from josefine_b200.raft import load_engine_library  # noqa: E402


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    lib = load_engine_library()
    dt = [("group", "<u4"), ("hdr", "<u4"), ("id0", "<u4"), ("addr", "<u4"), ("tok0", "<u8"), ("stride", "<u8")]
    kinds = np.tile(np.array([0, 1, 2], dtype=np.uint32), G)
    lead = np.zeros(3 * G, dtype=dt)
    lead["group"] = np.repeat(np.arange(G, dtype=np.uint32), 3)
    lead["hdr"] = kinds | (0 << 2) | (64 << 8)
    lead["id0"] = 1000
    lead["addr"] = np.where(kinds == 1, 3 << 16, 0)
    foll = np.zeros(G, dtype=dt)
    foll["group"] = np.arange(G, dtype=np.uint32)
    foll["hdr"] = 0 | (1 << 2) | (64 << 8)
    foll["id0"] = 990
    foll["addr"] = (1 << R) - 2
    rec = np.concatenate([lead, foll])
    n, p = len(rec), rec.ctypes.data
    ref = None
    for th in (1, 2, 4, 8, 16):
        applied = (C.c_uint32 * (G * R))()
        tot = (C.c_uint64 * 3)()
        assert lib.jr_fsm_fold_mt(C.c_void_p(p), C.c_size_t(n), G, R, applied, tot, th) == 0
        key = (list(tot), bytes(applied))
        ref = ref or key
        t = time.perf_counter()
        for _ in range(50):
            lib.jr_fsm_fold_mt(C.c_void_p(p), C.c_size_t(n), G, R, applied, tot, th)
        ms = (time.perf_counter() - t) / 50 * 1e3
        print(f"threads {th:2d}: {ms:.3f} ms per {n} records, same result as 1 thread: {key == ref}", flush=True)


if __name__ == "__main__":
    main()
