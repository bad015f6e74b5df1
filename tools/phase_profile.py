#!/usr/bin/env python
"""Warp-cycle attribution of the fused step kernel with the JR_PROFILE build
(josefine_b200/csrc/ab/lib_prof.so): where do a leader warp's and a follower warp's
cycles go per tick -- fetch, each Command kind, bookkeeping, barrier wait."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["JR_ENGINE_LIB"] = os.path.join(ROOT, "josefine_b200", "csrc", "ab", "lib_prof.so")
if not os.path.exists(os.environ["JR_ENGINE_LIB"]):
    raise SystemExit("build it first (here, no GPU needed): python __graft_entry__.py --profile-build")
import bench  # noqa: E402
from josefine_b200 import abi, RaftEngine  # noqa: E402

G, R, S = 65536, 5, 64
e = RaftEngine.create(G, R, seed=1, chain_capacity=1024, flags=abi.F_CAPTURE_FSM if os.environ.get('JR_BENCH_CAPTURE', '1') != '0' else 0, fsm_units=16)
e.step(0, flags=0, inject=bench.bootstrap_inject(G, R))
e.run(100, 100, 16, 1)
buf = (C.c_uint64 * 96)()
e._lib.jr_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
e._lib.jr_profile_read(e._h, buf)           # clear
now = 1700
for _ in range(5):
    e.run(now, 100, S, 1)
    e.truncate(8)
    e.discard_fsm(strict=False)
    now += 100 * S
e._lib.jr_profile_read(e._h, buf)
names = {1: "  drain: Leader::commit", 8: "  drain: mask+unit loads", 9: "  drain: advances", 0: "Tick", 2: "  replicate: first peer (scan)", 3: "  replicate: other peers (refs)", 4: "AppendEntries", 5: "AppendResponse", 6: "Heartbeat",
         7: "HeartbeatResponse", 10: "ClientRequest", 11: "fast AResp drain", 12: "WHOLE TICK", 13: "barrier wait", 14: "fetch (next_cmd)",
         15: "publish marks/counts"}
for role, rn in ((2, "LEADER warp"), (0, "FOLLOWER warp")):
    tick_n = buf[(role * 16 + 12) * 2 + 1] or 1
    print(f"== {rn}: cycles per warp-tick (count = events per warp-tick)")
    for slot in (12, 14, 8, 9, 1, 11, 2, 3, 0, 4, 5, 6, 7, 10, 2, 3, 15, 13):
        cyc, n = buf[(role * 16 + slot) * 2], buf[(role * 16 + slot) * 2 + 1]
        if n:
            print(f"  {names[slot]:22s} {cyc / tick_n:9.0f} cycles   x{n / tick_n:5.2f}   ({cyc / n:7.0f} per event)")
print("faults", e.fault_count())
