#!/usr/bin/env python
"""Small fused-kernel workload for compute-sanitizer (memcheck / racecheck / synccheck):
cold-start elections at R=3, bootstrapped steady state at R=5 and R=7 with tiny
shared-memory mailboxes (spill path), compact, kill_leaders, inject + capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from josefine_b200 import abi, RaftEngine  # noqa: E402
from tests.golden_scenarios import bootstrap  # noqa: E402

for R, G in ((3, 70), (5, 64), (7, 33)):
    e = RaftEngine.create(G, R, seed=R, flags=abi.F_CAPTURE_MESSAGES | abi.F_CAPTURE_FSM | abi.F_STREAM_DIGEST,
                          chain_capacity=128, fsm_units=128)
    if R == 3:
        e.run(100, 100, 40, 1)
    else:
        e.step(0, flags=0, inject=bootstrap(G, R))
        e.run(100, 100, 24, 1)
    e.step(9000, n_synth=1)
    e.kill_leaders(1, 300)
    e.compact()
    e.run(9100, 100, 8, 1)
    print(R, e.state_digest(), e.fault_count(), e.stream_digest()[2:])
print("sanitize workload done")
