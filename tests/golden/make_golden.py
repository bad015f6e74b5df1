#!/usr/bin/env python
"""Regenerates tests/golden/digests.json from the C++ restatement oracle.

The reference (Rust) cannot run here, so these are NOT outputs of josefine: they are
regression anchors of the restatement that the reference's own KATs pin
(tests/test_oracle_kat.py).  Each entry records a scenario and the normative digests
 after it; tests/test_golden.py replays the scenario on the oracle,
on the device code built for the host and (GPU suite) on the CUDA engine.

usage: python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from josefine_b200 import abi  # noqa: E402
from tests import golden_scenarios  # noqa: E402


def main():
    from oracle.restated import RestatedCluster
    out = {}
    for name, sc in golden_scenarios.SCENARIOS.items():
        eng = RestatedCluster.create(sc["G"], sc["R"], flags=abi.F_STREAM_DIGEST | sc.get("flags", 0),
                                     seed=sc["seed"], **sc.get("cfg", {}))
        golden_scenarios.play(eng, sc)
        out[name] = golden_scenarios.observe(eng)
    path = os.path.join(ROOT, "tests", "golden", "digests.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
