"""Committed golden digests (tests/golden/digests.json, made by tests/golden/make_golden.py
from the C++ restatement) replayed on every implementation."""
import json
import os

import pytest

from josefine_b200 import abi
from tests import golden_scenarios

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "digests.json")))


def _run(make, name):
    sc = golden_scenarios.SCENARIOS[name]
    eng = make(sc["G"], sc["R"], flags=abi.F_STREAM_DIGEST | sc.get("flags", 0), seed=sc["seed"], **sc.get("cfg", {}))
    golden_scenarios.play(eng, sc)
    assert golden_scenarios.observe(eng) == GOLDEN[name]


@pytest.mark.parametrize("name", sorted(golden_scenarios.SCENARIOS))
def test_oracle_matches_golden(name):
    from oracle.restated import RestatedCluster
    _run(lambda g, r, **kw: RestatedCluster.create(g, r, n_threads=4, **kw), name)


@pytest.mark.parametrize("name", ["churn_512x7", "strict_commit_key_256x3"])
def test_device_code_on_host_matches_golden(name):
    from tests.emu.emu import EmuEngine
    _run(lambda g, r, **kw: EmuEngine.create(g, r, **kw), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(golden_scenarios.SCENARIOS))
def test_cuda_engine_matches_golden(name):
    from josefine_b200 import RaftEngine
    _run(lambda g, r, **kw: RaftEngine.create(g, r, **kw), name)


def test_golden_covers_real_behaviour():
    assert GOLDEN["cold_1024x3_96ticks"]["groups_with_leader"] > 900
    assert GOLDEN["steady_2048x5_64ticks"]["max_commit"] > 50 and GOLDEN["steady_2048x5_64ticks"]["faulted"] == 0
    assert GOLDEN["churn_512x7"]["groups_with_leader"] < 512          # killed groups stay leaderless
    assert GOLDEN["strict_commit_key_256x3"]["faulted"] > 0           # D6: leaders hit the "commit" key
