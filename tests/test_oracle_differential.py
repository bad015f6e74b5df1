"""Differential test between the two independently written restatements of
josefine's src/raft: C++ (oracle/restated_raft.cpp, the parity oracle) and pure
Python (oracle/restated_raft.py).  The reference's own tests pin only single-voter
behaviour, so this is the defence against one shared misreading."""
import pytest

from josefine_b200 import abi
from oracle.restated import RestatedCluster
from oracle.restated_raft import PyCluster, election_timeout
from tests import kat_cases, parity

CAP = abi.F_CAPTURE_MESSAGES | abi.F_CAPTURE_FSM


def make_cpp(g, r, **kw):
    return RestatedCluster.create(g, r, **kw)


def make_py(g, r, **kw):
    return PyCluster.create(g, r, **kw)


class LitePair(parity.Pair):
    """The Python restatement has no digests / run(); compare states and streams only."""

    def finish(self):
        parity.compare_states(self.a, self.b, chain_ids=self.chain_ids, where="[final]")
        assert self.a.leader_table() == self.b.leader_table()
        assert self.a.fault_count() == self.b.fault_count()


@pytest.mark.parametrize("case", [c for c in kat_cases.ALL_KATS], ids=lambda f: f.__name__)
def test_reference_kat_on_python_restatement(case):
    case(make_py)


@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 7])
def test_cold_start(R):
    p = LitePair(make_cpp, make_py, 4, R, seed=10 + R, flags=CAP)
    parity.scenario_cold_start(p, steps=45)


@pytest.mark.parametrize("R", [3, 5, 7])
def test_steady(R):
    p = LitePair(make_cpp, make_py, 3, R, seed=2, flags=CAP)
    parity.scenario_steady(p, steps=20)


@pytest.mark.parametrize("seed", range(10))
@pytest.mark.parametrize("R", [3, 5])
def test_random_inject(R, seed):
    p = LitePair(make_cpp, make_py, 3, R, seed=seed, chain_capacity=64, flags=CAP)
    parity.scenario_random_inject(p, seed=1000 + seed * 3 + R, steps=40)


def test_random_inject_strict():
    p = LitePair(make_cpp, make_py, 3, 5, seed=4, chain_capacity=64, flags=CAP | abi.F_SLED_COMMIT_KEY_STRICT)
    parity.scenario_random_inject(p, seed=77, steps=40)


def test_timeout_function(oracle_lib):
    for args in [(0, 0, 1, 0, 500, 1000), (9, 123456789, 5, 77, 500, 1000), (2**63, 2**40, 7, 1000, 5, 6)]:
        assert election_timeout(*args) == oracle_lib.jro_election_timeout(*args)
