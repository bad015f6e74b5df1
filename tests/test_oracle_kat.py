"""Pins the C++ restatement (oracle/) against the reference's own known-answer
tests.  CPU only."""
import pytest

from josefine_b200 import abi
from oracle.restated import RestatedChain, RestatedCluster
from tests import kat_cases


def make_oracle(g, r, **kw):
    return RestatedCluster.create(g, r, **kw)


@pytest.mark.parametrize("case", kat_cases.ALL_KATS, ids=lambda f: f.__name__)
def test_reference_kat_on_oracle(case):
    case(make_oracle)


# ---- chain.rs:261-350, directly on the restated Chain ----------------------------
def test_chain_new():
    c = RestatedChain()
    assert c.get_commit() == 0 and c.get_head() == 0


def test_chain_append():
    c = RestatedChain()
    c.append()
    assert c.get_commit() == 0 and c.get_head() == 1


def test_chain_commit():
    c = RestatedChain()
    c.append()
    c.commit(1)
    assert c.get_commit() == 1 and c.get_head() == 1


def test_chain_extend():
    c = RestatedChain()
    c.extend(1, 0)
    assert c.get_commit() == 0 and c.get_head() == 1


def test_chain_range():
    c = RestatedChain()
    c.extend(1, 0)
    assert len(c) == 2  # chain.range(..).collect().len() == 2


def test_chain_has():
    c = RestatedChain()
    c.extend(1, 0)
    assert c.has(1)


def test_chain_compact():
    c = RestatedChain()
    for i, n in [(1, 0), (2, 1), (3, 2), (4, 3), (5, 3), (6, 5)]:
        c.extend(i, n)
    assert c.has(4)
    c.commit(6)
    c.compact()
    assert not c.has(4)
    assert all(c.has(i) for i in (0, 1, 2, 3, 5, 6))


def test_chain_extend_missing_parent_is_err():
    """chain.rs:180-185"""
    c = RestatedChain()
    c.extend(5, 3)
    assert c.fault == abi.FAULT_EXTEND_PARENT_MISSING


def test_chain_commit_missing_block_panics():
    """chain.rs:200-202"""
    c = RestatedChain()
    c.commit(9)
    assert c.fault == abi.FAULT_COMMIT_BLOCK_MISSING


def test_chain_commit_key_in_unbounded_range():
    """Deviation D6: sled's "commit" key ends an unbounded range with a bincode panic
    (chain.rs:198,213-226) -- only in strict mode."""
    for strict in (False, True):
        c = RestatedChain(strict=strict)
        c.append(1)
        c.append(2)
        assert c.range_from(0, 1, 5) == [1, 2] and c.fault == 0  # no commit key yet
        c.commit(1)
        assert c.range_from(0, 1, 1) == [1] and c.fault == 0     # nth(1) satisfied before the end
        got = c.range_from(0, 1, 5)
        if strict:
            assert c.fault == abi.FAULT_RANGE_COMMIT_KEY
        else:
            assert got == [1, 2] and c.fault == 0


# ---- progress.rs:242-275 (observable through a leader) --------------------------------
def test_progress_starts_in_probe_and_advances():
    api = RestatedCluster.create(1, 3, flags=kat_cases.CAPTURE)
    from josefine_b200 import Command
    api.step(0, flags=0, inject=[Command.timeout(0, 1), Command.vote_response(0, 1, 1, 2, True)])
    st = api.query(0, 1)
    assert st.role == abi.ROLE_LEADER
    assert st.progress_replicate == 0 and list(st.progress_head)[:3] == [0, 0, 0]  # progress.rs:248-254
    api.apply(Command.append_response(0, 1, node_id=2, term=1, head=666))          # progress.rs:257-262
    st = api.query(0, 1)
    assert st.progress_head[1] == 666 and st.progress_replicate == 0b010
    # unknown node: expect("the node does not exist"), progress.rs:43
    api.apply(Command.append_response(0, 1, node_id=9, term=1, head=1))
    assert api.query(0, 1).fault == abi.FAULT_PROGRESS_UNKNOWN_NODE


def test_election_timeout_range_and_determinism(oracle_lib):
    seen = set()
    for d in range(2000):
        t = oracle_lib.jro_election_timeout(7, 3, 2, d, 500, 1000)
        assert 500 <= t < 1000
        seen.add(t)
    assert len(seen) > 400
    assert oracle_lib.jro_election_timeout(7, 3, 2, 5, 500, 1000) == oracle_lib.jro_election_timeout(7, 3, 2, 5, 500, 1000)
