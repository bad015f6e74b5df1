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


# ---- hypothesis-driven differential runs ---------------------------------------------------------
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

from josefine_b200 import Command  # noqa: E402

_node = st.integers(1, 5)
_term = st.integers(0, 6)
_blk = st.integers(0, 12)


@st.composite
def _command(draw, G, R):
    g, to = draw(st.integers(0, G - 1)), draw(st.integers(1, R))
    kind = draw(st.sampled_from(["vreq", "vresp", "ae", "aresp", "hb", "hbresp", "timeout", "creq", "cresp", "tick", "noop"]))
    if kind == "vreq":
        return Command.vote_request(g, to, draw(_term), draw(_node), draw(_term), draw(_blk))
    if kind == "vresp":
        return Command.vote_response(g, to, draw(_term), draw(st.integers(1, R)), draw(st.booleans()))
    if kind == "ae":
        base = draw(st.integers(0, 10))
        n = draw(st.integers(0, 3))
        blocks = [(base + i + 1, base + i if draw(st.integers(0, 9)) else draw(_blk), 5000 + base * 8 + i) for i in range(n)]
        return Command.append_entries(g, to, draw(_term), draw(_node), blocks)
    if kind == "aresp":
        return Command.append_response(g, to, draw(st.integers(1, R)), draw(_term), draw(_blk))
    if kind == "hb":
        return Command.heartbeat(g, to, draw(_term), draw(_blk), draw(_node))
    if kind == "hbresp":
        return Command.heartbeat_response(g, to, draw(_blk), draw(st.booleans()))
    if kind == "timeout":
        return Command.timeout(g, to)
    if kind == "creq":
        return Command.client_request(g, to, draw(st.integers(1, 1 << 40)))
    if kind == "cresp":
        return Command.client_response(g, to, draw(st.integers(1, 99)))
    if kind == "tick":
        return Command.tick(g, to)
    return Command.noop(g, to)


@st.composite
def _script(draw):
    R = draw(st.sampled_from([2, 3, 5]))
    G = 2
    steps = []
    for _ in range(draw(st.integers(4, 18))):
        inj = draw(st.lists(_command(G, R), max_size=4))
        props = None
        if draw(st.integers(0, 3)) == 0:
            props = [(draw(st.integers(0, R)), draw(st.integers(1, 1 << 30))) for _ in range(G)]
        steps.append((inj, props, draw(st.sampled_from([0, 0, 1, 2])), draw(st.sampled_from([50, 100, 100, 400, 900]))))
    return R, G, draw(st.integers(0, 1 << 20)), steps


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(_script())
def test_hypothesis_scripts_cpp_vs_python(script):
    """Arbitrary short command scripts (mostly well-formed, some nonsense) on both restatements: every Message,
    Instruction and replica state must agree step by step."""
    R, G, seed, steps = script
    p = LitePair(make_cpp, make_py, G, R, seed=seed, chain_capacity=64, flags=CAP)
    now = 0
    for inj, props, n_synth, dt in steps:
        now += dt
        p.step(now, inject=inj, proposals=props, n_synth=n_synth)
    p.finish()


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(_script())
def test_hypothesis_scripts_device_code_vs_cpp(script):
    """The same scripts on the engine's device code (CPU emulation build) against the C++ restatement, digests included."""
    from tests.emu.emu import EmuEngine
    R, G, seed, steps = script
    p = parity.Pair(make_cpp, lambda g, r, **kw: EmuEngine.create(g, r, **kw), G, R, seed=seed, chain_capacity=64)
    now = 0
    for inj, props, n_synth, dt in steps:
        now += dt
        p.step(now, inject=inj, proposals=props, n_synth=n_synth)
    p.finish()
