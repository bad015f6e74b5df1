"""CPU suite: the engine's DEVICE CODE (compiled for the host through
tests/emu/cuda_emu.h) against the C++ restatement oracle, bit for bit.
The GPU suite (test_gpu_parity.py) runs the same scenarios through the real
CUDA library."""
import pytest

from josefine_b200 import abi
from oracle.restated import RestatedCluster
from tests import kat_cases, parity
from tests.emu.emu import EmuEngine


def make_oracle(g, r, **kw):
    return RestatedCluster.create(g, r, **kw)


def make_emu(g, r, **kw):
    return EmuEngine.create(g, r, **kw)


@pytest.mark.parametrize("case", kat_cases.ALL_KATS, ids=lambda f: f.__name__)
def test_reference_kat_on_device_code(case):
    case(make_emu)


@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 7])
def test_cold_start(R):
    p = parity.Pair(make_oracle, make_emu, 6, R, seed=R)
    parity.scenario_cold_start(p, steps=45)


@pytest.mark.parametrize("R", [3, 5, 7])
def test_steady_state(R):
    p = parity.Pair(make_oracle, make_emu, 5, R, seed=1)
    parity.scenario_steady(p, steps=24)


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("R", [3, 5])
def test_random_inject(R, seed):
    p = parity.Pair(make_oracle, make_emu, 3, R, seed=seed, chain_capacity=64)
    parity.scenario_random_inject(p, seed=seed * 7 + R, steps=50)


@pytest.mark.parametrize("R", [3, 5])
def test_random_inject_strict_commit_key(R):
    p = parity.Pair(make_oracle, make_emu, 3, R, seed=5, chain_capacity=64,
                    flags=parity.FULL | abi.F_SLED_COMMIT_KEY_STRICT)
    parity.scenario_random_inject(p, seed=99 + R, steps=50)


def test_run_equals_steps():
    """jr_run(n) == n x jr_step(DELIVER|TICK|SYNTH), bit for bit (ABI contract)."""
    a = make_emu(4, 3, seed=3, flags=parity.FULL)
    b = make_emu(4, 3, seed=3, flags=parity.FULL)
    a.run(100, 100, 30, 1)
    for k in range(30):
        b.step(100 + 100 * k, n_synth=1)
    parity.compare_states(a, b, chain_ids=40)
    parity.compare_digests(a, b)


@pytest.mark.parametrize("units,cache", [(1, 1), (2, 2), (3, 4), (0, 0)])
def test_spill_and_cache_paths(monkeypatch, units, cache):
    """Tiny shared-memory mailbox / table cache: units beyond Us spill to the global
    mailbox and the direct-mapped cache thrashes -- results must not change."""
    monkeypatch.setenv("JR_SMEM_UNITS", str(units))
    monkeypatch.setenv("JR_TABLE_CACHE", str(cache))
    p = parity.Pair(make_oracle, make_emu, 4, 5, seed=1)
    parity.scenario_steady(p, steps=20)
    p = parity.Pair(make_oracle, make_emu, 3, 3, seed=11, chain_capacity=64)
    parity.scenario_random_inject(p, seed=4242, steps=40)
    a = make_emu(4, 3, seed=3, flags=parity.FULL, fsm_units=256)
    b = make_oracle(4, 3, seed=3, flags=parity.FULL, fsm_units=256)
    a.run(100, 100, 40, 2)
    b.run(100, 100, 40, 2)
    parity.compare_states(a, b, chain_ids=90)
    parity.compare_digests(a, b)


def test_mailbox_beyond_index_range():
    """A leader that answers six HeartbeatResponse{!has} with six replicate() rounds
    emits > 31 units in one step: the delivery index overflows to the scan path and
    units spill past the shared-memory mailbox."""
    from josefine_b200 import Command
    p = parity.Pair(make_oracle, make_emu, 2, 7, seed=9, mailbox_units=128)
    parity.bootstrap_leaders(p, now=0)
    for k in range(4):
        p.step(100 * (k + 1), n_synth=1)
    inj = [Command.heartbeat_response(g, 1, commit=1, has_committed=False, from_=2 + (i % 6))
           for g in range(2) for i in range(6)]
    res = p.step(500, inject=inj, n_synth=1)
    per_sender = {}
    for m in res.messages:
        per_sender[(m.group, m.from_id)] = per_sender.get((m.group, m.from_id), 0) + 1 + m.n_blocks
    assert max(per_sender.values()) > 40
    for k in range(6):
        p.step(600 + 100 * k, n_synth=1)
    p.finish()


def test_engine_reset_is_a_fresh_engine():
    """jr_engine_reset == drop the handle and RaftHandle::new again (mod.rs:428-435)."""
    a = make_emu(5, 3, seed=8, flags=parity.FULL)
    a.run(100, 100, 25, 1)
    first = (a.state_digest(), a.leader_table())
    a.apply_reset = getattr(a, "_lib").jr_engine_reset
    a.apply_reset.argtypes = [__import__("ctypes").c_void_p]
    assert a.apply_reset(a._h) == 0
    fresh = make_emu(5, 3, seed=8, flags=parity.FULL)
    parity.compare_states(a, fresh, chain_ids=30)
    a.run(100, 100, 25, 1)
    assert (a.state_digest(), a.leader_table()) == first


def test_run_proposals_equals_steps_with_proposals():
    """jr_run_proposals(n) == n x jr_step(DELIVER|TICK, proposals[k]) (ABI contract), on the device
    code and on the oracle -- proposals aimed at leaders, followers (proxied) and nobody."""
    import random
    rng = random.Random(5)
    G, R, N = 6, 3, 36
    props = [[(rng.choice([0, 1, 2, 3]), 1000 * k + g + 1) for g in range(G)] for k in range(N)]
    a = make_emu(G, R, seed=4, flags=parity.FULL, fsm_units=256)
    b = make_emu(G, R, seed=4, flags=parity.FULL, fsm_units=256)
    o = make_oracle(G, R, seed=4, flags=parity.FULL, fsm_units=256)
    for eng in (a, o):
        eng.run(100, 100, 20, 0)                         # elect leaders first
        eng.run_proposals(2100, 100, props)
    b.run(100, 100, 20, 0)
    for k in range(N):
        b.step(2100 + 100 * k, proposals=props[k])
    parity.compare_states(a, b, chain_ids=60)
    parity.compare_digests(a, b)
    parity.compare_states(a, o, chain_ids=60)
    parity.compare_digests(a, o)
    assert max(c for (_, _, c) in a.leader_table()) > 5


@pytest.mark.parametrize("variant", ["plain", "sorted"])
def test_both_kernel_variants_match_the_oracle(monkeypatch, variant):
    """The role-sorted variant (leader -> warp 0; chosen when leaders sit on several replica
    indices) is pure scheduling: same results as the plain one, on scattered and on uniform leaders."""
    monkeypatch.setenv("JR_STEP_VARIANT", variant)
    from josefine_b200 import Command
    for R in (3, 5, 7):
        p = parity.Pair(make_oracle, make_emu, 7, R, seed=R)
        inj = []
        q = R // 2 + 1
        for g in range(7):                      # leader of group g on node g % R + 1
            n = g % R + 1
            inj.append(Command.timeout(g, n))
            for v in [v for v in range(1, R + 1) if v != n][:q - 1]:
                inj.append(Command.vote_response(g, n, 1, v, True))
        p.step(0, flags=0, inject=inj)
        for k in range(14):
            p.step(100 * (k + 1), n_synth=1)
        p.run(1500, 100, 12, 1)
        p.finish()
    p = parity.Pair(make_oracle, make_emu, 4, 3, seed=21, chain_capacity=64)
    parity.scenario_random_inject(p, seed=77, steps=40)


def test_three_single_node_engines_over_the_wire_match_resident_cluster():
    from tests.wire_cluster import run_networked_vs_resident
    frames, _ = run_networked_vs_resident(make_emu)
    assert frames > 100


def test_leader_routed_tokens():
    parity.scenario_leader_routed_tokens(make_emu, make_oracle)


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_split_launches_are_bit_identical(monkeypatch, parts):
    """launch_step may cut a block's fused ticks into consecutive tasks handed over through global memory
    ("Split launches"); any split must give the results of the unsplit launch."""
    import random
    monkeypatch.setenv("JR_PARTS", str(parts))
    split = make_emu(40, 3, seed=9, flags=parity.FULL, fsm_units=512)
    monkeypatch.setenv("JR_PARTS", "1")
    whole = make_emu(40, 3, seed=9, flags=parity.FULL, fsm_units=512)
    o = make_oracle(40, 3, seed=9, flags=parity.FULL, fsm_units=512)
    rng = random.Random(parts)
    props = [[(rng.choice([0, 1, 2, 3]), 7000 * k + g + 1) for g in range(40)] for k in range(19)]
    for eng in (split, whole, o):
        eng.run(100, 100, 23, 0)                      # odd tick counts: parts of unequal length, both mailbox parities
        eng.run(2400, 100, 17, 2)
        eng.run_proposals(4100, 100, props)
        eng.leader_table()
        eng.run_tokens(6000, 100, [[(k << 20) | (g + 1) for g in range(40)] for k in range(9)])
    fa, fb, fo = ([parity.fsm_tuple(f) for f in e.drain_fsm()] for e in (split, whole, o))
    assert fa == fb == fo
    parity.compare_states(split, whole, chain_ids=80)
    parity.compare_digests(split, whole)
    parity.compare_states(split, o, chain_ids=80)
    parity.compare_digests(split, o)
