"""Instruction-stream output path (jr_fsm_record / jr_fsm_records_* / jr_fsm_expand), log truncation (D7), node
restart and bulk introspection: cases run on the device code (CPU emulation) and on the GPU, against the oracle."""
from __future__ import annotations

import random

from josefine_b200 import abi, Command, fsm_tuple
from tests import parity

CAP = abi.F_CAPTURE_FSM | abi.F_STREAM_DIGEST


def _bootstrap(api, G, R, node=1):
    q = 0 if R == 1 else R // 2 + 1
    inj = []
    for g in range(G):
        inj.append(Command.timeout(g, node))
        for v in [v for v in range(1, R + 1) if v != node][:max(q - 1, 0)]:
            inj.append(Command.vote_response(g, node, 1, v, True))
    api.step(0, flags=0, inject=inj)


def strided_tokens(n_ticks, G, first_tick=0, stride=1 << 32):
    """tokens[k][g]: one proposal per group per tick, advancing by a constant stride per tick (what bench.py sends)."""
    return [[(first_tick + k + 1) * stride + g + 1 for g in range(G)] for k in range(n_ticks)]


def case_records_expand_to_the_oracle_stream(make, make_oracle, G=24, R=5, launches=4, ticks=48):
    """jr_run_tokens + jr_fsm_records_async/wait + jr_fsm_expand == the oracle's drain_fsm, launch after launch;
    and the steady state really is compact: O(1) records per replica per launch."""
    eng, ora = make(G, R, seed=1, flags=CAP, fsm_units=16), make_oracle(G, R, seed=1, flags=CAP)
    for api in (eng, ora):
        _bootstrap(api, G, R)
        api.run(100, 100, 16, 0)
        api.leader_table()
        api.drain_fsm()
    now, tick = 1700, 0
    per_launch = []
    for _ in range(launches):
        toks = strided_tokens(ticks, G, tick)
        eng.run_tokens(now, 100, toks)
        ora.run_tokens(now, 100, toks)
        now += 100 * ticks
        tick += ticks
        recs, batch = eng.fsm_records()
        want = [fsm_tuple(f) for f in ora.drain_fsm(cap=G * R * ticks * 4)]
        got = [fsm_tuple(f) for f in eng.fsm_expand(recs)]
        assert got == want
        assert batch.n_instructions == len(want) and batch.n_dropped == 0 and batch.n_records == len(recs)
        assert [batch.node_offset[r] for r in range(R + 1)] == sorted(batch.node_offset[r] for r in range(R + 1))
        keys = [(r.node, r.group) for r in recs]
        assert keys == sorted(keys)                       # sorted by (node, group), FIFO per replica
        per_launch.append(len(recs))
    # leader: 1 apply run + 1 notify run + ceil(instructions / 64) pattern words (+1 when a launch starts mid-word);
    # follower: 1 apply run
    leader_instr = 2 * ticks
    assert max(per_launch[1:]) <= G * (2 + (leader_instr + 63) // 64 + 1 + (R - 1)), per_launch   # (the first launch also holds the start-up)
    parity.compare_digests(eng, ora)


def case_irregular_streams_expand_exactly(make, make_oracle, seed=5):
    """Random tokens, proposals aimed at followers / nobody, elections and silenced leaders: every Instruction still
    comes back, in order."""
    rng = random.Random(seed)
    G, R, N = 10, 3, 30
    eng, ora = make(G, R, seed=seed, flags=CAP, fsm_units=256), make_oracle(G, R, seed=seed, flags=CAP)
    for api in (eng, ora):
        api.run(100, 100, 20, 0)
    now = 2100
    for rnd in range(4):
        props = [[(rng.choice([0, 1, 2, 3]), rng.getrandbits(48) + 1) for _ in range(G)] for _ in range(N)]
        for api in (eng, ora):
            api.run_proposals(now, 100, props)
        now += 100 * N
        if rnd == 1:
            assert eng.kill_leaders(9, 400) == ora.kill_leaders(9, 400)
        recs, _ = eng.fsm_records()
        assert [fsm_tuple(f) for f in eng.fsm_expand(recs)] == [fsm_tuple(f) for f in ora.drain_fsm(cap=1 << 16)]
    parity.compare_states(eng, ora, chain_ids=64)
    parity.compare_digests(eng, ora)


def case_expand_known_answers(lib):
    """jr_fsm_expand on hand-made records (no engine): the format is normative in the ABI header."""
    from josefine_b200.raft import expand_records

    def rec(group, node, kind, count, id0, addr=0, tok0=0, stride=0):
        r = abi.FsmRecord()
        r.group, r.hdr, r.id0, r.addr, r.tok0, r.stride = group, kind | ((node - 1) << 2) | (count << 8), id0, addr, tok0, stride
        return r

    A, N, P = abi.FSMR_APPLY, abi.FSMR_NOTIFY, abi.FSMR_PATTERN
    client = abi.ADDR_CLIENT << 16
    recs = [
        rec(1, 2, A, 3, 5, tok0=100, stride=10),                       # node 2 of group 1: blocks 5,6,7
        rec(0, 1, A, 1, 0, tok0=0, stride=0),                          # genesis block 0 -> 0 (count 1: next explicit)
        rec(0, 1, N, 2, 7, addr=client, tok0=70, stride=1),
        rec(0, 1, A, 2, 5, tok0=50, stride=1),
        rec(0, 1, P, 5, 0, tok0=0b01010),                              # stream: A N A N A
        rec(1, 2, A, 1, 9, tok0=7, stride=3),                          # single block 9 -> 3
    ]
    out = [fsm_tuple(f) for f in expand_records(lib, recs, 2, 3)]
    assert out == [
        (0, 1, abi.FSM_APPLY, 0, 0, 0, 0, 0),
        (0, 1, abi.FSM_NOTIFY, abi.ADDR_CLIENT, 0, 7, 0, 70),
        (0, 1, abi.FSM_APPLY, 0, 0, 5, 4, 50),
        (0, 1, abi.FSM_NOTIFY, abi.ADDR_CLIENT, 0, 8, 0, 71),
        (0, 1, abi.FSM_APPLY, 0, 0, 6, 5, 51),
        (1, 2, abi.FSM_APPLY, 0, 0, 5, 4, 100),
        (1, 2, abi.FSM_APPLY, 0, 0, 6, 5, 110),
        (1, 2, abi.FSM_APPLY, 0, 0, 7, 6, 120),
        (1, 2, abi.FSM_APPLY, 0, 0, 9, 3, 7),
    ]
    import pytest
    from josefine_b200 import RaftError
    with pytest.raises(RaftError):                                     # pattern marks 2 notifies, only 1 exists
        expand_records(lib, [rec(0, 1, N, 1, 1, addr=client), rec(0, 1, P, 2, 0, tok0=0b11)], 1, 1)
    with pytest.raises(RaftError):                                     # group out of range
        expand_records(lib, [rec(3, 1, A, 1, 1)], 2, 1)
    assert expand_records(lib, [], 4, 3) == []


def case_truncation_soak(make, make_oracle, G=16, R=5, cap=64, rounds=40, ticks=25, kill_at=12, compare_fsm=True):
    """D7: with jr_truncate after every launch a 64-id window carries a group through 1,000 ticks (ids up to
    ~1,000) with no reset and no fault; everything stays bit-equal to the oracle (which truncates its maps the
    same way), including after leaders are silenced and the window stops moving for those groups."""
    eng, ora = (m(G, R, seed=3, flags=CAP, chain_capacity=cap, fsm_units=max(32, ticks // 8 + 16)) for m in (make, make_oracle))
    for api in (eng, ora):
        _bootstrap(api, G, R)
    now = 100
    for rnd in range(rounds):
        for api in (eng, ora):
            api.run(now, 100, ticks, 1)
            api.truncate(4)
            if rnd % 5 == 4:
                api.compact()
        now += 100 * ticks
        if rnd == kill_at:
            assert eng.kill_leaders(5, 250) == ora.kill_leaders(5, 250)
        if compare_fsm:
            assert [fsm_tuple(f) for f in eng.drain_fsm(cap=1 << 16)] == [fsm_tuple(f) for f in ora.drain_fsm(cap=1 << 16)]
        else:                          # long soaks: the stream digests checked at the end cover the content
            assert eng.discard_fsm() == ora.discard_fsm()
        if rnd % 8 == 0:
            parity.compare_states(eng, ora, where=f"[round {rnd}]", chain_ids=0)
    st = eng.query_many([(g, 1) for g in range(G)])
    moved = [s for s in st if s.chain_floor > cap]
    assert len(moved) >= G // 2 and all(s.fault == 0 for s in moved)             # far past the window size, no fault
    assert any(s.head > 5 * cap for s in st)
    # block tables agree id for id inside the window (and are empty below the floor)
    reqs = [(g, n, max(int(eng.query(g, n).chain_floor) - 4, 0), cap + 8) for g in range(0, G, 5) for n in (1, 3)]
    assert eng.chain_read_many(reqs) == ora.chain_read_many(reqs)
    parity.compare_states(eng, ora, chain_ids=0)
    parity.compare_digests(eng, ora)


def case_window_limits(make, make_oracle):
    """Ids must stay inside [floor, floor + chain_capacity): appending past the end faults (as before); so does an
    AppendEntries that carries a block BELOW the floor (deviation D7) -- identically on both sides."""
    p = parity.Pair(make_oracle, make, 2, 3, seed=3, chain_capacity=16)
    parity.bootstrap_leaders(p, now=0)
    for k in range(10):
        p.step(100 * (k + 1), n_synth=1)
    p.both("truncate", 2)
    assert p.b.query(0, 2).chain_floor > 0
    floor = p.b.query(0, 2).chain_floor
    p.step(1100, inject=[Command.append_entries(0, 2, term=1, leader_id=1, blocks=[(floor - 1, floor, 5)])])
    assert p.b.query(0, 2).fault == abi.FAULT_ENGINE_CHAIN_CAPACITY
    for k in range(12, 40):
        p.step(100 * k, n_synth=1)          # group 1 runs into the end of its window: no truncate was called again
    assert p.b.query(1, 1).fault == abi.FAULT_ENGINE_CHAIN_CAPACITY
    p.finish()


def case_node_restart(make, make_oracle):
    """jr_node_restart = RaftHandle::new over a persisted chain (chain.rs:117-137): commit = head = id_gen = the
    persisted commit; State is the default.  N2: the restarted node faults on its first append once it
    leads, because id_gen.next() == commit is not > head."""
    p = parity.Pair(make_oracle, make, 1, 3, seed=7, chain_capacity=64)
    parity.bootstrap_leaders(p, now=0)
    for k in range(8):
        p.step(100 * (k + 1), n_synth=1)
    st = p.b.query(0, 2)
    blocks = [b for b in p.b.chain_read(0, 2, 0, 16) if b is not None]
    assert st.commit > 0 and len(blocks) >= st.commit
    p.both("node_restart", 0, 2, 900, blocks, st.commit)
    r = p.b.query(0, 2)
    assert (r.current_term, r.voted_for, r.role, r.head, r.commit, r.id_gen) == (0, 0, abi.ROLE_FOLLOWER, st.commit, st.commit, st.commit)
    assert r.election_time_ms == 900 and r.rng_draws == 1 and r.alive and r.fault == 0
    for k in range(9, 14):
        p.step(100 * k, n_synth=1)                       # it rejoins: heartbeats and AppendEntries reach it
    assert p.b.query(0, 2).leader_id == 1 and p.b.query(0, 2).head > st.commit
    # make the restarted node leader: its first append asserts id > head (chain.rs:163) -> fault 3
    p.both("set_alive", 0, 1, False)
    p.step(2000, flags=0, inject=[Command.timeout(0, 2)])
    if not p.b.handle(0, 2).is_leader():                 # voted_for = Some(1) after the heartbeats: Timeout is ignored (N1)
        p.both("node_restart", 0, 2, 2000, blocks, st.commit)
        p.step(2100, flags=0, inject=[Command.timeout(0, 2), Command.vote_response(0, 2, 1, 3, True)])
    assert p.b.handle(0, 2).is_leader()
    p.step(2200, flags=0, inject=[Command.client_request(0, 2, token=99)])
    assert p.b.query(0, 2).fault == abi.FAULT_APPEND_ID_NOT_GT_HEAD
    # a restart with commit == 0 runs Chain::init again: genesis block, id_gen = 1
    p.both("node_restart", 0, 3, 3000, [], 0)
    r = p.b.query(0, 3)
    assert (r.head, r.commit, r.id_gen, r.max_key) == (0, 0, 1, 0)
    assert p.b.chain_read(0, 3, 0, 4) == [(0, 0, 0), None, None, None]
    p.finish()


def case_bulk_introspection(make):
    """jr_query_many / jr_chain_read_many return exactly what the one-at-a-time calls return."""
    G, R = 37, 3
    api = make(G, R, seed=2, flags=CAP)
    api.run(100, 100, 30, 1)
    targets = [(g, n) for g in range(G) for n in range(1, R + 1)]
    many = [parity.state_tuple(s) for s in api.query_many(targets)]
    assert many == [parity.state_tuple(api.query(g, n)) for g, n in targets]
    reqs = [(g, 1 + g % R, g % 3, 5 + g % 7) for g in range(G)] + [(0, 1, 0, 0)]
    assert api.chain_read_many(reqs) == [api.chain_read(*r) for r in reqs]
    assert api.query_many([]) == []


def case_save_restore(make):
    """jr_engine_save / jr_engine_restore: a restored engine continues bit for bit (state, streams, Instructions)."""
    G, R = 9, 3
    a = make(G, R, seed=4, flags=CAP, chain_capacity=128)
    a.run(100, 100, 30, 1)
    a.leader_table()
    blob = a.save()
    b = make(G, R, seed=4, flags=CAP, chain_capacity=128)
    b.restore(blob)
    toks = strided_tokens(12, G)
    for api in (a, b):
        api.run_tokens(3100, 100, toks)            # uses the restored routing table
        api.run(4300, 100, 5, 1)
    assert [fsm_tuple(f) for f in a.drain_fsm(cap=1 << 16)] == [fsm_tuple(f) for f in b.drain_fsm(cap=1 << 16)]
    parity.compare_states(a, b, chain_ids=64)
    parity.compare_digests(a, b)
    import pytest
    from josefine_b200 import RaftError
    c = make(G + 1, R, seed=4, flags=CAP, chain_capacity=128)
    with pytest.raises(RaftError):
        c.restore(blob)


PAIRED = [case_records_expand_to_the_oracle_stream, case_irregular_streams_expand_exactly, case_truncation_soak,
          case_window_limits, case_node_restart]
SINGLE = [case_bulk_introspection, case_save_restore]
