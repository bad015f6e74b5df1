"""Engine-limit edge cases (deviation D4 and the JR_FAULT_ENGINE_* codes), run on every implementation."""
from josefine_b200 import abi, Command
from tests import kat_cases, parity


def case_chain_capacity_fault(make_a, make_b):
    """Block ids must stay below chain_capacity: the leader's append that would create id == capacity
    faults with JR_FAULT_ENGINE_CHAIN_CAPACITY on both sides, at the same proposal."""
    p = parity.Pair(make_a, make_b, 2, 3, seed=3, chain_capacity=8)
    parity.bootstrap_leaders(p, now=0)
    for k in range(14):
        p.step(100 * (k + 1), n_synth=1)
    p.finish()
    st = p.b.query(0, 1)
    assert st.fault == abi.FAULT_ENGINE_CHAIN_CAPACITY and st.head == 7 and st.id_gen == 9


def case_client_queue_overflow(make_a, make_b):
    """A follower without a leader queues ClientRequests (follower.rs:266); the engine bounds the
    queue at JR_CLIENT_QUEUE_CAP and faults on the next one."""
    p = parity.Pair(make_a, make_b, 1, 3, seed=1)
    for i in range(abi.CLIENT_QUEUE_CAP + 1):
        p.step(10 + i, flags=0, inject=[Command.client_request(0, 2, token=100 + i)])
    st = p.b.query(0, 2)
    assert st.fault == abi.FAULT_ENGINE_QUEUE_OVERFLOW and st.n_queued == abi.CLIENT_QUEUE_CAP
    p.finish()


def case_mailbox_overflow_faults_cleanly(make):
    """More units than mailbox_units in one step: the emitting replica gets
    JR_FAULT_ENGINE_MAILBOX_OVERFLOW; nobody else is disturbed (no oracle equivalent)."""
    api = make(3, 7, flags=kat_cases.CAPTURE, mailbox_units=8, seed=2)
    inj = []
    for g in range(3):
        inj.append(Command.timeout(g, 1))
        for v in (2, 3, 4):
            inj.append(Command.vote_response(g, 1, 1, v, True))
    api.step(0, flags=0, inject=inj)
    for k in range(1, 12):         # leader: heartbeat + 6 AppendEntries headers + their blocks soon exceed 8 units
        api.step(100 * k, n_synth=3)
        if api.fault_count():
            break
    for g in range(3):
        assert api.query(g, 1).fault == abi.FAULT_ENGINE_MAILBOX_OVERFLOW
        for n in range(2, 8):
            assert api.query(g, n).fault == 0
    api.step(5000)                 # the engine keeps stepping the healthy replicas
    assert api.fault_count() == 3


def case_fsm_fifo_overflow_never_touches_consensus(make):
    """A full Instruction FIFO is an observability limit: records are dropped and the drain says
    JR_E_CAPACITY, but the replica does not fault and its state equals that of an engine with a large FIFO."""
    import pytest
    from josefine_b200 import RaftError
    irregular = [[(1, 1000 + 37 * k * k)] for k in range(8)]    # tokens with no common stride: one record per Instruction
    small = make(1, 1, flags=abi.F_CAPTURE_FSM | abi.F_STREAM_DIGEST, fsm_units=4)
    big = make(1, 1, flags=abi.F_CAPTURE_FSM | abi.F_STREAM_DIGEST, fsm_units=64)
    for api in (small, big):
        api.apply(Command.timeout(0, 1))
        api.run_proposals(100, 100, irregular)    # 8 Notify + 8 Apply, 16 records > 4 units
    assert small.query(0, 1).fault == 0 and small.fault_count() == 0
    assert small.state_digest() == big.state_digest() and small.stream_digest() == big.stream_digest()
    with pytest.raises(RaftError) as err:
        small.drain_fsm()
    assert err.value.status == abi.E_CAPACITY
    assert len(big.drain_fsm()) == 16
    assert small.drain_fsm() == [] and big.drain_fsm() == []      # a drain takes what it returns
    # ... while a regular stream compresses: synthetic tokens advance by a constant stride, so 120 Instructions are
    # one APPLY run + one NOTIFY run + one PATTERN record
    small.run(1000, 100, 60, 1)
    assert len(small.drain_fsm(cap=1024)) == 120


def case_degenerate_calls(make):
    api = make(33, 3, flags=kat_cases.CAPTURE)      # ragged: not a multiple of the 32-group tile
    res = api.step(50, flags=0)                     # nothing to do
    assert res.messages == [] and res.fsm == []
    api.run(100, 100, 0, 0)                         # zero ticks
    assert api.query(32, 3).current_term == 0
    res = api.step(100, inject=[])                  # empty inject list
    assert all(m.group < 33 for m in res.messages)
