"""The reference's own known-answer tests for the Raft step path, written once
against the RaftApi interface so they run unchanged on the oracle (CPU) and on
the CUDA engine (GPU, through the C ABI).

Each case names the reference test it ports (paths relative to the reference).
Fixtures mirror src/raft/test/mod.rs:21-41: RaftConfig::default() has no peers,
so `new_follower()` is a single-voter group (R = 1); node id is 1 (the value
the reference's assertions expect, follower.rs:376).
"""
from josefine_b200 import abi, Command

CAPTURE = abi.F_CAPTURE_MESSAGES | abi.F_CAPTURE_FSM


def new_follower(make, **kw):
    """src/raft/test/mod.rs:21-30"""
    return make(1, 1, flags=CAPTURE, **kw)


def new_candidate_cluster(make, **kw):
    """A node in the Candidate role, as new_candidate() (test/mod.rs:32-41) builds by
    `Raft::from(follower)`.  The engine has no type-level conversion, so we use a
    3-replica group and let node 1 time out: it becomes Candidate and, with no
    votes delivered, stays one."""
    api = make(1, 3, flags=CAPTURE, **kw)
    api.apply(Command.timeout(0, 1))
    assert api.handle(0, 1).is_candidate()
    return api


# ---- follower.rs ---------------------------------------------------------------
def kat_follower_to_leader(make):
    """follower.rs:316-324: Timeout on a single voter => Leader"""
    api = new_follower(make)
    api.apply(Command.timeout(0, 1))
    assert api.handle(0, 1).is_leader()


def kat_follower_noop(make):
    """follower.rs:327-335"""
    api = new_follower(make)
    api.apply(Command.noop(0, 1))
    assert api.handle(0, 1).is_follower()


def kat_follower_apply_heartbeat(make):
    """follower.rs:338-358: apply_heartbeat(11, 12, BlockId 1)"""
    api = new_follower(make)
    res = api.apply(Command.heartbeat(0, 1, term=12, commit=1, leader_id=11))
    st = api.query(0, 1)
    assert st.voted_for == 11
    assert st.current_term == 12
    assert st.leader_id == 11
    assert len(res.messages) == 1
    m = res.messages[0]
    # "but we don't have block 1 in our chain"
    assert (m.kind, m.block, m.flag) == (abi.CMD_HEARTBEAT_RESPONSE, 0, 0)
    assert (m.to_kind, m.to_id) == (abi.ADDR_PEER, 11)


def kat_follower_apply_vote_request(make):
    """follower.rs:361-395: granted, then denied once voted_for is set"""
    api = new_follower(make)
    res = api.apply(Command.vote_request(0, 1, term=0, candidate_id=11, last_term=12, head=1))
    st = api.query(0, 1)
    assert st.voted_for == 11
    m = res.messages[0]
    assert (m.kind, m.term, m.node_id, m.flag) == (abi.CMD_VOTE_RESPONSE, 0, 1, 1)
    # the reference pokes voted_for = Some(2); any Some(_) denies -- ours is Some(11)
    res = api.apply(Command.vote_request(0, 1, term=0, candidate_id=11, last_term=12, head=1))
    m = res.messages[0]
    assert (m.kind, m.term, m.node_id, m.flag) == (abi.CMD_VOTE_RESPONSE, 0, 1, 0)


def kat_follower_apply_timeout(make):
    """follower.rs:398-403"""
    api = new_follower(make)
    api.apply(Command.timeout(0, 1))
    assert api.handle(0, 1).is_leader()


def kat_follower_apply_tick(make):
    """follower.rs:406-414: tick before the timeout keeps Follower, after it => Leader"""
    api = new_follower(make)
    timeout = api.query(0, 1).election_timeout_ms
    assert 500 <= timeout < 1000  # gen_range(500..1000), follower.rs:105-106
    api.apply(Command.tick(0, 1), now_ms=timeout)  # elapsed == timeout: strict `>` (mod.rs:354)
    assert api.handle(0, 1).is_follower()
    api.apply(Command.tick(0, 1), now_ms=timeout + 1)
    assert api.handle(0, 1).is_leader()


# ---- candidate.rs ----------------------------------------------------------------
def kat_candidate_apply_heartbeat(make):
    """candidate.rs:247-267: apply_heartbeat(term 11, leader 6, BlockId 1)"""
    api = new_candidate_cluster(make)
    res = api.apply(Command.heartbeat(0, 1, term=11, commit=1, leader_id=6))
    st = api.query(0, 1)
    assert st.role == abi.ROLE_FOLLOWER
    assert st.voted_for == 6
    assert st.current_term == 11
    m = res.messages[0]
    assert (m.kind, m.block, m.flag) == (abi.CMD_HEARTBEAT_RESPONSE, 0, 0)


# ---- leader.rs ---------------------------------------------------------------------
def kat_leader_apply_entry_single_node(make):
    """leader.rs:297-328: propose [123] on a lone leader => Notify then Apply{data 123}"""
    api = new_follower(make)
    api.apply(Command.timeout(0, 1))
    assert api.handle(0, 1).is_leader()
    magic = 123
    res = api.apply(Command.client_request(0, 1, token=magic))
    kinds = [f.kind for f in res.fsm]
    assert kinds == [abi.FSM_NOTIFY, abi.FSM_APPLY]
    assert res.fsm[0].block.id == 1 and res.fsm[0].block.data == magic
    assert res.fsm[0].client_kind == abi.ADDR_CLIENT
    assert (res.fsm[1].block.id, res.fsm[1].block.next, res.fsm[1].block.data) == (1, 0, magic)
    api.apply(Command.tick(0, 1))
    # chain.range(..).take(2).last() is the proposed block
    blocks = api.chain_read(0, 1, 0, 2)
    assert blocks[1] == (1, 0, magic)
    st = api.query(0, 1)
    assert (st.head, st.commit) == (1, 1)


# ---- mod.rs -------------------------------------------------------------------------
def kat_mod_term(make):
    """mod.rs:555-569: term(11) reaches the role (Follower: leader_id reset).
    Observable through the public step API as Heartbeat -> term(term)."""
    api = new_follower(make)
    api.apply(Command.heartbeat(0, 1, term=11, commit=0, leader_id=2))
    assert api.query(0, 1).current_term == 11


def kat_mod_need_election(make):
    """mod.rs:515-531: after election_timeout has elapsed needs_election() holds,
    so Tick turns the (single-voter) follower into a leader"""
    api = new_follower(make)
    t = api.query(0, 1).election_timeout_ms
    api.apply(Command.tick(0, 1), now_ms=t + 1)
    assert not api.handle(0, 1).is_follower()


# ---- server.rs -----------------------------------------------------------------------
def kat_server_event_loop(make):
    """server.rs:179-206: a lone node ticking every 100 ms is Leader within 2 s"""
    api = new_follower(make)
    for k in range(1, 21):
        api.step(k * 100, flags=abi.STEP_DELIVER | abi.STEP_TICK)
    assert api.handle(0, 1).is_leader()


# ---- chain.rs, through commands (direct Chain KATs are in test_oracle_kat.py) ------------
def kat_chain_new(make):
    """chain.rs:262-267"""
    api = new_follower(make)
    st = api.query(0, 1)
    assert (st.commit, st.head) == (0, 0)
    assert api.chain_read(0, 1, 0, 2) == [(0, 0, 0), None]  # genesis 0 -> 0, chain.rs:139-153


def kat_chain_extend_range_has(make):
    """chain.rs:289-324: extend(1 -> 0); head 1, commit 0; range(..) has 2 blocks; has(1)"""
    api = new_follower(make)
    res = api.apply(Command.append_entries(0, 1, term=0, leader_id=9, blocks=[(1, 0, 7)]))
    st = api.query(0, 1)
    assert (st.commit, st.head) == (0, 1)
    assert api.chain_read(0, 1, 0, 3) == [(0, 0, 0), (1, 0, 7), None]
    m = res.messages[0]  # follower.rs:163-172
    assert (m.kind, m.node_id, m.block, m.flag, m.to_id) == (abi.CMD_APPEND_RESPONSE, 1, 1, 1, 9)
    # has(1) observed as has_committed in the HeartbeatResponse
    res = api.apply(Command.heartbeat(0, 1, term=0, commit=1, leader_id=9))
    assert res.messages[0].flag == 1


def kat_chain_compact(make):
    """chain.rs:326-343: tree {(1,0),(2,1),(3,2),(4,3),(5,3),(6,5)}, commit 6 => block 4 removed"""
    api = new_follower(make)
    tree = [(1, 0), (2, 1), (3, 2), (4, 3), (5, 3), (6, 5)]
    api.apply(Command.append_entries(0, 1, term=0, leader_id=9, blocks=[(i, n, 0) for i, n in tree[:5]]))
    api.apply(Command.append_entries(0, 1, term=0, leader_id=9, blocks=[(i, n, 0) for i, n in tree[5:]]))
    assert api.chain_read(0, 1, 4, 1)[0] is not None
    res = api.apply(Command.heartbeat(0, 1, term=0, commit=6, leader_id=9))  # chain.commit(6) via follower.rs:200-203
    assert api.query(0, 1).commit == 6
    # follower applies range(prev..commit) = ids 0..5 in KEY order, dead branch included (follower.rs:204)
    assert [f.block.id for f in res.fsm] == [0, 1, 2, 3, 4, 5]
    api.compact()
    present = [b is not None for b in api.chain_read(0, 1, 0, 7)]
    assert present == [True, True, True, True, False, True, True]


def kat_follower_apply_append_entries(make):
    """follower.rs:416-425 `apply_append_entries`: despite its name the reference test is the election of a
    lone node after the timeout (same body as apply_tick, follower.rs:405-414)."""
    api = make(1, 1, flags=CAPTURE)
    timeout = api.query(0, 1).election_timeout_ms
    api.step(1)
    assert api.handle(0, 1).is_follower()
    api.step(timeout + 1)
    assert api.handle(0, 1).is_leader()


def kat_mod_send_all(make):
    """mod.rs:533-552 `send_all`: a broadcast is Message{from: Peer(id), to: Peers, command} (mod.rs:396-400).
    `send_all(Noop)` itself has no caller; the candidate's VoteRequest broadcast goes through the same function."""
    api = make(1, 3, flags=CAPTURE)
    res = api.apply(Command.timeout(0, 2))
    assert res.messages and all(m.kind == abi.CMD_VOTE_REQUEST for m in res.messages)
    for m in res.messages:
        assert (m.from_kind, m.from_id, m.to_kind) == (abi.ADDR_PEER, 2, abi.ADDR_PEERS)


def kat_config_validation(make):
    """config.rs:128-156: `default()` builds, `validation()` rejects heartbeat_timeout = 1 ms."""
    from josefine_b200 import RaftError
    make(1, 3)                                                   # defaults: 100 ms heartbeat, 500..1000 ms election
    for bad in (dict(heartbeat_ms=1), dict(election_min_ms=4, election_max_ms=9)):
        try:
            make(1, 3, **bad)
        except RaftError as e:
            assert e.status == abi.E_INVAL
        else:
            raise AssertionError(f"{bad} must be rejected")


ALL_KATS = [v for k, v in sorted(globals().items()) if k.startswith("kat_")]


# ---- the drop-in shape: ONE resident node per group, peers remote (RaftConfig::id + nodes) --------
def kat_single_resident_node_with_remote_peers(make):
    """What event_loop (server.rs:103-165) does for one josefine process: node 1 of a 3-node group is
    hosted here, nodes 2 and 3 are remote.  Their mail arrives as injected commands (tcp_rx arm,
    server.rs:127-137) and everything node 1 sends comes back through out_msgs (tcp_tx)."""
    api = make(1, 3, flags=CAPTURE, resident_mask=0b001)
    timeout = api.query(0, 1).election_timeout_ms
    res = api.step(timeout + 1)                                     # Tick: election timer fires
    assert api.handle(0, 1).is_candidate()
    vreq = [m for m in res.messages if m.kind == abi.CMD_VOTE_REQUEST]
    assert len(vreq) == 2 and all(m.to_kind == abi.ADDR_PEERS and m.from_id == 1 for m in vreq)   # N-1 broadcasts
    assert not [m for m in res.messages if m.from_id != 1]          # nodes 2 and 3 are not simulated
    res = api.step(timeout + 101, inject=[Command.vote_response(0, 1, term=0, from_=2, granted=True)])
    assert api.handle(0, 1).is_leader()
    hb = [m for m in res.messages if m.kind == abi.CMD_HEARTBEAT]
    assert hb and hb[0].term == 1
    res = api.step(timeout + 201, inject=[Command.client_request(0, 1, token=55)])
    ae = [m for m in res.messages if m.kind == abi.CMD_APPEND_ENTRIES]
    assert sorted(m.to_id for m in ae) == [2, 3]                    # replicate() to both remote peers
    assert all(m.n_blocks == 1 and m.blocks[0].data == 55 for m in ae)
    assert [f.kind for f in res.fsm] == [abi.FSM_NOTIFY]            # not committed yet: quorum needs a peer
    res = api.step(timeout + 301, inject=[Command.append_response(0, 1, node_id=2, term=1, head=1)])
    assert [(f.kind, f.block.id, f.block.data) for f in res.fsm] == [(abi.FSM_APPLY, 1, 55)]
    st = api.query(0, 1)
    assert (st.commit, st.progress_head[1], st.progress_replicate & 0b010) == (1, 1, 0b010)
    assert api.query(0, 2).alive == 0 and api.query(0, 3).alive == 0


def kat_follower_apply_append_entries(make):
    """follower.rs:416-425 `apply_append_entries`: despite its name the reference test is the election of a
    lone node after the timeout (same body as apply_tick, follower.rs:405-414)."""
    api = make(1, 1, flags=CAPTURE)
    timeout = api.query(0, 1).election_timeout_ms
    api.step(1)
    assert api.handle(0, 1).is_follower()
    api.step(timeout + 1)
    assert api.handle(0, 1).is_leader()


def kat_mod_send_all(make):
    """mod.rs:533-552 `send_all`: a broadcast is Message{from: Peer(id), to: Peers, command} (mod.rs:396-400).
    `send_all(Noop)` itself has no caller; the candidate's VoteRequest broadcast goes through the same function."""
    api = make(1, 3, flags=CAPTURE)
    res = api.apply(Command.timeout(0, 2))
    assert res.messages and all(m.kind == abi.CMD_VOTE_REQUEST for m in res.messages)
    for m in res.messages:
        assert (m.from_kind, m.from_id, m.to_kind) == (abi.ADDR_PEER, 2, abi.ADDR_PEERS)


def kat_config_validation(make):
    """config.rs:128-156: `default()` builds, `validation()` rejects heartbeat_timeout = 1 ms."""
    from josefine_b200 import RaftError
    make(1, 3)                                                   # defaults: 100 ms heartbeat, 500..1000 ms election
    for bad in (dict(heartbeat_ms=1), dict(election_min_ms=4, election_max_ms=9)):
        try:
            make(1, 3, **bad)
        except RaftError as e:
            assert e.status == abi.E_INVAL
        else:
            raise AssertionError(f"{bad} must be rejected")


ALL_KATS = [v for k, v in sorted(globals().items()) if k.startswith("kat_")]
