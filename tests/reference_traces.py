"""Hand-derived multi-replica traces of josefine's src/raft ("next" #6).

Every expected row below was derived BY READING THE RUST SOURCES (tychedelia/josefine @ 28b42c9, paths relative to
src/raft/) and carries the file:line that produces it in `why`.  Neither restatement under oracle/ nor the engine
was run to obtain these values; they are an independent pin for the multi-replica behaviour the reference's own
tests do not cover.  tests/test_reference_traces.py replays them on the C++ restatement, the
Python restatement's cluster-free subset, the device code on the CPU and the GPU.

What is NOT derivable from the Rust and therefore never asserted here: election timeout values (thread_rng in the
reference, deviation D2 here) and anything that depends on them.  All traces drive elections with injected
`Timeout` commands and keep ticks far apart from any timer expiry (timeouts are >= 500 ms, follower.rs:103-108).

The step schedule is the engine's (include/josefine_raft_abi.h, jr_step_args): per replica peer mail of the previous
step in ascending sender order, FIFO per sender; then injected commands; then the dense proposal; then Tick.
Messages are listed in the order jr_step returns them: sender ascending, FIFO per sender.

Trace format (also dumped to tests/golden/reference_traces/*.json by tests/golden/reference_traces/dump.py):
  {"name", "replicas", "config": {jr_config overrides}, "steps": [
     {"now": ms, "deliver": bool, "tick": bool, "inject": [command...], "proposals": [[node, token]] or None,
      "compact": bool (Chain::compact on every replica AFTER the step),
      "messages": [ {from, to ("peers" | node id | "client"), kind, fields..., why} ],
      "fsm": [ {node, kind ("apply"|"notify"), id, next | (client), data, why} ],
      "state": {node: {field: value, ..., "why": ...}},
      "chain": {node: [ids present]} } ]}
"""

# message / instruction helpers ------------------------------------------------------------------


def vreq(frm, term, head, why, copies=1):
    return [dict(kind="VoteRequest", **{"from": frm}, to="peers", term=term, node_id=frm, last_term=term, block=head, why=why)] * copies


def vresp(frm, to, term, granted, why):
    return dict(kind="VoteResponse", **{"from": frm}, to=to, term=term, node_id=frm, flag=int(granted), why=why)


def hb(frm, term, commit, why):
    return dict(kind="Heartbeat", **{"from": frm}, to="peers", term=term, block=commit, node_id=frm, why=why)


def hbresp(frm, to, commit, has, why):
    return dict(kind="HeartbeatResponse", **{"from": frm}, to=to, block=commit, flag=int(has), why=why)


def ae(frm, to, term, blocks, why):
    return dict(kind="AppendEntries", **{"from": frm}, to=to, term=term, node_id=frm, blocks=[list(b) for b in blocks], why=why)


def aresp(frm, to, term, head, why):
    return dict(kind="AppendResponse", **{"from": frm}, to=to, node_id=frm, term=term, block=head, flag=1, why=why)


def apply_(node, bid, nxt, data, why):
    return dict(node=node, kind="apply", id=bid, next=nxt, data=data, why=why)


def notify(node, bid, token, why):
    return dict(node=node, kind="notify", id=bid, data=token, client="client", why=why)


def each(nodes, fn):
    out = []
    for n in nodes:
        r = fn(n)
        out.extend(r if isinstance(r, list) else [r])
    return out


# ------------------------------------------------------------------------------------------------
# T1: R = 3, one candidate.  The N-1 duplicate VoteRequests (candidate.rs:30-37) and the follower's
# grant-then-deny answers (follower.rs:219-246); election by the first granted answer (election.rs:50).

T1 = dict(
    name="r3_election_duplicate_vote_requests", replicas=3, config={},
    steps=[
        dict(now=0, deliver=False, tick=False, inject=[("timeout", 1)],
             messages=vreq(1, 1, 0, "candidate.rs:24-37: voted_for=self, term 0+1, one send_all per entry of config.nodes (2 peers) -> "
                                    "two identical broadcasts; last_term = term (sic), head = chain head 0", copies=2),
             fsm=[],
             state={1: dict(role="candidate", current_term=1, voted_for=1, votes_seen=0b001, votes_granted=0b001,
                            why="follower.rs:248-256 voted_for None -> Candidate; candidate.rs:40-44 self VoteResponse -> election.rs:33-35; "
                                "status: 1 vote < quorum 2, total-votes 0 != 2 -> Voting (election.rs:50-56)"),
                    2: dict(role="follower", current_term=0, voted_for=0, why="nothing applied yet")}),
        dict(now=10, deliver=True, tick=False,
             messages=each((2, 3), lambda n: [
                 vresp(n, 1, 0, True, "follower.rs:97-101 can_vote: voted_for None, current_term 0 > last_term 1 false, commit 0 > head 0 false; "
                                      "follower.rs:225-233 reply carries the follower's OWN term 0, then voted_for = 1"),
                 vresp(n, 1, 0, False, "second copy of the same broadcast: voted_for is Some(1) now -> follower.rs:235-243 denied")]),
             fsm=[],
             state={2: dict(role="follower", current_term=0, voted_for=1, why="follower.rs:233; the request's term is ignored (follower.rs:52-57 `..`)"),
                    3: dict(role="follower", current_term=0, voted_for=1, why="same")}),
        dict(now=20, deliver=True, tick=False,
             messages=[hb(1, 1, 0, "candidate.rs:91-98: vote(2,true) -> 2 votes >= quorum 2 -> elect() candidate.rs:108-113 -> Leader, leader.rs:44-51 "
                                   "heartbeat{term 1, commit 0}; the three later answers reach a Leader, which ignores VoteResponse (leader.rs:263)")],
             fsm=[],
             state={1: dict(role="leader", current_term=1, voted_for=1, heartbeat_time_ms=20, progress_head=[0, 0, 0], progress_replicate=0,
                            why="candidate.rs:216-238: ReplicationProgress::new -> every node Probe, head 0 (progress.rs:15-23,155-162); heartbeat_time = now")}),
        dict(now=30, deliver=True, tick=False,
             messages=each((2, 3), lambda n: hbresp(n, 1, 0, True, "follower.rs:178-217: has(commit 0) true (genesis block), 0 > 0 false -> no commit; "
                                                                   "reply {own commit 0, has_committed true}")),
             fsm=[],
             state={2: dict(role="follower", current_term=1, voted_for=1, leader_id=1, commit=0,
                            why="follower.rs:184-187: term(1) clears voted_for/leader_id (mod.rs:360-365, follower.rs:27-29), then both = leader 1")}),
    ])

# ------------------------------------------------------------------------------------------------
# T2: R = 5 cold start, one candidate: NEVER elected with sender-major delivery, because each follower's grant is
# overwritten by its own later denials (election.rs:33-35 HashMap insert = last write wins) -- SURVEY note N3.

T2 = dict(
    name="r5_single_candidate_defeats_itself", replicas=5, config={},
    steps=[
        dict(now=0, deliver=False, tick=False, inject=[("timeout", 1)],
             messages=vreq(1, 1, 0, "candidate.rs:30-37: config.nodes has 4 entries -> 4 identical broadcasts", copies=4),
             fsm=[], state={1: dict(role="candidate", current_term=1, voted_for=1, votes_seen=1, votes_granted=1, why="as T1")}),
        dict(now=10, deliver=True, tick=False,
             messages=each((2, 3, 4, 5), lambda n: [vresp(n, 1, 0, True, "first copy: follower.rs:225-233")] +
                           [vresp(n, 1, 0, False, "copies 2-4: voted_for = Some(1) -> follower.rs:235-243")] * 3),
             fsm=[], state={5: dict(role="follower", voted_for=1, current_term=0, why="follower.rs:233")}),
        dict(now=20, deliver=True, tick=False,
             messages=[],
             fsm=[],
             state={1: dict(role="follower", current_term=1, voted_for=0, leader_id=0,
                            why="quorum = 5/2+1 = 3 (election.rs:66-73).  Answers arrive sender-major: node 2 grant (votes 2), then node 2's three denials "
                                "overwrite it (votes 1, total 2); node 3 grant (2 votes, total 3), denial (1 vote, total 3: 3-1 = 2 != 3 -> Voting); node 4 grant "
                                "(2, total 4), denial (1 vote, total 4: 4-1 = 3 == quorum -> Defeated, election.rs:52-53) -> candidate.rs:101-105: voted_for None, "
                                "Follower with leader_id None (candidate.rs:198-214).  The remaining answers reach a Follower: follower.rs:61 ignores them"),
                    2: dict(role="follower", current_term=0, voted_for=1,
                            why="never reset: this follower can neither vote again nor start an election (follower.rs:97-101,248-256) -- SURVEY N1")}),
    ])

# ------------------------------------------------------------------------------------------------
# T3: R = 5 steady state from a synthetic election through three commits: Probe <-> Replicate oscillation
# (progress.rs:76-94), leader apply range (leader.rs:87-99) vs follower apply range (follower.rs:198-207),
# head regression on a re-sent block (chain.rs:178-192).  heartbeat_ms = tick = 100: the strict `>`
# (leader.rs:78-80) makes the leader heartbeat every second tick.

_F = (2, 3, 4, 5)
T3 = dict(
    name="r5_steady_state_three_commits", replicas=5, config=dict(heartbeat_ms=100),
    steps=[
        dict(now=0, deliver=False, tick=False,
             inject=[("timeout", 1), ("vote_response", 1, 1, 2, True), ("vote_response", 1, 1, 3, True)],
             messages=vreq(1, 1, 0, "candidate.rs:30-37", copies=4) +
             [hb(1, 1, 0, "self vote + node 2 = 2 votes: Voting; node 3 -> 3 >= quorum 3 -> elect(), heartbeat (candidate.rs:91-113, leader.rs:44-51)")],
             fsm=[], state={1: dict(role="leader", current_term=1, heartbeat_time_ms=0, why="candidate.rs:216-238")}),
        dict(now=100, deliver=True, tick=True, proposals=[(1, 101)],
             messages=each(_F, lambda n: ae(1, n, 1, [(1, 0, 101)], "leader.rs:234-245: heartbeat_time.elapsed() = 100 > 100 false -> no heartbeat; replicate "
                                                                   "leader.rs:130-150: Probe, range(0..).nth(1) = block 1")) +
             each(_F, lambda n: [vresp(n, 1, 0, True, "follower.rs:225-233")] + [vresp(n, 1, 0, False, "follower.rs:235-243")] * 3 +
                  [hbresp(n, 1, 0, True, "follower.rs:178-217: term(1), leader 1, has(0), no commit")]),
             fsm=[notify(1, 1, 101, "leader.rs:177-188: chain.append -> id 1 (id_gen after genesis), next = head 0 (chain.rs:160-175); Notify{id, block_id 1, Client}")],
             state={1: dict(head=1, commit=0, id_gen=2, progress_head=[1, 0, 0, 0, 0], progress_replicate=0b00001,
                            why="leader.rs:190-196 self AppendResponse -> progress.rs:76-94 Probe.increment(1) true -> Replicate; committed_index = "
                                "[1,0,0,0,0] sorted desc, element [2] = 0 (progress.rs:48-60): no commit"),
                    2: dict(current_term=1, voted_for=1, leader_id=1, why="vote set voted_for = 1, the heartbeat cleared and set it again")}),
        dict(now=200, deliver=True, tick=True, proposals=[(1, 102)],
             messages=[hb(1, 1, 0, "leader.rs:237-240: elapsed 200 > 100 -> heartbeat{commit 0}, timer reset")] +
             each(_F, lambda n: ae(1, n, 1, [(1, 0, 101)], "peers still Probe/head 0 (no AppendResponse yet): nth(1) after 0 = block 1 again")) +
             each(_F, lambda n: aresp(n, 1, 1, 1, "follower.rs:130-176: voted_for Some(1) -> first branch skipped; extend block 1 (has(0)), head 1; AppendResponse{term 1, head 1}")),
             fsm=[notify(1, 2, 102, "append id 2 next 1")],
             state={1: dict(head=2, commit=0, progress_head=[2, 0, 0, 0, 0], heartbeat_time_ms=200,
                            why="the 4 VoteResponses and HeartbeatResponse{0,true} per follower change nothing (leader.rs:222-231,263)"),
                    3: dict(head=1, commit=0, why="chain.rs:178-192")}),
        dict(now=300, deliver=True, tick=True, proposals=[(1, 103)],
             messages=each(_F, lambda n: ae(1, n, 1, [(2, 1, 102), (3, 2, 103)], "peers Replicate/head 1: range(1..).skip(1).take(5) = blocks 2, 3 (leader.rs:152-157); "
                                                                                "no heartbeat (elapsed 100)")) +
             each(_F, lambda n: [hbresp(n, 1, 0, True, "follower.rs:178-217 again; commit 0"),
                                 aresp(n, 1, 1, 1, "block 1 re-sent: extend overwrites it, head stays 1")]),
             fsm=[apply_(1, 1, 0, 101, "AppendResponse(node 2, head 1): Probe -> Replicate, heads [2,1,0,0,0] -> [2] = 0; node 3: [2,1,1,0,0] -> 1 > commit 0: "
                                      "leader.rs:87-99 commit(1), range(0..=1).skip(1) = block 1"),
                  notify(1, 3, 103, "append id 3 next 2")],
             state={1: dict(head=3, commit=1, progress_head=[3, 1, 1, 1, 1], progress_replicate=0b11111, why="all four peers incremented 0 -> 1: Replicate")}),
        dict(now=400, deliver=True, tick=True, proposals=[(1, 104)],
             messages=[hb(1, 1, 1, "elapsed 200 > 100: heartbeat{commit 1}")] +
             each(_F, lambda n: ae(1, n, 1, [(2, 1, 102)], "AppendResponse(head 1) again: Replicate.increment(1) false (1 < 1) -> back to Probe (progress.rs:85-91); "
                                                          "Probe sends the single block after head 1")) +
             each(_F, lambda n: aresp(n, 1, 1, 3, "extend 2 then 3: head 3")),
             fsm=[notify(1, 4, 104, "append id 4 next 3")],
             state={1: dict(head=4, commit=1, progress_head=[4, 1, 1, 1, 1], progress_replicate=0b00001, why="progress.rs:85-91"),
                    4: dict(head=3, commit=0, why="heartbeat{commit 1} is still in flight")}),
        dict(now=500, deliver=True, tick=True, proposals=[(1, 105)],
             messages=each(_F, lambda n: ae(1, n, 1, [(4, 3, 104), (5, 4, 105)], "AppendResponse(head 3): Probe -> Replicate/head 3; blocks 4, 5")) +
             each(_F, lambda n: [hbresp(n, 1, 1, True, "follower.rs:198-207: has(1) and 1 > 0: commit(1); reply {commit 1, true}"),
                                 aresp(n, 1, 1, 2, "block 2 re-sent alone: chain.rs:188-190 head = block.id unconditionally -> head REGRESSES 3 -> 2")]),
             fsm=[apply_(1, 2, 1, 102, "node 3's AppendResponse: heads [4,3,3,1,1] -> [2] = 3 > 1: commit(3), range(1..=3).skip(1) = blocks 2, 3"),
                  apply_(1, 3, 2, 103, "same range"),
                  notify(1, 5, 105, "append id 5 next 4")] +
             each(_F, lambda n: apply_(n, 0, 0, 0, "follower.rs:203-206: range(prev 0 .. commit 1) is half-open -> the GENESIS block 0, not block 1 (SURVEY N4)")),
             state={1: dict(head=5, commit=3, progress_head=[5, 3, 3, 3, 3], progress_replicate=0b11111, why=""),
                    2: dict(head=2, commit=1, why="head regressed; blocks 1..3 are all still in the tree")},
             chain={2: [0, 1, 2, 3]}),
        dict(now=600, deliver=True, tick=True, proposals=[(1, 106)],
             messages=[hb(1, 1, 3, "heartbeat{commit 3}")] +
             each(_F, lambda n: ae(1, n, 1, [(4, 3, 104)], "AppendResponse(head 2) < 3: Replicate -> Probe; single block after 3")) +
             each(_F, lambda n: aresp(n, 1, 1, 5, "extend 4, 5")),
             fsm=[notify(1, 6, 106, "append id 6 next 5")],
             state={1: dict(head=6, commit=3, progress_head=[6, 3, 3, 3, 3], progress_replicate=0b00001, why="")}),
        dict(now=700, deliver=True, tick=True, proposals=[(1, 107)],
             messages=each(_F, lambda n: ae(1, n, 1, [(6, 5, 106), (7, 6, 107)], "Replicate/head 5: blocks 6, 7")) +
             each(_F, lambda n: [hbresp(n, 1, 3, True, "commit(3)"), aresp(n, 1, 1, 4, "block 4 re-sent: head 5 -> 4")]),
             fsm=[apply_(1, 4, 3, 104, "heads [6,5,5,3,3] -> 5 > 3: commit(5), range(3..=5).skip(1)"), apply_(1, 5, 4, 105, ""),
                  notify(1, 7, 107, "append id 7 next 6")] +
             each(_F, lambda n: [apply_(n, 1, 0, 101, "follower range(1..3): blocks 1, 2 -- block 3 itself waits for the next commit"),
                                 apply_(n, 2, 1, 102, "")]),
             state={1: dict(head=7, commit=5, why=""), 5: dict(head=4, commit=3, why="")}),
    ])

# ------------------------------------------------------------------------------------------------
# T4: R = 3 split vote: two candidates in the same term; one is elected by the third node's grant, the other is
# defeated by the exact-equality rule (election.rs:52); a candidate answers a same-term VoteRequest with a
# denial (candidate.rs:71-88) and ignores the term carried by VoteResponses (candidate.rs:91).

T4 = dict(
    name="r3_split_vote_defeat_by_equality", replicas=3, config={},
    steps=[
        dict(now=0, deliver=False, tick=False, inject=[("timeout", 1), ("timeout", 2)],
             messages=vreq(1, 1, 0, "candidate.rs:30-37", copies=2) + vreq(2, 1, 0, "candidate.rs:30-37", copies=2),
             fsm=[], state={1: dict(role="candidate", current_term=1, why=""), 2: dict(role="candidate", current_term=1, why="")}),
        dict(now=10, deliver=True, tick=False,
             messages=[vresp(1, 2, 1, False, "candidate.rs:71-88: request term 1 > own 1 false -> denied, reply carries term 1")] * 2 +
             [vresp(2, 1, 1, False, "same")] * 2 +
             [vresp(3, 1, 0, True, "node 3 hears sender 1 first: grant (follower.rs:225-233)"), vresp(3, 1, 0, False, "copy 2"),
              vresp(3, 2, 0, False, "then sender 2: voted_for = Some(1) -> denied"), vresp(3, 2, 0, False, "copy 2")],
             fsm=[], state={3: dict(voted_for=1, current_term=0, why="follower.rs:233")}),
        dict(now=20, deliver=True, tick=False,
             messages=[hb(1, 1, 0, "node 1: two denials from 2 -> votes {1:T,2:F}: 1 < 2, total-votes 1 != 2: Voting; node 3's grant -> 2 votes: elect()")],
             fsm=[],
             state={1: dict(role="leader", current_term=1, why="candidate.rs:108-113"),
                    2: dict(role="follower", current_term=1, voted_for=0, leader_id=0,
                            why="node 2: denials from 1 (votes {2:T,1:F}), then from 3: total 3, votes 1: 3-1 = 2 == quorum 2 -> Defeated (election.rs:52-53), "
                                "candidate.rs:101-105")}),
        dict(now=30, deliver=True, tick=False,
             messages=[hbresp(2, 1, 0, True, "follower.rs:178-217"), hbresp(3, 1, 0, True, "follower.rs:178-217")],
             fsm=[],
             state={2: dict(role="follower", current_term=1, voted_for=1, leader_id=1, why="follower.rs:184-187"),
                    3: dict(role="follower", current_term=1, voted_for=1, leader_id=1, why="term(1) raised node 3 from term 0")}),
    ])

# ------------------------------------------------------------------------------------------------
# T5: a follower accepts AppendEntries with blocks that form a BRANCHED chain (the tree of chain.rs:320-343), commits
# through a heartbeat (key-order apply of both branches), then Chain::compact removes the dead branch (chain.rs:239-253).

T5 = dict(
    name="r3_follower_branched_chain_commit_compact", replicas=3, config={},
    steps=[
        dict(now=0, deliver=False, tick=False,
             inject=[("append_entries", 2, 0, 3, [(1, 0, 11), (2, 1, 12), (3, 2, 13), (4, 3, 14), (5, 3, 15)]),
                     ("append_entries", 2, 0, 3, [(6, 5, 16)])],
             messages=[aresp(2, 3, 0, 5, "follower.rs:138-146: voted_for None and term 0 >= 0: term(0), leader = voted_for = 3; each block extends (parent present); "
                                         "head = last block 5; reply {term 0, head 5}"),
                       aresp(2, 3, 0, 6, "second AppendEntries: voted_for Some(3) == leader; extend 6 (parent 5)")],
             fsm=[], state={2: dict(head=6, commit=0, voted_for=3, leader_id=3, current_term=0, why="")},
             chain={2: [0, 1, 2, 3, 4, 5, 6]}),
        dict(now=10, deliver=False, tick=False, inject=[("heartbeat", 2, 0, 6, 3)], compact=True,
             messages=[hbresp(2, 3, 6, True, "follower.rs:198-215: has(6), 6 > 0: commit(6)")],
             fsm=[apply_(2, b, n, d, "follower.rs:203-206 range(0..6) in KEY order: both branches (4 and 5) are applied, block 6 is not")
                  for (b, n, d) in [(0, 0, 0), (1, 0, 11), (2, 1, 12), (3, 2, 13), (4, 3, 14), (5, 3, 15)]],
             state={2: dict(head=6, commit=6, why="")},
             chain={2: [0, 1, 2, 3, 5, 6]}),   # compact: ids [0,6) descending: 5 kept (first), 4 != next(5)=3 -> removed, expectation = next(4) = 3 -> 3,2,1,0 kept
    ])

# ------------------------------------------------------------------------------------------------
# T6: the panic sites a multi-replica group can reach, as sticky faults (deviation D3).

T6 = dict(
    name="r3_fault_sites", replicas=3, config={},
    steps=[
        dict(now=0, deliver=False, tick=False,
             inject=[("timeout", 1), ("vote_response", 1, 1, 2, True),            # node 1 leads term 1
                     ("append_entries", 1, 5, 2, []),                              # a higher-term AppendEntries reaches the leader
                     ("append_entries", 3, 0, 2, [(4, 3, 9)])],                    # a block whose parent is missing reaches follower 3
             messages=vreq(1, 1, 0, "", copies=2) + [hb(1, 1, 0, "elected by node 2's grant")],
             fsm=[],
             state={1: dict(role="leader", fault=6, current_term=5, voted_for=0,
                            why="leader.rs:200-208: term 5 > 1 -> mod.rs:360-365 sets voted_for None, current_term 5, then Role::term = unimplemented!() "
                                "(leader.rs:33-35) panics: JR_FAULT_LEADER_TERM_UNIMPLEMENTED with the two fields already written"),
                    3: dict(role="follower", fault=2, head=0, voted_for=2, leader_id=2,
                            why="follower.rs:138-146 accepted the leader, then chain.extend Err (chain.rs:180-185) propagates with `?`: JR_FAULT_EXTEND_PARENT_MISSING")}),
    ])

ALL_TRACES = [T1, T2, T3, T4, T5, T6]
