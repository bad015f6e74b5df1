"""FSM driver / client path cases, run on every implementation."""
from josefine_b200 import abi, Address, BatchedDriver, Command
from tests import kat_cases


class TestFsm:
    """src/raft/fsm.rs:101-131: state A/B set by the payload"""

    def __init__(self):
        self.state = "A"

    def transition(self, data: bytes) -> bytes:
        s = data.decode()
        assert s in ("A", "B")
        self.state = s
        return b""


def case_transition(make):
    """fsm.rs:133-159: an Apply{block 2 -> 1, data "B"} drives the Fsm to state B"""
    drv = BatchedDriver(lambda g, n: TestFsm(), {7: b"B"})
    ins = abi.FsmInstr()
    ins.group, ins.node, ins.kind = 0, 1, abi.FSM_APPLY
    ins.block.id, ins.block.next, ins.block.data = 2, 1, 7
    assert drv.feed([ins]) == []
    assert drv.fsm(0, 1).state == "B"


def case_block_zero_is_skipped(make):
    """fsm.rs:61-63 with the follower's half-open apply range (follower.rs:204), which
    starts at the previous commit = genesis block 0"""
    api = make(1, 1, flags=kat_cases.CAPTURE)
    api.apply(Command.append_entries(0, 1, term=0, leader_id=9, blocks=[(1, 0, 5)]))
    res = api.apply(Command.heartbeat(0, 1, term=0, commit=1, leader_id=9))
    assert [f.block.id for f in res.fsm] == [0]           # range(0..1) is just the genesis block
    drv = BatchedDriver(lambda g, n: TestFsm(), {5: b"B"})
    assert drv.feed(res.fsm) == [] and (0, 1) not in drv.fsms


def case_single_node_propose_completes(make):
    """leader.rs:297-328 + fsm.rs:57-77: Notify then Apply -> ClientResponse to Address::Client"""
    api = make(1, 1, flags=kat_cases.CAPTURE)
    api.apply(Command.timeout(0, 1))
    res = api.apply(Command.client_request(0, 1, token=123))
    drv = BatchedDriver(lambda g, n: TestFsm(), {123: b"B"})
    out = drv.feed(res.fsm)
    assert len(out) == 1 and out[0].to == Address.client() and out[0].request == 123 and out[0].result == b""
    assert drv.fsm(0, 1).state == "B"


def case_proxied_request_round_trip(make):
    """A client talks to a FOLLOWER: follower.rs:258-269 proxies to the leader with
    address = Peer(follower); the leader's driver answers Peer(follower) (fsm.rs:67-76);
    the follower relays to its client (follower.rs:271-282, server.rs:144-151)."""
    api = make(1, 3, flags=kat_cases.CAPTURE)
    api.step(0, flags=0, inject=[Command.timeout(0, 1), Command.vote_response(0, 1, 1, 2, True)])
    assert api.handle(0, 1).is_leader()
    api.step(100)                                            # heartbeat reaches the followers
    assert api.query(0, 2).leader_id == 1
    drv = BatchedDriver(lambda g, n: TestFsm(), {77: b"B"})
    responses, relayed = [], []
    res = api.step(200, inject=[Command.client_request(0, 2, token=77)])   # client -> follower 2
    fwd = [m for m in res.messages if m.kind == abi.CMD_CLIENT_REQUEST]
    assert len(fwd) == 1 and (fwd[0].from_id, fwd[0].to_id, fwd[0].client_kind, fwd[0].client_id) == (2, 1, abi.ADDR_PEER, 2)
    for k in range(3, 12):
        inject = [Command.client_response(r.group, r.to.id, r.request) for r in responses if r.to.kind == abi.ADDR_PEER]
        responses = []
        res = api.step(100 * k, inject=inject)
        relayed += [m for m in res.messages if m.kind == abi.CMD_CLIENT_RESPONSE and m.to_kind == abi.ADDR_CLIENT]
        responses = [r for r in drv.feed(res.fsm) if r.node == 1]           # only the leader holds the Notify
    assert len(relayed) == 1 and relayed[0].from_id == 2 and relayed[0].token == 77
    assert drv.fsm(0, 1).state == "B"
    # every replica that applied block 1 ran the same transition
    assert all(f.state == "B" for f in drv.fsms.values())


ALL_FSM_CASES = [v for k, v in sorted(globals().items()) if k.startswith("case_")]
