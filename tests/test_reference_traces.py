"""Replays the hand-derived reference traces (tests/reference_traces.py: every expected row cites the Rust line that
produces it and none was produced by running a restatement) on every implementation: the C++ restatement, the Python
restatement, the device code on the CPU, and the GPU."""
import json
import os

import pytest

from josefine_b200 import abi, Command
from tests import reference_traces as rt

CAP = abi.F_CAPTURE_MESSAGES | abi.F_CAPTURE_FSM
KINDS = {"VoteRequest": abi.CMD_VOTE_REQUEST, "VoteResponse": abi.CMD_VOTE_RESPONSE, "AppendEntries": abi.CMD_APPEND_ENTRIES,
         "AppendResponse": abi.CMD_APPEND_RESPONSE, "Heartbeat": abi.CMD_HEARTBEAT, "HeartbeatResponse": abi.CMD_HEARTBEAT_RESPONSE,
         "ClientRequest": abi.CMD_CLIENT_REQUEST, "ClientResponse": abi.CMD_CLIENT_RESPONSE}
ROLES = {"follower": abi.ROLE_FOLLOWER, "candidate": abi.ROLE_CANDIDATE, "leader": abi.ROLE_LEADER}


def _command(c):
    name, to, *rest = c
    if name == "timeout":
        return Command.timeout(0, to)
    if name == "vote_response":
        term, frm, granted = rest
        return Command.vote_response(0, to, term, frm, granted)
    if name == "append_entries":
        term, leader, blocks = rest
        return Command.append_entries(0, to, term, leader, blocks)
    if name == "heartbeat":
        term, commit, leader = rest
        return Command.heartbeat(0, to, term, commit, leader)
    if name == "client_request":
        return Command.client_request(0, to, rest[0])
    raise ValueError(name)


def _msg_view(m):
    to = "peers" if m.to_kind == abi.ADDR_PEERS else ("client" if m.to_kind == abi.ADDR_CLIENT else m.to_id)
    return {"from": m.from_id, "to": to, "kind": m.kind, "flag": m.flag, "node_id": m.node_id, "term": m.term,
            "last_term": m.last_term, "block": m.block, "token": m.token,
            "blocks": [[m.blocks[i].id, m.blocks[i].next, m.blocks[i].data] for i in range(m.n_blocks)]}


def _msg_want(e):
    w = {"from": e["from"], "to": e["to"], "kind": KINDS[e["kind"]], "flag": 0, "node_id": 0, "term": 0, "last_term": 0,
         "block": 0, "token": 0, "blocks": []}
    w.update({k: v for k, v in e.items() if k in w and k != "kind"})
    return w


def _fsm_view(f):
    if f.kind == abi.FSM_NOTIFY:
        return {"node": f.node, "kind": "notify", "id": f.block.id, "data": f.block.data,
                "client": "client" if f.client_kind == abi.ADDR_CLIENT else (f.client_kind, f.client_id)}
    return {"node": f.node, "kind": "apply", "id": f.block.id, "next": f.block.next, "data": f.block.data}


def replay(make, trace):
    R = trace["replicas"]
    api = make(1, R, seed=11, flags=CAP, chain_capacity=64, **trace["config"])
    for k, st in enumerate(trace["steps"]):
        where = f"{trace['name']} step {k}"
        flags = (abi.STEP_DELIVER if st["deliver"] else 0) | (abi.STEP_TICK if st["tick"] else 0)
        props = None
        if st.get("proposals"):
            props = [tuple(p) for p in st["proposals"]]
        res = api.step(st["now"], flags=flags, inject=[_command(c) for c in st.get("inject", [])], proposals=props)
        got = [_msg_view(m) for m in res.messages]
        want = [_msg_want(e) for e in st["messages"]]
        for i, (g, w) in enumerate(zip(got, want)):
            assert g == w, f"{where}: message #{i}\n  got  {g}\n  want {w}\n  why  {st['messages'][i]['why']}"
        assert len(got) == len(want), f"{where}: {len(got)} messages, expected {len(want)}; extra: {(got + want)[min(len(got), len(want))]}"
        gf = [_fsm_view(f) for f in res.fsm]
        wf = [{k2: v for k2, v in e.items() if k2 != "why"} for e in st["fsm"]]
        assert gf == wf, f"{where}: Instruction stream\n  got  {gf}\n  want {wf}"
        if st.get("compact"):
            api.compact()
        for node, fields in st.get("state", {}).items():
            s = api.query(0, int(node))
            for name, want_v in fields.items():
                if name == "why":
                    continue
                got_v = getattr(s, name)
                if name == "role":
                    want_v = ROLES[want_v]
                if name == "progress_head":
                    got_v = list(got_v)[:R]
                assert got_v == want_v, f"{where}: node {node} {name} = {got_v}, expected {want_v}\n  why {fields.get('why')}"
        for node, ids in st.get("chain", {}).items():
            present = [b[0] for b in api.chain_read(0, int(node), 0, 16) if b is not None]
            assert present == ids, f"{where}: node {node} holds blocks {present}, expected {ids}"


def _cpp(g, r, **kw):
    from oracle.restated import RestatedCluster
    return RestatedCluster.create(g, r, **kw)


def _py(g, r, **kw):
    from oracle.restated_raft import PyCluster
    return PyCluster.create(g, r, **kw)


def _emu(g, r, **kw):
    from tests.emu.emu import EmuEngine
    return EmuEngine.create(g, r, **kw)


def _gpu(g, r, **kw):
    from josefine_b200 import RaftEngine
    return RaftEngine.create(g, r, **kw)


IDS = [t["name"] for t in rt.ALL_TRACES]


@pytest.mark.parametrize("trace", rt.ALL_TRACES, ids=IDS)
def test_trace_on_cpp_restatement(trace):
    replay(_cpp, trace)


@pytest.mark.parametrize("trace", rt.ALL_TRACES, ids=IDS)
def test_trace_on_python_restatement(trace):
    replay(_py, trace)


@pytest.mark.parametrize("trace", rt.ALL_TRACES, ids=IDS)
def test_trace_on_device_code(trace):
    replay(_emu, trace)


@pytest.mark.gpu
@pytest.mark.parametrize("trace", rt.ALL_TRACES, ids=IDS)
def test_trace_on_gpu(trace):
    replay(_gpu, trace)


def test_json_fixtures_are_the_committed_dump():
    """tests/golden/reference_traces/*.json (for consumers that do not speak Python) are a plain dump of the module."""
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_traces")
    for t in rt.ALL_TRACES:
        with open(os.path.join(d, t["name"] + ".json")) as f:
            assert json.load(f) == json.loads(json.dumps(t))
