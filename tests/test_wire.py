"""Peer wire format (josefine_b200/wire.py; reference src/raft/tcp.rs:39-51,143-156).

Parity UNPINNED against josefine itself (no Rust toolchain): the expected strings below are
hand-derived from the serde derives in src/raft/rpc.rs:4-21, src/raft/mod.rs:145-227 and
src/raft/chain.rs:29-91 under serde_json's data model (externally tagged enums, declaration
field order, `serialize_bytes` -> array of numbers, compact separators).
"""
import json
import struct
import uuid

import pytest

from josefine_b200 import Address, Command, abi, msg_tuple
from josefine_b200.wire import Codec


def body(frame: bytes) -> str:
    (n,) = struct.unpack(">I", frame[:4])
    assert len(frame) == 4 + n                      # LengthDelimitedCodec: 4-byte big-endian length
    return frame[4:].decode()


def peer_msg(m, frm, to):
    m.from_kind, m.from_id, m.to_kind, m.to_id = abi.ADDR_PEER, frm, abi.ADDR_PEER, to
    return m


def test_heartbeat_bytes():
    # leader.rs:45-49: send_all(Heartbeat{term, commit, leader_id}) -> Message::new(Peer(id), Peers, ..) (mod.rs:390-400)
    m = Command.heartbeat(0, 0, term=12, commit=258, leader_id=1)
    m.from_kind, m.from_id, m.to_kind, m.to_id = abi.ADDR_PEER, 1, abi.ADDR_PEERS, 0
    assert body(Codec().encode_frame(m)) == (
        '{"from":{"Peer":1},"to":"Peers","command":'
        '{"Heartbeat":{"term":12,"commit":[0,0,0,0,0,0,1,2],"leader_id":1}}}')


def test_vote_request_and_response_bytes():
    # candidate.rs:31-36 (VoteRequest) / follower.rs VoteResponse reply
    vr = peer_msg(Command.vote_request(0, 2, term=3, candidate_id=1, last_term=2, head=7), 1, 2)
    assert body(Codec().encode_frame(vr)) == (
        '{"from":{"Peer":1},"to":{"Peer":2},"command":'
        '{"VoteRequest":{"term":3,"candidate_id":1,"last_term":2,"head":[0,0,0,0,0,0,0,7]}}}')
    vp = peer_msg(Command.vote_response(0, 1, term=3, from_=2, granted=True), 2, 1)
    assert body(Codec().encode_frame(vp)) == (
        '{"from":{"Peer":2},"to":{"Peer":1},"command":{"VoteResponse":{"term":3,"from":2,"granted":true}}}')


def test_append_entries_bytes():
    # leader.rs:144-165: AppendEntries{term, leader_id, blocks}; Block{id, next, data} chain.rs:86-91
    c = Codec()
    t1, t2 = c.intern_payload(b"\x01\x02"), c.intern_payload(b"")
    m = peer_msg(Command.append_entries(0, 3, term=5, leader_id=1, blocks=[(1, 0, t1), (2, 1, t2)]), 1, 3)
    assert body(c.encode_frame(m)) == (
        '{"from":{"Peer":1},"to":{"Peer":3},"command":{"AppendEntries":{"term":5,"leader_id":1,"blocks":['
        '{"id":[0,0,0,0,0,0,0,1],"next":[0,0,0,0,0,0,0,0],"data":[1,2]},'
        '{"id":[0,0,0,0,0,0,0,2],"next":[0,0,0,0,0,0,0,1],"data":[]}]}}}')


def test_append_and_heartbeat_response_bytes():
    ar = peer_msg(Command.append_response(0, 1, node_id=2, term=5, head=2, success=True), 2, 1)
    assert body(Codec().encode_frame(ar)) == (
        '{"from":{"Peer":2},"to":{"Peer":1},"command":'
        '{"AppendResponse":{"node_id":2,"term":5,"head":[0,0,0,0,0,0,0,2],"success":true}}}')
    hr = peer_msg(Command.heartbeat_response(0, 1, commit=1, has_committed=False, from_=2), 2, 1)
    assert body(Codec().encode_frame(hr)) == (
        '{"from":{"Peer":2},"to":{"Peer":1},"command":'
        '{"HeartbeatResponse":{"commit":[0,0,0,0,0,0,0,1],"has_committed":false}}}')


def test_unit_commands_and_addresses():
    for mk, name in ((Command.tick, "Tick"), (Command.timeout, "Timeout"), (Command.noop, "Noop")):
        m = mk(0, 1)
        m.from_kind, m.to_kind = abi.ADDR_LOCAL, abi.ADDR_LOCAL
        assert body(Codec().encode_frame(m)) == '{"from":"Local","to":"Local","command":"%s"}' % name
    m = Command.tick(0, 1)
    m.kind = abi.CMD_PROPOSE
    m.from_kind, m.to_kind = abi.ADDR_CLIENT, abi.ADDR_PEER
    assert body(Codec().encode_frame(m)) == '{"from":"Client","to":{"Peer":1},"command":"Propose"}'


def test_client_request_and_response_bytes():
    # follower.rs:258-263 forwards ClientRequest{id, address, proposal} (mod.rs:145-150) to the leader;
    # follower.rs:272-279 relays ClientResponse{id, res} (mod.rs:152-156)
    c = Codec()
    rid = uuid.UUID("00000000-0000-4000-8000-0000000000aa")
    tok = c.intern_request(b"hi", rid)
    rq = peer_msg(Command.client_request(0, 1, tok, Address.peer(2)), 2, 1)
    assert body(c.encode_frame(rq)) == (
        '{"from":{"Peer":2},"to":{"Peer":1},"command":{"ClientRequest":'
        '{"id":"00000000-0000-4000-8000-0000000000aa","address":{"Peer":2},"proposal":[104,105]}}}')
    c.responses[tok] = b"\x07"
    rp = peer_msg(Command.client_response(0, 2, tok), 1, 2)
    assert body(c.encode_frame(rp)) == (
        '{"from":{"Peer":1},"to":{"Peer":2},"command":{"ClientResponse":'
        '{"id":"00000000-0000-4000-8000-0000000000aa","res":{"Ok":[7]}}}}')
    c.responses[tok] = None                                       # rpc.rs:45-46: ResponseError {}
    assert '"res":{"Err":{}}' in body(c.encode_frame(rp))


def test_reference_tcp_tests_frame():
    # tcp.rs:171-194 `read_message` and tcp.rs:198-229 `send_message` both move
    # Message::new(Address::Peer(1), Address::Peer(2), Command::Tick) as serde_json::to_string in one frame
    c = Codec()
    m = peer_msg(Command.tick(0, 2), 1, 2)
    frame = c.encode_frame(m)
    assert body(frame) == '{"from":{"Peer":1},"to":{"Peer":2},"command":"Tick"}'
    back, rest = c.decode_frame(frame)
    assert rest == b"" and msg_tuple(back) == msg_tuple(m)


def all_kinds(c: Codec):
    t = c.intern_payload(b"abc")
    q = c.intern_request(b"req")
    return [
        peer_msg(Command.vote_request(3, 2, 9, 1, 8, 1 << 40), 1, 2),
        peer_msg(Command.vote_response(3, 1, 9, 2, False), 2, 1),
        peer_msg(Command.append_entries(3, 2, 9, 1, [(i + 1, i, t) for i in range(5)]), 1, 2),
        peer_msg(Command.append_entries(3, 2, 9, 1, []), 1, 2),
        peer_msg(Command.append_response(3, 1, 2, 9, 5, False), 2, 1),
        peer_msg(Command.heartbeat(3, 2, 9, 4, 1), 1, 2),
        peer_msg(Command.heartbeat_response(3, 1, 4, True, 2), 2, 1),
        peer_msg(Command.client_request(3, 1, q, Address.client()), 2, 1),
        peer_msg(Command.client_response(3, 2, q), 1, 2),
        Command.tick(3, 1), Command.timeout(3, 1), Command.noop(3, 1),
    ]


def test_round_trip_every_command():
    c = Codec()
    for m in all_kinds(c):
        frame = c.encode_frame(m)
        back, rest = c.decode_frame(frame, group=3)
        assert rest == b""
        # block payload tokens are re-interned on decode: compare the bytes they stand for
        for i in range(m.n_blocks):
            assert c.payloads[back.blocks[i].data] == c.payloads[m.blocks[i].data]
            back.blocks[i].data = m.blocks[i].data
        assert msg_tuple(back) == msg_tuple(m)
        assert c.encode_frame(back) == frame


def test_stream_of_frames_and_errors():
    c = Codec()
    msgs = all_kinds(c)
    stream = b"".join(c.encode_frame(m) for m in msgs)
    kinds = []
    while stream:
        m, stream = c.decode_frame(stream)
        kinds.append(m.kind)
    assert kinds == [m.kind for m in msgs]
    with pytest.raises(ValueError):
        c.decode_frame(b"\x00\x00")
    with pytest.raises(ValueError):
        c.decode_frame(struct.pack(">I", 100) + b"{}")
    six = {"from": "Local", "to": "Local", "command": {"AppendEntries": {"term": 1, "leader_id": 1, "blocks": [
        {"id": [0] * 8, "next": [0] * 8, "data": []}] * 6}}}
    with pytest.raises(ValueError):
        c.decode_json(six)


def test_block_id_is_big_endian():
    # chain.rs:63-66 BlockId::new(val) = val.to_be_bytes()
    m = peer_msg(Command.heartbeat(0, 2, 1, 0x0102030405060708, 1), 1, 2)
    v = json.loads(body(Codec().encode_frame(m)))
    assert v["command"]["Heartbeat"]["commit"] == [1, 2, 3, 4, 5, 6, 7, 8]


def test_three_processes_over_the_wire_match_resident_cluster():
    from oracle.restated import RestatedCluster
    from tests.wire_cluster import run_networked_vs_resident
    frames, nbytes = run_networked_vs_resident(RestatedCluster.create)
    assert frames > 100 and nbytes > frames * 60
