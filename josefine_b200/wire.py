"""Peer wire format of josefine's Raft transport, host side.

Reference: src/raft/tcp.rs:39-51 (receive) and 143-156 (send): every `Message`
(src/raft/rpc.rs:17-21) travels as one `LengthDelimitedCodec` frame -- a 4-byte big-endian
length prefix (tokio_util default) -- whose body is `serde_json` of the struct
(`tokio_serde::formats::SymmetricalJson`, compact output).

The JSON shape follows serde's data model for the derives in the reference:
  * `Message {from, to, command}`: object, fields in declaration order
  * `Address` (rpc.rs:5-14): unit variants are strings ("Peers", "Local", "Client"),
    `Peer(n)` is `{"Peer": n}`                                   (externally tagged enum)
  * `Command` (mod.rs:160-227): unit variants are strings ("Tick", "Propose", "Timeout",
    "Noop"); struct variants are `{"Variant": {fields...}}`; newtype variants
    `ClientRequest(..)` / `ClientResponse(..)` are `{"Variant": <inner>}`
  * `BlockId` (chain.rs:29-67) serialises with `serialize_bytes` of its 8 big-endian
    bytes, which serde_json writes as an array of numbers: id 1 -> [0,0,0,0,0,0,0,1]
  * `Block {id, next, data: Vec<u8>}`, `Proposal(Vec<u8>)`, `Response(Vec<u8>)`: byte
    vectors are arrays of numbers; `ClientRequestId` is a hyphenated Uuid string;
    `Result<Response, ResponseError>` is `{"Ok": [...]}` / `{"Err": {}}`.

**Parity unpinned.**  There is no Rust toolchain here, so these bytes were derived from the
serde / serde_json data model, not produced by josefine; tests/test_wire.py pins the
derivation (hand-written expected strings + round trips).  Also note: the
reference's own `deserialize_block_id` borrows `&[u8]`, which serde_json cannot provide
from a number array, so stock josefine probably cannot decode any frame that carries a
BlockId -- `decode_frame` here accepts the array form.

Payload bytes and request ids live on the host (deviation D5): `Codec` maps the engine's
64-bit tokens to them.
"""
from __future__ import annotations

import json
import struct
import uuid
from typing import Dict, List, Optional, Tuple

from . import abi

_UNIT_COMMANDS = {abi.CMD_TICK: "Tick", abi.CMD_PROPOSE: "Propose", abi.CMD_TIMEOUT: "Timeout", abi.CMD_NOOP: "Noop"}
_UNIT_BY_NAME = {v: k for k, v in _UNIT_COMMANDS.items()}


def _block_id(v: int) -> List[int]:
    return list(v.to_bytes(8, "big"))          # BlockId::new, chain.rs:63-66


def _from_block_id(a) -> int:
    return int.from_bytes(bytes(a), "big")


def _address(kind: int, node: int):
    if kind == abi.ADDR_PEER:
        return {"Peer": node}
    return {abi.ADDR_PEERS: "Peers", abi.ADDR_LOCAL: "Local", abi.ADDR_CLIENT: "Client"}[kind]


def _from_address(v) -> Tuple[int, int]:
    if isinstance(v, dict):
        return abi.ADDR_PEER, int(v["Peer"])
    return {"Peers": abi.ADDR_PEERS, "Local": abi.ADDR_LOCAL, "Client": abi.ADDR_CLIENT}[v], 0


class Codec:
    """Token <-> payload / request-id tables plus frame encode / decode."""

    def __init__(self):
        self.payloads: Dict[int, bytes] = {0: b""}
        self.requests: Dict[int, uuid.UUID] = {}
        self.responses: Dict[int, Optional[bytes]] = {}   # request token -> FSM answer; None = ResponseError
        self._by_content: Dict[bytes, int] = {b"": 0}
        self._next = 1

    # -- host-side tables (deviation D5) -------------------------------------------------
    def _new_token(self, data: bytes) -> int:
        tok = self._next
        self._next += 1
        self.payloads[tok] = bytes(data)
        self._by_content.setdefault(bytes(data), tok)
        return tok

    def intern_payload(self, data: bytes) -> int:
        """Token for block payload bytes; equal bytes share a token (followers only ever read them back)."""
        tok = self._by_content.get(bytes(data))
        return self._new_token(data) if tok is None else tok

    def intern_request(self, data: bytes, request_id: Optional[uuid.UUID] = None) -> int:
        """A client request always gets a token of its own: the token is what Notify / ClientResponse carry."""
        tok = self._new_token(data)
        self.requests[tok] = request_id or uuid.uuid4()
        return tok

    def _token_of_request(self, rid: uuid.UUID, data: bytes) -> int:
        for tok, r in self.requests.items():
            if r == rid:
                return tok
        return self.intern_request(data, rid)

    # -- Command <-> JSON value ------------------------------------------------------------
    def command_json(self, m: abi.Msg):
        k = m.kind
        if k in _UNIT_COMMANDS:
            return _UNIT_COMMANDS[k]
        if k == abi.CMD_VOTE_REQUEST:
            return {"VoteRequest": {"term": m.term, "candidate_id": m.node_id, "last_term": m.last_term,
                                    "head": _block_id(m.block)}}
        if k == abi.CMD_VOTE_RESPONSE:
            return {"VoteResponse": {"term": m.term, "from": m.node_id, "granted": bool(m.flag)}}
        if k == abi.CMD_APPEND_ENTRIES:
            blocks = [{"id": _block_id(m.blocks[i].id), "next": _block_id(m.blocks[i].next),
                       "data": list(self.payloads.get(m.blocks[i].data, b""))} for i in range(m.n_blocks)]
            return {"AppendEntries": {"term": m.term, "leader_id": m.node_id, "blocks": blocks}}
        if k == abi.CMD_APPEND_RESPONSE:
            return {"AppendResponse": {"node_id": m.node_id, "term": m.term, "head": _block_id(m.block),
                                       "success": bool(m.flag)}}
        if k == abi.CMD_HEARTBEAT:
            return {"Heartbeat": {"term": m.term, "commit": _block_id(m.block), "leader_id": m.node_id}}
        if k == abi.CMD_HEARTBEAT_RESPONSE:
            return {"HeartbeatResponse": {"commit": _block_id(m.block), "has_committed": bool(m.flag)}}
        if k == abi.CMD_CLIENT_REQUEST:   # mod.rs:145-150
            return {"ClientRequest": {"id": str(self.requests.setdefault(m.token, uuid.uuid4())),
                                      "address": _address(m.client_kind, m.client_id),
                                      "proposal": list(self.payloads.get(m.token, b""))}}
        if k == abi.CMD_CLIENT_RESPONSE:  # mod.rs:152-156; the result bytes are the FSM's answer, kept by the host
            res = self.responses.get(m.token, b"")
            return {"ClientResponse": {"id": str(self.requests.setdefault(m.token, uuid.uuid4())),
                                       "res": {"Err": {}} if res is None else {"Ok": list(res)}}}
        raise ValueError(f"unknown command kind {k}")

    def message_json(self, m: abi.Msg) -> dict:
        return {"from": _address(m.from_kind, m.from_id), "to": _address(m.to_kind, m.to_id),
                "command": self.command_json(m)}

    def encode_frame(self, m: abi.Msg) -> bytes:
        """tcp.rs:143-156: one length-delimited JSON frame."""
        body = json.dumps(self.message_json(m), separators=(",", ":")).encode()
        return struct.pack(">I", len(body)) + body

    # -- decode ------------------------------------------------------------------------------
    def decode_json(self, v: dict, group: int = 0) -> abi.Msg:
        m = abi.Msg()
        m.group = group
        m.from_kind, m.from_id = _from_address(v["from"])
        m.to_kind, m.to_id = _from_address(v["to"])
        c = v["command"]
        if isinstance(c, str):
            m.kind = _UNIT_BY_NAME[c]
            return m
        (name, f), = c.items()
        if name == "VoteRequest":
            m.kind, m.term, m.node_id, m.last_term, m.block = (abi.CMD_VOTE_REQUEST, f["term"], f["candidate_id"],
                                                               f["last_term"], _from_block_id(f["head"]))
        elif name == "VoteResponse":
            m.kind, m.term, m.node_id, m.flag = abi.CMD_VOTE_RESPONSE, f["term"], f["from"], int(f["granted"])
        elif name == "AppendEntries":
            m.kind, m.term, m.node_id = abi.CMD_APPEND_ENTRIES, f["term"], f["leader_id"]
            if len(f["blocks"]) > abi.MAX_AE_BLOCKS:
                raise ValueError("more than MAX_INFLIGHT=5 blocks in one AppendEntries (progress.rs:117)")
            m.n_blocks = len(f["blocks"])
            for i, b in enumerate(f["blocks"]):
                m.blocks[i].id, m.blocks[i].next = _from_block_id(b["id"]), _from_block_id(b["next"])
                m.blocks[i].data = self.intern_payload(bytes(b["data"]))
        elif name == "AppendResponse":
            m.kind, m.node_id, m.term, m.block, m.flag = (abi.CMD_APPEND_RESPONSE, f["node_id"], f["term"],
                                                          _from_block_id(f["head"]), int(f["success"]))
        elif name == "Heartbeat":
            m.kind, m.term, m.block, m.node_id = abi.CMD_HEARTBEAT, f["term"], _from_block_id(f["commit"]), f["leader_id"]
        elif name == "HeartbeatResponse":
            m.kind, m.block, m.flag = abi.CMD_HEARTBEAT_RESPONSE, _from_block_id(f["commit"]), int(f["has_committed"])
        elif name == "ClientRequest":
            m.kind = abi.CMD_CLIENT_REQUEST
            m.client_kind, m.client_id = _from_address(f["address"])
            m.token = self._token_of_request(uuid.UUID(f["id"]), bytes(f["proposal"]))
        elif name == "ClientResponse":
            m.kind = abi.CMD_CLIENT_RESPONSE
            m.token = self._token_of_request(uuid.UUID(f["id"]), b"")
            self.responses[m.token] = bytes(f["res"]["Ok"]) if "Ok" in f["res"] else None
        else:
            raise ValueError(f"unknown command {name}")
        return m

    def decode_frame(self, frame: bytes, group: int = 0) -> Tuple[abi.Msg, bytes]:
        """tcp.rs:39-51: one frame off the front of `frame`; returns (message, rest)."""
        if len(frame) < 4:
            raise ValueError("short frame")
        (n,) = struct.unpack(">I", frame[:4])
        if len(frame) < 4 + n:
            raise ValueError("truncated frame")
        return self.decode_json(json.loads(frame[4:4 + n]), group), frame[4 + n:]
