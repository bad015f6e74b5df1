"""Host-side FSM driver for the batched engine -- mirror of josefine's src/raft/fsm.rs.

The engine emits `Instruction`s (fsm.rs:19-29) per replica: `Apply{block}` for every
committed block in the reference's order and `Notify{id, client_address, block_id}` when a
leader appends a client request.  The reference runs one `Driver` per node (fsm.rs:31-93);
`BatchedDriver` keeps one logical driver per (group, node):

  * Apply of block 0 (the genesis block) is skipped            fsm.rs:61-63
  * the block's payload goes through `Fsm::transition`          fsm.rs:15-17,90-92
  * if a Notify is registered for the block id, a ClientResponse is produced for
    the recorded address                                        fsm.rs:66-76
    (`Address::Client` completes the request on this host; `Address::Peer(n)` is a
    proxied request: the response must be delivered to node n, which relays it to
    its client -- follower.rs:271-282; feed it back with `Command.client_response`).

Payload bytes never go to the GPU (deviation D5): blocks carry a 64-bit token that
`payloads` maps to the bytes the client proposed.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Protocol, Tuple

from . import abi
from .raft import Address


class Fsm(Protocol):
    """fsm.rs:15-17"""

    def transition(self, data: bytes) -> bytes: ...


@dataclass
class ClientResponse:
    """mod.rs:152-156 (`res` = Ok(bytes) or an Exception for ResponseError)"""
    group: int
    node: int            # the replica whose driver produced it
    to: Address          # where the reference's rpc_tx message would go
    request: int         # ClientRequest.id token
    result: object


class BatchedDriver:
    def __init__(self, make_fsm: Callable[[int, int], Fsm], payloads: Dict[int, bytes]):
        self._make, self.payloads = make_fsm, payloads
        self.fsms: Dict[Tuple[int, int], Fsm] = {}
        self.notifications: Dict[Tuple[int, int], Dict[int, Tuple[Address, int]]] = {}

    def fsm(self, group: int, node: int) -> Fsm:
        key = (group, node)
        if key not in self.fsms:
            self.fsms[key] = self._make(group, node)
        return self.fsms[key]

    def feed_records(self, lib, records, n_groups: int, n_replicas: int) -> List[ClientResponse]:
        """Consume a batch in the engine's compact form (jr_fsm_record, as jr_fsm_records_wait returns it): the batch is
        expanded by jr_fsm_expand into the exact Instruction order jr_step would have returned and fed as usual.
        (A host that only needs the per-replica apply watermark folds the batch with jr_fsm_fold instead -- one linear
        pass, no expansion; that is what bench.py's end-to-end leg does.)"""
        from .raft import expand_records
        return self.feed(expand_records(lib, list(records), n_groups, n_replicas))

    def feed(self, instructions: List[abi.FsmInstr]) -> List[ClientResponse]:
        """Consume Instructions in emission order (the engine returns them group-major, node
        ascending, FIFO per node -- per-node order is what the reference guarantees)."""
        out: List[ClientResponse] = []
        for ins in instructions:
            key = (ins.group, ins.node)
            notes = self.notifications.setdefault(key, {})
            if ins.kind == abi.FSM_NOTIFY:                       # fsm.rs:78-81
                notes[ins.block.id] = (Address(ins.client_kind, ins.client_id), ins.block.data)
                continue
            if ins.block.id == 0:                                # fsm.rs:61-63
                continue
            try:
                res: object = self.fsm(*key).transition(self.payloads.get(ins.block.data, b""))
            except Exception as e:  # ResponseError, fsm.rs:73
                res = e
            hit: Optional[Tuple[Address, int]] = notes.pop(ins.block.id, None)
            if hit is not None:                                  # fsm.rs:67-76
                out.append(ClientResponse(ins.group, ins.node, hit[0], hit[1], res))
        return out
