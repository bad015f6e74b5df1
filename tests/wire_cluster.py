"""Three josefine 'processes' -- each hosting ONE node of the same group (resident_mask, INTEGRATION.md
section 1) -- exchanging length-delimited JSON frames (josefine_b200/wire.py), compared step by step
with the all-resident 3-replica cluster.  Frames are delivered in the engine's own mail order
(ascending sender, FIFO), so the two arrangements must agree exactly.

Reference shape: src/raft/server.rs:103-165 event_loop (tcp_rx arm -> raft.apply, rpc_rx arm -> tcp_tx)
with src/raft/tcp.rs framing in between.
"""
from josefine_b200 import abi
from josefine_b200.wire import Codec

CAPTURE = abi.F_CAPTURE_MESSAGES | abi.F_CAPTURE_FSM
FIELDS = ("current_term", "voted_for", "leader_id", "head", "commit", "id_gen", "role", "fault",
          "votes_granted", "progress_head", "progress_replicate", "n_queued", "election_time_ms", "rng_draws")


def view(api, node):
    d = api.query(0, node).as_dict()
    return tuple(tuple(d[f]) if isinstance(d[f], list) else d[f] for f in FIELDS)


def run_networked_vs_resident(make, n_steps=60, R=3, seed=11):
    whole = make(1, R, flags=CAPTURE, seed=seed)
    procs = {n: make(1, R, flags=CAPTURE, seed=seed, resident_mask=1 << (n - 1)) for n in range(1, R + 1)}
    codec = Codec()                                  # one token space for the test: see wire.Codec.intern_payload
    wires = {n: [] for n in procs}                   # node -> frames in flight (bytes), in delivery order
    now, n_frames, n_bytes, applied = 0, 0, 0, 0
    for step in range(n_steps):
        now += 100
        prop = None
        if step >= 25 and step % 3 == 0:             # a client proposal at a rotating node (follower or leader)
            node = 1 + step % R
            prop = [(node, codec.intern_request(b"payload-%d" % step))]
        whole.step(now, proposals=prop)
        outs = {}
        for n, p in procs.items():
            inbox = []
            for frame in wires[n]:
                m, rest = codec.decode_frame(frame)
                assert rest == b""
                m.to_kind, m.to_id = abi.ADDR_PEER, n     # a `Peers` broadcast is applied by each receiver
                inbox.append(m)
            mine = prop if prop and prop[0][0] == n else None
            outs[n] = p.step(now, inject=inbox, proposals=mine)
            applied += sum(1 for f in outs[n].fsm if f.kind == abi.FSM_APPLY)
        wires = {n: [] for n in procs}
        for sender in sorted(outs):                  # ascending sender, FIFO within a sender
            for m in outs[sender].messages:
                assert m.from_id == sender
                frame = codec.encode_frame(m)
                n_frames, n_bytes = n_frames + 1, n_bytes + len(frame)
                dests = [m.to_id] if m.to_kind == abi.ADDR_PEER else [n for n in procs if n != sender]
                for d in dests:
                    wires[d].append(frame)
        for n, p in procs.items():
            assert view(p, n) == view(whole, n), (step, n)
            head = whole.query(0, n).head
            assert p.chain_read(0, n, 0, head + 2) == whole.chain_read(0, n, 0, head + 2), (step, n)
    roles = [whole.query(0, n).role for n in procs]
    assert roles.count(abi.ROLE_LEADER) == 1 and whole.query(0, 1).commit > 0 and applied > 0
    return n_frames, n_bytes
