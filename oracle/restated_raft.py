"""Second, independently written RESTATEMENT of josefine's src/raft -- pure Python.

TEST INFRASTRUCTURE ONLY.  It exists because the reference's own tests pin only
single-voter behaviour: everything multi-replica would
otherwise rest on ONE reading of the Rust source.  This file was written from
the reference again, role-per-class like the reference's typestate
(Raft<Follower> / Raft<Candidate> / Raft<Leader>), with plain dicts and sorted
key lists instead of the C++ oracle's std::map objects, and is compared
differentially against the C++ restatement in tests/test_oracle_differential.py.
It is not josefine and is never on the product path.

Citations are file:line in the reference checkout.
"""
from __future__ import annotations

import bisect
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

from josefine_b200 import abi

MASK64 = (1 << 64) - 1


def mix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & MASK64
    return x ^ (x >> 31)


def fold(h: int, w: int) -> int:
    return mix64((h ^ w) & MASK64)


def election_timeout(seed, group, node, draw, mn, mx) -> int:
    """Deviation D2, normative text in include/josefine_raft_abi.h."""
    x = mix64(seed ^ 0x6A09E667F3BCC908)
    x = mix64((x + group) & MASK64)
    x = mix64((x + ((node << 32) | draw)) & MASK64)
    return mn + (((x >> 32) * (mx - mn)) >> 32)


class Panic(Exception):
    def __init__(self, code):
        self.code = code


# --------------------------------------------------------------------------- chain.rs
class Chain:
    """chain.rs:99-253 over a dict + sorted key list (sled's ordered keyspace)."""

    def __init__(self, capacity: int, strict: bool):
        self.blocks: Dict[int, Tuple[int, int]] = {}  # id -> (next, data)
        self.keys: List[int] = []
        self.capacity, self.strict = capacity, strict
        self.has_commit_key = False
        self.id_gen = 0
        self.commit_id = 0
        self.head = 0
        first = self._next_id()          # chain.rs:139-153
        assert first == 0
        self._insert(0, 0, 0)

    def _next_id(self) -> int:           # chain.rs:24-26
        v = self.id_gen
        self.id_gen += 1
        return v

    def _insert(self, bid, nxt, data):
        if bid not in self.blocks:
            bisect.insort(self.keys, bid)
        self.blocks[bid] = (nxt, data)

    def has(self, bid) -> bool:          # chain.rs:155-157
        return bid in self.blocks

    def append(self, data) -> int:       # chain.rs:160-175
        bid = self._next_id()
        if not bid > self.head:
            raise Panic(abi.FAULT_APPEND_ID_NOT_GT_HEAD)
        if bid >= self.capacity:
            raise Panic(abi.FAULT_ENGINE_CHAIN_CAPACITY)
        self._insert(bid, self.head, data)
        self.head = bid
        return bid

    def extend(self, bid, nxt, data):    # chain.rs:178-192
        if not self.has(nxt):
            raise Panic(abi.FAULT_EXTEND_PARENT_MISSING)
        if bid >= self.capacity:
            raise Panic(abi.FAULT_ENGINE_CHAIN_CAPACITY)
        self._insert(bid, nxt, data)
        self.head = bid

    def commit(self, bid):               # chain.rs:195-205
        if bid not in self.blocks:
            raise Panic(abi.FAULT_COMMIT_BLOCK_MISSING)
        self.has_commit_key = True
        self.commit_id = bid

    def keys_between(self, lo, hi, inclusive):  # chain.rs:208-228, bounded forms
        i = bisect.bisect_left(self.keys, lo)
        j = bisect.bisect_right(self.keys, hi) if inclusive else bisect.bisect_left(self.keys, hi)
        return self.keys[i:j]

    def from_skip_take(self, lo, skip, take):
        """range(lo..).skip(skip).take(take): pulls skip+take items; running off the last
        block meets the "commit" key and panics in bincode (chain.rs:198,221-226) -- D6."""
        i = bisect.bisect_left(self.keys, lo)
        avail = self.keys[i:i + skip + take]
        if len(avail) < skip + take and self.strict and self.has_commit_key:
            raise Panic(abi.FAULT_RANGE_COMMIT_KEY)
        return avail[skip:]

    def compact(self):                   # chain.rs:239-253
        expected = None
        for bid in reversed(self.keys_between(0, self.commit_id, False)):
            nxt = self.blocks[bid][0]
            if expected is not None and bid != expected:
                del self.blocks[bid]
                self.keys.remove(bid)
            expected = nxt


# --------------------------------------------------------------------------- shared state
@dataclass
class Shared:
    """Raft<T> fields common to all roles (mod.rs:326-341) + State (mod.rs:271-287)."""
    id: int
    peers: List[int]
    seed: int
    group: int
    emin: int
    emax: int
    hb: int
    chain: Chain
    term: int = 0
    voted_for: Optional[int] = None
    election_time: int = 0
    election_timeout: int = 0
    draws: int = 0
    now: int = 0
    rpc: List[tuple] = field(default_factory=list)   # (to_kind, to_id, cmd dict)
    fsm: List[tuple] = field(default_factory=list)

    def send(self, to_kind, to_id, **cmd):            # mod.rs:390-400
        self.rpc.append((to_kind, to_id, cmd))

    def needs_election(self) -> bool:                 # mod.rs:352-357
        return max(self.now - self.election_time, 0) > self.election_timeout

    def set_election_timeout(self):                   # follower.rs:103-113
        self.election_timeout = election_timeout(self.seed, self.group, self.id, self.draws, self.emin, self.emax)
        self.draws += 1
        self.election_time = self.now


def cmd(kind, **kw):
    d = dict(kind=kind, term=0, node_id=0, last_term=0, block=0, flag=False, blocks=[], token=0,
             client=(abi.ADDR_PEERS, 0))
    d.update(kw)
    return d


# --------------------------------------------------------------------------- roles
class Follower:
    role = abi.ROLE_FOLLOWER

    def __init__(self, s: Shared, queued=None):
        self.s = s
        self.leader_id: Optional[int] = None
        self.queued: List[tuple] = list(queued or [])   # (token, (addr kind, id))

    def term(self, t):                                # mod.rs:360-365 + follower.rs:27-29
        self.s.voted_for = None
        self.s.term = t
        self.leader_id = None

    def apply(self, c):                               # follower.rs:38-63
        k, s = c["kind"], self.s
        if k == abi.CMD_TICK:                         # follower.rs:121-128
            return self.timeout() if s.needs_election() else self
        if k == abi.CMD_TIMEOUT:
            return self.timeout()
        if k == abi.CMD_APPEND_ENTRIES:               # follower.rs:130-176
            ldr = c["node_id"]
            if s.voted_for is None and c["term"] >= s.term:
                self.term(c["term"])
                s.election_time = s.now
                self.leader_id = ldr
                s.voted_for = ldr
            if s.voted_for is not None and s.voted_for != ldr and c["term"] < s.term:
                raise Panic(abi.FAULT_AE_STALE_LEADER)
            if c["blocks"]:
                for (bid, nxt, data) in c["blocks"]:
                    s.chain.extend(bid, nxt, data)
                s.send(abi.ADDR_PEER, ldr, kind=abi.CMD_APPEND_RESPONSE, node_id=s.id, term=s.term,
                       block=s.chain.head, flag=True)
            return self
        if k == abi.CMD_HEARTBEAT:                    # follower.rs:178-217
            ldr = c["node_id"]
            s.set_election_timeout()
            self.term(c["term"])
            self.leader_id = ldr
            s.voted_for = ldr
            q, self.queued = self.queued, []
            for (tok, addr) in q:
                s.send(abi.ADDR_PEER, ldr, kind=abi.CMD_CLIENT_REQUEST, token=tok, client=addr)
            has = s.chain.has(c["block"])
            if has and c["block"] > s.chain.commit_id:
                prev = s.chain.commit_id
                s.chain.commit(c["block"])
                for bid in s.chain.keys_between(prev, c["block"], False):   # prev..commit
                    nxt, data = s.chain.blocks[bid]
                    s.fsm.append(("apply", bid, nxt, data))
            s.send(abi.ADDR_PEER, ldr, kind=abi.CMD_HEARTBEAT_RESPONSE, block=s.chain.commit_id, flag=has)
            return self
        if k == abi.CMD_VOTE_REQUEST:                 # follower.rs:97-101,219-246
            can = not (s.voted_for is not None or s.term > c["last_term"] or s.chain.commit_id > c["block"])
            s.send(abi.ADDR_PEER, c["node_id"], kind=abi.CMD_VOTE_RESPONSE, term=s.term, node_id=s.id, flag=can)
            if can:
                s.voted_for = c["node_id"]
            return self
        if k == abi.CMD_CLIENT_REQUEST:               # follower.rs:258-269
            addr = (abi.ADDR_PEER, s.id)
            if self.leader_id is not None:
                s.send(abi.ADDR_PEER, self.leader_id, kind=abi.CMD_CLIENT_REQUEST, token=c["token"], client=addr)
            else:
                if len(self.queued) >= abi.CLIENT_QUEUE_CAP:
                    raise Panic(abi.FAULT_ENGINE_QUEUE_OVERFLOW)
                self.queued.append((c["token"], addr))
            return self
        if k == abi.CMD_CLIENT_RESPONSE:              # follower.rs:271-282
            s.send(abi.ADDR_CLIENT, 0, kind=abi.CMD_CLIENT_RESPONSE, token=c["token"])
            return self
        return self

    def timeout(self):                                # follower.rs:248-256
        s = self.s
        if s.voted_for is not None:
            return self
        s.set_election_timeout()
        return Candidate(s).seek_election()           # follower.rs:285-304: queue not carried


class Candidate:
    role = abi.ROLE_CANDIDATE

    def __init__(self, s: Shared):
        self.s = s
        self.votes: Dict[int, bool] = {}              # election.rs:8
        self.queued: List[tuple] = []

    def term(self, t):                                # candidate.rs:161-163
        self.s.voted_for = None
        self.s.term = t
        self.votes.clear()

    def status(self):                                 # election.rs:37-73
        n = len(self.s.peers) + 1
        quorum = 0 if n == 1 else n // 2 + 1
        yes = sum(1 for v in self.votes.values() if v)
        if yes >= quorum:
            return "elected"
        if len(self.votes) - yes == quorum:
            return "defeated"
        return "voting"

    def seek_election(self):                          # candidate.rs:24-45
        s = self.s
        s.voted_for = s.id
        s.term += 1
        for _ in s.peers:
            s.send(abi.ADDR_PEERS, 0, kind=abi.CMD_VOTE_REQUEST, term=s.term, node_id=s.id, last_term=s.term,
                   block=s.chain.head)
        return self.apply(cmd(abi.CMD_VOTE_RESPONSE, node_id=s.id, term=s.term, flag=True))

    def to_follower(self):                            # candidate.rs:198-214
        return Follower(self.s, queued=self.queued)

    def apply(self, c):                               # candidate.rs:170-196
        k, s = c["kind"], self.s
        if k == abi.CMD_TICK:                         # candidate.rs:48-68
            if not s.needs_election():
                return self
            if self.status() == "elected":
                raise Panic(abi.FAULT_CANDIDATE_TICK_ELECTED)
            s.voted_for = None
            return self.to_follower().apply(cmd(abi.CMD_TIMEOUT))
        if k == abi.CMD_VOTE_REQUEST:                 # candidate.rs:71-88
            if c["term"] > s.term:
                self.term(c["term"])
                return self.to_follower()
            s.send(abi.ADDR_PEER, c["node_id"], kind=abi.CMD_VOTE_RESPONSE, node_id=s.id, term=s.term, flag=False)
            return self
        if k == abi.CMD_VOTE_RESPONSE:                # candidate.rs:91-113
            self.votes[c["node_id"]] = bool(c["flag"])
            st = self.status()
            if st == "elected":
                ldr = Leader(s)
                ldr.heartbeat()
                return ldr
            if st == "defeated":
                s.voted_for = None
                return self.to_follower()
            return self
        if k == abi.CMD_APPEND_ENTRIES:               # candidate.rs:116-134
            return self.to_follower() if c["term"] >= s.term else self
        if k == abi.CMD_HEARTBEAT:                    # candidate.rs:137-157
            has = s.chain.has(c["block"])
            own = s.chain.commit_id
            self.term(c["term"])
            s.voted_for = c["node_id"]
            f = self.to_follower()
            s.send(abi.ADDR_PEER, c["node_id"], kind=abi.CMD_HEARTBEAT_RESPONSE, block=own, flag=has)
            return f
        if k == abi.CMD_CLIENT_REQUEST:               # candidate.rs:189-192
            if len(self.queued) >= abi.CLIENT_QUEUE_CAP:
                raise Panic(abi.FAULT_ENGINE_QUEUE_OVERFLOW)
            self.queued.append((c["token"], c["client"]))
            return self
        return self


class Leader:
    role = abi.ROLE_LEADER

    def __init__(self, s: Shared):                    # candidate.rs:216-238
        self.s = s
        # progress.rs:15-23: every node (peers + self) starts in Probe at head 0
        self.progress: Dict[int, List] = {n: [0, False] for n in s.peers + [s.id]}  # [head, replicate?]
        self.heartbeat_time = s.now

    def term(self, t):                                # mod.rs:360-365 then leader.rs:33-35
        self.s.voted_for = None
        self.s.term = t
        raise Panic(abi.FAULT_LEADER_TERM_UNIMPLEMENTED)

    def heartbeat(self):                              # leader.rs:44-51
        s = self.s
        s.send(abi.ADDR_PEERS, 0, kind=abi.CMD_HEARTBEAT, term=s.term, block=s.chain.commit_id, node_id=s.id)

    def commit(self):                                 # leader.rs:87-99, progress.rs:48-60
        s = self.s
        heads = sorted((p[0] for p in self.progress.values()), reverse=True)
        q = heads[len(heads) // 2]
        if q > s.chain.commit_id:
            prev = s.chain.commit_id
            s.chain.commit(q)
            for bid in s.chain.keys_between(prev, q, True)[1:]:          # (prev..=new).skip(1)
                nxt, data = s.chain.blocks[bid]
                s.fsm.append(("apply", bid, nxt, data))

    def replicate(self):                              # leader.rs:124-174
        s = self.s
        for peer in s.peers:
            head, repl = self.progress[peer]
            ids = s.chain.from_skip_take(head, 1, abi.MAX_AE_BLOCKS if repl else 1)
            blocks = [(b, s.chain.blocks[b][0], s.chain.blocks[b][1]) for b in ids]
            s.send(abi.ADDR_PEER, peer, kind=abi.CMD_APPEND_ENTRIES, term=s.term, node_id=s.id, blocks=blocks)

    def apply(self, c):                               # leader.rs:248-266
        k, s = c["kind"], self.s
        if k == abi.CMD_TICK:                         # leader.rs:234-245
            if max(s.now - self.heartbeat_time, 0) > s.hb:
                self.heartbeat()
                self.heartbeat_time = s.now
            self.replicate()
        elif k == abi.CMD_HEARTBEAT_RESPONSE:         # leader.rs:222-231
            if not c["flag"] and c["block"] > 0:
                self.replicate()
        elif k == abi.CMD_APPEND_RESPONSE:            # leader.rs:211-219, progress.rs:42-46,76-94
            p = self.progress.get(c["node_id"])
            if p is None:
                raise Panic(abi.FAULT_PROGRESS_UNKNOWN_NODE)
            if p[0] < c["block"]:
                p[0], p[1] = c["block"], True
            else:
                p[1] = False
            self.commit()
        elif k == abi.CMD_APPEND_ENTRIES:             # leader.rs:200-208
            if c["term"] > s.term:
                self.term(c["term"])
        elif k == abi.CMD_CLIENT_REQUEST:             # leader.rs:177-197
            t = s.term
            bid = s.chain.append(c["token"])
            s.fsm.append(("notify", bid, c["client"], c["token"]))
            return self.apply(cmd(abi.CMD_APPEND_RESPONSE, node_id=s.id, term=t, flag=True, block=s.chain.head))
        return self


# --------------------------------------------------------------------------- harness
class PyNode:
    def __init__(self, cfg: abi.Config, g: int, node: int):
        R = cfg.n_replicas
        chain = Chain(cfg.chain_capacity, bool(cfg.flags & abi.F_SLED_COMMIT_KEY_STRICT))
        self.s = Shared(id=node, peers=[p for p in range(1, R + 1) if p != node], seed=cfg.seed,
                        group=cfg.group_offset + g, emin=cfg.election_min_ms, emax=cfg.election_max_ms,
                        hb=cfg.heartbeat_ms, chain=chain)
        self.s.set_election_timeout()                 # follower.rs:93-95
        self.h = Follower(self.s)
        self.fault = 0
        self.alive = not cfg.resident_mask or bool((cfg.resident_mask >> (node - 1)) & 1)
        self.prev_out: List[tuple] = []

    def apply(self, c, now):
        if not self.alive or self.fault:
            return
        self.s.now = now
        try:
            self.h = self.h.apply(c)
        except Panic as p:
            self.fault = p.code


class PyCluster:
    """Same surface as RaftApi (the parts the differential test uses)."""

    def __init__(self, cfg: abi.Config):
        # RaftConfig::validate, config.rs:70-75 (heartbeat / election timeout floors); empty gen_range, follower.rs:103-108
        if cfg.heartbeat_ms < 5 or cfg.election_min_ms < 5 or cfg.election_max_ms <= cfg.election_min_ms:
            from josefine_b200.raft import RaftError
            raise RaftError(abi.E_INVAL, "PyCluster", "configuration does not validate")
        self.cfg = cfg
        self.n_groups, self.n_replicas = cfg.n_groups, cfg.n_replicas
        self.nodes = [[PyNode(cfg, g, n) for n in range(1, cfg.n_replicas + 1)] for g in range(cfg.n_groups)]
        self.step_index = 0

    @classmethod
    def create(cls, g, r, **kw):
        return cls(abi.default_config(g, r, **kw))

    @staticmethod
    def _from_msg(m: abi.Msg):
        return cmd(m.kind, term=m.term, node_id=m.node_id, last_term=m.last_term, block=m.block, flag=bool(m.flag),
                   blocks=[(m.blocks[i].id, m.blocks[i].next, m.blocks[i].data) for i in range(m.n_blocks)],
                   token=m.token, client=(m.client_kind, m.client_id))

    def step(self, now_ms, flags=abi.STEP_DELIVER | abi.STEP_TICK, inject=(), proposals=None, n_synth=0, **_):
        from josefine_b200.raft import StepResult
        R = self.n_replicas
        buckets: Dict[tuple, list] = {}
        for m in inject:
            buckets.setdefault((m.group, m.to_id), []).append(self._from_msg(m))
        out_msgs, out_fsm = [], []
        for g, grp in enumerate(self.nodes):
            gg = self.cfg.group_offset + g
            for n in grp:
                me = n.s.id
                if flags & abi.STEP_DELIVER:
                    for sender in grp:
                        if sender.s.id == me:
                            continue
                        for (tk, tid, c) in sender.prev_out:
                            if tk == abi.ADDR_PEERS or (tk == abi.ADDR_PEER and tid == me):
                                n.apply(c, now_ms)
                for c in buckets.get((g, me), []):
                    n.apply(c, now_ms)
                if proposals is not None and proposals[g][0] == me:
                    n.apply(cmd(abi.CMD_CLIENT_REQUEST, token=proposals[g][1], client=(abi.ADDR_CLIENT, 0)), now_ms)
                for i in range(n_synth):
                    if n.h.role != abi.ROLE_LEADER:
                        break
                    tok = (((self.step_index * 8 + i + 1) << 32) | (gg & 0xFFFFFFFF)) & MASK64
                    n.apply(cmd(abi.CMD_CLIENT_REQUEST, token=tok, client=(abi.ADDR_CLIENT, 0)), now_ms)
                if flags & abi.STEP_TICK:
                    n.apply(cmd(abi.CMD_TICK), now_ms)
            for n in grp:
                for (tk, tid, c) in n.s.rpc:
                    out_msgs.append(self._to_msg(g, n.s.id, tk, tid, c))
                for f in n.s.fsm:
                    out_fsm.append(self._to_fsm(g, n.s.id, f))
                n.prev_out, n.s.rpc, n.s.fsm = n.s.rpc, [], []
        self.step_index += 1
        return StepResult(out_msgs, out_fsm)

    def apply(self, c: abi.Msg, now_ms: int = 0):
        return self.step(now_ms, flags=0, inject=[c])

    def handle(self, g, node):
        from josefine_b200.raft import ReplicaHandle
        return ReplicaHandle(self, g, node)

    @staticmethod
    def _to_msg(g, sender, tk, tid, c) -> abi.Msg:
        m = abi.Msg()
        m.group, m.from_kind, m.from_id, m.to_kind, m.to_id, m.kind = g, abi.ADDR_PEER, sender, tk, tid, c["kind"]
        k = c["kind"]
        if k == abi.CMD_VOTE_REQUEST:
            m.term, m.node_id, m.last_term, m.block = c["term"], c["node_id"], c["last_term"], c["block"]
        elif k == abi.CMD_VOTE_RESPONSE:
            m.term, m.node_id, m.flag = c["term"], c["node_id"], int(c["flag"])
        elif k == abi.CMD_APPEND_ENTRIES:
            m.term, m.node_id, m.n_blocks = c["term"], c["node_id"], len(c["blocks"])
            for i, (b, nx, d) in enumerate(c["blocks"]):
                m.blocks[i].id, m.blocks[i].next, m.blocks[i].data = b, nx, d
        elif k == abi.CMD_APPEND_RESPONSE:
            m.node_id, m.term, m.block, m.flag = c["node_id"], c["term"], c["block"], int(c["flag"])
        elif k == abi.CMD_HEARTBEAT:
            m.term, m.block, m.node_id = c["term"], c["block"], c["node_id"]
        elif k == abi.CMD_HEARTBEAT_RESPONSE:
            m.block, m.flag = c["block"], int(c["flag"])
        elif k == abi.CMD_CLIENT_REQUEST:
            m.token, m.client_kind, m.client_id = c["token"], c["client"][0], c["client"][1]
        elif k == abi.CMD_CLIENT_RESPONSE:
            m.token = c["token"]
        return m

    @staticmethod
    def _to_fsm(g, node, f) -> abi.FsmInstr:
        o = abi.FsmInstr()
        o.group, o.node = g, node
        if f[0] == "apply":
            o.kind = abi.FSM_APPLY
            o.block.id, o.block.next, o.block.data = f[1], f[2], f[3]
        else:
            o.kind = abi.FSM_NOTIFY
            o.block.id, o.block.next, o.block.data = f[1], 0, f[3]
            o.client_kind, o.client_id = f[2]
        return o

    def query(self, g, node) -> abi.ReplicaState:
        n = self.nodes[g][node - 1]
        s, h = n.s, n.h
        st = abi.ReplicaState()
        st.current_term, st.voted_for = s.term, s.voted_for or 0
        st.leader_id = (h.leader_id or 0) if h.role == abi.ROLE_FOLLOWER else 0
        st.election_time_ms, st.election_timeout_ms, st.rng_draws = s.election_time, s.election_timeout, s.draws
        st.head, st.commit, st.id_gen = s.chain.head, s.chain.commit_id, s.chain.id_gen
        st.max_key = s.chain.keys[-1] if s.chain.keys else 0
        if h.role == abi.ROLE_LEADER:
            st.heartbeat_time_ms = h.heartbeat_time
            for nid, (head, repl) in h.progress.items():
                st.progress_head[nid - 1] = head
                if repl:
                    st.progress_replicate |= 1 << (nid - 1)
        if h.role == abi.ROLE_CANDIDATE:
            for nid, v in h.votes.items():
                st.votes_seen |= 1 << (nid - 1)
                if v:
                    st.votes_granted |= 1 << (nid - 1)
        st.role, st.fault, st.alive = h.role, n.fault, int(n.alive)
        st.n_queued = len(getattr(h, "queued", []))
        return st

    def chain_read(self, g, node, first, n):
        ch = self.nodes[g][node - 1].s.chain
        return [(b, ch.blocks[b][0], ch.blocks[b][1]) if b in ch.blocks else None for b in range(first, first + n)]

    def compact(self):
        for grp in self.nodes:
            for n in grp:
                if n.alive and not n.fault:
                    n.s.chain.compact()

    def set_alive(self, g, node, alive):
        self.nodes[g][node - 1].alive = bool(alive)

    def fault_count(self):
        return sum(1 for grp in self.nodes for n in grp if n.fault)

    def leader_table(self):
        out = []
        for grp in self.nodes:
            e = (0, 0, 0)
            for n in grp:
                if n.alive and not n.fault and n.h.role == abi.ROLE_LEADER:
                    if e[1] == 0 or n.s.term >= e[0]:
                        e = (n.s.term, n.s.id, n.s.chain.commit_id)
            out.append(e)
        return out
