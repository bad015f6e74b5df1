"""Shared parity harness: drive two RaftApi implementations (oracle vs engine)
with the same calls and compare everything observable, bit for bit."""
from __future__ import annotations

import random
from typing import Callable, List, Optional

from josefine_b200 import abi, Command, fsm_tuple, msg_tuple

FULL = abi.F_CAPTURE_MESSAGES | abi.F_CAPTURE_FSM | abi.F_STREAM_DIGEST


def state_tuple(st: abi.ReplicaState) -> tuple:
    d = st.as_dict()
    return tuple((k, tuple(v) if isinstance(v, list) else v) for k, v in d.items())


def compare_states(a, b, groups=None, chain_ids=0, where=""):
    G, R = a.n_groups, a.n_replicas
    for g in (groups if groups is not None else range(G)):
        for n in range(1, R + 1):
            sa, sb = state_tuple(a.query(g, n)), state_tuple(b.query(g, n))
            assert sa == sb, f"{where} state differs g={g} node={n}:\n  A={dict(sa)}\n  B={dict(sb)}"
            if chain_ids:
                ca, cb = a.chain_read(g, n, 0, chain_ids), b.chain_read(g, n, 0, chain_ids)
                assert ca == cb, f"{where} chain differs g={g} node={n}:\n  A={ca}\n  B={cb}"


def compare_step(ra, rb, where=""):
    ma, mb = [msg_tuple(m) for m in ra.messages], [msg_tuple(m) for m in rb.messages]
    if ma != mb:
        for i, (x, y) in enumerate(zip(ma, mb)):
            assert x == y, f"{where} message #{i} differs:\n  A={x}\n  B={y}"
        assert len(ma) == len(mb), f"{where} message count {len(ma)} vs {len(mb)}; extra: {(ma + mb)[min(len(ma), len(mb))]}"
    fa, fb = [fsm_tuple(f) for f in ra.fsm], [fsm_tuple(f) for f in rb.fsm]
    if fa != fb:
        for i, (x, y) in enumerate(zip(fa, fb)):
            assert x == y, f"{where} fsm #{i} differs:\n  A={x}\n  B={y}"
        assert len(fa) == len(fb), f"{where} fsm count {len(fa)} vs {len(fb)}"


def compare_digests(a, b, where=""):
    assert a.state_digest() == b.state_digest(), f"{where} state digest differs"
    assert a.stream_digest() == b.stream_digest(), f"{where} stream digest differs: {a.stream_digest()} vs {b.stream_digest()}"
    assert a.fault_count() == b.fault_count(), f"{where} fault count differs"


class Pair:
    """Applies every call to both implementations and checks equality as it goes."""

    def __init__(self, make_a: Callable, make_b: Callable, G: int, R: int, check_states_every: int = 1,
                 chain_ids: int = 48, **cfg):
        cfg.setdefault("flags", FULL)
        self.a, self.b = make_a(G, R, **cfg), make_b(G, R, **cfg)
        self.G, self.R = G, R
        self.k = 0
        self.every = check_states_every
        self.chain_ids = chain_ids

    def step(self, now, **kw):
        ra, rb = self.a.step(now, **kw), self.b.step(now, **kw)
        where = f"[step {self.k} now={now}]"
        compare_step(ra, rb, where)
        self.k += 1
        if self.every and self.k % self.every == 0:
            compare_states(self.a, self.b, chain_ids=self.chain_ids, where=where)
        return ra

    def run(self, now0, dt, n, n_synth=0):
        self.a.run(now0, dt, n, n_synth)
        self.b.run(now0, dt, n, n_synth)
        fa, fb = [fsm_tuple(f) for f in self.a.drain_fsm()], [fsm_tuple(f) for f in self.b.drain_fsm()]
        assert fa == fb, f"[run {n} steps] fsm streams differ"
        self.k += n

    def both(self, name, *args):
        ra, rb = getattr(self.a, name)(*args), getattr(self.b, name)(*args)
        assert ra == rb, f"{name}{args}: {ra} vs {rb}"
        return ra

    def finish(self):
        compare_states(self.a, self.b, chain_ids=self.chain_ids, where="[final]")
        compare_digests(self.a, self.b, "[final]")
        assert self.a.leader_table() == self.b.leader_table()


# ---- scenarios (used for both emu-vs-oracle on CPU and engine-vs-oracle on GPU) -----------------

def bootstrap_leaders(p: Pair, now=0, node=1):
    """Synthetic vote trace that makes `node` the leader of every group: Timeout on
    `node`, then enough injected granted VoteResponses for a quorum (needed for R >= 5,
    where sender-major delivery of the duplicated VoteRequests never elects)."""
    inj = []
    quorum = 0 if p.R == 1 else p.R // 2 + 1
    voters = [v for v in range(1, p.R + 1) if v != node]
    for g in range(p.G):
        inj.append(Command.timeout(g, node))
        for v in voters[:max(quorum - 1, 0)]:
            inj.append(Command.vote_response(g, node, term=1, from_=v, granted=True))
    return p.step(now, flags=0, inject=inj)


def scenario_cold_start(p: Pair, steps=40, dt=100, proposals_after=15, n_synth=1):
    """Config #2 shape: cold start -> seeded timeouts -> elections -> proposals."""
    for k in range(steps):
        p.step((k + 1) * dt, n_synth=n_synth if k >= proposals_after else 0)
    p.finish()


def scenario_steady(p: Pair, steps=24, dt=100, n_synth=1):
    """Config #3 shape: pre-elected leaders (node 1), steady AppendEntries."""
    bootstrap_leaders(p, now=0)
    for k in range(steps):
        p.step((k + 1) * dt, n_synth=n_synth)
    p.finish()


def scenario_random_inject(p: Pair, seed=0, steps=60, dt=100, per_step=3, max_node=None):
    """Randomised differential test: arbitrary (mostly well-formed) commands injected
    into random replicas on top of normal delivery + ticks."""
    rng = random.Random(seed)
    R, G = p.R, p.G
    max_node = max_node or R
    tok = 1000
    for k in range(steps):
        inj = []
        for _ in range(rng.randint(0, per_step)):
            g, to = rng.randrange(G), rng.randint(1, R)
            kind = rng.choice(["vreq", "vresp", "ae", "aresp", "hb", "hbresp", "timeout", "creq", "cresp", "noop", "tick"])
            term = rng.randint(0, 6)
            node = rng.randint(1, max_node)
            blk = rng.randint(0, 12)
            if kind == "vreq":
                inj.append(Command.vote_request(g, to, term, node, rng.randint(0, 6), blk))
            elif kind == "vresp":
                inj.append(Command.vote_response(g, to, term, node, rng.random() < 0.6))
            elif kind == "ae":
                nb = rng.randint(0, 3)
                base = rng.randint(0, 10)
                blocks = [(base + i + 1, base + i if rng.random() < 0.85 else rng.randint(0, 12), tok + i) for i in range(nb)]
                tok += nb
                inj.append(Command.append_entries(g, to, term, node, blocks))
            elif kind == "aresp":
                inj.append(Command.append_response(g, to, node, term, blk))
            elif kind == "hb":
                inj.append(Command.heartbeat(g, to, term, blk, node))
            elif kind == "hbresp":
                inj.append(Command.heartbeat_response(g, to, blk, rng.random() < 0.5))
            elif kind == "timeout":
                inj.append(Command.timeout(g, to))
            elif kind == "creq":
                tok += 1
                inj.append(Command.client_request(g, to, tok))
            elif kind == "cresp":
                inj.append(Command.client_response(g, to, rng.randint(1, 99)))
            elif kind == "noop":
                inj.append(Command.noop(g, to))
            else:
                inj.append(Command.tick(g, to))
        props = None
        if rng.random() < 0.3:
            tok += G
            props = [(rng.randint(0, R), tok + g) for g in range(G)]
        p.step((k + 1) * dt, inject=inj, proposals=props, n_synth=rng.choice([0, 0, 1, 2]))
        if rng.random() < 0.05:
            p.both("compact")
        if rng.random() < 0.03:
            g, n = rng.randrange(G), rng.randint(1, R)
            p.both("set_alive", g, n, False)
    p.finish()


def scenario_leader_routed_tokens(make_a, make_b, G=12, R=3, seed=6):
    """jr_run_tokens on both implementations: routes before any announce (dropped), fresh routes, routes gone
    stale after a leader is silenced (token lands on a dead node / a follower), re-announce after failover."""
    cfg = dict(seed=seed, flags=FULL, fsm_units=256, fsm_host_records=G * R * 64)   # holes in the token grid: short runs
    a, b = make_a(G, R, **cfg), make_b(G, R, **cfg)
    now = [0]

    def both(name, *args):
        ra, rb = getattr(a, name)(*args), getattr(b, name)(*args)
        assert ra == rb, name
        return ra

    def tokens(n, salt):
        return [[0 if (g + k) % 5 == 0 else (salt << 32) | (k << 16) | (g + 1) for g in range(G)] for k in range(n)]

    def run_tokens(n, salt):
        t = tokens(n, salt)
        a.run_tokens(now[0] + 100, 100, t)
        b.run_tokens(now[0] + 100, 100, t)
        now[0] += 100 * n
        fa, fb = [fsm_tuple(f) for f in a.drain_fsm()], [fsm_tuple(f) for f in b.drain_fsm()]
        assert fa == fb
        compare_states(a, b, chain_ids=64, where=f"[tokens salt={salt}]")
        return fa

    both("run", 100, 100, 20, 0)                      # elections
    now[0] = 2000
    assert run_tokens(4, 1) == []                     # no announce yet: every token is dropped
    table = both("leader_table")
    assert sum(1 for (_, lid, _) in table if lid) * 10 >= 8 * G      # (a split vote can leave a group leaderless)
    fsm = run_tokens(8, 2)
    assert any(f[2] == abi.FSM_NOTIFY for f in fsm) and any(f[2] == abi.FSM_APPLY for f in fsm)
    commit0 = max(c for (_, _, c) in both("leader_table"))
    assert commit0 > 0
    killed = both("kill_leaders", 77, 500)            # routes to those leaders are now stale
    assert 0 < killed < G
    run_tokens(6, 3)
    both("leader_table")                              # groups without a live leader announce 0 -> dropped again
    run_tokens(6, 4)
    # equivalence with explicit proposals (the ABI's definition), on implementation a
    c = make_a(G, R, **cfg)
    c.run(100, 100, 20, 0)
    route = [lid for (_, lid, _) in c.leader_table()]
    t = tokens(8, 2)
    c.run_proposals(2500, 100, [[(route[g] if tok else 0, tok) for g, tok in enumerate(tick)] for tick in t])
    d = make_a(G, R, **cfg)
    d.run(100, 100, 20, 0)
    d.leader_table()
    d.run_tokens(2500, 100, t)
    compare_states(c, d, chain_ids=64, where="[tokens == proposals]")
    compare_digests(c, d, "[tokens == proposals]")
    compare_digests(a, b, "[final]")
