"""Scenarios behind tests/golden/digests.json (see tests/golden/make_golden.py)."""
from josefine_b200 import abi, Command

SCENARIOS = {
    # BASELINE config #2 shape: cold start, seeded timeouts, elections, then proposals
    "cold_1024x3_96ticks": dict(G=1024, R=3, seed=0, kind="cold", ticks=96, cfg=dict(chain_capacity=256, fsm_units=256)),
    # BASELINE config #3 shape (reduced G): pre-elected leaders, steady AppendEntries
    "steady_2048x5_64ticks": dict(G=2048, R=5, seed=1, kind="steady", ticks=64, cfg=dict(chain_capacity=160, fsm_units=192)),
    # BASELINE config #5 shape: 7 replicas, leader loss, compact
    "churn_512x7": dict(G=512, R=7, seed=2, kind="churn", ticks=40, cfg=dict(chain_capacity=256, fsm_units=192)),
    # deviation D6 reproduced: leaders panic on the sled "commit" key
    "strict_commit_key_256x3": dict(G=256, R=3, seed=3, kind="cold", ticks=64, flags=abi.F_SLED_COMMIT_KEY_STRICT,
                                    cfg=dict(chain_capacity=128, fsm_units=192)),
}


def bootstrap(G, R, node=1):
    q = 0 if R == 1 else R // 2 + 1
    inj = []
    for g in range(G):
        inj.append(Command.timeout(g, node))
        for v in [v for v in range(1, R + 1) if v != node][:max(q - 1, 0)]:
            inj.append(Command.vote_response(g, node, 1, v, True))
    return inj


def play(eng, sc):
    G, R = sc["G"], sc["R"]
    if sc["kind"] == "cold":
        eng.run(100, 100, sc["ticks"], 1)
    elif sc["kind"] == "steady":
        eng.step(0, flags=0, inject=bootstrap(G, R))
        eng.run(100, 100, sc["ticks"], 1)
    elif sc["kind"] == "churn":
        eng.step(0, flags=0, inject=bootstrap(G, R))
        now = 100
        for rnd in range(3):
            eng.run(now, 100, sc["ticks"], 1)
            now += 100 * sc["ticks"]
            eng.kill_leaders(rnd, 100)
            eng.compact()


def observe(eng):
    md, fd, nm, nf = eng.stream_digest()
    leaders = eng.leader_table()
    return {"state_digest": eng.state_digest(), "msg_digest": md, "fsm_digest": fd, "n_msgs": nm, "n_fsm": nf,
            "faulted": eng.fault_count(), "groups_with_leader": sum(1 for (_, l, _) in leaders if l),
            "max_commit": max(c for (_, _, c) in leaders)}
