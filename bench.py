#!/usr/bin/env python
"""bench.py -- Raft-group ticks/sec of the batched Chained-Raft step path.

Metric (BASELINE.json): Raft-group ticks/sec @ 64Ki groups x 5 replicas.
One group-tick = all R replicas of one group each drain the peer mail of the
previous step, take the step's client proposal (leader) and apply one
Command::Tick (the reference defines no such unit -- its
Tick is a 100 ms wall-clock interval, src/raft/server.rs:25).

Workload (config.workload): BASELINE config #3 -- 65,536 groups x 5 replicas per
GPU, leaders pre-elected on node 1 by a synthetic vote trace, then steady state:
one client proposal per group per tick, AppendEntries / AppendResponse /
Heartbeat / HeartbeatResponse traffic between the co-resident replicas.

A bench "step" = TICKS_PER_STEP consecutive group-ticks of every group (one
jr_run).  L2 is flushed between timed steps (the per-GPU working set, ~0.1 GB,
is smaller than the 126 MB L2, so without the flush the state would be served
from L2 forever); inside a step the ticks run back to back as they do in
production.  Device time is taken with CUDA events on the engine's stream,
per step, flush excluded; max over ranks.

Arms:
  (default)          the CUDA engine.  `value` = device-resident throughput;
                     `e2e` = the same workload driven tick by tick through the
                     C-ABI jr_step with HOST buffers: proposals H2D from pinned
                     memory and the per-group {term, leader, commit} table D2H
                     every tick.
  --impl reference   the CPU comparator: the C++ RESTATEMENT of josefine's
                     src/raft (oracle/; josefine itself is Rust and cannot be
                     built here) on all host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from josefine_b200 import abi, Command  # noqa: E402

METRIC = "Raft-group ticks/sec @ 64Ki groups x 5 replicas"   # BASELINE.json `metric`
UNIT = "group-ticks/s"
GROUPS_PER_GPU = 65536
REPLICAS = 5
TICKS_PER_STEP = 64
DT_MS = 100
SEED = 1
L2_FLUSH_BYTES = 256 << 20
CHAIN_CAP_MAX = 6144        # block-table ids per replica (12 B x 327,680 replicas each = 3.9 MB per id, 24 GB total)


def bootstrap_inject(G, R, node=1, scattered=False):
    """Synthetic vote trace: Timeout on `node` (or on node g % R + 1 with `scattered`), plus
    quorum-1 granted VoteResponses."""
    q = 0 if R == 1 else R // 2 + 1
    inj = []
    for g in range(G):
        n = (g % R) + 1 if scattered else node
        inj.append(Command.timeout(g, n))
        for v in [v for v in range(1, R + 1) if v != n][:max(q - 1, 0)]:
            inj.append(Command.vote_response(g, n, 1, v, True))
    return inj


def algorithmic_bytes_per_group_tick(make, R, ticks=64):
    """Minimum bytes one steady-state group-tick must move with THIS layout
    (DESIGN.md "Algorithmic bytes"), measured on a small captured run:
      state   : follower 52 B (P0,P1,P2 + max_key), leader 116 B (+P3, 2 progress planes) -- read AND written
      mailbox : every 16 B unit written once and read once per addressee (+4 B count, written and read);
                AppendEntries of one sender that carry the same block run share the block units
      blocks  : 12 B written per block appended/extended, 12 B read per block shipped or applied, 4 B has() probe
      fsm     : 16 B per Instruction
    """
    G = 32
    eng = make(G, R, seed=SEED, flags=abi.F_CAPTURE_MESSAGES | abi.F_CAPTURE_FSM, chain_capacity=ticks * 2 + 64,
               fsm_units=16)
    eng.step(0, flags=0, inject=bootstrap_inject(G, R))
    for k in range(16):  # reach the steady regime
        eng.step((k + 1) * DT_MS, n_synth=1)
    tot = 0
    for k in range(16, 16 + ticks):
        res = eng.step((k + 1) * DT_MS, n_synth=1)
        b = 0
        b += G * ((R - 1) * 52 + 116) * 2                    # state read + written
        b += G * R * 4 * 2 + G * R * (R - 1) * 4             # mailbox counts: reset/written, read by each peer
        seen_vreq, seen_runs = set(), set()
        for m in res.messages:
            readers = (R - 1) if m.to_kind == abi.ADDR_PEERS else 1
            if m.kind == abi.CMD_VOTE_REQUEST:                # N-1 copies share one unit
                key = (m.group, m.from_id)
                if key in seen_vreq:
                    continue
                seen_vreq.add(key)
            b += 16 * (1 + readers)                           # header unit: written once, read per addressee
            if m.kind == abi.CMD_APPEND_ENTRIES and m.n_blocks:
                run = (m.group, m.from_id, tuple(m.blocks[i].id for i in range(m.n_blocks)))
                if run not in seen_runs:                      # identical block runs of one sender are emitted once
                    seen_runs.add(run)
                    b += m.n_blocks * (12 + 16)               # leader reads the table entries, writes the block units
                b += m.n_blocks * (16 + 4 + 12)               # each follower reads the units, probes has(next), writes its table
        for f in res.fsm:
            b += 16
            b += 12                                           # Notify: block written by append; Apply: block read
        tot += b
    return tot / (G * ticks)


class ClockSampler:
    """nvidia-smi clocks and throttle reasons (profiling recipe's clocks line).  Started before
    the warm-up so nvidia-smi's start-up latency is absorbed; `window()` then keeps the samples
    that arrived inside the timed region (or, if the region was shorter than the sampling
    period, the ones closest to it -- the GPU is under the same load during warm-up)."""

    def __init__(self, index):
        self.index = index
        self.rows = []      # (arrival time, fields)
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def window(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        time.sleep(0.1)
        self.proc.terminate()
        good = [(t, r) for t, r in self.rows if len(r) >= 6 and r[0].isdigit()]
        inside = [r for t, r in good if t0 <= t <= t1 + 0.03]
        where = "timed region"
        if not inside and good:   # region shorter than the sampling period: nearest samples under the same load
            good.sort(key=lambda tr: min(abs(tr[0] - t0), abs(tr[0] - t1)))
            inside = [r for _, r in good[:5]]
            where = "nearest to the timed region (region shorter than the 20 ms sampling period)"
        sm = [int(r[0]) for r in inside]
        mx = [int(r[1]) for r in inside if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside for i in range(4) if r[2 + i].startswith("Active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "sampled": where}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def effective_cores():
    """Host threads this process may really use: affinity mask and cgroup CPU quota, not just cpu_count()."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return n


def best_cpu_threads(R, cores):
    """The restatement allocates heavily; more threads than the allocator / cgroup can feed
    makes it SLOWER.  Probe a few counts on a small sample and keep the fastest, so the CPU
    arm is measured at its best."""
    best, best_rate = 1, 0.0
    cand = sorted({1, 4, 8, 16, 32, 64, cores} & set(range(1, cores + 1)))
    for th in cand:
        g = max(512, 32 * th)
        c = cpu_reference_run(g, R, 80, th)
        t0 = time.perf_counter()
        c.run(DT_MS * 17, DT_MS, 32, 1)
        rate = g * 32 / (time.perf_counter() - t0)
        if rate > best_rate:
            best, best_rate = th, rate
    return best


def cpu_reference_run(G, R, ticks, threads):
    """Times the C++ restatement on `threads` host cores: steady state after bootstrap."""
    from oracle.restated import RestatedCluster
    c = RestatedCluster.create(G, R, n_threads=threads, seed=SEED, chain_capacity=ticks * 8 + 4096)
    c.step(0, flags=0, inject=bootstrap_inject(G, R))
    c.run(DT_MS, DT_MS, 16, 1)
    return c


def run_reference(args):
    """--impl reference: the C++ restatement of src/raft on the host cores (bounded sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    R = REPLICAS
    cores = best_cpu_threads(R, effective_cores())
    G = max(4096, 64 * cores)
    span = 16 * TICKS_PER_STEP            # ticks one cluster lives before it is rebuilt (bounds the std::map chains)
    st = {"c": None, "left": 0, "now": 0}

    def step():
        if st["left"] < TICKS_PER_STEP:   # untimed rebuild, like the GPU arm's rebase
            st["c"] = cpu_reference_run(G, R, span + 64, cores)
            st["left"], st["now"] = span, DT_MS * 17
            return 0.0
        t0 = time.perf_counter()
        st["c"].run(st["now"], DT_MS, TICKS_PER_STEP, 1)
        st["now"] += DT_MS * TICKS_PER_STEP
        st["left"] -= TICKS_PER_STEP
        return time.perf_counter() - t0

    def timed_step():
        d = step()
        return d if d > 0.0 else step()

    for _ in range(args.warmup):
        timed_step()
    dt = sum(timed_step() for _ in range(args.steps))
    value = G * TICKS_PER_STEP * args.steps / dt
    sample = (f"{G} groups x {R} replicas x {TICKS_PER_STEP} ticks per step, {cores} threads (fastest of a probe over "
              f"1..{effective_cores()} usable host threads; groups partitioned statically)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": "BASELINE config #3 steady-state AppendEntries, bounded sample: " + sample,
                   "comparator": "C++ restatement of josefine src/raft (oracle/), NOT josefine itself (Rust, unbuildable here)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--groups", type=int, default=GROUPS_PER_GPU, help="groups per GPU")
    ap.add_argument("--replicas", type=int, default=REPLICAS)
    ap.add_argument("--scattered-leaders", action="store_true",
                    help="diagnostic: leader of group g on node g %% R + 1 instead of node 1 (not the BASELINE workload)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from josefine_b200 import RaftEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    G, R = args.groups, args.replicas
    S = TICKS_PER_STEP
    total_ticks = S * (max(args.warmup, 20) + args.steps) + 64
    stream = torch.cuda.Stream()          # explicit non-default stream: handle 0 would mean "engine's own"
    torch.cuda.set_stream(stream)

    def make(g, r, **kw):
        kw.setdefault("device", local)
        return RaftEngine.create(g, r, **kw)

    # ---------------- device-resident arm ----------------
    # The block table grows by one id per tick (12 B x R x G each), so the engine is sized for
    # at most CHAIN_CAP_MAX ids and REBASED (jr_engine_reset + the bootstrap trace + 16 warm ticks)
    # between timed steps when it runs low -- never inside a timed event pair.
    cap = min(total_ticks + 64, CHAIN_CAP_MAX)
    eng = make(G, R, seed=SEED, group_offset=rank * G, chain_capacity=cap,
               flags=abi.F_CAPTURE_FSM, fsm_units=2 * S + 8, mailbox_units=64)
    eng.set_stream(stream.cuda_stream)
    boot = bootstrap_inject(G, R, scattered=args.scattered_leaders)
    state = {"ticks_left": 0, "now": 0}

    def rebase(e):
        e.reset()
        e.step(0, flags=0, inject=boot)
        e.run(DT_MS, DT_MS, 16, 1)
        state["ticks_left"] = cap - 32 - 16
        state["now"] = DT_MS * 17

    rebase(eng)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    leaders = torch.empty(G * 16, dtype=torch.uint8, device="cuda")
    gathered = torch.empty(world * G * 16, dtype=torch.uint8, device="cuda") if world > 1 else None
    side = torch.cuda.Stream() if world > 1 else None
    announce_done = [None]

    def one_step():
        eng.run(state["now"], DT_MS, S, 1)
        state["now"] += DT_MS * S
        state["ticks_left"] -= S
        if world > 1:
            # the one cross-shard exchange: leader announce, once per step (every 64 ticks).  The table is packed on
            # the engine stream; the NCCL all-gather runs on a side stream and overlaps the next step's kernel.
            if announce_done[0] is not None:
                stream.wait_event(announce_done[0])            # previous announce must be over before `leaders` is rewritten
            eng.leader_table_device(leaders.data_ptr())
            packed = torch.cuda.Event()
            packed.record(stream)
            with torch.cuda.stream(side):
                side.wait_event(packed)
                dist.all_gather_into_tensor(gathered, leaders)
                ev = torch.cuda.Event()
                ev.record(side)
            announce_done[0] = ev

    def between_steps():
        if state["ticks_left"] < S:
            rebase(eng)
            if world > 1:      # the rebase is untimed host work of uneven length: re-align the ranks before timing again
                torch.cuda.synchronize()
                dist.barrier()
        flush.fill_(1)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 20)):   # >= 20 untimed steps: also covers nvidia-smi's start-up
        between_steps()
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_region0 = time.time()
    evs = []
    for i in range(args.steps):
        between_steps()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        one_step()
        if world > 1 and i == args.steps - 1:
            stream.wait_event(announce_done[0])                # the last announce is not hidden by a next step: time it
        b.record(stream)
        evs.append((a, b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.window(t_region0, time.time()) if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    faults = eng.fault_count()
    value = world * G * S * args.steps / (ms * 1e-3)
    launches = args.steps * (2 if world > 1 else 1)   # one fused step_kernel launch per step (+ leader_table_kernel)

    # ---------------- end-to-end arm (host buffers through the C ABI) ----------------
    # Every step: jr_run_tokens copies that step's 64-bit payload tokens [S][G] from PINNED host memory
    # (H2D, on the engine's copy stream), addresses each to the leader the previous step's
    # jr_leader_table_async announced, and runs the S ticks fused; jr_leader_table_async copies
    # the per-group {term, leader, commit} result back (D2H).  Two steps are in flight: the host
    # submits step k+1, then waits for and reads step k's result -- copy-in, kernels and copy-out
    # of neighbouring steps overlap.
    e2e = None
    if not args.no_e2e:
        del eng
        torch.cuda.empty_cache()
        e2 = make(G, R, seed=SEED, group_offset=rank * G, chain_capacity=cap, mailbox_units=64)
        e2.set_stream(stream.cuda_stream)
        rebase(e2)
        NB = 2
        prop = torch.zeros(NB, S, G, dtype=torch.int64).pin_memory()      # tokens[NB][S][G], one proposal per group-tick
        table = torch.zeros(NB, G, 2, dtype=torch.int64).pin_memory()     # jr_leader_entry[NB][G]
        prop[...] = ((torch.arange(NB * S, dtype=torch.int64).view(NB, S, 1) + 1) << 32) + torch.arange(G, dtype=torch.int64)
        lib = e2._lib
        pstride, tstride = S * G * 8, G * 16
        e2.leader_table()                                                 # first announce: where the tokens go
        checks = []

        excluded = [0.0]

        def submit(i):
            if state["ticks_left"] < S:     # block table full: rebase, OUTSIDE the timed region (clock paused)
                e2.sync()
                tp = time.perf_counter()
                rebase(e2)
                e2.leader_table()
                excluded[0] += time.perf_counter() - tp
            tn = state["now"]
            state["now"] += DT_MS * S
            state["ticks_left"] -= S
            st = lib.jr_run_tokens(e2._h, C.c_uint64(tn), C.c_uint32(DT_MS), C.c_uint32(S),
                                   C.cast(prop.data_ptr() + (i % NB) * pstride, C.POINTER(C.c_uint64)))   # H2D + route + fused kernel
            assert st == 0, st
            st = lib.jr_leader_table_async(e2._h, C.cast(table.data_ptr() + (i % NB) * tstride,
                                                         C.POINTER(abi.LeaderEntry)))   # result D2H
            assert st == 0, st

        def consume(i):
            st = lib.jr_leader_table_wait(e2._h)
            assert st == 0, st
            checks.append(int(table[i % NB, 0, 1].item() >> 32))        # read the step's result: commit of group 0

        def e2e_steps(n):
            submit(0)
            for i in range(n):
                if i + 1 < n:
                    submit(i + 1)
                consume(i)

        e2e_steps(max(args.warmup, 2))
        e2.sync()
        if world > 1:
            dist.barrier()
        excluded[0] = 0.0
        t0 = time.perf_counter()
        e2e_steps(args.steps)
        e2.sync()
        dt = time.perf_counter() - t0 - excluded[0]
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * G * S * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": S * G * 8,
               "d2h_bytes_per_step": G * 16, "ms_per_step": dt * 1e3 / args.steps,
               "api": "per step: jr_run_tokens(pinned u64 tokens[64][G], routed to the last announced leader) + jr_leader_table_async(pinned jr_leader_entry[G]) "
                      "+ jr_leader_table_wait; two steps in flight",
               "commit_last": checks[-1], "faulted_replicas": e2.fault_count(),
               "timing": "host wall clock around all timed steps incl. the final sync, max over ranks"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline + cpu baseline (rank 0) ----------------
    abytes = algorithmic_bytes_per_group_tick(make, R)
    peak, peak_src = measured_peak()
    avg_launch_s = (ms * 1e-3) / args.steps            # one step_kernel launch = S fused ticks of all G groups
    achieved = abytes * G * S / avg_launch_s / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "step_kernel_latest.json")
    if os.path.exists(prof):   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this
        pj = json.load(open(prof))           # command, per launch (one launch = TICKS_PER_STEP ticks)
        if pj.get("ticks_per_launch") == S and pj.get("groups") == G:
            traffic = pj["dram_bytes_per_launch"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "step_kernel<5>", "algorithmic_bytes_per_group_tick": abytes,
                "avg_launch_us": avg_launch_s * 1e6, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": abytes * G * S,
                "note": f"one launch = {S} fused group-ticks of all {G} groups; replica state stays in registers and the "
                        "mailboxes in shared memory across those ticks, so DRAM traffic is far below the algorithmic "
                        "bytes (which count every tick's state + mailbox movement); the engine runs a block's ticks as "
                        "two ticket-ordered tasks when that fills the last wave of CTAs (DESIGN.md, Split launches); "
                        "see profiles/"}
    cpu = None
    if not args.no_cpu:
        cores = best_cpu_threads(R, effective_cores())
        cg, ct = max(4096, 64 * cores), 256
        c = cpu_reference_run(cg, R, ct + 32, cores)
        t0 = time.perf_counter()
        c.run(DT_MS * 17, DT_MS, ct, 1)
        cdt = time.perf_counter() - t0
        cpu = {"value": cg * ct / cdt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"C++ restatement of josefine src/raft: {cg} groups x {R} replicas x {ct} ticks, {cores} threads "
                         f"(fastest of a probe over 1..{effective_cores()} usable host threads)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"BASELINE config #3: {G} groups x {R} replicas per GPU, pre-elected leaders, steady-state "
                               f"AppendEntries, 1 proposal/group/tick",
                   "groups_per_gpu": G, "replicas": R, "ticks_per_step": S, "tick_ms": DT_MS, "seed": SEED,
                   "l2": f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write); ticks inside a step run back to back",
                   "parallelism": f"groups sharded over {world} GPU(s); leader-announce all_gather once per step" if world > 1
                   else "single GPU", "faulted_replicas": faults,
                   "untimed_steps_before_timing": max(args.warmup, 20)},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
