"""Instruction-stream output path, truncation (D7), restart, bulk introspection, checkpoint: the device code on the
CPU (tests/emu) and on the GPU, against the C++ oracle.  Cases live in tests/stream_cases.py."""
import pytest

from tests import stream_cases


def _oracle(g, r, **kw):
    from oracle.restated import RestatedCluster
    return RestatedCluster.create(g, r, **kw)


def _emu(g, r, **kw):
    from tests.emu.emu import EmuEngine
    return EmuEngine.create(g, r, **kw)


def _gpu(g, r, **kw):
    from josefine_b200 import RaftEngine
    return RaftEngine.create(g, r, **kw)


@pytest.mark.parametrize("case", stream_cases.PAIRED, ids=lambda f: f.__name__)
def test_paired_on_device_code(case):
    case(_emu, _oracle)


@pytest.mark.parametrize("case", stream_cases.SINGLE, ids=lambda f: f.__name__)
def test_single_on_device_code(case):
    case(_emu)


def test_single_cases_hold_on_the_oracle_too():
    stream_cases.case_bulk_introspection(_oracle)     # (query_many / chain_read_many fall back to loops there)


def test_expand_known_answers_host_code():
    """jr_fsm_expand is pure host code in the engine library: it runs without a GPU."""
    import ctypes as C
    from josefine_b200.raft import ENGINE_LIB_PATH, _bind
    lib = C.CDLL(ENGINE_LIB_PATH)
    _bind(lib, "jr_")
    stream_cases.case_expand_known_answers(lib)


def test_expand_known_answers_emulation_build():
    from tests.emu import emu
    stream_cases.case_expand_known_answers(emu.load())


@pytest.mark.gpu
@pytest.mark.parametrize("case", stream_cases.PAIRED, ids=lambda f: f.__name__)
def test_paired_on_gpu(case):
    case(_gpu, _oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("case", stream_cases.SINGLE, ids=lambda f: f.__name__)
def test_single_on_gpu(case):
    case(_gpu)


@pytest.mark.gpu
def test_truncation_soak_50k_ticks_on_gpu():
    """>= 50,000 ticks in a 4,096-id window, no reset, no fault, digests equal to the oracle."""
    stream_cases.case_truncation_soak(_gpu, _oracle, G=64, R=5, cap=4096, rounds=50, ticks=1000, kill_at=30, compare_fsm=False)


@pytest.mark.gpu
def test_full_size_stream_equals_oracle_65536x5_256_ticks():
    """jr_run_tokens + the batched drain at 65,536 x 5 for 256 ticks.  Every launch's records are
    expanded by jr_fsm_expand and compared BYTE FOR BYTE with the oracle's jro_drain_fsm (~25M Instructions per launch);
    BatchedDriver fed from the records answers exactly what BatchedDriver fed from the oracle's Instructions answers."""
    import ctypes as C

    import numpy as np

    from josefine_b200 import BatchedDriver, abi
    from josefine_b200.raft import load_engine_library
    from oracle.restated import load as load_oracle
    from tests.stream_cases import _bootstrap, CAP
    G, R, S, LAUNCHES = 65536, 5, 64, 4
    eng = _gpu(G, R, seed=1, flags=CAP, fsm_units=16, chain_capacity=512)
    ora = RestatedOracle = None
    from oracle.restated import RestatedCluster
    ora = RestatedCluster.create(G, R, n_threads=min(16, __import__("os").cpu_count() or 1), seed=1, flags=CAP, chain_capacity=512)
    for api in (eng, ora):
        _bootstrap(api, G, R)
        api.run(100, 100, 16, 0)
        api.leader_table()
        api.discard_fsm(strict=False)
    lib, olib = load_engine_library(), load_oracle()
    cap = G * (R + 1) * S + 4 * G * R
    out_e, out_o = (abi.FsmInstr * cap)(), (abi.FsmInstr * cap)()

    class CountingFsm:
        def __init__(self):
            self.applied = 0

        def transition(self, data):
            self.applied += 1
            return data

    sub = 48                                                        # groups whose client round trip is checked through BatchedDriver
    drv_e, drv_o = (BatchedDriver(lambda g, n: CountingFsm(), {}) for _ in range(2))
    now, tick, total = 1700, 0, 0
    for launch in range(LAUNCHES):
        toks = ((np.arange(tick + 1, tick + S + 1, dtype=np.uint64)[:, None] << np.uint64(32)) +
                np.arange(1, G + 1, dtype=np.uint64)[None, :]).copy()
        ptr = toks.ctypes.data_as(C.POINTER(C.c_uint64))
        assert lib.jr_run_tokens(eng._h, C.c_uint64(now), C.c_uint32(100), C.c_uint32(S), ptr) == 0
        assert lib.jr_engine_sync(eng._h) == 0
        assert olib.jro_run_tokens(ora._h, C.c_uint64(now), C.c_uint32(100), C.c_uint32(S), ptr) == 0
        for api in (eng, ora):
            api.truncate(8)
        now += 100 * S
        tick += S
        assert lib.jr_fsm_records_async(eng._h) == 0
        recs, batch = C.POINTER(abi.FsmRecord)(), abi.FsmBatch()
        assert lib.jr_fsm_records_wait(eng._h, C.byref(recs), C.byref(batch)) == 0
        n_e, n_o = C.c_size_t(0), C.c_size_t(0)
        assert lib.jr_fsm_expand(recs, C.c_size_t(batch.n_records), G, R, out_e, C.c_size_t(cap), C.byref(n_e)) == 0
        assert olib.jro_drain_fsm(ora._h, out_o, C.c_size_t(cap), C.byref(n_o)) == 0
        assert n_e.value == n_o.value == batch.n_instructions and n_e.value > G * R * S * 0.9
        a = np.frombuffer(out_e, dtype=np.uint8, count=n_e.value * C.sizeof(abi.FsmInstr))
        b = np.frombuffer(out_o, dtype=np.uint8, count=n_o.value * C.sizeof(abi.FsmInstr))
        assert np.array_equal(a, b), f"launch {launch}: expanded stream differs from the oracle's"
        assert batch.n_records < (12 if launch == 0 else 8) * G      # compact: O(1) records per replica per launch (the first also holds the start-up)
        total += n_e.value
        # client path on a subset: records -> BatchedDriver vs oracle Instructions -> BatchedDriver
        sub_recs = [abi.FsmRecord.from_buffer_copy(recs[i]) for i in range(batch.n_records) if recs[i].group < sub]
        first = next(i for i in range(n_o.value) if out_o[i].group >= sub)
        ra = drv_e.feed_records(lib, sub_recs, G, R)
        rb = drv_o.feed([out_o[i] for i in range(first)])
        key = lambda r: (r.group, r.node, r.to, r.request)   # noqa: E731
        assert [key(r) for r in ra] == [key(r) for r in rb] and len(ra) >= sub * (S - 4)
    assert total > G * (R + 1) * S * LAUNCHES * 0.95
    assert eng.state_digest() == ora.state_digest() and eng.stream_digest() == ora.stream_digest()


def test_fold_mt_equals_fold_on_every_thread_count():
    """jr_fsm_fold_mt (a pool of spinning-then-sleeping workers, groups partitioned over threads) == jr_fsm_fold, call after
    call, with the thread count changing between calls and pauses long enough for the workers to fall asleep."""
    import ctypes as C
    import random
    import time
    from josefine_b200 import abi
    from josefine_b200.raft import load_engine_library
    lib = load_engine_library()
    G, R = 6000, 5
    recs = (abi.FsmRecord * (G * 4))()
    n = 0
    for node in range(1, R + 1):
        for g in range(G):
            if node == 1:
                for kind, cnt in ((0, 64), (1, 64)):
                    r = recs[n]
                    r.group, r.hdr, r.id0, r.addr = g, kind | (node << 2) | (cnt << 8), 100 + g % 7, 0
                    n += 1
            elif node == 2:                                   # one masked APPLY record for all four followers
                r = recs[n]
                r.group, r.hdr, r.id0, r.addr = g, 0 | (node << 2) | (60 << 8), 90 + g % 5, 0b11110
                n += 1

    def run(threads):
        applied, tot = (C.c_uint32 * (G * R))(), (C.c_uint64 * 3)()
        assert lib.jr_fsm_fold_mt(C.cast(recs, C.c_void_p), C.c_size_t(n), G, R, applied, tot, threads) == 0
        return list(tot), list(applied)

    ref = run(1)
    rng = random.Random(5)
    for rep in range(60):
        assert run(rng.choice([2, 3, 4, 8])) == ref
        if rep % 20 == 19:
            time.sleep(0.01)                                  # > the workers' spin window: the next call has to wake them
