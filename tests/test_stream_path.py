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
