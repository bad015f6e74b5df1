import pytest

from tests import limit_cases


def _oracle(g, r, **kw):
    from oracle.restated import RestatedCluster
    return RestatedCluster.create(g, r, **kw)


def _emu(g, r, **kw):
    from tests.emu.emu import EmuEngine
    return EmuEngine.create(g, r, **kw)


def _gpu(g, r, **kw):
    from josefine_b200 import RaftEngine
    return RaftEngine.create(g, r, **kw)


PAIRED = [limit_cases.case_chain_capacity_fault, limit_cases.case_client_queue_overflow]
SINGLE = [limit_cases.case_mailbox_overflow_faults_cleanly, limit_cases.case_fsm_fifo_overflow_never_touches_consensus,
          limit_cases.case_degenerate_calls]


@pytest.mark.parametrize("case", PAIRED, ids=lambda f: f.__name__)
def test_limits_device_code_vs_oracle(case):
    case(_oracle, _emu)


@pytest.mark.parametrize("case", SINGLE, ids=lambda f: f.__name__)
def test_limits_device_code(case):
    case(_emu)


@pytest.mark.gpu
@pytest.mark.parametrize("case", PAIRED, ids=lambda f: f.__name__)
def test_limits_gpu_vs_oracle(case):
    case(_oracle, _gpu)


@pytest.mark.gpu
@pytest.mark.parametrize("case", SINGLE, ids=lambda f: f.__name__)
def test_limits_gpu(case):
    case(_gpu)
