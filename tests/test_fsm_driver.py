import pytest

from tests import fsm_cases


def _oracle(g, r, **kw):
    from oracle.restated import RestatedCluster
    return RestatedCluster.create(g, r, **kw)


def _emu(g, r, **kw):
    from tests.emu.emu import EmuEngine
    return EmuEngine.create(g, r, **kw)


def _gpu(g, r, **kw):
    from josefine_b200 import RaftEngine
    return RaftEngine.create(g, r, **kw)


@pytest.mark.parametrize("case", fsm_cases.ALL_FSM_CASES, ids=lambda f: f.__name__)
def test_fsm_path_on_oracle(case):
    case(_oracle)


@pytest.mark.parametrize("case", fsm_cases.ALL_FSM_CASES, ids=lambda f: f.__name__)
def test_fsm_path_on_device_code(case):
    case(_emu)


@pytest.mark.gpu
@pytest.mark.parametrize("case", fsm_cases.ALL_FSM_CASES, ids=lambda f: f.__name__)
def test_fsm_path_on_gpu(case):
    case(_gpu)
