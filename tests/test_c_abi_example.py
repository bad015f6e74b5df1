"""The C ABI from plain C: examples/multi_node.c (BASELINE config #1's shape) is compiled with
gcc against include/josefine_raft_abi.h and the in-tree CUDA library, then run."""
import os
import subprocess

import pytest

from josefine_b200.raft import ENGINE_LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, name="multi_node", lib=ENGINE_LIB_PATH):
    exe = tmp_path / name
    subprocess.check_call(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", name + ".c"), lib,
                           "-Wl,-rpath," + os.path.dirname(lib), "-o", str(exe)])
    return exe


def test_example_compiles_against_the_header(tmp_path):
    _build(tmp_path)      # CPU: the header is valid C and every symbol the example uses links
    _build(tmp_path, "batched_quantum")


def test_batched_quantum_example_runs_on_the_device_code(tmp_path):
    """examples/batched_quantum.c -- the hot path's loop from plain C -- linked against the host build of
    the device code (tests/emu): same C ABI, same kernels, no GPU."""
    from tests.emu.emu import LIB_PATH as EMU_LIB_PATH
    exe = _build(tmp_path, "batched_quantum", EMU_LIB_PATH)
    out = subprocess.run([str(exe), "96", "4"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("OK") and "96 of 96 groups elected a leader" in out.stdout


@pytest.mark.gpu
def test_batched_quantum_example_runs(tmp_path):
    exe = _build(tmp_path, "batched_quantum")
    out = subprocess.run([str(exe), "4096", "8"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("OK") and "4096 of 4096 groups elected a leader" in out.stdout


@pytest.mark.gpu
def test_multi_node_example_runs(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "is leader of term" in out.stdout and out.stdout.strip().endswith("OK")
    # leader applies all 3; followers apply the half-open range prev..commit (follower.rs:204): 2 each
    assert out.stdout.count("applies block") == 7


@pytest.mark.gpu
def test_api_misuse_returns_status_codes():
    """jr_status is for API misuse only (config.rs:60-84 style validation), never consensus outcomes."""
    import ctypes as C
    from josefine_b200 import abi, Command, RaftEngine, RaftError
    with pytest.raises(RaftError) as e:
        RaftEngine.create(4, 9)                       # R > JR_MAX_REPLICAS
    assert e.value.status == abi.E_INVAL
    with pytest.raises(RaftError):
        RaftEngine.create(4, 3, election_min_ms=1000, election_max_ms=1000)   # empty gen_range
    eng = RaftEngine.create(4, 3)
    with pytest.raises(RaftError) as e:
        eng.apply(Command.tick(0, 4))                 # node outside the group
    assert e.value.status == abi.E_UNKNOWN_NODE
    with pytest.raises(RaftError) as e:
        eng.apply(Command.tick(9, 1))                 # group out of range
    assert e.value.status == abi.E_INVAL
    with pytest.raises(RaftError) as e:
        eng.step(100, proposals=[(7, 1)] * 4)         # proposal to an unknown node
    assert e.value.status == abi.E_UNKNOWN_NODE
    a = abi.StepArgs()
    buf = (abi.Msg * 4)()
    a.out_msgs, a.cap_msgs = buf, 4                   # capture buffer without the capture flag
    assert eng._lib.jr_step(eng._h, C.byref(a)) == abi.E_INVAL
    cap = RaftEngine.create(2, 3, flags=abi.F_CAPTURE_MESSAGES)
    cap.apply(Command.timeout(0, 1))
    a = abi.StepArgs()
    a.flags = abi.STEP_DELIVER | abi.STEP_TICK
    a.out_msgs, a.cap_msgs = buf, 1                   # too small: count is still reported
    st = cap._lib.jr_step(cap._h, C.byref(a))
    assert st == abi.E_CAPACITY and a.n_msgs > 1
