"""Race check of the split-launch hand-over ("Split launches") with ThreadSanitizer.

tests/emu/tsan_split.cpp runs the engine's device code compiled for the host with two CTAs at a time, so a
block's part k runs while its part k-1 is still going and has to wait for the release/acquire flag.  The
reference gets this class of safety from Rust ownership (`apply(self)`, src/raft/mod.rs:488); here it is
checked: the real protocol must be race-free AND the detector must catch the protocol with the ordering
removed (negative control), otherwise a clean run would prove nothing.
"""
import os
import shutil
import subprocess

import pytest

EMU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
CXX = shutil.which("g++")


def build(target):
    r = subprocess.run(["make", "-C", EMU, f"CXX={CXX}", target], capture_output=True, text=True)
    if r.returncode != 0:
        if "tsan" in (r.stderr + r.stdout).lower():
            pytest.skip("no ThreadSanitizer runtime for this compiler")
        raise AssertionError(r.stderr[-2000:])
    return os.path.join(EMU, target)


def run(exe):
    env = dict(os.environ, JR_EMU_CTAS="2", TSAN_OPTIONS="halt_on_error=0 exitcode=66")
    env.pop("JR_PARTS", None)
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=600)
    return r.returncode, r.stdout, (r.stdout + r.stderr).count("WARNING: ThreadSanitizer")


@pytest.mark.skipif(CXX is None, reason="needs g++")
def test_split_handover_is_race_free_and_the_detector_has_teeth():
    rc, out, races = run(build("tsan_split"))
    assert "split == whole" in out and races == 0 and rc == 0, (rc, races, out[-500:])
    rc, out, races = run(build("tsan_split_broken"))       # flag accesses made relaxed: must be reported
    assert races > 0 and rc == 66, (rc, races, out[-500:])
