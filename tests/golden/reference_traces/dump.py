"""Serialises tests/reference_traces.py (hand-derived, Rust-cited expected rows) to JSON, one file per trace.
Pure serialisation: no restatement, oracle or engine is imported or run."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))   # tests/
import importlib.util  # noqa: E402
spec = importlib.util.spec_from_file_location("reference_traces", os.path.join(os.path.dirname(os.path.dirname(HERE)), "reference_traces.py"))
rt = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rt)

for t in rt.ALL_TRACES:
    with open(os.path.join(HERE, t["name"] + ".json"), "w") as f:
        json.dump(t, f, indent=1)
        f.write("\n")
print(len(rt.ALL_TRACES), "traces written")
