mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sym_fold.py tests/test_stream_path.py -m gpu -x -q --tb=short 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/step_breakdown.py 60 --only 3 2>&1 | tee gpurun_out/ab_view.txt
