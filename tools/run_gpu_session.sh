python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/sweep_opts.py "" 2>&1 | tee gpurun_out/sweep_default.txt
