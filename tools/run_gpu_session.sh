python tools/sweep_group.py 2>&1 | tee gpurun_out/sweep_group.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
