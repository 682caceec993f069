# round 2, session 4: tests incl. the grid, fixed latency breakdown, grid vs scan
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/s4_pytest.log
python tools/section_clocks.py A 2 > gpurun_out/s4_section_clocks.log 2>&1
python tools/section_clocks.py A 0 >> gpurun_out/s4_section_clocks.log 2>&1
cat gpurun_out/s4_section_clocks.log
python tools/grid_bench.py > gpurun_out/s4_grid_bench.log 2>&1; cat gpurun_out/s4_grid_bench.log
python bench.py > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/s4_bench.err
