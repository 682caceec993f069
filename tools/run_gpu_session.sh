python tools/sweep_opts.py "" "-DPT_BLOCKS_PER_SM=3" "-DPT_NOINLINE_HIT" "-DPT_NOINLINE_HIT -DPT_BLOCKS_PER_SM=3" 2>&1 | tee gpurun_out/sweep_blocks_b.txt
