timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "several_output_ranges or tiled_tcgen05" > gpurun_out/r2e_pytest_tdot.log 2>&1; tail -3 gpurun_out/r2e_pytest_tdot.log
timeout 900 python tools/bench_sym22.py --shares 4 --kch 32 --reps 2 --overlap 0,1,0,1 > gpurun_out/r2e_sym22_overlap.log 2>&1; grep -E "overlap|share" gpurun_out/r2e_sym22_overlap.log | cut -c1-110
