set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "baseline_configs or random_systems or list_reuse_matches" 2>&1 | tail -3 > gpurun_out/r02_gputest_w.log; cat gpurun_out/r02_gputest_w.log
timeout 200 python tools/time_kernels.py C4 10 2>&1 | tail -1 | cut -c1-200
