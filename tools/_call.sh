set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_gputest_i.log; cat gpurun_out/r02_gputest_i.log
timeout 120 python tools/time_kernels.py C3 20 2>&1 | tail -1 > gpurun_out/r02_time_i.log
timeout 300 python tools/time_kernels.py C5 5 2>&1 | tail -1 >> gpurun_out/r02_time_i.log
cat gpurun_out/r02_time_i.log
timeout 300 python tools/parity_detail.py C3 > gpurun_out/r02_parity_detail_c3.log 2>&1; tail -9 gpurun_out/r02_parity_detail_c3.log
