set -x
timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_gputest_peer.log; cat gpurun_out/r02_gputest_peer.log
