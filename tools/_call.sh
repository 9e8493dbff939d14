# scratch script for `gpurun -- 'bash tools/_call.sh'`: what a round-end check runs
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gputest.log; cat gpurun_out/gputest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -3 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
