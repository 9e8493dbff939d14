set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_gputest_final.log; cat gpurun_out/r02_gputest_final.log
NBS_CHUNK_TILES=4 timeout 300 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q -k "lockstep or column" 2>&1 | tail -3 > gpurun_out/r02_gputest_chunk4.log; cat gpurun_out/r02_gputest_chunk4.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_final.json 2> gpurun_out/r02_bench_c3_final.err; tail -3 gpurun_out/r02_bench_c3_final.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 1 --steps 10 --warmup 5 > gpurun_out/r02_bench_c5_n1_final.json 2> gpurun_out/r02_bench_c5_n1_final.err; tail -3 gpurun_out/r02_bench_c5_n1_final.err
timeout 120 python tools/time_kernels.py C3 20 2>&1 | tail -1 > gpurun_out/r02_time_final.log
timeout 120 python tools/time_kernels.py C3 20 forces 2>&1 | tail -1 >> gpurun_out/r02_time_final.log
NBS_LIST_SKIN=0 timeout 120 python tools/time_kernels.py C3 20 2>&1 | tail -1 >> gpurun_out/r02_time_final.log
timeout 200 python tools/time_kernels.py C4 10 2>&1 | tail -1 >> gpurun_out/r02_time_final.log
cat gpurun_out/r02_time_final.log
python tools/one_eval.py C3 6 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_c3_final.csv python tools/one_eval.py C3 6 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair -s 3 -c 1 -f -o gpurun_out/r02_k_pair_final python tools/one_eval.py C3 6 > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
