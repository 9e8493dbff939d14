set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reference_fixture or baseline_configs or deterministic or list_reuse or nonperiodic" 2>&1 | tail -15 > gpurun_out/r02_gputest_c.log; cat gpurun_out/r02_gputest_c.log
rm -f gpurun_out/r02_time_c.log
for mc in 2 3; do
NBS_LIST_SKIN=0 NBS_PAIR_MINCTAS=$mc timeout 120 python tools/time_kernels.py C3 20 2>&1 | tail -1 | sed "s/^/skin0 mc=$mc /" >> gpurun_out/r02_time_c.log
NBS_LIST_SKIN=0 NBS_PAIR_MINCTAS=$mc timeout 120 python tools/time_kernels.py C3 20 forces 2>&1 | tail -1 | sed "s/^/skin0 mc=$mc /" >> gpurun_out/r02_time_c.log
done
NBS_LIST_SKIN=0 NBS_CHUNK_TILES=2 timeout 120 python tools/time_kernels.py C3 20 2>&1 | tail -1 | sed "s/^/skin0 ct=2 /" >> gpurun_out/r02_time_c.log
timeout 200 python tools/time_moving.py C3 100 0.0009 0.07 2>&1 | tail -2 >> gpurun_out/r02_time_c.log
timeout 200 python tools/time_kernels.py C5 5 2>&1 | tail -1 >> gpurun_out/r02_time_c.log
cat gpurun_out/r02_time_c.log
