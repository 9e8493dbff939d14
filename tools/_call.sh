set -x
for b in 0 4 8 16; do NBS_PME_BATCH=$b timeout 120 python tools/time_kernels.py C3 20 2>&1 | tail -1 >> gpurun_out/r02_time_p.log; done
for b in 0 8 32; do NBS_PME_BATCH=$b timeout 200 python tools/time_kernels.py C4 10 2>&1 | tail -1 >> gpurun_out/r02_time_p.log; done
for b in 16 32; do NBS_PME_BATCH=$b timeout 300 python tools/time_kernels.py C5 5 2>&1 | tail -1 >> gpurun_out/r02_time_p.log; done
cat gpurun_out/r02_time_p.log
