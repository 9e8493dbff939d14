set -x
V=openmm-nonbonded-slicing_b200/csrc/variants
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reference_fixture or baseline_configs or deterministic or list_reuse or nonperiodic or random_systems" 2>&1 | tail -15 > gpurun_out/r02_gputest_c.log; cat gpurun_out/r02_gputest_c.log
rm -f gpurun_out/r02_time_e.log
for lib in default exactinline warps4 warps6; do
  if [ $lib = default ]; then unset NBS_B200_LIBRARY; else export NBS_B200_LIBRARY=$PWD/$V/lib_$lib.so; fi
  NBS_LIST_SKIN=0 timeout 120 python tools/time_kernels.py C3 20 2>&1 | tail -1 >> gpurun_out/r02_time_e.log
  NBS_LIST_SKIN=0 timeout 120 python tools/time_kernels.py C3 20 forces 2>&1 | tail -1 >> gpurun_out/r02_time_e.log
done
unset NBS_B200_LIBRARY
cat gpurun_out/r02_time_e.log
NBS_LIST_SKIN=0 python tools/one_eval.py C3 3 > gpurun_out/plain.log 2>&1 && NBS_LIST_SKIN=0 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 2 -c 1 -f -o gpurun_out/r02_k_pair_b python tools/one_eval.py C3 3 > gpurun_out/ncu.log 2>&1; tail -3 gpurun_out/ncu.log
