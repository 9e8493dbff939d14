V=$PWD/openmm-nonbonded-slicing_b200/csrc/variants
out=gpurun_out/r02_occupancy_variants.log
rm -f $out
for lib in default w6m3 w4m5; do
  if [ $lib = default ]; then unset NBS_B200_LIBRARY; else export NBS_B200_LIBRARY=$V/lib_$lib.so; fi
  echo "== $lib" >> $out
  timeout 120 python tools/time_kernels.py C3 20 2>&1 | tail -1 | sed 's/.*C3 /C3 /' | cut -c1-110 >> $out
  timeout 120 python tools/time_kernels.py C3 20 forces 2>&1 | tail -1 | sed 's/.*C3 /C3 /' | cut -c1-110 >> $out
  timeout 300 python tools/time_kernels.py C5 4 2>&1 | tail -1 | sed 's/.*C5 /C5 /' | cut -c1-110 >> $out
done
cat $out
