timeout 300 python tools/parity_detail.py C2 double > gpurun_out/r02_parity_c2_double.log 2>&1; cat gpurun_out/r02_parity_c2_double.log
timeout 300 python tools/parity_detail.py C2 mixed > gpurun_out/r02_parity_c2_mixed.log 2>&1; cat gpurun_out/r02_parity_c2_mixed.log
