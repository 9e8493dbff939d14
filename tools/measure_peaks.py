"""Prints the measured FP32 FMA and rsqrt rates of cuda:0 (nbs_measure_peaks) as one JSON line."""
import ctypes as C
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nbs = importlib.import_module("openmm-nonbonded-slicing_b200")
lib = nbs.abi.load_library()
out = (C.c_double*4)()
nbs.abi.check(lib.nbs_measure_peaks(0, out))
dp = (C.c_double*8)()
nbs.abi.check(lib.nbs_measure_dp_rates(0, dp))
print(json.dumps({"warp_instr_per_clk_per_sm": {"dfma": round(dp[0], 3), "i32_to_f64": round(dp[1], 3), "f32_to_f64": round(dp[2], 3),
                                                "f64_to_f32": round(dp[3], 3), "i64_to_f64": round(dp[4], 3)}}))
print(json.dumps({"fp32_fma_tflops": round(out[0], 2), "rsqrt_gops": round(out[1], 1), "sms": int(out[2]),
                  "nominal_sm_mhz": out[3], "derived_fp32_tflops_at_nominal": round(out[2]*128*2*out[3]*1e-6, 2)}))
