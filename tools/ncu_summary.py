"""Summarises one kernel of an `ncu --set full --import-source on` capture as text: key metrics, stall reasons, opcode
mix per pair-kernel step (when the step count is given) and the hottest source lines.
usage: ncu_summary.py capture.ncu-rep [warp_steps] > profiles/<name>_summary.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None


def page(*args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("--page", "raw")
names, units, values = raw[0], raw[1], raw[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_active", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max"]
print(f"# {rep}")
for w in want:
    for i, n in enumerate(names):
        if n == w:
            print(f"{n:75s} {values[i]:>16s} {units[i]}")
src = page("--page", "source", "--print-source", "sass")
hdr = src[1]
stall = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tot = {hdr[i]: 0 for i in stall}
i_src, i_inst = hdr.index("Source"), hdr.index("Instructions Executed")
mix = {}
for r in src[2:]:
    for i in stall:
        try:
            tot[hdr[i]] += int(r[i])
        except (ValueError, IndexError):
            pass
    try:
        parts = r[i_src].split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        mix[op] = mix.get(op, 0) + int(r[i_inst])
    except (ValueError, IndexError):
        pass
total = sum(tot.values())
print("\n# warp stall samples (all samples)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"{k:28s} {v:7d} {100*v/max(total, 1):5.1f} %")
total_inst = sum(mix.values())
print(f"\n# opcode mix ({total_inst} warp instructions" + (f", {total_inst/steps:.0f} per warp-step of 32 pair evaluations" if steps else "") + ")")
for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:24]:
    print(f"{k:10s} {100*v/total_inst:5.1f} %" + (f" {v/steps:7.1f} per step" if steps else ""))
both = page("--page", "source", "--print-source", "sass,cuda")
hdr = both[2]
i_samp, i_inst, i_wave = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared")
lines = {}
for r in both[3:]:
    if r and r[0] != "":
        try:
            a = lines.setdefault(int(r[0]), [r[1], 0, 0, 0])
            a[1] += int(r[i_samp]); a[2] += int(r[i_inst]); a[3] += int(r[i_wave] or 0)
        except (ValueError, IndexError):
            pass
ti, ts, tw = (sum(v[k] for v in lines.values()) for k in (2, 1, 3))
print("\n# hottest source lines (share of instructions / stall samples / shared-memory wavefronts)")
for ln, (text, smp, inst, wave) in sorted(sorted(lines.items(), key=lambda kv: -kv[1][2])[:28]):
    print(f"{ln:5d} {100*inst/max(ti, 1):5.1f} % {100*smp/max(ts, 1):5.1f} % {100*wave/max(tw, 1):5.1f} %  {text.strip()[:100]}")
