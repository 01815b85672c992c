"""Text digest of an `ncu --set full` report for profiles/: python scripts/ncu_digest.py rep [units] [title]

Sections: key rows of the details page, pipe utilisation + DRAM bytes from the raw page, opcode mix, and the source lines
with the most stall samples (`units` = work items of the launch, e.g. circuit instances, for per-unit instruction counts).
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
title = sys.argv[3] if len(sys.argv) > 3 else rep


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


KEEP = ("Duration", "Elapsed Cycles", "DRAM Throughput", "Executed Ipc Active", "Issue Slots Busy", "Mem Busy",
        "Max Bandwidth", "L1/TEX Hit Rate", "L2 Hit Rate", "One or More Eligible", "No Eligible",
        "Active Warps Per Scheduler", "Eligible Warps Per Scheduler", "Warp Cycles Per Issued Instruction",
        "Executed Instructions", "Registers Per Thread", "Dynamic Shared Memory Per Block",
        "Static Shared Memory Per Block", "Block Size", "Grid Size", "Waves Per SM", "Block Limit Registers",
        "Block Limit Shared Mem", "Block Limit Warps", "Theoretical Active Warps per SM", "Theoretical Occupancy",
        "Achieved Occupancy", "SM Frequency", "DRAM Frequency", "Compute (SM) Throughput", "Memory Throughput",
        "SM Busy", "Shared Memory Configuration Size")
print(f"# {title}")
rows = list(csv.reader(ncu("--page", "details", "--csv").splitlines()))
name = None
for r in rows[1:]:
    if len(r) > 14 and r[12] in KEEP and r[13]:
        if name is None:
            name = r[4]
            print(f"# kernel: {name}")
        print(f"| {r[11][:32]:32s} | {r[12][:44]:44s} | {r[13]:15s} | {r[14]}")

RAW = ("sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_tensor_op_dmma.sum",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
       "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "gpu__time_duration.sum", "smsp__inst_executed.sum")
raw = list(csv.reader(ncu("--page", "raw", "--csv").splitlines()))
print("\n# raw counters")
if len(raw) >= 3:
    hdr, unit, val = raw[0], raw[1], raw[2]
    for i, h in enumerate(hdr):
        if h in RAW or any(h.startswith(p) for p in ("sm__pipe_tensor", "sm__inst_executed_pipe_tensor")):
            print(f"{h:78s} {val[i]:>18s} {unit[i]}")
    try:
        rd = float(val[hdr.index("dram__bytes_read.sum")].replace(",", ""))
        wr = float(val[hdr.index("dram__bytes_write.sum")].replace(",", ""))
        ur, uw = unit[hdr.index("dram__bytes_read.sum")], unit[hdr.index("dram__bytes_write.sum")]
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = rd * mult.get(ur, 1.0) + wr * mult.get(uw, 1.0)
        dur = float(val[hdr.index("gpu__time_duration.sum")].replace(",", ""))
        ud = unit[hdr.index("gpu__time_duration.sum")]
        sec = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(ud, 1e-9)
        print(f"dram bytes read+write per launch = {tot:.0f} B; {tot / sec / 1e9:.1f} GB/s over {sec * 1e6:.1f} us (under ncu)")
    except (ValueError, IndexError):
        pass

src = list(csv.reader(ncu("--page", "source", "--csv").splitlines()))
if len(src) > 2:
    hdr = src[1]
    data = [r for r in src[2:] if len(r) == len(hdr)]
    try:
        isrc, ist, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        tot_s = sum(int(r[ist] or 0) for r in data) or 1
        tot_e = sum(int(r[iex] or 0) for r in data) or 1
        print(f"\n# opcode mix\nstatic instrs {len(data)} executed {tot_e} ({tot_e / units:.0f} per unit) samples {tot_s}")
        h, hs = collections.Counter(), collections.Counter()
        for r in data:
            t = r[isrc].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            h[op] += int(r[iex] or 0)
            hs[op] += int(r[ist] or 0)
        for op, c in h.most_common(16):
            print(f"{op:10s} exec {c / tot_e * 100:5.1f}%  stall {hs[op] / tot_s * 100:5.1f}%")
    except ValueError:
        pass

out = ncu("--page", "source", "--print-source", "cuda,sass", "--csv")
cur, lines = None, []
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name"):
        continue
    if r[0] != "" and r[2] == "-":
        try:
            lines.append((int(r[7]) / units, int(r[4]), cur, r[0], r[1].strip()[:100]))
        except ValueError:
            pass
if lines:
    ts = sum(x[1] for x in lines) or 1
    print(f"\n# instructions per unit and stall share per source line (top 14 by stall samples)")
    for x in sorted(lines, key=lambda x: -x[1])[:14]:
        print(f"{x[0]:9.1f} {100 * x[1] / ts:5.1f}% {x[2]}:{x[3]}  {x[4]}")
