"""Executed instructions / stall samples per CUDA source line: python scripts/ncu_lines.py rep units [top]"""
import csv, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; rows = []
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name"): continue
    if r[0] != "" and r[2] == "-":
        try: rows.append((int(r[7]) / units, int(r[4]), cur, r[0], r[1].strip()[:95]))
        except ValueError: pass
tot = sum(x[0] for x in rows); ts = sum(x[1] for x in rows)
print(f"instructions per unit: {tot:.0f}; stall samples {ts}")
for x in sorted(rows, key=lambda x: -x[1])[:top]:
    print(f"{x[0]:8.1f} {100*x[1]/ts:5.1f}% {x[2]}:{x[3]}  {x[4]}")
