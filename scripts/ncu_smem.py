"""Shared-memory wavefronts (total / excessive) per CUDA source line: python scripts/ncu_smem.py rep units [top]"""
import csv, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; rows = []
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 25 or r[0] in ("Line No", "Function Name"): continue
    if r[0] != "" and r[2] == "-":
        try: rows.append((int(r[19]) / units, int(r[18]) / units, int(r[24]) / units, cur, r[0], r[1].strip()[:90]))
        except ValueError: pass
print(f"per unit: shared wavefronts {sum(x[0] for x in rows):.0f}, excessive {sum(x[1] for x in rows):.0f}, local sectors {sum(x[2] for x in rows):.0f}")
for x in sorted(rows, key=lambda x: -x[0])[:top]:
    print(f"{x[0]:8.1f} exc {x[1]:7.1f} loc {x[2]:7.1f} {x[3]}:{x[4]}  {x[5]}")
