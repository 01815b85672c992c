"""Print the useful rows of an ncu report: python scripts/ncu_summary.py file.ncu-rep [--source N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
for r in list(csv.reader(out.splitlines()))[1:]:
    if len(r) > 14 and r[13]:
        print(f"{r[4][:40]:40s} | {r[11][:32]:32s} | {r[12][:52]:52s} | {r[13]:12s} | {r[14]}")
if "--source" in sys.argv:
    n = int(sys.argv[sys.argv.index("--source") + 1])
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    print(hdr)
