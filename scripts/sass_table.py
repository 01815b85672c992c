"""SASS mnemonic table of the shipped library (profiles/r02/sass_table.txt): python scripts/sass_table.py > out.txt
Runs cuobjdump here (no GPU needed)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "matchcake_b200/libmcb200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
archs = collections.Counter(re.findall(r"arch = (sm_\w+)", sass))
print(f"# cuobjdump -sass {lib} | grep -c <mnemonic>  (cubins by arch: {dict(archs)})")
names = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", "UTMALDG", "UTCMMA|UTCHMMA|UTCQMMA", "LDTM", "CREDUX|REDUX", "SHFL",
         "LDS", "STS", "LDG", "STG", "MUFU"]
per_kernel = collections.OrderedDict()
cur = None
tot = collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per_kernel[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        per_kernel[cur][op] += 1
        tot[op] += 1
for n in names:
    c = sum(v for k, v in tot.items() if any(k.startswith(p) for p in n.split("|")))
    print(f"{n:28s} {c:8d}")
print("\n# per kernel: DMMA / DFMA+DMUL+DADD / UBLKCP")
for k, c in sorted(per_kernel.items()):
    fp = sum(v for o, v in c.items() if o in ("DFMA", "DMUL", "DADD"))
    print(f"{c['DMMA']:6d} {fp:6d} {c['UBLKCP']:4d}  {k}")
