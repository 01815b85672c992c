"""Opcode histogram (executed / stall samples) from an ncu report's source page."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
isrc = hdr.index("Source"); ist = hdr.index("Warp Stall Sampling (All Samples)"); iex = hdr.index("Instructions Executed")
tot_s = sum(int(r[ist] or 0) for r in data); tot_e = sum(int(r[iex] or 0) for r in data)
print("static instrs", len(data), "executed", tot_e, "samples", tot_s)
h = collections.Counter(); hs = collections.Counter()
for r in data:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    h[op] += int(r[iex] or 0); hs[op] += int(r[ist] or 0)
for op, c in h.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 18):
    print(f"{op:10s} exec {c/tot_e*100:5.1f}%  stall {hs[op]/tot_s*100:5.1f}%")
