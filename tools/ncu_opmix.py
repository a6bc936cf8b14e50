"""Dynamic opcode mix and hottest source lines of one kernel launch from an .ncu-rep.  Usage: ncu_opmix.py rep rays [launch_index]"""
import csv, subprocess, sys, collections, re
rep, rays = sys.argv[1], float(sys.argv[2])
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] , capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name"')[1:]
blk = '"Kernel Name"' + blocks[which]
rows = list(csv.reader(blk.splitlines()))
print(rows[0][1][:80])
hdr = rows[1]
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
ithr = hdr.index("Thread Instructions Executed")
ops = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) <= iex or not r[iex].isdigit(): continue
    t = r[isrc].split()
    if not t: continue
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += int(r[iex]); tot += int(r[iex])
print("warp instructions", tot, " per ray (x32):", tot * 32 / rays)
for k, v in ops.most_common(28):
    print(f"  {k:10s} {v / tot * 100:5.1f}%   per ray {v * 32 / rays:7.1f}")
