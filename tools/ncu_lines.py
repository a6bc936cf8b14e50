"""Per-source-line executed-instruction totals of one kernel from an .ncu-rep (needs -lineinfo and --import-source).
Usage: ncu_lines.py rep rays [top]"""
import csv, subprocess, sys, io
rep, rays = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname, rows, iex, ithr = None, [], None, None
for r in csv.reader(io.StringIO(raw)):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        iex, ithr = r.index("Instructions Executed"), r.index("Thread Instructions Executed")
        continue
    if iex is None or len(r) <= ithr or r[2] != "-":          # cuda rows carry "-" in the Address column
        continue
    try:
        n, th = int(r[iex]), int(r[ithr])
    except ValueError:
        continue
    if n:
        rows.append((n, th, fname, r[0], r[1].strip()[:105]))
tot = sum(r[0] for r in rows)
print(f"total warp instructions {tot}  (x32 / rays = {tot * 32 / rays:.0f}); thread instructions / ray = {sum(r[1] for r in rows) / rays:.0f}")
for n, th, f, ln, src in sorted(rows, reverse=True)[:top]:
    print(f"{n * 32 / rays:8.1f} {n / tot * 100:5.1f}% lanes {th / n:5.1f}  {f}:{ln:>5}  {src}")
