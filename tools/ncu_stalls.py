"""Per-source-line warp-stall samples of one kernel from an .ncu-rep (-lineinfo, --import-source on).
Usage: ncu_stalls.py rep [top]   -> samples, share, dominant stall reasons, instructions executed per source line"""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname, hdr, rows = None, None, []
for r in csv.reader(io.StringIO(raw)):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or r[2] != "-":          # cuda rows carry "-" in the Address column
        continue
    d = dict(zip(hdr, r))
    try:
        n = int(d["# Samples"])
    except ValueError:
        continue
    st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
    rows.append((n, int(d["Instructions Executed"] or 0), fname, r[0], r[1].strip()[:90], st))
tot = sum(r[0] for r in rows)
print(f"total samples {tot}")
for n, ie, f, ln, src, st in sorted(rows, key=lambda r: -r[0])[:top]:
    why = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{n:7d} {n / tot * 100:5.1f}%  inst {ie:9d}  {f}:{ln:>4}  {src}\n{'':16}[{why}]")
