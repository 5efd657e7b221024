"""Top stall sites of one launch in an ncu --set full --import-source report: SASS lines ranked by warp-stall samples.
usage: python tools/ncu_hot.py <rep> <launch-skip> [top]"""
import csv, io, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1",
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] in ("Address", "Line No")][0]
h = rows[hi]
print(rows[1][1][:120] if len(rows) > 1 and len(rows[1]) > 1 else "")
isrc = h.index("Source")
isamp = h.index("# Samples")
iexe = h.index("Instructions Executed")
stalls = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Address", "Line No", "Kernel Name"):
        break
    if len(r) > max(isamp, iexe):
        data.append(r)
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot)
rank = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:top]
for i in sorted(rank):
    r = data[i]
    st = sorted(((int(r[j] or 0), h[j][6:]) for j in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[isamp]):6d} {100*int(r[isamp])/max(tot,1):5.1f}% exe={r[iexe]:>8s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}  {r[isrc][:90]}")
