"""Top warp-stall reasons + instruction mix per kernel of ncu --set full reports (first launch of each kernel name).
usage: python tools/ncu_stalls.py <reports...>"""
import csv
import io
import subprocess
import sys


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(paths):
    for path in paths:
        rows = raw(path)
        if len(rows) < 3:
            print(path, "no data")
            continue
        h = rows[0]
        kn = h.index("Kernel Name")
        stalls = [k for k in h if "smsp__average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio")]
        print(f"== {path}")
        seen = set()
        for r in rows[2:]:
            name = r[kn].split("(")[0][:70]
            if name in seen:
                continue
            seen.add(name)
            ss = sorted(((float(r[h.index(k)].replace(",", "")), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                         for k in stalls), reverse=True)[:5]
            print(f"  {name:70s} " + "  ".join(f"{n}={v:.2f}" for v, n in ss))


if __name__ == "__main__":
    main(sys.argv[1:])
