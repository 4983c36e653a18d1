"""Hot source lines of an ncu report: `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`
aggregated per CUDA source line (warp-stall samples, executed instructions).

usage: python scripts/ncu_source_hot.py gpurun_out/X.ncu-rep [top]
"""
import csv
import io
import subprocess
import sys


def main(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fpath, func, hdr = None, None, None
    agg = {}
    total = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1].split("(")[0]
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or r[0] == "" or not r[0].isdigit():
            continue
        try:
            s = int(r[4])
            ins = int(r[hdr["Instructions Executed"]]) if r[hdr["Instructions Executed"]].isdigit() else 0
        except (ValueError, IndexError):
            continue
        key = (func, fpath, int(r[0]), r[1].strip()[:110])
        a = agg.setdefault(key, [0, 0])
        a[0] += s
        a[1] += ins
        total += s
    print(f"# {rep}: {total} stall samples")
    for (func, fp, ln, src), (s, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100.0 * s / max(total, 1):6.2f}%  {ins:>10d} inst  {fp}:{ln:<4d} {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
