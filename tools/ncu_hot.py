"""Hottest SASS instructions of one kernel of an ncu report (warp-stall samples per instruction, with the dominant stall
reasons) -- `ncu -i rep --page source --csv --kernel-id ::regex:<name>:<n>` condensed.

    python tools/ncu_hot.py gpurun_out/x.ncu-rep 'regex:k_tc_gemm' 3 [top]
"""
import csv, subprocess, sys


def main(rep, kernel, nth, top=30):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::{kernel}:{nth}"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stalls = [(h, hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    num = lambda v: int(v) if v.isdigit() else 0
    tot = sum(num(r[si]) for r in data) or 1
    print("#", rows[0][1][:100] if rows[0] else "", "| samples", tot, "| instructions", len(data))
    agg = {h: sum(num(r[i]) for r in data) for h, i in stalls}
    print("# stall reasons:", ", ".join(f"{h[6:]} {100 * v / tot:.0f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for idx, r in sorted(enumerate(data), key=lambda t: -num(t[1][si]))[:top]:
        st = sorted(((h[6:], num(r[i])) for h, i in stalls if num(r[i]) > 0), key=lambda kv: -kv[1])[:3]
        print(f"{idx:5d} {100 * num(r[si]) / tot:5.1f}%  exec {r[ie]:>8s}  {r[src].strip()[:64]:64s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 30)
