"""Top-N SASS instructions by stall samples from an `ncu --page source --csv` export (one table per kernel)."""
import csv, sys
lines = open(sys.argv[1]).read().splitlines()
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
secs = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
for a, b in zip(secs[:-1], secs[1:]):
    rows = list(csv.reader(lines[a + 1:b]))
    hdr, data = rows[0], rows[1:]
    si, ii, xi = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(int(d[si]) for d in data)
    tot_inst = sum(int(d[xi]) for d in data)
    print("=====", lines[a][:100], "samples", tot, "warp-instr", tot_inst)
    # opcode histogram weighted by executed count
    hist = {}
    for d in data:
        op = d[ii].split()[0] if not d[ii].strip().startswith("@") else d[ii].split()[1]
        op = op.split(".")[0]
        hist[op] = hist.get(op, 0) + int(d[xi])
    print("  executed mix:", ", ".join(f"{k}:{100*v/tot_inst:.1f}%" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:14]))
    for n, d in sorted(enumerate(data), key=lambda nd: -int(nd[1][si]))[:topn]:
        print(f"  {n:5d} {100*int(d[si])/tot:5.1f}%  exec {int(d[xi]):9d}  {d[ii].strip()[:80]}")
