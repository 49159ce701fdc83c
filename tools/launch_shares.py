"""Per-kernel shares of ONE forward from an `ncu --metrics gpu__time_duration.sum --csv` launch list: the launches between two
consecutive k_pack_luma (first kernel of a forward) are one forward; prints symbol | launches | us | share."""
import collections, csv, re, sys


def main(path, which=1):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    starts = [i for i, r in enumerate(data) if r[ki].startswith("k_pack_luma") or "k_pack_luma" in r[ki]]
    if len(starts) < which + 2:
        which = 0
    seg = data[starts[which]:starts[which + 1]] if len(starts) > which + 1 else data[starts[which]:]
    agg = collections.OrderedDict()
    for r in seg:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("rf::", "")
        name = re.sub(r"\((int|bool)\)", "", name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi]) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"# one forward: {len(seg)} launches, sum {tot / 1e3:.3f} ms (cold-cache, serialised under ncu: compare SHARES with bench.py's "
          f"kernel_ms_per_step / symbol_ms_per_step)")
    print("# kernel | launches | us | share")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k} | {n} | {us:.1f} | {us / tot:.3f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
