"""Summarise the source page of an ncu report of viterbi_pair_kernel: where the warp samples fall
(ACS loop vs traceback vs rest), per-region issue fraction and top stall reasons.

usage: python profiles/ncu_source_summary.py report.ncu-rep [--dump-region trace|loop] [--top N]
"""
import collections
import csv
import subprocess
import sys


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    return {h: i for i, h in enumerate(hdr)}, hdr, rows[2:]


def main():
    path = sys.argv[1]
    ix, hdr, data = load(path)
    keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    groups = collections.defaultdict(list)
    for r in data:
        groups[int(r[ix["Instructions Executed"]])].append(r)
    print("total samples", tot)
    for e, g in sorted(groups.items(), key=lambda kv: -sum(int(r[ix["# Samples"]]) for r in kv[1]))[:6]:
        n = sum(int(r[ix["# Samples"]]) for r in g)
        sel = sum(int(r[ix["stall_selected"]]) for r in g)
        agg = {k: sum(int(r[ix[k]]) for r in g) for k in keys}
        top = [(k[6:], "%.1f%%" % (100 * v / max(n, 1))) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:6]]
        print("exec=%d ninstr=%d samples=%d (%.1f%%) issue=%.3f %s" % (e, len(g), n, 100 * n / tot, sel / max(n, 1), top))
    if "--dump" in sys.argv:
        e = int(sys.argv[sys.argv.index("--dump") + 1])
        for r in groups[e]:
            st = sorted(((k[6:], int(r[ix[k]])) for k in keys if int(r[ix[k]]) > 0), key=lambda x: -x[1])
            print(r[ix["Address"]][-5:], r[ix["Source"]].strip()[:72].ljust(72), r[ix["# Samples"]].rjust(5), st[:3])
    if "--hot" in sys.argv:
        n = int(sys.argv[sys.argv.index("--hot") + 1])
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
            st = sorted(((k[6:], int(r[ix[k]])) for k in keys if int(r[ix[k]]) > 0), key=lambda x: -x[1])
            print(r[ix["Address"]][-5:], r[ix["Instructions Executed"]].rjust(8), r[ix["Source"]].strip()[:64].ljust(64),
                  r[ix["# Samples"]].rjust(5), st[:3])


if __name__ == "__main__":
    main()
