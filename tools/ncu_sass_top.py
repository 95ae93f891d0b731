#!/usr/bin/env python
"""Top stalled SASS instructions of the first kernel in an ncu report, with their stall reasons.

    python tools/ncu_sass_top.py <report.ncu-rep> [top=30] [context-index ...]

Looks for what a per-line summary hides: local-memory reloads (LDL), slow-path calls (CALL.REL), waits
on a single late load.  Extra integer arguments print +-18 instructions of context around that index.
"""
import csv
import subprocess
import sys


def load(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, out, name = None, [], ""
    for r in rows:
        if r and r[0] == "Kernel Name" and hdr is None:
            name = r[1]
        if r and r[0] == "Address":
            if hdr is not None:
                break
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            out.append(r)
    return name, hdr, out


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    name, hdr, out = load(rep)
    ix = {n: k for k, n in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in out)
    inst = sum(int(r[ix["Instructions Executed"]]) for r in out)
    print("%s\nSASS %d, warp-instructions %d, samples %d" % (name, len(out), inst, tot))
    stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    agg = {n: sum(int(r[ix[n]]) for r in out) for n in stalls}
    print("stall mix: " + ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / tot) for n, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
    for k in sorted(range(len(out)), key=lambda k: -int(out[k][ix["# Samples"]]))[:top]:
        r = out[k]
        s = sorted(((n[6:], int(r[ix[n]])) for n in stalls if int(r[ix[n]]) > 0), key=lambda x: -x[1])[:3]
        print("%5d %5.1f%% x%-8s %-58s %s" % (k, 100.0 * int(r[ix["# Samples"]]) / tot, r[ix["Instructions Executed"]],
                                             r[ix["Source"]].strip()[:58], s))
    special = [k for k, r in enumerate(out) if any(t in r[ix["Source"]] for t in ("LDL", "STL", "CALL.REL"))
               and int(r[ix["Instructions Executed"]]) > 0]
    print("local-memory / call instructions executed:")
    for k in special:
        r = out[k]
        print("%5d %5.1f%% x%-8s %s" % (k, 100.0 * int(r[ix["# Samples"]]) / tot, r[ix["Instructions Executed"]],
                                       r[ix["Source"]].strip()[:70]))
    for c in sys.argv[3:]:
        c = int(c)
        print("-- context %d" % c)
        for k in range(max(0, c - 18), min(len(out), c + 18)):
            r = out[k]
            print("%5d %5s x%-8s %s" % (k, r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:100]))


if __name__ == "__main__":
    main()
