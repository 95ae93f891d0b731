#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction samples with CUDA source lines (via nvdisasm -g).

    python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-mangled-substring> [top]

ncu's CSV source page carries metrics only per SASS instruction; nvdisasm supplies the line of
each instruction offset.  Prints instructions-executed and stall-sample shares per source line and
per coarse phase (line ranges given in PHASES below).
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


CSRC_FILES = ("td_common.cuh", "td_rng.cuh", "td_rules.cuh", "td_obs.cuh", "td_kernels.cuh")


def sass_lines(lib, kernel):
    """SASS offset -> (source file base name, line) of one kernel, from nvdisasm -g."""
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    out = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, active = None, False
        for ln in txt.splitlines():
            if ln.startswith(".text."):
                active = kernel in ln
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
            if m:
                out[int(m.group(1), 16)] = cur
    return out


def main():
    rep, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lines = sass_lines(lib, kernel)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    names = [rows[i - 1][1] if i > 0 else "" for i in starts]
    i0 = starts[0]
    end = starts[1] - 1 if len(starts) > 1 else len(rows)
    hdr = rows[i0]
    data = [r for r in rows[i0 + 1:end] if len(r) == len(hdr)]
    base = int(data[0][0], 16)
    si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
    per_line_i, per_line_s = collections.Counter(), collections.Counter()
    for r in data:
        off = int(r[0], 16) - base
        ln = lines.get(off)
        per_line_i[ln] += int(r[ie] or 0)
        per_line_s[ln] += int(r[si] or 0)
    ti, ts = sum(per_line_i.values()), sum(per_line_s.values())
    print("kernel:", names[0][:80], "sass:", len(data), "inst:", ti, "samples:", ts)
    csrc = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc")
    src = {f: open(os.path.join(csrc, f)).read().splitlines() for f in CSRC_FILES}
    print("-- top lines by stall samples")
    for key, s_ in per_line_s.most_common(top):
        f, ln = key if key else ("?", 0)
        text = src[f][ln - 1].strip()[:100] if f in src and 0 < ln <= len(src[f]) else "?"
        print("%-14s %5s  samp %5.1f%%  inst %5.1f%%  | %s" % (f, ln, 100.0 * s_ / ts, 100.0 * per_line_i[key] / ti, text))
    # coarse phases by function: find "__device__ ... name(" definitions
    funcs = {}
    for f, text in src.items():
        funcs[f] = []
        for n, t in enumerate(text, 1):
            m = re.match(r"\s*(?:template.*>\s*)?(?:__global__|__device__).*?\b(\w+)\s*\(", t)
            if m and not t.strip().endswith(";"):
                funcs[f].append((n, m.group(1)))

    def func_of(key):
        if not key:
            return "?"
        f, ln = key
        if f not in funcs:
            return "<" + f + ">"
        name = "?"
        for n, fn in funcs[f]:
            if n <= ln:
                name = fn
        return name

    fi, fs = collections.Counter(), collections.Counter()
    for key in per_line_i:
        fi[func_of(key)] += per_line_i[key]
        fs[func_of(key)] += per_line_s[key]
    print("-- per function")
    for f, s_ in fs.most_common():
        print("%-22s samp %5.1f%%  inst %5.1f%%  (%d warp-inst)" % (f, 100.0 * s_ / ts, 100.0 * fi[f] / ti, fi[f]))


if __name__ == "__main__":
    main()
