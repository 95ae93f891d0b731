"""Static SASS size of one kernel broken down by the source function each instruction is attributed to.

    cuobjdump -xelf all gym_td_b200/libtd_b200.so && nvdisasm -g td_engine.sm_100a.cubin > all.txt
    python tools/sass_by_function.py all.txt 'td_step_kernelILi1ELb0ELi100'
"""
import os
import re
import sys
from collections import Counter

SRCS = ["gym_td_b200/csrc/%s" % f for f in ("td_common.cuh", "td_rng.cuh", "td_rules.cuh", "td_obs.cuh", "td_kernels.cuh")]


def function_spans(path):
    """(first line, last line, name) of every device function / kernel; works inside `namespace td {`."""
    spans, depth, pending = [], 0, None
    for i, ln in enumerate(open(path).read().split("\n"), 1):
        if pending is None and "#define" not in ln:
            m = re.search(r"(?:__device__|__global__)[^;]*?(\w+)\s*\([^;]*$", ln) or re.match(r"^(td_\w+)\(", ln)
            if m:
                pending = (m.group(1), i, depth)
        depth += ln.count("{") - ln.count("}")
        if pending and "}" in ln and depth == pending[2]:
            spans.append((pending[1], i, pending[0]))
            pending = None
    return spans


def main():
    dis, key = sys.argv[1], sys.argv[2]
    spans = {os.path.basename(f): function_spans(f) for f in SRCS}
    counts, cur, active, total = Counter(), "?", False, 0
    for ln in open(dis):
        if ln.startswith("//---") and ".text." in ln:
            active = key in ln
            cur = "?"
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            base = m.group(1).split("/")[-1]
            if base in spans:
                n = int(m.group(2))
                cur = next((f for a, b, f in spans[base] if a <= n <= b), "%s line %d" % (base, n))
            else:
                cur = "<" + m.group(1).split("/")[-1] + ">"
            continue
        if re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s", ln):
            counts[cur] += 1
            total += 1
    for f, c in counts.most_common(40):
        print("%6d  %5.1f%%  %s" % (c, 100.0 * c / total, f))
    print("%6d  total" % total)


if __name__ == "__main__":
    main()
