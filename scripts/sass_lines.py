#!/usr/bin/env python
"""Static SASS accounting for one kernel of an (in-tree) cubin / .so: instruction counts per opcode and per source line
(the kernels are compiled with -lineinfo).  The transform kernels are instruction-issue bound and their hot loops are
fully unrolled, so the static count of the loop body is the dynamic count per tile - this is the offline metric the
kernel work is steered by between GPU runs.

    python scripts/sass_lines.py <kernel-name-substring> [--so PATH] [--top N] [--inline]
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel")
    ap.add_argument("--so", default=os.path.join(ROOT, "xai-audio-deepfakes_b200", "libaddvisor_sm100.so"))
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--inline", action="store_true", help="attribute to the innermost inlined frame's caller chain")
    ap.add_argument("--range", default=None, help="only count lines lo-hi of the kernel's own file")
    args = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="sass_")
    if args.so.endswith(".cubin"):
        cubins = [args.so]
    else:
        subprocess.run(["cuobjdump", "-xelf", "all", args.so], cwd=tmp, capture_output=True)
        cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    for cb in cubins:
        syms = subprocess.run(["readelf", "-sW", cb], capture_output=True, text=True).stdout
        for line in syms.splitlines():
            f = line.split()
            if len(f) >= 8 and f[3] == "FUNC" and args.kernel in f[-1]:
                idx = int(f[0].rstrip(":"))
                out = subprocess.run(["nvdisasm", "-gi" if args.inline else "-g", "-fun", str(idx), cb],
                                     capture_output=True, text=True).stdout
                report(f[-1], out, args)
    return 0


def report(name, text, args):
    cur = None
    per_line, per_op = collections.Counter(), collections.Counter()
    for line in text.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            if args.inline:
                chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
                if chain:
                    cur = cur + tuple((os.path.basename(a), int(b)) for a, b in chain)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
        if m:
            per_line[cur] += 1
            per_op[m.group(1)] += 1
    total = sum(per_op.values())
    print(f"== {name}: {total} instructions")
    print("  " + "  ".join(f"{k}:{v}" for k, v in per_op.most_common(24)))
    for k, v in per_line.most_common(args.top):
        print(f"  {v:5d}  {k}")


if __name__ == "__main__":
    sys.exit(main())
