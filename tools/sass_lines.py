#!/usr/bin/env python
"""Static SASS instruction counts per source line of one kernel, without a GPU (nvdisasm line info of the built library).

usage: tools/sass_lines.py <kernel-name-substring> [--lib path.so] [--by outer|inner|both] [--top N] [extra nvcc flags ...]

  outer  instructions attributed to the line of the KERNEL BODY they were inlined into (where in the pass the code sits)
  inner  instructions attributed to the innermost function line (which primitive costs what: normalize, division, unpack ...)

With extra flags (-DHK_...=1) the .cu file that holds the kernel is recompiled into a scratch object first, so that a tuning
variant can be compared with the default build line by line.  Static counts are not time — a loop body counts once — but for
the ALU-bound kernels (k_spatial, k_denoise: DESIGN.md 4) the instruction count of the per-neighbour / per-tap body is what
there is to remove, and this shows where it is."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bevy_hikari_b200 import build as B  # noqa: E402

SOURCES = ["csrc/kernels_light.cu", "csrc/kernels_post.cu", "csrc/kernels_upscale.cu"]


def disassemble(kernel, flags):
    with tempfile.TemporaryDirectory() as d:
        texts = []
        if flags:
            for src in SOURCES:
                obj = os.path.join(d, os.path.basename(src) + ".o")
                r = subprocess.run([B.NVCC] + B.NVCC_FLAGS + flags + ["-c", os.path.join(B.HERE, src), "-o", obj], capture_output=True, text=True)
                if r.returncode:
                    sys.exit(r.stderr)
                subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
        else:
            subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(B.HERE, "libhikari_b200.so")], cwd=d, capture_output=True)
        for f in sorted(os.listdir(d)):
            if f.endswith(".cubin"):
                texts.append(subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(d, f)], capture_output=True, text=True).stdout)
    return "\n".join(texts)


def main():
    args = sys.argv[1:]
    if not args:
        sys.exit(__doc__)
    kernel = args[0]
    by, top, flags = "both", 25, []
    i = 1
    while i < len(args):
        if args[i] == "--by":
            by = args[i + 1]; i += 2
        elif args[i] == "--top":
            top = int(args[i + 1]); i += 2
        else:
            flags.append(args[i]); i += 1
    text = disassemble(kernel, flags)
    cur = None
    outer, inner, total, opcodes = collections.Counter(), collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    chain = []
    for line in text.splitlines():
        m = re.match(r"\.text\.(\S+):", line)
        if m:
            cur = m.group(1) if kernel in m.group(1) else None
            chain = []
            continue
        if cur is None:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            if not getattr(main, "_in_chain", False):
                chain = []
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            main._in_chain = True
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            main._in_chain = False
            if chain:
                inner[(cur, chain[0])] += 1
                outer[(cur, chain[-1])] += 1
                opcodes[(cur, chain[-1])][m.group(1).split(".")[0]] += 1
            total[cur] += 1
    for k in sorted(total):
        print(f"== {k}: {total[k]} instructions")
        for name, table in (("outer", outer), ("inner", inner)):
            if by not in (name, "both"):
                continue
            rows = sorted(((n, loc) for (kk, loc), n in table.items() if kk == k), reverse=True)[:top]
            print(f"  -- by {name} line")
            for n, (f, ln) in rows:
                extra = ""
                if name == "outer":
                    extra = "  " + " ".join(f"{op}:{c}" for op, c in opcodes[(k, (f, ln))].most_common(6))
                print(f"  {n:6d}  {100.0 * n / total[k]:5.1f} %  {f}:{ln}{extra}")


if __name__ == "__main__":
    main()
