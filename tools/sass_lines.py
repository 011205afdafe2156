#!/usr/bin/env python
"""Instruction count per source line of every kernel in an object file (needs -lineinfo).

    python tools/sass_lines.py path/to/file.o [top]
Used to keep the latency-bound kernels small: they run right after the sweep has flushed L2, so their code is
fetched cold and size is time.
"""
import collections, os, re, subprocess, sys, tempfile

obj = os.path.abspath(sys.argv[1]); top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout
fn = cur = None
cnt = collections.Counter(); tot = collections.Counter()
for line in sass.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\.text\.(\S+):', line)
    if m: fn = m.group(1); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line) and fn:
        tot[fn] += 1
        if cur: cnt[(fn[:48], cur)] += 1
for f, n in tot.most_common(): print(f"{n:7d} instr  {n * 16 / 1024:7.1f} KB  {f}")
for (f, c), n in cnt.most_common(top): print(f"{n:6d}  {f}  {c[0]}:{c[1]}")
