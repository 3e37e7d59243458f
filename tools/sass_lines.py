"""Static SASS instruction counts per source line of one kernel (nvdisasm -g on the cubin inside
libfluxb200.so).  For straight-line code executed once per batch the static count IS the per-batch count;
loops show their body size.  No GPU needed.

    python tools/sass_lines.py trace_kernelIfLb0ELb1E [first_line last_line]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    pat = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
    tmp = tempfile.mkdtemp()
    subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'fluxpy_b200', 'libfluxb200.so')],
                          cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.startswith('fluxb200.') and f.endswith('.cubin')][0]
    text = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    counts = collections.Counter()
    inside, cur = False, None
    total = 0
    for line in text.splitlines():
        if line.startswith('//---') and '.text.' in line:
            inside = pat in line
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', line) and cur:
            counts[cur] += 1
            total += 1
    print(f'{pat}: {total} instructions')
    src = {}
    for (f, ln), c in sorted(counts.items()):
        if f not in src:
            path = os.path.join(ROOT, 'fluxpy_b200', 'csrc', f)
            src[f] = open(path).read().splitlines() if os.path.exists(path) else []
        if f == 'assemble.cuh' and not (lo <= ln <= hi):
            continue
        text_line = src[f][ln - 1].strip()[:100] if ln - 1 < len(src[f]) else ''
        print(f'{c:5d}  {f}:{ln:<5d} {text_line}')


if __name__ == '__main__':
    main()
