"""Where a kernel's spill / local-memory instructions sit (static SASS, per source line).

    python tools/sass_spills.py trace2_kernelIfE
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass(pat):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'fluxpy_b200', 'libfluxb200.so')],
                          cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.startswith('fluxb200.') and f.endswith('.cubin')][0]
    text = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    inside, cur, out = False, None, []
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
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m:
            out.append((m.group(1), cur, m.group(2)))
    return out


def main():
    out = sass(sys.argv[1])
    print(len(out), 'instructions')
    c = collections.Counter()
    for a, cur, i in out:
        if re.search(r'\b(STL|LDL)', i):
            c[cur] += 1
    src = {}
    for (f, ln), v in sorted(c.items()):
        path = os.path.join(ROOT, 'fluxpy_b200', 'csrc', f)
        if f not in src:
            src[f] = open(path).read().splitlines() if os.path.exists(path) else []
        t = src[f][ln - 1].strip()[:90] if ln - 1 < len(src[f]) else ''
        print(f'{v:3d} {f}:{ln} {t}')
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write('\n'.join(f'{a} {c2[0]}:{c2[1]} {i}' for a, c2, i in out))


if __name__ == '__main__':
    main()
