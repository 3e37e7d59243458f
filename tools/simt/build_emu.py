"""Build ``libfluxb200_emu.so``: the library's own CUDA sources compiled for the HOST on top of the
SIMT emulator in ``tools/simt/emu`` -- TEST INFRASTRUCTURE, never loaded by the product.

What it is for: running the kernels' real source (LBVH build, trace, CSR fill, queries, SpMV, block
ops) and the real host orchestration of ``fluxb200.cu`` on a machine without a GPU, so that the CPU test
tier can compare them with the oracle and so that a kernel change can be checked for logic and
warp-synchronisation errors before any GPU time is spent.  It says nothing about speed and nothing about
races that depend on the hardware's memory model.

The sources are copied into ``tools/simt/_build/`` with two mechanical rewrites (the product files are
not touched):

* ``kernel<<<grid, block, smem, stream>>>(args)``  ->  ``emu::Launch(grid, block, smem, stream)(kernel, args)``
* ``extern __shared__ T name[];``                  ->  ``T *name = emu::dyn_smem<T>();``

Everything else is handled by the stand-in ``cuda_runtime.h`` (qualifiers, intrinsics, atomics, a
synchronous runtime API) and by the few ``#ifdef FB_EMU`` alternatives to inline PTX in the sources.
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'fluxpy_b200', 'csrc')
BUILD = os.path.join(HERE, '_build')
SO_PATH = os.path.join(BUILD, 'libfluxb200_emu.so')

_NAME = r'([A-Za-z_][\w:]*(?:\s*<[^<>();{}]*>)?)'
_LAUNCH = re.compile(_NAME + r'\s*<<<(.*?)>>>\s*\(', re.S)
_DYN = re.compile(r'extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w:<> ]+?)\s+(\w+)\s*\[\s*\]\s*;')


def rewrite(text):
    def launch(m):
        return f'emu::Launch({m.group(2)})({m.group(1)}, '
    text = _LAUNCH.sub(launch, text)
    text = re.sub(r'(emu::Launch\([^;]*?\)\([^,;()]+), \s*\)', r'\1)', text)   # kernels without arguments
    text = re.sub(r'\b__noinline__\b', 'EMU_NOINLINE', text)
    text = _DYN.sub(lambda m: f'{m.group(1)} *{m.group(2)} = emu::dyn_smem<{m.group(1)}>();', text)
    return text


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.cpp', '.h')))


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in sources()] + \
        [os.path.join(HERE, 'emu', f) for f in os.listdir(os.path.join(HERE, 'emu'))] + \
        [os.path.join(ROOT, 'include', 'fluxb200.h'), os.path.abspath(__file__)]
    defs = os.environ.get('FLUXB200_EMU_DEFS', '').split()   # e.g. "-DFB_KEXPAND=12 -DFB_EXPAND_LEAVES=16": tuning variants
    stamp = os.path.join(BUILD, 'defs.txt')
    if (open(stamp).read() if os.path.exists(stamp) else '') != ' '.join(defs):
        force = True
    if (not force and os.path.exists(SO_PATH)
            and os.path.getmtime(SO_PATH) >= max(os.path.getmtime(d) for d in deps)):
        return SO_PATH
    dst = os.path.join(BUILD, 'fluxpy_b200', 'csrc')
    os.makedirs(dst, exist_ok=True)
    os.makedirs(os.path.join(BUILD, 'include'), exist_ok=True)
    shutil.copy(os.path.join(ROOT, 'include', 'fluxb200.h'), os.path.join(BUILD, 'include', 'fluxb200.h'))
    for f in sources():
        with open(os.path.join(CSRC, f)) as fh:
            text = fh.read()
        if f.endswith(('.cu', '.cuh')):
            text = rewrite(text)
            if '<<<' in text or 'extern __shared__' in text:
                raise RuntimeError(f'{f}: a launch or dynamic shared declaration was not rewritten')
        with open(os.path.join(dst, f), 'w') as fh:
            fh.write(text)
    cxx = os.environ.get('CXX', 'g++')
    flags = ['-std=c++17', '-O2', '-g1', '-fPIC', '-pthread', '-ffp-contract=off', '-fno-fast-math', '-mfma',
             '-fno-strict-aliasing', '-Wall', '-Wno-unknown-pragmas', '-Wno-unused-variable', '-Wno-unused-function',
             '-Wno-unused-but-set-variable', '-I', os.path.join(HERE, 'emu')] + defs
    objs = []
    for src, extra in ((os.path.join(dst, 'fluxb200.cu'), ['-x', 'c++']),
                       (os.path.join(dst, 'host_expand.cpp'), []),
                       (os.path.join(HERE, 'emu', 'simt_engine.cpp'), [])):
        obj = os.path.join(BUILD, os.path.basename(src) + '.o')
        cmd = [cxx] + flags + extra + ['-c', src, '-o', obj]
        out = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or out.returncode:
            print(' '.join(cmd))
            print(out.stdout + out.stderr)
        if out.returncode:
            raise RuntimeError(f'compiling {src} for the SIMT emulator failed')
        objs.append(obj)
    out = subprocess.run([cxx, '-shared', '-o', SO_PATH] + objs + ['-lpthread'], capture_output=True, text=True)
    if out.returncode:
        print(out.stdout + out.stderr)
        raise RuntimeError('linking libfluxb200_emu.so failed')
    with open(stamp, 'w') as fh:
        fh.write(' '.join(defs))
    return SO_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
