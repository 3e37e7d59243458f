"""CPU tier: the SIMT emulator (tools/simt, test infrastructure) checked on its own -- warp collectives,
barriers, atomics and dynamic shared memory give CUDA's results, and the three error classes it exists to
catch are caught: a collective not every named lane reaches, a barrier not every live thread reaches, a
write outside a device allocation."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT = os.path.join(ROOT, 'tools', 'simt')


@pytest.fixture(scope='module')
def selftest(tmp_path_factory):
    sys.path.insert(0, SIMT)
    import build_emu
    out = tmp_path_factory.mktemp('emu_selftest')
    with open(os.path.join(SIMT, 'selftest', 'selftest.cu')) as f:
        text = build_emu.rewrite(f.read())
    assert '<<<' not in text and 'extern __shared__' not in text
    src = os.path.join(out, 'selftest.cpp')
    with open(src, 'w') as f:
        f.write(text)
    exe = os.path.join(out, 'selftest')
    subprocess.check_call([os.environ.get('CXX', 'g++'), '-std=c++17', '-O1', '-pthread', '-ffp-contract=off',
                           '-Wno-unknown-pragmas', '-I', os.path.join(SIMT, 'emu'), src,
                           os.path.join(SIMT, 'emu', 'simt_engine.cpp'), '-o', exe])
    return exe


def run(exe, mode, threads=None):
    env = dict(os.environ)
    if threads:
        env['SIMT_EMU_THREADS'] = str(threads)
    return subprocess.run([exe, mode], capture_output=True, text=True, timeout=120, env=env)


@pytest.mark.parametrize('threads', [1, 4])
def test_collectives_barriers_atomics(selftest, threads):
    out = run(selftest, 'ops', threads)
    assert out.returncode == 0 and out.stdout.strip() == 'ok', out.stdout + out.stderr


@pytest.mark.parametrize('mode, message', [('divergent', 'DEADLOCK'), ('barrier', 'DEADLOCK'),
                                           ('oob', 'out-of-bounds write ABOVE')])
def test_errors_are_caught(selftest, mode, message):
    out = run(selftest, mode)
    assert out.returncode != 0 and 'not detected' not in out.stdout, out.stdout
    assert message in out.stderr, out.stderr
