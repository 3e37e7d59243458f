import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    if os.environ.get('FLUXB200_TEST_EMU') == '1':
        # Test infrastructure only: run the gpu-marked tests against the library's own CUDA sources
        # compiled for the host on the SIMT emulator (tools/simt).  The package itself has no such
        # switch -- the emulated library is swapped in from here, from the test side.
        sys.path.insert(0, os.path.join(ROOT, 'tools', 'simt'))
        import build_emu
        from fluxpy_b200 import _lib
        _lib.SO_PATH = build_emu.build()
        _lib._lib = None
    zone = os.environ.get('FLUXB200_TEST_HORIZON')
    if zone:
        # Run whatever is selected with the trace kernel's horizon skip switched on for every shape model
        # (value = leaves per near zone): the whole gpu tier then doubles as the check of that variant,
        # on the emulator and on a B200 alike.
        from fluxpy_b200 import shape
        plain_init = shape.CudaTrimeshShapeModel.__init__

        def init_with_horizon(self, *args, **kwargs):
            plain_init(self, *args, **kwargs)
            if int(zone) > 0:
                self.set_option('horizon_zone', int(zone))
            self.set_option('horizon_skip', 1 if int(zone) > 0 else 0)   # 0: the skip (default on) switched off
        shape.CudaTrimeshShapeModel.__init__ = init_with_horizon


    variant = os.environ.get('FLUXB200_TEST_VARIANT')
    if variant:
        # the same for the trace kernel's generation (1: per-lane stacks, 2: warp-shared queue, the default)
        from fluxpy_b200 import shape
        prev_init = shape.CudaTrimeshShapeModel.__init__

        def init_with_variant(self, *args, **kwargs):
            prev_init(self, *args, **kwargs)
            self.set_option('trace_variant', int(variant))
        shape.CudaTrimeshShapeModel.__init__ = init_with_variant


# gpu-tier tests that hand torch CUDA tensors (device pointers) to the library or need NCCL: no emulator run
NEEDS_REAL_DEVICE = ('test_device_resident_radiosity_and_steady_state', 'test_block_extraction_and_lowrank_feed',
                     'test_two_rank_sharded_assembly_and_solve', 'test_ingersoll_device_resident_solver')
# too large for the emulator's ~2 M rays/s (full 50k / 82k / 200k-face matrices)
TOO_LARGE_FOR_EMU = ('test_closed_cratered_body_82k_sampled_rows', 'test_full_50k_matrix_properties',
                     'test_slab_properties_at_full_size', 'test_culling_structures_are_conservative_at_scale')


def pytest_collection_modifyitems(config, items):
    if os.environ.get('FLUXB200_TEST_EMU') != '1':
        return
    for item in items:
        name = item.originalname if hasattr(item, 'originalname') else item.name
        if name in NEEDS_REAL_DEVICE:
            item.add_marker(pytest.mark.skip(reason='needs a real CUDA device (torch tensors / NCCL)'))
        elif name in TOO_LARGE_FOR_EMU and os.environ.get('FLUXB200_TEST_EMU_LARGE') != '1':
            item.add_marker(pytest.mark.skip(reason='too large for the SIMT emulator (FLUXB200_TEST_EMU_LARGE=1 to run)'))


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def digests():
    import json
    with open(os.path.join(GOLDEN, 'digests.json')) as f:
        return json.load(f)
