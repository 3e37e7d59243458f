"""A stand-in for the ``embree`` Python module (sampotter/python-embree) --
TEST INFRASTRUCTURE ONLY, used in the build container by
``oracle/make_golden.py`` so that the reference's *unmodified*
``EmbreeTrimeshShapeModel`` (src/flux/shape.py:295-421) can run without the
Embree binaries, which are not installable offline.

Only the API surface that shape.py touches is provided (shape.py:307-344,
375-390, 409-419).  ``Scene.intersect1M`` / ``occluded1M`` call the C oracle's
closest-hit restatement of Embree's robust-mode kernels
(``oracle/ff_oracle.c``); everything on the Python side of that call -- ray
construction, masking, hit test -- is then the reference's own code.
"""
import enum
import sys
import types

import numpy as np

from . import oracle as _oracle

INVALID_GEOMETRY_ID = np.uint32(0xFFFFFFFF)

#: last ray batch handed to intersect1M (org, dir), for ray set-up checks
last_rays = {}
#: brute force (False) or BVH (True) closest hit
use_bvh = True


class GeometryType(enum.Enum):
    Triangle = 0


class BufferType(enum.Enum):
    Index = 0
    Vertex = 1


class Format(enum.Enum):
    Uint3 = 0
    Float3 = 1


class SceneFlags(enum.Flag):
    Robust = 4


class BuildQuality(enum.Enum):
    High = 2


class IntersectContextFlags(enum.Flag):
    INCOHERENT = 0
    COHERENT = 1


class IntersectContext:
    def __init__(self):
        self.flags = IntersectContextFlags.INCOHERENT


class _Geometry:
    def __init__(self):
        self.buffers = {}

    def set_new_buffer(self, buf_type, slot, fmt, byte_stride, item_count):
        dt = {Format.Float3: np.float32, Format.Uint3: np.uint32}[fmt]
        assert byte_stride == 3 * np.dtype(dt).itemsize and slot == 0
        buf = np.zeros((item_count, 3), dtype=dt)
        self.buffers[buf_type] = buf
        return buf

    def set_build_quality(self, q):
        pass

    def commit(self):
        pass

    def release(self):
        pass


class _Scene:
    def __init__(self):
        self.flags = None
        self._geom = None
        self._scene = None

    def set_flags(self, flags):
        self.flags = flags

    def set_build_quality(self, q):
        pass

    def attach_geometry(self, geometry):
        self._geom = geometry

    def commit(self):
        assert self.flags == SceneFlags.Robust, 'oracle restates robust mode only'
        self._scene = _oracle.OracleScene(
            self._geom.buffers[BufferType.Vertex].copy(),
            self._geom.buffers[BufferType.Index].copy())

    def intersect1M(self, context, rayhit):
        last_rays['org'] = rayhit.org.copy()
        last_rays['dir'] = rayhit.dir.copy()
        self._scene.intersect1M(rayhit.org, rayhit.dir, rayhit.tnear, rayhit.tfar,
                                rayhit.prim_id, rayhit.geom_id, use_bvh=use_bvh)

    intersectNp = intersect1M

    def occluded1M(self, context, ray):
        self._scene.occluded1M(ray.org, ray.dir, ray.tnear, ray.tfar, use_bvh=use_bvh)


class Device:
    def make_geometry(self, geometry_type):
        assert geometry_type == GeometryType.Triangle
        return _Geometry()

    def make_scene(self):
        return _Scene()


class Ray1M:
    def __init__(self, n):
        self.org = np.zeros((n, 3), np.float32)
        self.dir = np.zeros((n, 3), np.float32)
        self.tnear = np.zeros((n,), np.float32)
        self.tfar = np.zeros((n,), np.float32)
        self.time = np.zeros((n,), np.float32)
        self.mask = np.zeros((n,), np.uint32)
        self.id = np.zeros((n,), np.uint32)
        self.flags = np.zeros((n,), np.uint32)


class RayHit1M(Ray1M):
    def __init__(self, n):
        super().__init__(n)
        self.prim_id = np.zeros((n,), np.uint32)
        self.geom_id = np.zeros((n,), np.uint32)


def install():
    """Register this module as ``embree`` so ``import embree`` in
    src/flux/shape.py:1-4 succeeds."""
    mod = types.ModuleType('embree')
    for k, v in globals().items():
        if not k.startswith('_') or k in ('_Scene', '_Geometry'):
            setattr(mod, k, v)
    mod.__standin__ = sys.modules[__name__]
    sys.modules['embree'] = mod
    return mod
