"""The C-ABI shared library loads and exports every symbol include/*.h declares; struct layouts agree with the
ctypes mirrors; without a GPU every compute entry point fails loudly (no CPU fallback).  No GPU needed."""
import ctypes as C
import os
import re

import pytest

import helpers
from miro_b200 import capi

ROOT = helpers.ROOT


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(miro_(?:gpu|host)_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("header", ["miro_gpu.h", "miro_host.h"])
def test_every_declared_symbol_is_exported(header):
    L = capi.lib()
    names = declared_functions(header)
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/{header} but not exported by libmiro_gpu.so"
    listed = set(capi.GPU_SYMBOLS + capi.HOST_SYMBOLS)
    assert set(names) <= listed, set(names) - listed


def test_struct_sizes_match_the_header():
    # sizes stated in include/miro_gpu.h comments; compiled check lives in csrc (static_assert)
    assert C.sizeof(capi.Ray) == 48 and C.sizeof(capi.Hit) == 20
    assert C.sizeof(capi.Node) == 128 and C.sizeof(capi.Tri) == 48 and C.sizeof(capi.MBTri) == 96
    assert C.sizeof(capi.Instance) == 64 and C.sizeof(capi.Prim) == 48
    assert C.sizeof(capi.Material) == 128 and C.sizeof(capi.Light) == 64
    L = capi.lib()
    assert L.miro_gpu_abi_version() == 2
    if hasattr(L, "miro_gpu_sizeof"):
        L.miro_gpu_sizeof.argtypes = [C.c_int]; L.miro_gpu_sizeof.restype = C.c_size_t
        for k, t in enumerate([capi.Ray, capi.Hit, capi.Node, capi.Tri, capi.MBTri, capi.Instance, capi.Prim, capi.Material,
                               capi.Light, capi.Texture, capi.SceneDesc, capi.Camera, capi.RenderParams, capi.Counters]):
            assert L.miro_gpu_sizeof(k) == C.sizeof(t), (k, t)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = capi.lib()
    ctx = C.c_void_p()
    rc = L.miro_gpu_create(C.byref(ctx), 0)
    assert rc == capi.ENODEVICE and not ctx.value
    assert b"no CPU fallback" in L.miro_gpu_last_error(None)
    grp = C.c_void_p()
    ids = (C.c_int * 2)(0, 1)
    assert L.miro_gpu_group_create(C.byref(grp), ids, 2) == capi.ENODEVICE and not grp.value      # several GPUs behind one caller: the same
    fx = helpers.Fixture(helpers.fixture_path("c1_cornell"))
    sc = fx.scene()
    with pytest.raises(Exception) as e:
        sc.attach(0)
    assert "miro_gpu_create" in str(e.value)
    with pytest.raises(Exception):
        sc.render()
    sc.close()
