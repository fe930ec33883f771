"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from a_dive_into_ray_tracing_b200 import capi, ctypes_defs as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rt_capi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    capi.build_library()
    lib = C.CDLL(capi.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 18
    assert sorted(capi.SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), s
    assert capi.load_library().rt_version() == 5


def test_struct_sizes_match_header():
    # include/rt_capi.h: plain structs, natural alignment
    assert C.sizeof(D.RtConfig) == 24
    assert C.sizeof(D.RtCamera) == 96
    assert C.sizeof(D.RtStats) == 96
    assert D.SPHERE_DT.itemsize == 48 and D.TRIANGLE_DT.itemsize == 52 and D.QUAD_DT.itemsize == 28
    assert D.MATERIAL_DT.itemsize == 40 and D.BVH_NODE_DT.itemsize == 32
    assert C.sizeof(D.RtSceneDesc) == 4 * 16 + 96 + 32 + 3 * 16 + 2 * 16
    assert D.GROUP_DT.itemsize == 32 and D.INSTANCE_DT.itemsize == 64
    assert D.MEDIUM_DT.itemsize == 56 and D.PERLIN_DT.itemsize == 6144 and C.sizeof(D.RtImage) == 16


def test_struct_sizes_against_the_c_compiler(tmp_path):
    """sizeof() of every struct of the header as gcc sees it == the ctypes / numpy mirrors"""
    import subprocess
    names = ["rt_config", "rt_sphere", "rt_triangle", "rt_quad", "rt_material", "rt_perlin", "rt_image", "rt_medium",
             "rt_camera", "rt_scene_desc", "rt_bvh_node", "rt_stats_t", "rt_group", "rt_instance"]
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "rt_capi.h"\nint main(void) {\n' +
                   "".join('  printf("%%zu\\n", sizeof(%s));\n' % n for n in names) + "  return 0;\n}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(D.RtConfig), D.SPHERE_DT.itemsize, D.TRIANGLE_DT.itemsize, D.QUAD_DT.itemsize,
            D.MATERIAL_DT.itemsize, D.PERLIN_DT.itemsize, C.sizeof(D.RtImage), D.MEDIUM_DT.itemsize,
            C.sizeof(D.RtCamera), C.sizeof(D.RtSceneDesc), D.BVH_NODE_DT.itemsize, C.sizeof(D.RtStats),
            D.GROUP_DT.itemsize, D.INSTANCE_DT.itemsize]
    assert got == want, list(zip(names, got, want))


def test_no_cpu_fallback():
    from tests.conftest import HAS_GPU
    if HAS_GPU:
        pytest.skip("GPU present")
    assert capi.device_count() == 0
    with pytest.raises(capi.RtError) as e:
        capi.Context(profile=0)
    assert e.value.code == 5  # RT_ERR_NODEVICE


def test_product_never_touches_the_oracle():
    """Nothing under the package or include/ may reference oracle/ (the checker)."""
    pkg = os.path.join(ROOT, "a_dive_into_ray_tracing_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "libref_l0" not in txt, f
                assert "libemu" not in txt and "tests.emu" not in txt, f


def test_graft_entry_build_runs():
    """the driver's "does it build" entry point (kept in step with the ABI version)"""
    import __graft_entry__ as g
    g.build()


def test_scene_size_limit_is_rejected_not_wrapped():
    """Node links and strides are signed 32-bit byte offsets (8 orderings x 2n nodes x 32 B): a scene past
    RT_MAX_PRIMS must be refused with RT_ERR_INVALID at flatten time, not wrap around. The count is
    checked before any array is read, so a lying count with a tiny array is enough to exercise it - through the
    host build of the very same flatten_scene() (tests/emu)."""
    from a_dive_into_ray_tracing_b200 import scenes
    from tests.emu import pyemu
    pyemu.build()
    L = C.CDLL(pyemu.PATH)
    L.emu_create.restype = C.c_void_p
    L.emu_create.argtypes = [C.POINTER(D.RtSceneDesc), C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_char_p,
                             C.c_int]
    sc = scenes.weekend(64, 36)
    d = sc.desc()
    err = C.create_string_buffer(256)
    d.n_spheres = (1 << 22)  # > RT_MAX_PRIMS = 2^22 - 1
    assert not L.emu_create(C.byref(d), 0, 1, 0.3, 3, 0, 4, err, 256)
    assert b"too many primitives" in err.value, err.value
