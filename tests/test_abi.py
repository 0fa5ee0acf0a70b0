"""The C-ABI library loads and exports exactly what include/usrt.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "usrt.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(usrt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from unitysimpleraytracing_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), "libusrt_b200.so does not export %s" % name
    assert sorted(_lib.SIGNATURES) == names, "python SIGNATURES and include/usrt.h disagree"


def test_exports_match_nm():
    from unitysimpleraytracing_b200 import _lib
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = sorted(set(re.findall(r" T (usrt_[a-z0-9_]+)", out)))
    assert exported == declared_symbols()


def test_struct_layouts_match_reference_sizes():
    # MeshBufferContainer.cs:98-106 checks Triangle == 128 and AABB == 32; Constants.cginc gives the rest
    from unitysimpleraytracing_b200 import scene_types as T
    assert T.Triangle.itemsize == 128 and T.AABB.itemsize == 32
    assert T.InternalNode.itemsize == 24 and T.LeafNode.itemsize == 8 and T.RaycastResult.itemsize == 16
    assert T.Triangle.fields["b"][1] == 16 and T.Triangle.fields["c"][1] == 32
    assert T.Triangle.fields["a_uv"][1] == 48 and T.Triangle.fields["a_normal"][1] == 80
    assert T.AABB.fields["max"][1] == 16
    # MAX_FLOAT is the integer literal 0x7F7FFFFF converted to float (Constants.cginc:7)
    assert np.float32(T.MAX_FLOAT).view(np.uint32) == 0x4EFF0000


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "usrt.h"\n_Static_assert(sizeof(usrt_triangle)==128,"t");\n'
                   '_Static_assert(sizeof(usrt_aabb)==32,"a");_Static_assert(sizeof(usrt_internal_node)==24,"i");\n'
                   '_Static_assert(sizeof(usrt_leaf_node)==8,"l");_Static_assert(sizeof(usrt_raycast_result)==16,"r");\n'
                   'int main(void){return 0;}\n')
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from unitysimpleraytracing_b200 import _lib, host
    with pytest.raises(_lib.UsrtError):
        host.Context(1024)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "unitysimpleraytracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "usrt_oracle" not in text and "np_oracle" not in text and "from oracle" not in text, f
