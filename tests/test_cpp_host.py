"""The compiled-language host: host/cpp/usrt_host.hpp mirrors the reference's C# dispatch classes over the C ABI
(the reference's own host language, C#, has no toolchain here; host/csharp/ holds that source)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from unitysimpleraytracing_b200 import _lib
    exe = str(tmp_path / "drawer_main")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "host", "cpp", "drawer_main.cpp"),
                           "-o", exe, "-L" + libdir, "-lusrt_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_host_compiles_and_links_against_the_abi(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
def test_cpp_host_runs_the_reference_sequence(tmp_path, oracle):
    from unitysimpleraytracing_b200 import meshes
    exe = _build(tmp_path)
    tris = meshes.uniform_soup(30000, seed=77); cam = meshes.SCENE_SOUP_CAMERA
    tris.tofile(tmp_path / "tris.bin")
    w, h = 200, 120
    m = [repr(float(x)) for x in np.asarray(cam["cam_to_world"], np.float32).reshape(16)]
    args = [exe, str(tmp_path / "tris.bin"), str(w), str(h), repr(float(cam["near"])), repr(float(cam["tan_half_fov"]))] + m + [str(tmp_path / "hits.bin")]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(tmp_path / "hits.bin", dtype=oracle.RAYCAST_RESULT)
    want = oracle.Scene(tris).trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], threads=8)
    assert got.tobytes() == want.tobytes()
