"""The C++ host side above the C-ABI (stratusgfx_b200/host/stratus_gi_backend.h: the reference's method names over
libstratusgi.so) and its driver program host_frame.cpp.  CPU: it builds, links against the in-tree libraries and fails
loudly without a CUDA device.  GPU: driven from C++ in the reference's call order it produces the bytes of the Python host path."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = ["0", "160", "90", "32", "2", "0x57A7", "4"]          # kind, W, H, VPLs, frames, seed, denoise iterations of the `tiny` workload


@pytest.fixture(scope="module")
def host_bin():
    """The in-tree binary as it travelled with the repository; built only when it (or a library it links) is missing, so that
    a GPU box never recompiles CUDA because of file times."""
    from stratusgfx_b200.build import HOST_BIN, LIB, build_host, build_native
    from stratusgfx_b200.scene import build as build_scene
    build_scene()
    if not os.path.exists(LIB):
        build_native()
    if not os.path.exists(HOST_BIN):
        build_host()
    return HOST_BIN


def _read(path, frames, W, H):
    raw = np.fromfile(path, np.uint8)
    out, off = [], 0
    for _ in range(frames):
        v = int(raw[off:off + 4].view(np.int32)[0]); off += 4
        vis = raw[off:off + 4 * v].view(np.int32).copy(); off += 4 * v
        img = raw[off:off + W * H * 8].view(np.uint16).reshape(H, W, 4).copy(); off += W * H * 8
        out.append((v, vis, img))
    assert off == raw.size
    return out


def test_cpp_host_builds_and_has_no_cpu_path(host_bin, tmp_path):
    from stratusgfx_b200.build import build_host
    assert build_host() == host_bin and os.access(host_bin, os.X_OK)          # up to date with its sources (g++ only)
    src = open(os.path.join(ROOT, "stratusgfx_b200", "host", "stratus_gi_backend.h")).read()
    for name in ("UpdatePointLights_", "PerformVirtualPointLightCullingStage1_", "PerformVirtualPointLightCullingStage2_",
                 "ComputeVirtualPointLightGlobalIllumination_"):          # StratusRendererBackend.h:657-669
        assert name in src
    assert "oracle" not in src and "oracle" not in open(os.path.join(ROOT, "stratusgfx_b200", "host", "host_frame.cpp")).read()
    r = subprocess.run([host_bin, "bad"], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the no-device behaviour cannot be shown here")
    r = subprocess.run([host_bin] + TINY + ["0", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr and not (tmp_path / "o.bin").exists()


@pytest.mark.gpu
def test_cpp_host_reproduces_the_python_host_path(host_bin, tiny, tmp_path):
    from tests.parity_util import run_cuda_staged
    out = str(tmp_path / "cpp.bin")
    r = subprocess.run([host_bin] + TINY + ["0", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "host_frame ok" in r.stdout
    cpp = _read(out, 2, tiny.W, tiny.H)
    py = run_cuda_staged(tiny)
    for f in range(2):
        v, vis, img = cpp[f]
        assert v == py[f]["V"] and np.array_equal(vis, py[f]["visible"][1:v + 1])
        assert np.array_equal(img, py[f]["combined"].view(np.uint16)), f


@pytest.mark.gpu
def test_cpp_host_with_device_prepass_reproduces_the_python_prepass_path(host_bin, tiny, tmp_path):
    from stratusgfx_b200.prepass import light_records
    from tests.parity_util import make_pass
    out = str(tmp_path / "cpp_pp.bin")
    r = subprocess.run([host_bin] + TINY + ["1", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cpp = _read(out, 2, tiny.W, tiny.H)
    planes = np.tile(np.array([0, 0, 0, 1], np.float32), (6, 1))           # host_frame's all-inclusive frustum
    gp = make_pass(tiny)
    upper = gp.set_lights(light_records(tiny.vpl_data, tiny.atlas_by_vpl))
    for f in range(2):
        fp = tiny.fp[f].copy(); fp.totalNumLights = upper
        gp.set_gbuffer(tiny.gb[f], tiny.gb[f - 1]); gp.bind()
        gp.vpl_prepass(list(fp.viewPosition), planes)
        gp.cull(fp); gp.color(fp); gp.tile_lists(fp); gp.gather(fp); gp.denoise(fp)
        v = gp.visible_count()
        assert cpp[f][0] == v and np.array_equal(cpp[f][1], gp.output("vplVisibleIndices")[1:v + 1])
        assert np.array_equal(cpp[f][2], gp.output("lightingOut").view(np.uint16)), f
    gp.close()
