"""The drop-in boundary: wire-format sizes, header <-> ctypes layout, exported symbols, no CPU fallback."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from stratusgfx_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "stratusgi.h")


def test_wire_format_sizes_match_reference_static_asserts():
    # Source/Engine/StratusGpuCommon.h:297-306
    assert C.sizeof(_abi.VplData) == 64
    assert C.sizeof(_abi.VplStage1PerTileOutputs) == 32
    assert C.sizeof(_abi.VplStage2PerTileOutputs) == 52
    assert C.sizeof(_abi.AtlasEntry) == 8
    assert _abi.MAX_TOTAL_VPLS_PER_FRAME == 4480 and _abi.MAX_TOTAL_VPLS_BEFORE_CULLING == 10000 and _abi.MAX_VPLS_PER_TILE == 12


def test_ctypes_layout_matches_header(tmp_path):
    """Compile the header with gcc and compare sizeof/offsetof with the ctypes mirror."""
    probes = {
        "SgiCreateParams": (_abi.CreateParams, ["width", "rows", "denoise_iters", "device", "stream"]),
        "SgiFrameParams": (_abi.FrameParams, ["viewPosition", "time", "deltaSeconds", "minGiOcclusionFactor", "fogColor",
                                              "infiniteLightDirection", "cascadeProjViews", "cascadePlanes", "totalNumLights", "gatherMode"]),
        "SgiBuffers": (_abi.Buffers, ["depth", "prevNormal", "ssao", "lightingOut", "vplData", "atlasEntryByVpl", "shadowCubeMaps",
                                      "diffuseCubeMaps", "infiniteLightShadowMap", "vplCountDevice"]),
        "SgiLightRecord": (_abi.LightRecord, ["position", "radius", "farPlane", "intensity", "atlasIndex", "atlasLayer"]),
        "SgiPointLight": (_abi.PointLight, ["position", "color", "radius", "farPlane"]),
        "SgiDirectParams": (_abi.DirectParams, ["numLights", "numShadowLights", "sunEnabled", "nonShadowCasters", "shadowCasters", "shadowIndices",
                                                "emissionStrength", "infiniteLightDepthBias", "infiniteLightZnear", "infiniteLightZfar",
                                                "pointLightAmbientIntensity", "screenOut"]),
        "SgiSsaoParams": (_abi.SsaoParams, ["aspectRatio", "projPlaneZDist", "intensity", "structureBuffer", "rotationLookup", "occlusionScratch", "ssaoOut"]),
        "SgiTaaParams": (_abi.TaaParams, ["screen", "prevScreen", "velocity", "colorOut", "historyOut"]),
        "SgiPrepassParams": (_abi.PrepassParams, ["viewPosition", "numLights", "frustumPlanes", "lights", "vplDataOut", "atlasEntryOut",
                                                  "orderOut", "countOut"]),
        "SgiVplData": (_abi.VplData, ["color", "specularPosition", "radius", "intensity"]),
        "SgiBandParams": (_abi.BandParams, ["compute_y0", "compute_rows", "rank", "world", "owns_frame_image", "halo_mode"]),
        "SgiPeerInfo": (_abi.PeerInfo, ["mem_handle", "local_base", "off_res", "off_gi", "off_hist", "off_flags", "off_image",
                                        "y_offset", "rows", "compute_y0", "compute_rows", "rank", "device", "pid"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    for cname, (_, fields) in probes.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append('return 0;}')
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).strip().splitlines())
    for cname, (ctype, fields) in probes.items():
        assert int(got[cname]) == C.sizeof(ctype), cname
        for f in fields:
            assert int(got[f"{cname}.{f}"]) == getattr(ctype, f).offset, f"{cname}.{f}"


def _declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"\b(sgi_[a-z_]+)\s*\(", text)))


def test_header_declares_exactly_the_python_symbol_list():
    assert _declared_symbols() == sorted(_abi.ABI_SYMBOLS)


def test_integration_notes_name_every_entry_point():
    """INTEGRATION.md section 1 is the table a maintainer reads: every entry point of the header has a row there."""
    text = open(os.path.join(os.path.dirname(HEADER), "..", "INTEGRATION.md")).read()
    missing = [sym for sym in _declared_symbols() if sym not in text]
    assert not missing, missing


@pytest.fixture(scope="module")
def native_lib():
    from stratusgfx_b200.build import build_native
    return build_native()     # nvcc cross-compiles for sm_100a without a GPU


def test_library_loads_and_exports_every_declared_symbol(native_lib):
    lib = C.CDLL(native_lib)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    lib.sgi_version.restype = C.c_char_p
    assert b"sm_100a" in lib.sgi_version()


def test_library_contains_only_sm100a_code(native_lib):
    out = subprocess.check_output(["cuobjdump", "-lelf", native_lib], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_a_device(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; this checks the no-device behaviour")
    from stratusgfx_b200.gipass import load_library, StratusGiError, VplGiPass
    L = load_library()
    cp = _abi.CreateParams(160, 90, 0, 90, 32, 1024, 4, 0, None)
    h = C.c_void_p()
    assert L.sgi_create(C.byref(cp), C.byref(h)) == _abi.SGI_ERR_CUDA
    assert not h.value
    with pytest.raises(StratusGiError):
        VplGiPass(160, 90)


def test_invalid_arguments_are_rejected(native_lib):
    from stratusgfx_b200.gipass import load_library
    L = load_library()
    h = C.c_void_p()
    for bad in (_abi.CreateParams(0, 90, 0, 90, 32, 1024, 4, 0, None), _abi.CreateParams(160, 90, 80, 20, 32, 1024, 4, 0, None),
                _abi.CreateParams(160, 90, 0, 90, 32, 1024, 0, 0, None)):
        assert L.sgi_create(C.byref(bad), C.byref(h)) == _abi.SGI_ERR_INVALID_ARG
    assert L.sgi_create(None, C.byref(h)) == _abi.SGI_ERR_INVALID_ARG
    assert L.sgi_destroy(None) == _abi.SGI_ERR_INVALID_ARG


def test_missing_library_fails_loudly(monkeypatch):
    import stratusgfx_b200.gipass as g
    monkeypatch.setattr(g, "_lib", None)
    monkeypatch.setattr(g, "_LIB_PATH", "/nonexistent/libstratusgi.so")
    with pytest.raises(g.StratusGiError):
        g.load_library()


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under stratusgfx_b200/ may import, link or dlopen it."""
    pkg = os.path.join(ROOT, "stratusgfx_b200")
    offenders = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(d, f), errors="ignore").read()
                for m in re.finditer(r"^.*\boracle\b.*$", text, re.M):
                    line = m.group(0).strip()
                    if line.startswith(("//", "#", "*", '"""')) or "oracle" not in line.split("//")[0].split("#")[0]:
                        continue
                    if re.search(r"(import|include|CDLL|dlopen|from)\b.*oracle", line):
                        offenders.append(f"{f}: {line}")
    assert not offenders, offenders
