"""Torch-free check of the C++ host path on a GPU box: host_frame on `tiny` (with and without the device pre-pass) against the CPU
oracle — visible lists exact, lit image within the parity bar.  (tests/test_zz_cpp_host.py is the full test against the Python path.)"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py                                      # noqa: E402
from oracle.oracle_py import OracleFrame                          # noqa: E402
from stratusgfx_b200.prepass import light_records                 # noqa: E402
from tests.parity_util import SceneData, compare_radiance         # noqa: E402
from tests.test_zz_cpp_host import TINY, _read                    # noqa: E402

BIN = os.path.join(ROOT, "stratusgfx_b200", "host", "host_frame")
d = SceneData("tiny", frames=2)
s = d.scene
ok = True
for prepass in (0, 1):
    out = f"/tmp/cpp_host_{prepass}.bin"
    r = subprocess.run([BIN] + TINY + [str(prepass), out], capture_output=True, text=True)
    print(r.stdout.strip(), r.stderr.strip(), "rc", r.returncode)
    if r.returncode != 0:
        ok = False
        continue
    cpp = _read(out, 2, d.W, d.H)
    vd, at = d.vpl_data, d.atlas_by_vpl
    if prepass:
        planes = np.tile(np.array([0, 0, 0, 1], np.float32), (6, 1))
        vd, at, order = oracle_py.vpl_prepass(light_records(d.vpl_data, d.atlas_by_vpl), list(d.fp[0].viewPosition), planes)
    o = OracleFrame(d.W, d.H, s.cube_res, s.csm_res, s.denoise_iters)
    for f in range(2):
        fp = d.fp[f].copy(); fp.totalNumLights = len(vd)
        o.frame(fp, d.gb[f], d.gb[f - 1], np.ascontiguousarray(vd), np.ascontiguousarray(at), d.csm[f], d.shadow_atlas, d.diffuse_atlas)
        V = int(o.visible[0])
        vis_ok = cpp[f][0] == V and np.array_equal(cpp[f][1], o.visible[1:V + 1])
        rep = compare_radiance(cpp[f][2].view(np.float16), o.combined)
        print(f"prepass={prepass} frame={f} V={V} visible_exact={vis_ok} bit_exact={rep['bit_exact']:.5f} psnr={rep['psnr_db']:.1f} ok={rep['ok']}")
        ok = ok and vis_ok and rep["ok"] and rep["bit_exact"] > 0.99
print("CPP HOST", "OK" if ok else "FAIL")
