"""A/B helper for the TMA-staged a-trous kernel (SGI_ATROUS_TMA=1): dumps the planes of one frame so two runs can be compared.
    python tools/debug/atrous_tma_check.py <workload> <out.npz>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from stratusgfx_b200 import _abi
from tests.parity_util import SceneData, make_pass

w, out = sys.argv[1], sys.argv[2]
d = SceneData(w, frames=1)
gp = make_pass(d)
gp.set_gbuffer(d.gb[0], d.gb[-1]); gp.bind()
gp.frame(d.fp[0])
np.savez(out, lit=gp.output("lightingOut"), a=gp.plane(_abi.PLANE_DENOISE_RES_A), b=gp.plane(_abi.PLANE_DENOISE_RES_B))
gp.close()
