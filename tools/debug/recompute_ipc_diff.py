"""Debug: three band processes on one GPU in SGI_HALO_RECOMPUTE mode (tests/test_multigpu_nccl.py's worker); prints where the
assembled frames differ from the one-GPU frames."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch.multiprocessing as mp
from tests.test_multigpu_nccl import _p2p_worker, _free_port
from tests.parity_util import SceneData, run_cuda_staged

if __name__ == "__main__":
    world, frames, workload, yaw = int(sys.argv[1]) if len(sys.argv) > 1 else 3, 3, "c1", 0.25
    mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    tmp = tempfile.mkdtemp()
    out_path = os.path.join(tmp, "f%d.npy")
    mp.spawn(_p2p_worker, args=(world, _free_port(), workload, out_path, True, frames, yaw, None, mode), nprocs=world, join=True)
    single = run_cuda_staged(SceneData(workload, frames=frames, yaw_deg_per_frame=yaw))
    for f in range(frames):
        a = np.load(out_path % f).view(np.uint16); b = single[f]["combined"].view(np.uint16)
        bad = np.any(a != b, axis=(1, 2))
        rows = np.nonzero(bad)[0]
        print("frame", f, "differing rows", len(rows), (rows.min(), rows.max()) if len(rows) else None, "pixels", int((a != b).any(axis=2).sum()))
        if len(rows):
            d = (a != b).any(axis=2)
            cols = np.nonzero(d.any(axis=0))[0]
            ys, xs = np.nonzero(d)
            print("   cols", int(cols.min()), int(cols.max()), "first pixels", [(int(y), int(x)) for y, x in zip(ys[:6], xs[:6])])
            af = np.load(out_path % f).astype(np.float32); bf = single[f]["combined"].astype(np.float32)
            print("   max abs diff", float(np.abs(af - bf).max()), "sample", af[ys[0], xs[0]], bf[ys[0], xs[0]])
            print("   x mod 8 histogram", np.bincount(xs % 8, minlength=8).tolist(), "y mod 8", np.bincount(ys % 8, minlength=8).tolist())
            # contiguous runs
            runs = np.split(rows, np.nonzero(np.diff(rows) > 1)[0] + 1)
            print("   runs", [(int(r[0]), int(r[-1])) for r in runs][:12])
