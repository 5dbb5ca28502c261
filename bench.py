#!/usr/bin/env python
"""bench.py — VPL GI + denoise throughput of the CUDA path (and, with --impl reference, of the CPU
restatement of the reference's shaders) on BASELINE.json's metric.

A "step" is one frame of the hot path over one batch of synthetic input:
    cull + colour + tile lists (stage1+stage2) + stochastic gather + à-trous/temporal denoise.
Metric: Gpixel·VPL/s = W·H·N_vpl / t_frame / 1e9, ms_per_step = ms/frame.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4] [--impl ours|reference]

N > 1 is launched by torchrun, one rank per GPU; the frame is split into horizontal row bands
 — strong scaling of one frame.  Every step rank 0 broadcasts the VPL table and
the frame's refreshed cube maps over NCCL (one packed buffer, issued a frame ahead); each rank runs the pass
on its band.  The reservoir rows a band's à-trous taps read beyond its own rows are, by default, computed by
the band itself (`--bands recompute`, SGI_HALO_RECOMPUTE: the gather and the non-final passes run on 28 / 24 /
16 halo rows as well; same bytes, no exchange inside the frame); `--bands p2p` is the exchange (the
neighbours' kernels store every pass's boundary rows through NVLink peer stores ordered by epoch flags,
csrc/sgi_band.cu).  In both, the final pass's GI rows go to the neighbours (next frame's reprojection) and the
lit rows of every band travel to rank 0's frame image by copy-engine transfers.  After the timed region rank 0
replays the same frames on one whole-frame handle and compares the assembled image with it byte for byte
("image_equals_single_gpu").  `--bands apron` is the plain-handle comparison (30 redundant rows per inner
edge for every pass + an NCCL gather), `--split views` gives every GPU its own camera view.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "vpl_gi_denoise_throughput"
UNIT = "Gpixel*VPL/s"
CPU_SAMPLE_ROWS = 270  # bounded CPU sample: one 1/8-frame band of a 4K frame


def workload_text(s):
    kind = "Cornell box" if s.kind == 0 else "Sponza-scale hall"
    return f"{s.workload}: {s.width}x{s.height} {kind}, {s.num_vpls} VPLs, cube shadow maps 32^2x6, {s.denoise_iters} denoise passes, tile lists on"


def bench_config(scene, world, args):
    """The `config` object of the bench line: what is measured, as a function of the workload and the command line only — so that
    `--impl reference`, launched with the same flags, prints the SAME object (the CPU arm times a bounded sample of this workload; what
    the sample was is in its `cpu_baseline.sample`).  Everything a run measures or decides (visible VPLs, band rows after balancing,
    wait time-outs, the image check) is in the line's `run` object instead."""
    from stratusgfx_b200.bands import APRON, halo_rows, recompute_rows
    iters = scene.denoise_iters
    p2p = world > 1 and args.bands in ("p2p", "recompute")
    recompute = world > 1 and args.bands == "recompute"
    device_maps = args.maps == "device"
    use_early = not args.no_early
    prefetch = world > 1 and not args.no_prefetch
    cfg = {"workload": workload_text(scene), "l2": "inputs_exceed_l2 (per-frame working set > 126 MB)", "storage": "f16/unorm8 planes"}
    if world > 1 and args.split == "views":
        cfg["parallelism"] = f"{world} independent camera views, one whole view per GPU, no exchange (the batch axis of configs[4]); a step = one frame of every view"
        return cfg
    if world == 1:
        cfg["parallelism"] = "single GPU"
    elif p2p:
        halo = recompute_rows(iters) if recompute else halo_rows(iters)
        cfg["parallelism"] = (
            f"{world} row bands (cuts balanced by {args.balance} timed pre-pass round(s); rows in run.band_rows) + {halo}-row halos " +
            ("on which the band recomputes the gather and the non-final passes itself (28/24/16 rows; only the final pass's GI rows travel, NVLink peer stores)" if recompute
             else "filled by the neighbours' kernels through NVLink peer stores (CUDA IPC)") +
            f", one NCCL broadcast of the VPL table{'' if device_maps else ' + 5 cube maps'} per frame ({'issued one frame ahead' if (prefetch or use_early) else 'at the start of the frame'}), "
            "lit rows copied into rank 0's frame image after the final pass (copy engine over NVLink, overlapping the next frame; the last frame's image is awaited inside the timed region)")
    else:
        cfg["parallelism"] = f"{world} row bands + {APRON}-row apron recomputed redundantly, one NCCL broadcast of VPL table + 5 cube maps, NCCL gather of lit bands"
    cfg["look_ahead"] = (("cull + colouring" + (" + table broadcast" if world > 1 else "") + " + cube-map refresh of frame i+1 on a second stream under frame i's denoise passes "
                          "(sgi_early_begin / sgi_cull_early; the reference culls at BE:1395, long before the GI pass at BE:2043)") if use_early else "off: everything on one stream")
    cfg["refreshed_cube_maps"] = ("rendered on the device every frame (sgi_render_vpl_cubemaps, 5 maps)" if device_maps
                                  else ("copied from the per-frame broadcast (5 maps)" if world > 1 else "resident (uploaded once)"))
    return cfg


# ---- clocks during the timed region (B200_PROFILING.md) -----------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons while the timed region runs.  Two samplers run side by side: NVML (nvidia_ml_py) polled
    every ~2 ms from a thread — the timed region of a 1.3-ms frame is short, nvidia-smi's 100-ms loop would see one or two
    samples of it — and `nvidia-smi -lms 100` as the fallback when NVML is unusable.  Any failure degrades to "unavailable";
    it never breaks the bench."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}     # nvmlClocksEventReason*

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.nvml_sm, self.nvml_max, self.nvml_reasons, self._stop, self.nvml_thread = [], None, 0, False, None
        self._begin, self._end = 0, None

    def mark_begin(self):          # the timed region starts: only NVML samples taken between the marks are reported
        self._begin = len(self.nvml_sm)
        self.nvml_reasons = 0

    def mark_end(self):
        self._end = len(self.nvml_sm)

    def _nvml_loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._stop:
                self.nvml_sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.nvml_reasons |= int(get_reasons(h))
                time.sleep(0.002)
        except Exception:
            self.nvml_sm = []          # NVML unusable: the nvidia-smi rows are used instead

    def start(self):
        try:
            self.nvml_thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.nvml_thread.start()
        except Exception:
            self.nvml_thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        self._stop = True
        if self.nvml_thread is not None:
            self.nvml_thread.join(timeout=1.0)
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.nvml_sm:
            inside = self.nvml_sm[self._begin:self._end] or self.nvml_sm
            reasons = sorted(n for n in self.NAMES if self.nvml_reasons & self.BITS[n])
            return {"sm_mhz": float(np.median(inside)), "sm_max_mhz": self.nvml_max, "samples": len(inside), "reasons": reasons,
                    "source": "nvml, 2 ms poll, samples inside the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


# ---- algorithmic bytes per stage, for the roofline -----------------------------------------------
def algorithmic_bytes(P, T, N, V, iters):
    return {
        "cull": 64 * N + 64 * V + 4 * (V + 1),
        "color": (64 + 32 + 8) * V,
        "stage1": (4 + 3) * P + 32 * T,
        "stage2": (32 + 52) * T + 20 * V,
        "gather": 30 * P,
        "denoise_pass": 23 * P,          # non-final pass
        "denoise_final": 76 * P,
        "frame": (4 + 3) * P + 32 * T + (32 + 52) * T + 30 * P + 23 * (iters - 1) * P + 76 * P,
    }


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(stage):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a stage's kernel(s) from the committed
    ncu --set full capture of this command (profiles/ncu_traffic.json, c4 on one GPU)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    v = d.get("stages", {}).get(stage)
    return (int(v), d.get("source")) if v is not None else (None, None)


# ---- CPU arm: the oracle on a bounded band sample -------------------------------------------------------------------
def cpu_sample_run(scene, steps, warmup):
    """Times the CPU restatement (oracle/, kind "port") on rows [y0, y0+rows) of the workload's frame, all host threads."""
    from oracle.oracle_py import OracleFrame, lib
    H = scene.height
    rows = min(CPU_SAMPLE_ROWS, H)
    y0 = ((H - rows) // 2) // 5 * 5
    fp = scene.frame_params(0)
    gb, pgb = scene.gbuffer(0, y0, rows), scene.gbuffer(-1, y0, rows)
    vd, at = scene.vpls()
    sh, df = scene.cubemaps()
    csm = scene.csm(0)
    o = OracleFrame(scene.width, rows, scene.cube_res, scene.csm_res, scene.denoise_iters, y_offset=y0, full_height=H)
    cores = int(lib().orc_num_threads())
    times = []
    for i in range(warmup + steps):
        fp.time = 1000.0 + 16.0 * i
        t = time.perf_counter()
        o.frame(fp, gb, pgb, vd, at, csm, sh, df)
        times.append(time.perf_counter() - t)
    t_step = float(np.mean(times[warmup:]))
    value = scene.width * rows * scene.num_vpls / t_step / 1e9
    # kind "port": oracle/oracle.cpp, which tests/test_reference_glsl.py holds bit for bit against the reference's own shader text
    # compiled with its glm (oracle/make_ref_glsl.py); the compiled text itself is single-threaded (shader globals), so the port is timed
    return dict(value=value, unit=UNIT, cores=cores, kind="port",
                sample=f"rows {y0}..{y0 + rows} of {scene.width}x{H} ({rows}/{H} of the frame), whole VPL set, {steps} step(s) after {warmup} warm-up, mean",
                ms_per_step=t_step * 1e3, V=int(o.visible[0]))


def cpu_reference_text_run(scene, steps, warmup):
    """Times the REFERENCE'S OWN SHADER TEXT (oracle/_ref/libref_glsl.so: the six shaders of the path compiled from their GLSL against
    the reference's glm by oracle/make_ref_glsl.py, kind "reference") on rows [y0, y0+rows) of the workload's frame, fragments and
    work groups spread over all host threads.  The cull and the colouring run on the whole VPL set, the tile stages on the band's
    tile rows, the gather and the four denoise passes on the band's rows of full-height planes (taps that leave the band read
    rows this sample never computed: same work, and the results are not used — correctness of this code against the oracle and the
    CUDA path is the business of tests/test_reference_glsl.py)."""
    from oracle import oracle_py as O
    ref = O.reference_glsl()
    if ref is None:
        raise RuntimeError("oracle/_ref/libref_glsl.so is not there")
    W, H, N, iters = scene.width, scene.height, scene.num_vpls, scene.denoise_iters
    if (W // 5) % 32 or (H // 5) % 2:
        raise RuntimeError("the reference's stage-2 dispatch needs a tile grid of whole 32x2 groups")
    rows = min(CPU_SAMPLE_ROWS, H) // 10 * 10
    y0 = ((H - rows) // 2) // 10 * 10
    y1 = y0 + rows
    fp = scene.frame_params(0)
    gb, pgb = scene.gbuffer(0), scene.gbuffer(-1)
    vd, at = scene.vpls()
    sh, df = scene.cubemaps()
    csm = scene.csm(0)
    cores = int(O.lib().orc_num_threads())
    half = lambda shape: np.zeros(shape, np.float16)
    prev_gi, prev_hist = half((H, W, 4)), half((H, W))
    gi_color, gi_res = half((H, W, 4)), half((H, W, 4))
    gi_color[..., 3] = 1.0
    fbo = [dict(gi=half((H, W, 4)), reservoir=half((H, W, 4))) for _ in range(2)]
    for b in fbo:
        b["gi"][..., 3] = 1.0
    times, V = [], 0
    for i in range(warmup + steps):
        fp.time = 1000.0 + 16.0 * i
        t = time.perf_counter()
        _, vis = O.reference_glsl_cull(ref, fp, vd, csm, scene.csm_res)
        V = int(vis[0])
        idx = np.sort(vis[1:V + 1])
        visible = np.concatenate([[V], idx]).astype(np.int32)
        updated = np.zeros((4480, 16), np.float32); updated[:V] = vd[idx]
        entries = np.zeros((4480, 2), np.int32); entries[:V] = at[idx]                 # the host's shadowDiffuseIndices (BE:1690-1701)
        updated = O.reference_glsl_color(ref, fp, updated, visible, entries, sh, df, scene.cube_res)
        stage1 = O.reference_glsl_stage1(ref, fp, gb, tile_rows=(y0 // 5, y1 // 5))
        O.reference_glsl_stage2(ref, fp, W, H, vd, stage1, visible, tile_rows=(y0 // 5, y1 // 5))
        color, res = O.reference_glsl_gather(ref, fp, gb, updated, visible, entries, sh, scene.cube_res, rows=(y0, y1))
        with np.errstate(over="ignore"):
            gi_color[y0:y1, :, :3] = color[y0:y1]; gi_res[y0:y1] = res[y0:y1]
        illum, shadows = gi_color, gi_res
        for p in range(iters):                                                         # the host loop of BE:1879-1952
            a = O.RefDenoiseArgs()
            a.W, a.H, a.multiplier, a.rowBegin, a.rowEnd = W, H, (1 << p) - 1, y0, y1
            a.mergeReservoirs, a.final_ = (1 if p < 1 else 0), (1 if p + 1 == iters else 0)
            a.time, a.framesPerSecond = float(fp.time), float(np.float32(1.0 / fp.deltaSeconds))
            nv = np.array([V], np.int32)
            combined, gi = np.zeros((H, W, 3), np.float32), np.zeros((H, W, 3), np.float32)
            rsv, hist = np.zeros((H, W, 4), np.float32), np.zeros((H, W), np.float32)
            (a.numVisible, a.screenH4, a.velocityH2, a.normal4, a.ids, a.depth, a.prevNormal4, a.prevIds, a.prevDepth, a.indirectIlluminationH4,
             a.indirectShadowsH4, a.prevIndirectIlluminationH4, a.historyDepthH) = [x.ctypes.data for x in (
                nv, gb.screen, gb.velocity, gb.normal, gb.id, gb.depth, pgb.normal, pgb.id, pgb.depth, illum, shadows, prev_gi, prev_hist)]
            a.combinedOut3, a.giOut3, a.reservoirOut4, a.historyOut = combined.ctypes.data, gi.ctypes.data, rsv.ctypes.data, hist.ctypes.data
            ref.refglsl_denoise_pass(ctypes.addressof(a))
            out = fbo[p % 2]
            with np.errstate(over="ignore"):
                out["gi"][y0:y1, :, :3] = gi[y0:y1]; out["reservoir"][y0:y1] = rsv[y0:y1]
                if p + 1 == iters:
                    prev_hist[y0:y1] = hist[y0:y1]
            illum, shadows = out["gi"], out["reservoir"]
        prev_gi[y0:y1] = illum[y0:y1]
        times.append(time.perf_counter() - t)
    t_step = float(np.mean(times[warmup:]))
    return dict(value=W * rows * N / t_step / 1e9, unit=UNIT, cores=cores, kind="reference",
                sample=f"the reference's own shader text (compiled: oracle/make_ref_glsl.py) on rows {y0}..{y1} of {W}x{H} ({rows}/{H} of the frame), whole VPL set, "
                       f"{steps} step(s) after {warmup} warm-up, mean",
                ms_per_step=t_step * 1e3, V=V)


def cpu_arm(scene, steps, warmup):
    """The CPU arm: the reference's compiled shader text when oracle/_ref/libref_glsl.so travelled with the repository (kind
    "reference"), else the oracle port (kind "port")."""
    try:
        r = cpu_reference_text_run(scene, steps, warmup)
        try:        # for transparency: the hand-written port of the same shaders on the same band (it is ~5x faster than the compiled text)
            r["port_value"] = cpu_sample_run(scene, 1, 1)["value"]
        except Exception:
            r["port_value"] = None
        return r
    except Exception as e:          # never let the baseline break the bench: fall back to the port and say why
        r = cpu_sample_run(scene, steps, warmup)
        r["sample"] += f" (oracle port; the compiled reference shader text was not usable: {type(e).__name__}: {e})"
        return r


def set_host_threads(share=1):
    """torchrun exports OMP_NUM_THREADS=1 to every rank.  The CPU arm is one process on rank 0 and uses every core this process may
    run on (share = 1); the ranks of a multi-GPU run split the cores between them for the procedural scene generation (share =
    world), which is host-side set-up outside every timed region.  Returns the thread count."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = max(1, n // max(1, share))
    os.environ["OMP_NUM_THREADS"] = str(n)          # read by libgomp when the host libraries are loaded
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)      # and for a runtime that is already initialised
    except OSError:
        pass
    return n


def use_all_host_cores():
    return set_host_threads(1)


def run_reference(args, rank):
    if rank != 0:
        return
    use_all_host_cores()
    from stratusgfx_b200.scene import Scene
    scene = Scene(args.workload)
    r = cpu_arm(scene, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(scene, max(args.gpus, 1), args), "run": {"sample": r["sample"], "visible_vpls": r.get("V")},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "port_value") if k in r},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- our arm ----------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from stratusgfx_b200.bands import (APRON, balance_cuts, bands_from_cuts, broadcast_frame_inputs, connect_bands, gather_bands, halo_rows, recompute_rows,
                                       plan_bands)
    from stratusgfx_b200.gipass import VplGiPass
    from stratusgfx_b200.scene import Scene

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        set_host_threads(world)
    scene = Scene(args.workload)
    W, H, N, iters = scene.width, scene.height, scene.num_vpls, scene.denoise_iters
    p2p = world > 1 and args.bands in ("p2p", "recompute")
    recompute = world > 1 and args.bands == "recompute"      # SGI_HALO_RECOMPUTE: no exchange inside the frame (include/stratusgi.h)
    halo = recompute_rows(iters) if recompute else halo_rows(iters)
    isolated = world > 1 and args.bands == "isolated"     # diagnosis: band geometry and halo rows, but no neighbour attached (wrong seams)
    apron = halo if (p2p or isolated) else APRON
    bands = plan_bands(H, world, apron)
    vd, at = scene.vpls()
    sh, df = scene.cubemaps()
    csm = scene.csm(0)
    balance_log = []
    if world > 1 and args.balance > 0:
        # measured load balancing, before the timed region: every rank times the row-dependent stages (tile lists, gather,
        # denoise) on its own rows with a plain handle, the times are all-gathered and the cuts moved to equalise them
        for it in range(args.balance):
            b = bands[rank]
            # recompute mode: the calibration handle is a band handle of the same geometry without peers, so that the rows a band
            # recomputes beyond its inner edges (one edge for the first and the last band, two for the others) are in its time
            c0, c1 = (b.y0, b.y0 + b.rows) if recompute else (b.b0, b.b1)
            cal = VplGiPass(W, H, c0, c1 - c0, scene.cube_res, scene.csm_res, iters, device=local_rank)
            cal.set_gbuffer(scene.gbuffer(0, c0, c1 - c0), scene.gbuffer(-1, c0, c1 - c0))
            cal.set_vpls(vd, at); cal.set_atlases(sh, df); cal.set_csm(csm)
            if recompute:
                cal.band_setup(b.b0, b.b1 - b.b0, b.rank, world, owns_frame_image=False, halo_mode=1)
            cal.bind()
            cfp = scene.frame_params(0)
            cs = torch.cuda.ExternalStream(cal.stream_ptr(), device=dev)
            # the row-dependent cost of a band = its whole frame (tile lists run next to the gather on the pass's side stream,
            # exactly as in the timed region) minus the part every band pays whatever its height (cull + colour)
            evs, fixed = [], []
            for k in range(8):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record(cs)
                cal.cull(cfp); cal.color(cfp)
                e1.record(cs)
                cal.frame(cfp)
                e2.record(cs)
                evs.append((e1, e2)); fixed.append((e0, e1))
            cal.synchronize()
            t_fixed = float(np.median([a.elapsed_time(bb) for a, bb in fixed[2:]]))
            t_mine = max(1e-3, float(np.median([a.elapsed_time(bb) for a, bb in evs[2:]])) - 0.8 * t_fixed)
            cal.close()
            times = [None] * world
            dist.all_gather_object(times, t_mine)
            cuts = balance_cuts(bands, times, H, apron)
            balance_log.append({"rows": [b.b1 - b.b0 for b in bands], "ms": [round(t, 4) for t in times]})
            bands = bands_from_cuts(cuts, H, apron)
    me = bands[rank]
    band, y0, rows = me.b1 - me.b0, me.y0, me.rows

    fp = scene.frame_params(0)
    gb, pgb = scene.gbuffer(0, y0, rows), scene.gbuffer(-1, y0, rows)

    side = torch.cuda.Stream(device=dev)     # everything (kernels, copies, NCCL, events) runs on this one stream
    torch.cuda.synchronize(dev)
    ctx = torch.cuda.stream(side)
    ctx.__enter__()
    gp = VplGiPass(W, H, y0, rows, scene.cube_res, scene.csm_res, iters, device=local_rank, use_torch_stream=True)
    gp.set_gbuffer(gb, pgb)
    gp.set_vpls(vd, at)
    gp.set_atlases(sh, df)
    gp.set_csm(csm)
    gp.bind()

    # per-step NCCL traffic of the band split: VPL table + atlas entries + the <= 5 cube maps refreshed this frame, packed
    # into ONE byte buffer (one broadcast; NCCL has no 16-bit integer type) that the pass reads in place
    refreshed = [int(at[i, 0]) * 320 + int(at[i, 1]) for i in range(min(5, N))]
    device_maps = args.maps == "device"        # the refreshed maps are rendered on every GPU instead of travelling with the tables
    if world > 1:
        R = scene.cube_res
        nb_v, nb_a, nb_m = N * 64, N * 8, (0 if device_maps else len(refreshed) * 6 * R * R * 2)
        packed = torch.empty(nb_v + nb_a + nb_m, dtype=torch.uint8, device=dev)
        v_view = packed[:nb_v].view(torch.float32).view(N, 16)
        a_view = packed[nb_v:nb_v + nb_a].view(torch.int32).view(N, 2)
        maps_buf = packed[nb_v + nb_a:].view(torch.int16).view(-1, 6, R, R)
        v_view.copy_(gp.t["vplData"]); a_view.copy_(gp.t["atlasEntryByVpl"])
        gp.t["vplData"], gp.t["atlasEntryByVpl"] = v_view, a_view
        sa = gp.t["shadowAtlas"].view(-1, 6, R, R)
        if not device_maps:
            maps_buf.copy_(torch.stack([sa[s] for s in refreshed]))
        refreshed_idx = torch.tensor(refreshed, dtype=torch.int64, device=dev)
        gp.bind()
        if p2p:
            connect_bands(dist, gp, me, world, halo_mode=1 if recompute else 0)           # band_setup + CUDA IPC attach of the neighbours and of rank 0's frame image
        elif isolated:
            gp.band_setup(me.b0, me.b1 - me.b0, me.rank, world, owns_frame_image=False)
        else:
            full_out = torch.empty((H, W, 4), dtype=torch.float16, device=dev) if rank == 0 else None

    # What the renderer produces BEFORE the GI pass of a frame — the VPL table (CPU side), the refreshed cube maps and the
    # visibility cull (BE:1982 -> BE:1395, long before BE:2043-2044) — is prepared for frame i+1 on a second, higher-priority
    # stream while frame i runs its denoise passes: the table broadcast lands in place, the maps are refreshed (copied from the
    # broadcast, or rendered on the device with --maps device), then sgi_cull_early culls and colours.  The main stream holds
    # nothing but sgi_frame, which starts at the gather.  --no-early keeps everything on the main stream (round-1 behaviour).
    use_early = not args.no_early
    prefetch = world > 1 and not args.no_prefetch
    pending = [None]
    early = torch.cuda.Stream(device=dev, priority=-1) if use_early else None
    early_ptr = early.cuda_stream if use_early else None
    fp_next = fp.copy()
    if device_maps:
        gp.set_scene_boxes(scene.boxes())
        refreshed_vpls = torch.arange(min(5, N), dtype=torch.int32, device=dev)       # the lights popped from lightsToUpdate this frame
        torch.cuda.synchronize(dev)
    if world > 1:
        src_tables = packed.clone()          # rank 0's "CPU-produced" tables of the next frame (synthetic: the same every frame)
        if not use_early and prefetch:
            stage = torch.empty_like(packed)

    def refresh_maps(stream_ptr):
        if device_maps:
            gp.render_vpl_cubemaps(refreshed_vpls, stream=stream_ptr)
        elif world > 1:
            sa.index_copy_(0, refreshed_idx, maps_buf)        # the refreshed cube maps into their atlas layers: one scatter

    def issue_broadcast(buf):
        if rank == 0:
            buf.copy_(src_tables)
        pending[0] = dist.broadcast(buf, 0, async_op=True)

    def prepare_next(i_next):
        """Everything frame i_next needs before its gather, on the early stream (enqueued right after frame i_next - 1)."""
        with torch.cuda.stream(early):
            gp.early_begin(early_ptr)                          # frame i_next - 1 has read the tables and the atlases
            if world > 1:
                if pending[0] is None:
                    issue_broadcast(packed)
                pending[0].wait()                              # stream-ordered: the tables of frame i_next are in `packed`
            refresh_maps(early_ptr)
            fp_next.time = 1000.0 + 16.0 * i_next
            gp.cull_early(fp_next, early_ptr, True)
            if world > 1:
                issue_broadcast(packed)                        # the tables of frame i_next + 1 travel while frame i_next computes

    def step(i):
        fp.time = 1000.0 + 16.0 * i
        if use_early:
            gp.frame(fp)
            prepare_next(i + 1)
        else:
            if world > 1:
                if prefetch:
                    if pending[0] is None:
                        issue_broadcast(stage)
                    pending[0].wait()                              # stream-ordered: the pass's stream waits for NCCL's
                    packed.copy_(stage)
                    issue_broadcast(stage)                         # the next frame's tables travel while this frame computes
                else:
                    broadcast_frame_inputs(dist, [packed])
            refresh_maps(None)
            gp.frame(fp)
        if world > 1 and not p2p and not isolated:
            gather_bands(dist, gp.t["lightingOut"][me.owned], full_out, bands, rank)
        # p2p: every band's lit rows travel to rank 0's frame image after its final pass (copy engine, during the next frame);
        # rank 0's stream waits for them one frame late, and for the last frame in finish()

    def finish():
        if p2p and rank == 0:
            gp.frame_image_wait()
        if pending[0] is not None:
            if use_early:
                with torch.cuda.stream(early):
                    pending[0].wait()
            else:
                pending[0].wait()
            pending[0] = None
        if use_early:
            side.wait_stream(early)           # the look-ahead work of the frame after the last one is inside the timed region too

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    finish()
    fence()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = gp.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if p2p:
        gp.band_stats()                       # restart the account of in-kernel waits: only the timed region is reported
    fence()
    if rank == 0:
        sampler.mark_begin()
    e0.record()
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps     # host time to enqueue one step (no synchronisation inside)
    finish()                                  # the last frame's image is assembled (and the look-ahead broadcast drained) inside the timed region
    e1.record()
    fence()
    if rank == 0:
        sampler.mark_end()
    ms_total = e0.elapsed_time(e1)
    launches = gp.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = W * H * N / (ms_step * 1e-3) / 1e9
    V = gp.visible_count()
    band_timeouts = gp.band_timeouts() if p2p else 0
    band_spin = None
    if p2p:
        # in-kernel waits of the timed region that found a neighbour's flag behind, per rank: count per frame, summed and longest
        # wait in microseconds (waits of different CTAs overlap, so the sum is an upper bound of the time lost)
        st = gp.band_stats()
        khz = (clocks or {}).get("sm_mhz") or 1965.0
        mine = {"rank": rank, "waits_per_frame": round(st["waits_that_spun"] / args.steps, 2),
                "spin_sum_us_per_frame": round(st["spin_sum_kcycles"] * 1024 / khz / args.steps, 2),
                "spin_max_us": round(st["spin_max_kcycles"] * 1024 / khz, 2), "host_enqueue_ms": round(host_enqueue_ms, 4)}
        band_spin = [None] * world
        dist.all_gather_object(band_spin, mine)

    # ---- the assembled image against one GPU: rank 0 replays the same frames (same `time` sequence, hence the same temporal
    # history) on a whole-frame handle and compares the lit image byte for byte ----
    image_equal = None
    if world > 1 and not args.no_image_check and not isolated:
        if rank == 0:
            set_host_threads(1)               # the other ranks idle at the barrier below
            img = (gp.frame_image() if p2p else full_out).clone()
            one = VplGiPass(W, H, 0, H, scene.cube_res, scene.csm_res, iters, device=local_rank, use_torch_stream=True)
            one.set_gbuffer(scene.gbuffer(0), scene.gbuffer(-1)); one.set_vpls(vd, at); one.set_atlases(sh, df); one.set_csm(csm)
            if device_maps:
                one.set_scene_boxes(scene.boxes())
            one.bind()
            if device_maps:
                one.render_vpl_cubemaps(refreshed_vpls)          # the same five maps, rendered the same way (idempotent)
            fp1 = scene.frame_params(0)
            for i in range(args.warmup + args.steps):
                fp1.time = 1000.0 + 16.0 * i
                one.frame(fp1)
            one.synchronize(); torch.cuda.synchronize(dev)
            ref_img = one.t["lightingOut"]
            a, b = img.view(torch.int16)[..., :3], ref_img.view(torch.int16)[..., :3]
            diff = int((a != b).any(dim=-1).sum().item())
            image_equal = {"equal": diff == 0, "pixels_differing": diff, "frames_replayed": args.warmup + args.steps}
            one.close()
            del img
        fence()

    # ---- per-stage device times (CUDA events on the launching stream), for the roofline ----
    stages, stage_launches = {}, {}
    order = [("cull", lambda: gp.cull(fp)), ("color", lambda: gp.color(fp)), ("tile_lists", lambda: gp.tile_lists(fp)), ("gather", lambda: gp.gather(fp))]
    order += [(f"denoise_pass{i}", (lambda i=i: gp.denoise_pass(fp, i))) for i in range(iters)]
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(order) + 1)] for _ in range(args.steps)]
    fence()
    for k in range(args.steps):
        fp.time = 1000.0 + 16.0 * (args.warmup + args.steps + k)
        evs[k][0].record()
        for j, (name, fn) in enumerate(order):
            n0 = gp.launch_count()
            fn()
            stage_launches[name] = gp.launch_count() - n0
            evs[k][j + 1].record()
    fence()
    for j, (name, _) in enumerate(order):
        stages[name] = float(np.mean([evs[k][j].elapsed_time(evs[k][j + 1]) for k in range(args.steps)]))
    # a stage that launched nothing was fused into an earlier kernel (denoise pass 0 is written by the gather): its
    # event interval is only the record overhead, so it is reported as fused and carries no rate of its own
    fused = {name for name, n in stage_launches.items() if n == 0}
    for name in fused:
        stages[name] = 0.0

    # ---- end to end through the public host-buffer API (stratusgfx_b200/hostpipe.py): every step copies that frame's
    # inputs from pinned host memory to the device, runs the pass and copies the lit image back to the host; three
    # device slots let the copy-in of frame i+1 overlap the pass of frame i (previous-frame planes stay on the device)
    from stratusgfx_b200.hostpipe import HostFramePipeline, pin_frame, pin_maps
    R = scene.cube_res
    pipe = HostFramePipeline(gp)
    host = pin_frame(torch, gb, vd, at, csm)
    first_prev = {k: torch.from_numpy(np.ascontiguousarray(getattr(pgb, k))).pin_memory() for k in ("normal", "id", "depth")}
    maps = pin_maps(torch, refreshed, sh.reshape(-1, 6, R, R)[refreshed], df.reshape(-1, 6, R, R, 4)[refreshed])
    for i in range(3):
        pipe.submit(fp, host, maps, first_prev)
    pipe.result(2)
    fence()
    t0 = time.perf_counter()
    base = pipe.count
    for i in range(args.steps):
        fp.time = 1000.0 + 16.0 * i
        tk = pipe.submit(fp, host, maps)
        if tk - 2 >= base:
            pipe.result(tk - 2)               # the consumer takes frame i-2 while i-1 and i are in flight
    for tk in range(max(base, pipe.count - 2), pipe.count):
        pipe.result(tk)
    fence()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    # ---- what bounds e2e: the host -> device link(s).  Every rank copies exactly its per-frame input set (pinned host memory ->
    # its device slot) five times, all ranks at once, nothing else running: the aggregate rate is the ceiling of `e2e`, and
    # bytes / rate is the time the copies alone need per frame ----
    probe_bytes = sum(v.numel() * v.element_size() for v in host.values())
    probe_stream = torch.cuda.Stream(device=dev)
    fence()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(probe_stream):
        pe0.record()
        for _ in range(5):
            for k, v in host.items():
                pipe.slots[0][k].copy_(v, non_blocking=True)
        pe1.record()
    probe_stream.synchronize()
    probe_ms = pe0.elapsed_time(pe1) / 5.0
    fence()
    probe = {"bytes_per_rank": int(probe_bytes), "ms_per_frame_copy": round(probe_ms, 4), "gbs_per_rank": round(probe_bytes / (probe_ms * 1e-3) / 1e9, 2)}
    if world > 1:
        allp = [None] * world
        dist.all_gather_object(allp, probe)
        probe = {"bytes_per_rank": [q["bytes_per_rank"] for q in allp], "ms_per_frame_copy": [q["ms_per_frame_copy"] for q in allp],
                 "gbs_per_rank": [q["gbs_per_rank"] for q in allp],
                 "aggregate_gbs": round(sum(q["bytes_per_rank"] for q in allp) / (max(q["ms_per_frame_copy"] for q in allp) * 1e-3) / 1e9, 2)}
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    if world > 1:
        t = torch.tensor([float(h2d), float(d2h)], device=dev)
        dist.all_reduce(t)
        h2d, d2h = int(t[0].item()), int(t[1].item())

    ctx.__exit__(None, None, None)
    if rank == 0:
        P, T = W * rows, (W // 5) * (rows // 5)
        ab = algorithmic_bytes(P, T, N, V, iters)
        peak, peak_src = load_peaks()
        per_stage = {}
        for name, ms in stages.items():
            key = {"tile_lists": None, "cull": "cull", "color": "color", "gather": "gather"}.get(name, "denoise_final" if name == f"denoise_pass{iters - 1}" else "denoise_pass")
            b = ab["stage1"] + ab["stage2"] if name == "tile_lists" else ab[key]
            if name in fused:
                per_stage[name] = {"ms": 0.0, "fused_into": "gather", "algorithmic_bytes": int(b),
                                   "note": "no launch of its own: the gather kernel writes this pass's output; its bytes count in frame_algorithmic_bytes only"}
                continue
            per_stage[name] = {"ms": round(ms, 5), "algorithmic_bytes": int(b), "gbs": round(b / (ms * 1e-3) / 1e9, 2), "hbm_frac": round(b / (ms * 1e-3) / 1e9 / peak, 4)}
        dom = max(stages, key=stages.get)
        d = per_stage[dom]
        traffic, traffic_src = load_traffic(dom) if (world == 1 and scene.workload == "c4") else (None, None)
        roofline = {"kernel": dom, "bound": "hbm", "achieved": d["gbs"], "peak": peak, "unit": "GB/s", "frac": d["hbm_frac"], "traffic": traffic,
                    "traffic_source": traffic_src,
                    "peak_source": peak_src, "share_of_step": round(stages[dom] / sum(stages.values()), 4),
                    "frame_algorithmic_bytes": int(ab["frame"]), "frame_hbm_frac": round(ab["frame"] / (ms_step * 1e-3) / 1e9 / peak, 4),
                    "note": "the dominant kernel is FP32-issue/L1 bound, not HBM bound (DESIGN.md section 3); frac is its algorithmic bytes against the measured HBM peak"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": bench_config(scene, world, args),
                "run": {"visible_vpls": V, "band_wait_timeouts": band_timeouts, "image_equals_single_gpu": (image_equal or {}).get("equal"),
                        "image_check": image_equal, "band_spin": band_spin, "band_rows": [b.b1 - b.b0 for b in bands], "band_balancing": balance_log,
                        "host_enqueue_ms_per_step": round(host_enqueue_ms, 4)},
                "image_equals_single_gpu": (image_equal or {}).get("equal"),
                "clocks": clocks,
                "e2e": {"value": W * H * N / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "how": "HostFramePipeline: pinned host -> device copy of the frame's G-buffer/VPL table/CSM/5 cube maps, sgi_frame, lit image -> pinned host, every step; 3 device slots overlap copies with the pass",
                        "h2d_probe": probe,
                        "bound": "host->device link: the probe copies each rank's per-frame input set with nothing else running, all ranks at once; e2e cannot be shorter than ms_per_frame_copy (max over ranks)"},
                "gpu_launches": int(launches),
                "roofline": roofline,
                "stages": per_stage}
        if world == 1 and not args.no_cpu_baseline:
            use_all_host_cores()
            r = cpu_arm(scene, 1, 1)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "port_value") if k in r}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    gp.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()



# ---- second split: independent camera views, one whole view per GPU, no exchange at all ---------------
def run_views(args, rank, world, local_rank):
    """`--split views`: every rank renders its OWN whole view of the workload (a different camera yaw per rank) — the batch
    axis of BASELINE.json's configs[4].  No data-path collective; weak scaling: value = views x pixels x VPLs / time."""
    import torch
    import torch.distributed as dist
    from stratusgfx_b200.gipass import VplGiPass
    from stratusgfx_b200.scene import Scene

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        set_host_threads(world)
    scene = Scene(args.workload, yaw_deg_per_frame=3.0)
    view = rank                                       # frame index = view index: 3 degrees of yaw apart
    W, H, N, iters = scene.width, scene.height, scene.num_vpls, scene.denoise_iters
    vd, at = scene.vpls(); sh, df = scene.cubemaps(); csm = scene.csm(view)
    fp = scene.frame_params(view)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize(dev)
    with torch.cuda.stream(side):
        gp = VplGiPass(W, H, 0, H, scene.cube_res, scene.csm_res, iters, device=local_rank, use_torch_stream=True)
        gp.set_gbuffer(scene.gbuffer(view), scene.gbuffer(view - 1)); gp.set_vpls(vd, at); gp.set_atlases(sh, df); gp.set_csm(csm); gp.bind()

        def fence():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        for i in range(args.warmup):
            fp.time = 1000.0 + 16.0 * i
            gp.frame(fp)
        fence()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start(); time.sleep(0.25)
        launches0 = gp.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fence()
        if rank == 0:
            sampler.mark_begin()
        e0.record()
        for i in range(args.steps):
            fp.time = 1000.0 + 16.0 * (args.warmup + i)
            gp.frame(fp)
        e1.record()
        fence()
        if rank == 0:
            sampler.mark_end()
        ms_total = e0.elapsed_time(e1)
        launches = gp.launch_count() - launches0
        V = gp.visible_count()
        if world > 1:
            t = torch.tensor([ms_total], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    if rank == 0:
        peak, peak_src = load_peaks()
        ab = algorithmic_bytes(W * H, (W // 5) * (H // 5), N, V, iters)
        line = {"metric": METRIC, "value": world * W * H * N / (ms_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": bench_config(scene, world, args), "run": {"visible_vpls_view0": V},
                "clocks": clocks, "e2e": None, "gpu_launches": int(launches),
                "roofline": {"kernel": "frame", "bound": "hbm", "achieved": round(ab["frame"] / (ms_step * 1e-3) / 1e9, 2), "peak": peak, "unit": "GB/s",
                             "frac": round(ab["frame"] / (ms_step * 1e-3) / 1e9 / peak, 4), "traffic": None, "peak_source": peak_src,
                             "note": "whole frame of one view: algorithmic bytes of SURVEY.md 8(d) over the frame time, per GPU"},
                "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    gp.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 100 frames for the CUDA path, a timed region of about 0.1 s on c4; 20 bounded samples for --impl reference)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c4", choices=["tiny", "c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--balance", type=int, default=2, help="measured band-balancing rounds before the timed region (N > 1)")
    ap.add_argument("--no-early", action="store_true", help="keep the table broadcast, the cube-map refresh and the cull on the main stream (no look-ahead stream)")
    ap.add_argument("--maps", default="broadcast", choices=["broadcast", "device"], help="the <= 5 cube maps refreshed per frame: copied from rank 0's broadcast (N > 1; N = 1: resident) or rendered on every GPU by sgi_render_vpl_cubemaps")
    ap.add_argument("--no-prefetch", action="store_true", help="broadcast the per-frame tables at the start of the frame instead of one frame ahead (N > 1)")
    ap.add_argument("--no-image-check", action="store_true", help="N > 1: skip the byte comparison of the assembled image with a single-GPU replay")
    ap.add_argument("--split", default="bands", choices=["bands", "views"], help="N > 1: row bands of one frame (default, strong scaling) or one independent camera view per GPU (weak scaling)")
    ap.add_argument("--bands", default="recompute", choices=["recompute", "p2p", "apron", "isolated"],
                    help="multi-GPU halo strategy (N > 1). recompute (default): SGI_HALO_RECOMPUTE, the bands compute the non-final passes on their halo rows themselves; p2p: SGI_HALO_EXCHANGE, the neighbours' kernels store every pass's boundary rows; apron: plain handles, NCCL gather; isolated: band handles without neighbours, a timing diagnosis whose seams are wrong")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 100 if args.impl == "ours" else 20
    args.steps = max(args.steps, 1)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        if world != max(args.gpus, 1) and world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torchrun (one rank per GPU)")
        if args.split == "views":
            run_views(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
