#!/usr/bin/env python
"""bench.py - headline benchmark of the sliding-window WVM->SVM landmark detector hot path.

Default workload (BASELINE.json north_star / configs[3]): a batch of 256 synthetic 640x480 1-channel frames per GPU through
ALL 15 ffpDetectApp landmark detectors (WVM->SVM five-stage cascades, hq64 features, full pyramids, step 1): every frame is
scanned by every detector, 4 302 040 classified windows per frame. One "step" = one pass over one such batch per GPU.
Metric: classified patches (windows) per second, whole job. The line also carries a `facefrontal` block: BASELINE
configs[1] (256 frames, FaceFrontal cascade alone, HOG SVM stage) measured in the same run, for round-over-round comparison.

  value   frames already resident in HBM; timed with CUDA events on the library's stream
  e2e     the same through the C ABI call a user makes (fdb_detector_set_detect_batch) with HOST (pinned) frames: H2D copy
          of the frames and D2H of the results inside the timed region
  roofline.achieved  algorithmic bytes (W*H + 8 B per window, per frame) of the dominant kernel (the fused
          hq64 + WVM window kernel, all its launches of a step) / its CUDA-event duration
  cpu_baseline       the reference's own classes (oracle/_ref) timed on this box's host cores on a bounded sample

--impl reference times the reference's own CPU implementation (oracle/_ref when built, else the oracle port) on the same
workload with all host cores: (frame, detector) pairs are spread over one process per core; pyramids come from cv2's SIMD
cv::resize / cv::pyrDown (what the reference links), frames are generated outside the timed region.
Multi-GPU (torchrun): frames are sharded over ranks (weak scaling: --frames per GPU; --frames-total T: strong scaling, T
frames split over the ranks), models are replicated, the only exchange is ONE NCCL gather of the detection records at the
end of the job.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WINDOW_BYTES = 8  # dense stage-1 record {f32 fout, i32 level}
W, H = 640, 480
CFG = "FaceFrontal"


# ------------------------------------------------------------------------------------------------
# CPU arms (oracle port / compiled reference) - run in worker processes, one per host core
# ------------------------------------------------------------------------------------------------
_worker_state = {}


def feature_svm_model(syn, feature, det_kw, layers, extract):
    """The SVM of a feature-space cascade: 1024 support vectors blended from feature vectors of seeded
    windows of frames 1000..1003; extract(frame, layer_x_y) -> [n, dim] is the product's extractor (b200 arm) or the
    oracle's (CPU arms) - the two agree bit for bit for the histogram features (tests/test_gpu_features.py)."""
    pw, ph = det_kw["patch_width"], det_kw["patch_height"]
    vecs = []
    for k in range(4):
        lxy = syn.feature_sample_windows(layers, pw, ph, 24, seed=9000 + k)
        vecs.append(extract(syn.synthetic_frame(1000 + k), lxy))
    return syn.make_feature_svm(np.concatenate(vecs), seed=300, num_sv=1024, gamma=syn.FEATURE_GAMMA[feature], center=True)


def cascade_names(workload):
    from featuredetection_b200 import synthetic as syn
    return [CFG] if workload == "facefrontal" else [c[0] for c in syn.LANDMARK_CONFIGS]


def _cpu_worker_init(kind, profile, feature, names):
    """one process per host core (the reference objects are not thread-safe): every worker holds
    the classifiers of every detector of the workload"""
    from featuredetection_b200 import synthetic as syn
    from oracle import fdoracle as fo
    try:
        import cv2
        cv2.setNumThreads(1)
        have_cv2 = True
    except Exception:
        have_cv2 = False
    use_ref = kind == "reference"
    models = {}
    for nm in names:
        det_kw, wvm, svm = syn.landmark_models(nm, profile)
        feat = None
        if feature != "hq64":
            feat = fo.Features(syn.feature_desc(kind=feature), det_kw["patch_width"], det_kw["patch_height"])
            r = fo.detect_frame(det_kw, fo.Wvm(wvm), None, syn.synthetic_frame(0), stage=1, want_dense=False)
            svm = feature_svm_model(syn, feature, det_kw, r["layers"], lambda fr, lxy: feat.extract(det_kw, fr, lxy))
        models[nm] = (det_kw, fo.Wvm(wvm, use_ref=use_ref), fo.Svm(svm, use_ref=use_ref), feat)
    _worker_state.update(models=models, use_ref=use_ref, fo=fo, pyramid_impl="cv2" if (use_ref and have_cv2) else "restated")


def _cpu_worker_item(item):
    """one (frame, detector) pair: Detector::detect of that detector on that frame"""
    frame, name = item
    st = _worker_state
    fo = st["fo"]
    det_kw, wvm, svm, feat = st["models"][name]
    t0 = time.perf_counter()
    if st["use_ref"]:
        r = fo.ref_detect_frame(det_kw, wvm, svm, frame, want_dense=False, svm_features=feat, pyramid_impl=st["pyramid_impl"])
        split = list(r["timing"])
    else:
        r = fo.detect_frame(det_kw, wvm, svm, frame, want_dense=False, svm_features=feat, timing=True)
        split = list(r["timing"])[:5]
    return r["windows"], time.perf_counter() - t0, split, st["pyramid_impl"]


class CpuArm:
    """The reference CPU path on all host cores: (frame, detector) pairs, longest first, over one process per core."""

    SPLIT = ("pyramid", "extract+hq64", "wvm", "overlap_elimination", "svm+nms")

    def __init__(self, workload, profile, feature="hq64"):
        from oracle import fdoracle as fo
        from featuredetection_b200 import synthetic as syn
        fo.build()
        self.kind = "reference" if fo.ref_available() else "port"
        self.cores = os.cpu_count() or 1
        self.names = cascade_names(workload)
        self.syn = syn
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_worker_init,
                                                 initargs=(self.kind, profile, feature, self.names))
        self.pyramid_impl = None

    def frames(self, first, count):
        """generated OUTSIDE the timed region (the GPU arm pre-generates its frames too)"""
        return [self.syn.synthetic_frame(first + k) for k in range(count)]

    def run(self, frames):
        # longest first: windows per frame of each cfg at 640x480, only used to order the work
        cost = {"FaceFrontal": 16, "FaceLeftProfile": 60, "FaceRightProfile": 60, "RightEyeCenter": 158}
        items = [(f, nm) for f in frames for nm in self.names]
        items.sort(key=lambda it: -cost.get(it[1], 365))
        t0 = time.perf_counter()
        res = list(self.pool.imap_unordered(_cpu_worker_item, items, chunksize=1))
        wall = time.perf_counter() - t0
        split = np.sum([r[2] for r in res], axis=0)
        self.pyramid_impl = res[0][3]
        return sum(r[0] for r in res), wall, {k: float(v) for k, v in zip(self.SPLIT, split)}

    def describe(self):
        return ("cv2 %s (SIMD cv::resize / cv::pyrDown, one pyramid per detector and frame as the reference builds them)" % __import__("cv2").__version__
                if self.pyramid_impl == "cv2" else "scalar restatement (oracle/fd_oracle.c)")

    def close(self):
        self.pool.close()
        self.pool.join()


# ------------------------------------------------------------------------------------------------
# supervised-descent regressor (BASELINE configs[4]): --workload sdm
# ------------------------------------------------------------------------------------------------
SDM_L, SDM_STEPS, SDM_BOX = 68, 5, (220, 140, 200, 200)


def _sdm_worker_init(kind):
    from featuredetection_b200 import synthetic as syn
    from oracle import fdoracle as fo
    _worker_state.update(sdm=fo.Sdm(syn.make_sdm(SDM_L, SDM_STEPS, 500), use_ref=(kind == "reference")), syn=syn)


def _sdm_worker_run(frame_ids):
    st = _worker_state
    sdm, syn = st["sdm"], st["syn"]
    frames = {k: syn.synthetic_frame(k) for k in set(i % 8 for i in frame_ids)}
    start = sdm.align_rigid(SDM_BOX)
    t0 = time.perf_counter()
    for i in frame_ids:
        sdm.optimize(frames[i % 8], start)
    return len(frame_ids), time.perf_counter() - t0


class SdmCpuArm:
    """fdo_sdm_optimize per face on all host cores; kind "reference" = descriptors by the reference's own hog.c (oracle/_ref),
    the crop / resize / regressor glue restated (DescriptorExtractor.hpp needs OpenCV)."""

    def __init__(self):
        from oracle import fdoracle as fo
        fo.build()
        self.kind = "reference" if fo.ref_available() else "port"
        self.cores = os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_sdm_worker_init, initargs=(self.kind,))

    def run(self, faces_per_core, first=0):
        chunks = [list(range(first + c * faces_per_core, first + (c + 1) * faces_per_core)) for c in range(self.cores)]
        t0 = time.perf_counter()
        res = self.pool.map(_sdm_worker_run, chunks)
        return sum(r[0] for r in res), time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def sdm_config(args, world, sample=None):
    K = SDM_L * 279
    cfg = {"workload": "BASELINE configs[4]: libSupervisedDescent SDM fit (alignRigid + optimize), %d landmarks, %d cascade steps, "
                       "vlhog-uoctti adaptive descriptors (30x30 patch, 3x3 cells, 279 values per landmark), regressors %d x %d float32 "
                       "~N(0, 1e-3); one face (fixed 200x200 box) per 640x480 1-channel frame, %d faces per GPU per step"
                       % (SDM_L, SDM_STEPS, K + 1, 2 * SDM_L, args.frames),
           "faces_per_gpu": args.frames, "global_faces": args.frames * world, "landmarks": SDM_L, "cascade_steps": SDM_STEPS,
           "parallelism": "face-sharded dp%d" % world,
           "l2": "frames (%.0f MB) + descriptor rows (%.0f MB per step) exceed the 126 MB L2; a 256 MB scratch write also flushes L2 "
                 "between timed steps" % (args.frames * W * H / 1e6, args.frames * K * 4 / 1e6)}
    if sample:
        cfg["sample"] = sample
    return cfg


def run_reference_sdm(args, rank, world):
    if rank != 0:
        return
    arm = SdmCpuArm()
    per_core = max(1, args.ref_frames_per_core * 8)
    for _ in range(args.warmup):
        arm.run(1)
    faces, total = 0, 0.0
    for s in range(args.steps):
        f, wall = arm.run(per_core, first=s * per_core * arm.cores)
        faces += f
        total += wall
    arm.close()
    value = faces / total
    sample = "%d faces per step (%d per core x %d processes), %d steps" % (per_core * arm.cores, per_core, arm.cores, args.steps)
    emit({
        "impl": "reference", "metric": "sdm_fitted_faces_per_s", "value": value, "unit": "faces/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic", "config": sdm_config(args, world, sample),
        "cpu_baseline": {"value": value, "unit": "faces/s", "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def run_sdm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from featuredetection_b200 import synthetic as syn
    from featuredetection_b200.detector import Context, SdmLandmarkModel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    ctx = Context(local_rank)
    sdm = SdmLandmarkModel(ctx, syn.make_sdm(SDM_L, SDM_STEPS, 500))
    n, N = args.frames, 2 * SDM_L
    base = syn.synthetic_frames(rank % 89, 8)
    host_frames = torch.from_numpy(np.concatenate([base] * ((n + 7) // 8))[:n]).pin_memory()
    dev_frames = host_frames.to(device)
    start = np.tile(sdm.align_rigid(np.array([SDM_BOX], np.int32)), (n, 1))
    host_start = torch.from_numpy(start).pin_memory()
    dev_start = host_start.to(device)
    dev_shapes = torch.empty_like(dev_start)
    dev_status = torch.zeros(n, dtype=torch.int32, device=device)
    dev_ff = torch.arange(n, dtype=torch.int32, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    gathered = [torch.empty_like(dev_shapes) for _ in range(world)] if world > 1 else None
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def flush_l2():
        flush.fill_(1)
        torch.cuda.synchronize()

    def step_resident():
        dev_shapes.copy_(dev_start)
        torch.cuda.synchronize()
        ctx.timer_start()
        sdm.optimize_device(dev_frames.data_ptr(), W, H, n, dev_ff.data_ptr(), n, dev_shapes.data_ptr(), dev_status.data_ptr())
        return ctx.timer_stop()

    def step_e2e():
        t0 = time.perf_counter()
        shapes, status = sdm.optimize(host_frames.numpy(), host_start.numpy())
        return 1e3 * (time.perf_counter() - t0), shapes, status

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = ctx.launch_count()
    if rank == 0:
        sampler.start()
    step_ms = []
    for _ in range(args.steps):
        flush_l2()
        step_ms.append(step_resident())
    if world > 1:  # the only exchange: fitted shapes of all ranks (the path itself has no collective)
        dist.all_gather(gathered, dev_shapes)
    barrier()
    launches = ctx.launch_count() - launches0
    prof = []
    for _ in range(max(3, min(args.steps, 5))):
        flush_l2()
        dev_shapes.copy_(dev_start)
        torch.cuda.synchronize()
        prof.append(sdm.profile_device(dev_frames.data_ptr(), W, H, n, dev_ff.data_ptr(), n, dev_shapes.data_ptr(), dev_status.data_ptr()))
    prof = {k: float(np.mean([p[k] for p in prof])) for k in prof[0]}
    for _ in range(2):
        step_e2e()
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        flush_l2()
        ms, shapes_e2e, status_e2e = step_e2e()
        e2e_ms.append(ms)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    tot = torch.tensor([sum(step_ms), sum(e2e_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms, e2e_total = float(tot[0].item()), float(tot[1].item())
    # the resident and the end-to-end fits must agree bit for bit (same kernels, same inputs)
    same = bool(np.array_equal(dev_shapes.cpu().numpy(), shapes_e2e))

    if rank == 0:
        faces_step = n * world
        value = faces_step * args.steps / (total_ms * 1e-3)
        e2e_value = faces_step * args.steps / (e2e_total * 1e-3)
        peak, peak_src = hbm_peak()
        K = SDM_L * 279
        # dominant kernel: sdm_hog_kernel, one launch per cascade step over all faces. Algorithmic bytes per launch: every
        # descriptor window read once as u8 (mean window side over the steps ~ 2 * window half) + the 279 float32 it writes.
        side = 2 * np.array([_sdm_window_half(start[0], s) for s in range(SDM_STEPS)])
        algo_bytes = float(np.mean(side.astype(np.float64) ** 2 + 279 * 4)) * SDM_L * n
        hog_ms = prof["hog"] / SDM_STEPS
        achieved = algo_bytes / (hog_ms * 1e-3) / 1e9
        gemm_flops = 2.0 * n * K * N
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is a rank-0, N=1 datum
            arm = SdmCpuArm()
            arm.run(1)
            per_core = max(8, args.cpu_frames_per_core)
            f, wall = arm.run(per_core)
            arm.close()
            cpu = {"value": f / wall, "unit": "faces/s", "cores": arm.cores, "kind": arm.kind,
                   "sample": "%d faces of the same workload (%d per core x %d processes), %.1f s wall" % (f, per_core, arm.cores, wall)}
        emit({
            "metric": "sdm_fitted_faces_per_s", "value": value, "unit": "faces/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic", "config": sdm_config(args, world), "clocks": clocks,
            "landmark_fits_per_s": value * SDM_L,
            "e2e": {"value": e2e_value, "unit": "faces/s", "h2d_bytes_per_step": int(n * W * H + n * N * 4 + n * 4),
                    "d2h_bytes_per_step": int(n * N * 4 + n * 4), "ms_per_step": e2e_total / args.steps,
                    "identical_to_resident": same},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "sdm_hog_kernel (crop + float32 resize + VLFeat HOG per landmark; one launch per cascade step)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": int(algo_bytes), "kernel_ms": hog_ms,
                         "launches_per_step": SDM_STEPS,
                         "note": "instruction bound (~11.7e3 warp instructions per descriptor: 900 pixels x 9 orientation scores + exact-order cell sums); the regressor "
                                 "product is reported under `gemm`"},
            "gemm": {"kernel": "sdm_gemm_dmma_kernel (FP64 tensor cores: float32 operands, float64 accumulation like cv::gemm on CV_32F)", "m": n, "n": N, "k": K,
                     "ms": prof["gemm"] / SDM_STEPS, "tflops_fp64": gemm_flops / (prof["gemm"] / SDM_STEPS * 1e-3) / 1e12},
            "kernel_ms_per_fit": prof, "faces_out_of_image": int((status_e2e != 0).sum()),
            "cpu_baseline": cpu})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _sdm_window_half(shape, step):
    """window half size of a cascade step for the START shape (bench accounting only; SdmLandmarkModel.hpp:212-229)"""
    L = SDM_L
    a1 = np.array([(shape[8] + shape[9]) / 2, (shape[8 + L] + shape[9 + L]) / 2])
    a2 = np.array([(shape[11] + shape[12]) / 2, (shape[11 + L] + shape[12 + L]) / 2])
    wsh = int(round(float(np.linalg.norm(a1 - a2)) / 4 * (1 / (1 + np.exp((step + 1) - SDM_STEPS)))))
    return wsh + 3 - wsh % 3


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _loop_nvml(self):
        """NVML polls in well under a millisecond, so even a 40 ms timed region gets several samples; any failure
        (module missing, driver call refused) drops back to the nvidia-smi loop"""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        mx = str(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self.stop_flag:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = int(get_reasons(h))
            self.samples.append([str(sm), mx] + ["Active" if r & b else "Not Active" for _, b in bits])
            time.sleep(0.005)

    def _loop(self):
        try:
            self._loop_nvml()
            return
        except Exception:
            pass
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=10)
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            if len(s) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
WINDOWS_PER_FRAME = {"facefrontal": 16185, "landmarks15": 4302040}  # SURVEY.md section 8 (checked against the plan at run time)


def run_reference(args, rank, world):
    """the reference's own CPU implementation of the workload on all host cores of this box (rank 0 only)"""
    if rank != 0:
        return
    arm = CpuArm(args.workload, args.profile, args.feature)
    nfr = args.ref_frames_per_step
    warm = arm.frames(0, 1)
    for _ in range(max(args.warmup, 1)):
        arm.run(warm)
    times, windows, split = [], 0, None
    for s in range(args.steps):
        frames = arm.frames(100 + s * nfr, nfr)   # outside the timed region
        w, wall, sp = arm.run(frames)
        windows += w
        times.append(wall)
        split = sp if split is None else {k: split[k] + sp[k] for k in sp}
    pyr = arm.describe()
    arm.close()
    total = sum(times)
    value = windows / total
    wpf = WINDOWS_PER_FRAME[args.workload]
    sample = "%d frames x %d detectors per step = %d (frame, detector) items over %d processes, %d steps" % (
        nfr, len(arm.names), nfr * len(arm.names), arm.cores, args.steps)
    line = {
        "impl": "reference", "metric": "classified_patches_per_s", "value": value, "unit": "patches/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.frames_total else "weak", "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": arm.cores, "kind": arm.kind, "sample": sample,
                         "pyramid": pyr, "split_core_seconds": split},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s": value / wpf,
    }
    emit(line)


def workload_config(args, world, extra=None):
    frames = args.frames
    if args.workload == "facefrontal":
        fdesc = ("hq64 u8 features for both stages, as ffpDetectApp wires it" if args.feature == "hq64" else
                 "WVM on hq64 patches, RBF-SVM on %s features (adaptiveTrackingApp/default.cfg parameters: 9 unsigned bins, "
                 "interpolated binning, cell 5, block 1, l2norm -> 144-d float32)" % args.feature.upper() if args.feature == "hog" else
                 "WVM on hq64 patches, RBF-SVM on %s features" % args.feature.upper())
        wl = ("BASELINE configs[1]: %d-frame batch per GPU, 640x480 1-channel, FaceFrontal WVM->SVM five-stage cascade "
              "(%s), full pyramid, step 1x1" % (frames, fdesc))
        wpf, pyr = 16185, 1931000
    else:
        wl = ("BASELINE north_star / configs[3]: %d-frame batch per GPU, 640x480 1-channel, all 15 ffpDetectApp landmark detectors on every "
              "frame (WVM->SVM five-stage cascades, hq64 u8 features for both stages as ffpDetectApp wires them, full pyramids, step 1x1)" % frames)
        wpf, pyr = 4302040, 1110000
    cfg = {"workload": wl, "frames_per_gpu": frames, "global_frames": frames * world, "windows_per_frame": wpf,
           "threshold_profile": args.profile, "parallelism": "frame-sharded dp%d" % world,
           "l2": "per-step working set (frames + materialised pyramids + dense records = %.0f MB) exceeds the 126 MB L2; a 256 MB "
                 "scratch write also flushes L2 between timed steps" % ((W * H + pyr + 8 * wpf) * frames / 1e6)}
    if extra:
        cfg["sample"] = extra
    return cfg


# ------------------------------------------------------------------------------------------------
# `single` psvm detector (ffpDetectApp.cpp:427-500; SURVEY 8(d) "every window goes to the SVM"): --workload single-psvm
# ------------------------------------------------------------------------------------------------
SINGLE_STRIDE = 16  # the CPU arm classifies every 16th window of its frames (215 us per window and core)


def _single_models():
    """FaceFrontal geometry + its 1024-support-vector u8 RBF SVM (threshold 0; run_single raises it to the 99.9 % quantile of
    frame 0's distances so that ~0.1 % of the windows are positive - the CPU arm times classification, not thresholding)"""
    from featuredetection_b200 import synthetic as syn
    det_kw, _, svm = syn.landmark_models(CFG)
    return det_kw, svm


def _single_worker_init(kind):
    from featuredetection_b200 import synthetic as syn
    from oracle import fdoracle as fo
    det_kw, svm = _single_models()
    _worker_state.update(fo=fo, syn=syn, kw=det_kw, svm=fo.Svm(svm, use_ref=(kind == "reference")))


def _single_worker_run(frame_ids):
    """timed: HistEq64 + SVM distance of the sampled windows (the per-window work of SlidingWindowDetector::detect);
    untimed: the pyramid and slicing the sampled windows out of it (python)"""
    st = _worker_state
    fo, syn, kw = st["fo"], st["syn"], st["kw"]
    n, t = 0, 0.0
    for k in frame_ids:
        frame = syn.synthetic_frame(k % 8)
        _, layers = fo.pyramid(frame, kw["incremental_scale_factor"], kw["min_scale_factor"], kw["max_scale_factor"])
        raw, w0 = [], 0
        for _, _, img in layers:
            wx, wy = img.shape[1] - 20, img.shape[0] - 20  # strict < loop bounds (DirectPyramidFeatureExtractor.cpp:101-103)
            for w in range((-w0) % SINGLE_STRIDE, wx * wy, SINGLE_STRIDE):
                y, x = divmod(w, wx)
                raw.append(img[y:y + 20, x:x + 20])
            w0 += wx * wy
        t0 = time.perf_counter()
        eq = np.stack([fo.hq64(p).ravel() for p in raw])
        st["svm"].eval(eq)
        t += time.perf_counter() - t0
        n += len(raw)
    return n, t


class SingleCpuArm:
    """HistEq64 + SvmClassifier::computeHyperplaneDistance per window on all host cores (kind "reference": the compiled
    reference classes of oracle/_ref), on every SINGLE_STRIDE-th window of the frames"""

    def __init__(self):
        from oracle import fdoracle as fo
        fo.build()
        self.kind = "reference" if fo.ref_available() else "port"
        self.cores = os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_single_worker_init, initargs=(self.kind,))

    def run(self, frames_per_core, first=0):
        chunks = [list(range(first + c * frames_per_core, first + (c + 1) * frames_per_core)) for c in range(self.cores)]
        res = self.pool.map(_single_worker_run, chunks)
        return sum(r[0] for r in res), max(r[1] for r in res)  # the slowest core's classification time

    def close(self):
        self.pool.close()
        self.pool.join()


def single_config(args, world, wpf, extra=None):
    cfg = {"workload": "ffpDetectApp `single` detector with a psvm classifier (SURVEY 8(d) no-exit case of BASELINE configs[3]: every window "
                       "goes to the RBF-SVM): %d-frame batch per GPU, 640x480 1-channel, FaceFrontal pyramid (13 layers), HistEq64 20x20 "
                       "patches, 1024 u8 support vectors" % args.frames,
           "frames_per_gpu": args.frames, "global_frames": args.frames * world, "windows_per_frame": wpf, "support_vectors": 1024,
           "parallelism": "frame-sharded dp%d" % world,
           "l2": "a 256 MB scratch write flushes L2 between timed steps"}
    if extra:
        cfg["sample"] = extra
    return cfg


def run_reference_single(args, rank, world):
    if rank != 0:
        return
    arm = SingleCpuArm()
    for _ in range(args.warmup):
        arm.run(1)
    windows, total = 0, 0.0
    for s in range(args.steps):
        w, wall = arm.run(1, first=s * arm.cores)
        windows += w
        total += wall
    arm.close()
    value = windows / total
    sample = "every %dth window of %d frames per step (1 per core x %d processes), %d steps" % (SINGLE_STRIDE, arm.cores, arm.cores, args.steps)
    emit({"impl": "reference", "metric": "classified_patches_per_s", "value": value, "unit": "patches/s", "n_gpus": args.gpus,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "u8/s32/f64", "data": "synthetic",
          "config": single_config(args, world, 16185, sample),
          "cpu_baseline": {"value": value, "unit": "patches/s", "cores": arm.cores, "kind": arm.kind, "sample": sample},
          "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def run_single(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from featuredetection_b200 import synthetic as syn, sharding
    from featuredetection_b200.detector import Context, SlidingWindowCascade, DetectorSet, DETECTION_DTYPE

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    ctx = Context(local_rank)
    det_kw, svm = _single_models()
    n = args.frames
    probe = SlidingWindowCascade(ctx, det_kw, None, svm)
    probe.prepare(W, H, 1)
    _, d0 = probe.detect_single(syn.synthetic_frames(0, 1))
    svm.threshold = float(np.float32(np.quantile(d0, 0.999)))   # ~0.1 % of the windows positive
    del probe
    casc = SlidingWindowCascade(ctx, det_kw, None, svm)
    casc.prepare(W, H, n)
    if not casc.single_dense:
        raise SystemExit("bench.py: the tensor-core SVM path is not available for this model")
    wpf = casc.windows_per_frame
    lo, hi = sharding.shard_range(n * world, rank, world)
    base = syn.synthetic_frames(lo % 97, min(8, n))
    host_frames = torch.from_numpy(np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]).pin_memory()
    dev_frames = host_frames.to(device)
    dev_dist = torch.empty((n, wpf), dtype=torch.float64, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    det_cap = 64 * n
    gather_cap = max(256, 32 * n)
    gather = sharding.DetectionGather(gather_cap, dist, device) if world > 1 else None
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def flush_l2():
        flush.fill_(1)
        torch.cuda.synchronize()

    def exchange(dets):
        if world > 1:  # the only exchange: fixed-size blocks of positives (the path itself has no collective)
            gather.collect(gather.submit(dets[:gather_cap], lo))

    def step_resident():
        dets = casc.detect_single_device(dev_frames.data_ptr(), n, dev_dist.data_ptr(), det_cap=det_cap)
        exchange(dets)
        return dets

    def step_e2e():
        dets, _ = casc.detect_single(host_frames.numpy(), want_distances=False, det_cap=det_cap)
        exchange(dets)
        return dets

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = ctx.launch_count()
    if rank == 0:
        sampler.start()
    step_ms, kern_ms, kern_n = [], 0.0, 0
    for _ in range(args.steps):
        flush_l2()
        ctx.timer_start()
        dets = step_resident()
        step_ms.append(ctx.timer_stop())
        ms, k = casc.single_dense_profile()
        kern_ms += ms
        kern_n += k
    barrier()
    launches = ctx.launch_count() - launches0
    for _ in range(2):
        step_e2e()
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        flush_l2()
        ctx.timer_start()
        dets_e2e = step_e2e()
        e2e_ms.append(ctx.timer_stop())
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    tot = torch.tensor([sum(step_ms), sum(e2e_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms, e2e_total = float(tot[0].item()), float(tot[1].item())
    same = bool(np.array_equal(dets["window"], dets_e2e["window"]) and np.array_equal(dets["svm_distance"], dets_e2e["svm_distance"]))

    if rank == 0:
        windows_step = wpf * n * world
        value = windows_step * args.steps / (total_ms * 1e-3)
        e2e_value = windows_step * args.steps / (e2e_total * 1e-3)
        # dominant kernel: svm_dense_kernel, one launch per chunk of frames. Algorithmic work per launch:
        # 2 * windows * support vectors * patch pixels integer operations of the u8 matrix product.
        kernel_ms = kern_ms / kern_n
        ops = 2.0 * wpf * (n * args.steps / kern_n) * svm.sv.shape[0] * svm.sv.shape[1]
        achieved = ops / (kernel_ms * 1e-3) / 1e12
        mp_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        bf16 = json.load(open(mp_path))["bf16_tflops"] if os.path.exists(mp_path) else 1638.6  # burst: the kernel is timed alone
        peak = 2.0 * bf16  # kind::i8 runs at twice the bf16 rate on sm_100a; no measured int8 figure exists on this pool
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is a rank-0, N=1 datum
            arm = SingleCpuArm()
            arm.run(1)
            wcpu, wall = arm.run(8)
            arm.close()
            cpu = {"value": wcpu / wall, "unit": "patches/s", "cores": arm.cores, "kind": arm.kind,
                   "sample": "every %dth window of %d frames (8 per core x %d processes), %.1f s on the slowest core" % (SINGLE_STRIDE, 8 * arm.cores, arm.cores, wall)}
        emit({
            "metric": "classified_patches_per_s", "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/s32/f64", "data": "synthetic", "config": single_config(args, world, wpf),
            "frames_per_s": value / wpf, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": int(W * H * n),
                    "d2h_bytes_per_step": int(len(dets_e2e) * 16 + 4 * kern_n // args.steps), "ms_per_step": e2e_total / args.steps,
                    "identical_to_resident": same},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "svm_dense_kernel (HistEq64 producers + tcgen05.mma kind::i8 [windows x 400] . [400 x 1024] + float64 "
                                                      "RBF epilogue; one launch per %d-frame chunk)" % (n * args.steps // kern_n),
                         "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": "2 x bf16_tflops (burst) of MEASURED_PEAKS.json: kind::i8 runs at twice the bf16 rate; no int8 figure is measured on this pool (ncu: tensor pipe 13 % active)",
                         "algorithmic_ops_per_launch": ops, "kernel_ms": kernel_ms, "launches_per_step": kern_n / args.steps,
                         "note": "the tensor pipe is ~13 % busy (profiles/svmd_*): the launch is bound by the float64 exp epilogue "
                                 "(7 float64 operations per window x support vector) and by streaming the support vectors from L2 "
                                 "(426 KB per 128-window tile)"},
            "detections_per_step": int(len(dets)) * world, "cpu_baseline": cpu})
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private copy of fd 1 for it and point fd 1 at stderr, so that whatever a
    library prints (NCCL's version banner goes to stdout whatever NCCL_DEBUG_FILE says) cannot end up in front of it"""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    _claim_stdout()
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


# ------------------------------------------------------------------------------------------------
# detection::AggregatedFeaturesDetector: --workload aggdet
# ------------------------------------------------------------------------------------------------
AGG = dict(cell=4, window=(10, 10), octave_layer_count=5, unsigned_bins=9, bias=0.0, quantile=0.999)


def _aggdet_model():
    """linear SVM of a FHOG face-sized window (10 x 10 cells of 4 px, 31 features per cell), seeded; the threshold lets
    0.1 % of the windows of frame 0 through (a detector at its operating point scores few windows above the threshold)"""
    rng = np.random.default_rng(77)
    kh, kw = AGG["window"]
    return rng.normal(0, 0.05, (kh, kw, 3 * AGG["unsigned_bins"] + 4)).astype(np.float32)


def _aggdet_cpu_init():
    from oracle import fdoracle as fo
    fo.build()
    _worker_state.update(fo=fo, w=_aggdet_model())


def _aggdet_cpu_item(item):
    frame, thr = item
    fo = _worker_state["fo"]
    t0 = time.perf_counter()
    _, _, maps = fo.aggregated_features_detect(frame, _worker_state["w"], AGG["bias"], thr, cell=AGG["cell"], octave_layer_count=AGG["octave_layer_count"],
                                               want_scores=True)
    return sum(m.size for m in maps), time.perf_counter() - t0


def run_aggdet(args, rank, world, local_rank):
    from featuredetection_b200 import synthetic as syn
    n = args.frames
    w = _aggdet_model()
    cfg = {"workload": "detection::AggregatedFeaturesDetector (SURVEY 8(f) rank 2): %d-frame batch per GPU, 640x480 1-channel, GrayscaleFilter + FhogFilter "
                       "(cell %d, 9 unsigned bins, interpolated cells), %dx%d-cell linear-SVM window, %d pyramid layers per octave, IoU suppression 0.3"
                       % (n, AGG["cell"], AGG["window"][1], AGG["window"][0], AGG["octave_layer_count"]),
           "frames_per_gpu": n, "global_frames": n * world, "parallelism": "frame-sharded dp%d" % world,
           "l2": "a 256 MB scratch write flushes L2 between timed steps"}
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import fdoracle as fo
        fo.build()
        cores = os.cpu_count() or 1
        pool = mp.get_context("spawn").Pool(cores, initializer=_aggdet_cpu_init)
        frames = [syn.synthetic_frame(100 + k) for k in range(cores * 2)]
        pool.map(_aggdet_cpu_item, [(frames[0], 1e9)] * cores)
        tot_w, tot_t = 0, 0.0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            res = pool.map(_aggdet_cpu_item, [(f, 1e9) for f in frames], chunksize=1)
            tot_t += time.perf_counter() - t0
            tot_w += sum(r[0] for r in res)
        pool.close(); pool.join()
        value = tot_w / tot_t
        sample = "%d frames per step over %d processes (oracle restatement: C FHOG + numpy correlation), %d steps" % (len(frames), cores, args.steps)
        emit({"impl": "reference", "metric": "scored_windows_per_s", "value": value, "unit": "windows/s", "n_gpus": args.gpus, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "u8/f32", "data": "synthetic", "config": cfg,
              "cpu_baseline": {"value": value, "unit": "windows/s", "cores": cores, "kind": "port", "sample": sample},
              "e2e": {"value": value, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    import torch
    import torch.distributed as dist
    from featuredetection_b200.detector import Context, AggregatedFeaturesDetector
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    ctx = Context(local_rank)
    base = syn.synthetic_frames((rank * n) % 97, min(8, n))
    probe = AggregatedFeaturesDetector(ctx, w, AGG["bias"], 1e9, cell=AGG["cell"], octave_layer_count=AGG["octave_layer_count"])
    probe.prepare(W, H, 1)
    allsc = np.concatenate([sc for _, _, sc in probe.score_maps(base[0])])
    thr = float(np.quantile(allsc, AGG["quantile"]))
    det = AggregatedFeaturesDetector(ctx, w, AGG["bias"], thr, cell=AGG["cell"], octave_layer_count=AGG["octave_layer_count"])
    det.prepare(W, H, n)
    layers = det.layers()
    npos = det.positions_per_frame
    host_frames = torch.from_numpy(np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]).pin_memory()
    dev_frames = host_frames.to(device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); ctx.synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        l0 = ctx.launch_count()
        ms, out = [], None
        for _ in range(args.steps):
            flush.fill_(1); torch.cuda.synchronize()
            ctx.timer_start()
            out = fn()
            ms.append(ctx.timer_stop())
        barrier()
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), out, ctx.launch_count() - l0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    tot, out, launches = timed(lambda: det.detect_device(dev_frames.data_ptr(), n))
    prof = [det.profile_device(dev_frames.data_ptr(), n) for _ in range(3)]
    e2e_tot, out_e2e, _ = timed(lambda: det.detect(host_frames.numpy()))
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        value = npos * n * world * args.steps / (tot * 1e-3)
        e2e = npos * n * world * args.steps / (e2e_tot * 1e-3)
        pk = {k: float(np.mean([p[k] for p in prof])) for k in prof[0]}
        peak, peak_src = hbm_peak()
        # FHOG (histogram + descriptor kernels): every layer pixel read once, every feature written once (31 x 4 bytes per cell)
        algo = (sum(l["width"] * l["height"] for l in layers) + sum(l["cells_x"] * l["cells_y"] for l in layers) * 31 * 4) * n
        achieved = algo / ((pk["histograms"] + pk["descriptors"]) * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic("aggdet_hist_kernel")
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            pool = mp.get_context("spawn").Pool(cores, initializer=_aggdet_cpu_init)
            frames = [syn.synthetic_frame(100 + k) for k in range(cores * 2)]
            pool.map(_aggdet_cpu_item, [(frames[0], 1e9)] * cores)
            t0 = time.perf_counter()
            res = pool.map(_aggdet_cpu_item, [(f, 1e9) for f in frames], chunksize=1)
            wall = time.perf_counter() - t0
            pool.close(); pool.join()
            cpu = {"value": sum(r[0] for r in res) / wall, "unit": "windows/s", "cores": cores, "kind": "port",
                   "sample": "%d frames over %d processes (oracle restatement: C FHOG + numpy correlation), %.1f s wall" % (len(frames), cores, wall)}
        emit({"metric": "scored_windows_per_s", "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
              "config": dict(cfg, windows_per_frame=npos, pyramid_layers=len(layers)),
              "frames_per_s": value / npos, "clocks": clocks,
              "e2e": {"value": e2e, "unit": "windows/s", "h2d_bytes_per_step": int(W * H * n), "d2h_bytes_per_step": int(len(out_e2e[1]) * 24 + 4),
                      "ms_per_step": e2e_tot / args.steps, "frames_per_s": e2e / npos},
              "gpu_launches": int(launches),
              "roofline": {"bound": "hbm", "kernel": "aggdet_hist_kernel + aggdet_desc_kernel (FHOG feature maps of every pyramid layer)", "achieved": achieved,
                           "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                           "algorithmic_bytes_per_step": int(algo), "kernel_ms_per_step": pk["histograms"] + pk["descriptors"]},
              "kernel_ms": pk, "detections_per_step": int(len(out[1])) * world, "cpu_baseline": cpu})
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# `single` detector in a feature space (ffpDetectApp.cpp:445-454 -> FilteringPyramidFeatureExtractor: the patch filter chain runs
# on EVERY window, then the psvm): --workload single-whi / single-hog (BASELINE configs[2] names WHI features)
# ------------------------------------------------------------------------------------------------
def _single_feature_models(feature):
    from featuredetection_b200 import synthetic as syn
    from oracle import fdoracle as fo
    det_kw, wvm, _ = syn.landmark_models(CFG)
    feat = fo.Features(syn.feature_desc(kind=feature), det_kw["patch_width"], det_kw["patch_height"])
    r = fo.detect_frame(det_kw, fo.Wvm(wvm), None, syn.synthetic_frame(0), stage=1, want_dense=False)
    svm = feature_svm_model(syn, feature, det_kw, r["layers"], lambda fr, lxy: feat.extract(det_kw, fr, lxy))
    return det_kw, svm, feat


def _single_feature_cpu_init(feature):
    from oracle import fdoracle as fo
    fo.build()
    det_kw, svm, feat = _single_feature_models(feature)
    _worker_state.update(fo=fo, kw=det_kw, svm=fo.Svm(svm), feat=feat)


def _single_feature_cpu_item(crop):
    st = _worker_state
    t0 = time.perf_counter()
    r = st["fo"].detect_frame(st["kw"], None, st["svm"], crop, want_dense=False, svm_features=st["feat"])
    return r["windows"], time.perf_counter() - t0


def _single_feature_cpu(feature, steps, frames_per_core=1):
    """bounded sample: the top-left 320x240 quarter of frames (all pyramid layers of the quarter, every window classified)"""
    from featuredetection_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    pool = mp.get_context("spawn").Pool(cores, initializer=_single_feature_cpu_init, initargs=(feature,))
    crops = [np.ascontiguousarray(syn.synthetic_frame(100 + k)[:240, :320]) for k in range(cores * frames_per_core)]
    pool.map(_single_feature_cpu_item, crops[:cores])
    tot_w, tot_t = 0, 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        res = pool.map(_single_feature_cpu_item, crops, chunksize=1)
        tot_t += time.perf_counter() - t0
        tot_w += sum(r[0] for r in res)
    pool.close(); pool.join()
    return tot_w, tot_t, cores, "%d quarter frames (320x240) per step over %d processes, every window through the %s chain + the SVM" % (
        len(crops), cores, feature.upper())


def run_single_feature(args, rank, world, local_rank):
    from featuredetection_b200 import synthetic as syn
    feature = args.workload.split("-")[1]
    n = args.frames
    cfg = {"workload": "ffpDetectApp `single` detector in a feature space (ffpDetectApp.cpp:445-454; BASELINE configs[2] names WHI): %d-frame batch per GPU, "
                       "640x480 1-channel, FaceFrontal geometry, the %s patch-filter chain on EVERY window (16 185 per frame), then the RBF-SVM "
                       "(1024 float32 support vectors)" % (n, feature.upper()),
           "frames_per_gpu": n, "global_frames": n * world, "windows_per_frame": 16185, "parallelism": "frame-sharded dp%d" % world,
           "l2": "a 256 MB scratch write flushes L2 between timed steps"}
    if args.impl == "reference":
        if rank != 0:
            return
        w_, t_, cores, sample = _single_feature_cpu(feature, args.steps)
        value = w_ / t_
        emit({"impl": "reference", "metric": "classified_patches_per_s", "value": value, "unit": "patches/s", "n_gpus": args.gpus, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": 1e3 * t_ / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "u8/f32/f64", "data": "synthetic", "config": cfg,
              "cpu_baseline": {"value": value, "unit": "patches/s", "cores": cores, "kind": "port", "sample": sample},
              "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    import torch
    import torch.distributed as dist
    from featuredetection_b200.detector import Context, SlidingWindowCascade
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    ctx = Context(local_rank)
    det_kw, wvm, _ = syn.landmark_models(CFG)
    fdesc = syn.feature_desc(kind=feature)
    probe = SlidingWindowCascade(ctx, det_kw, wvm, None, feature=fdesc)
    probe.prepare(W, H, 1)
    svm = feature_svm_model(syn, feature, det_kw, probe.layers(), probe.extract_features)
    del probe
    single = SlidingWindowCascade(ctx, det_kw, None, svm, feature=fdesc)
    single.prepare(W, H, 1)
    _, d0 = single.detect_single(syn.synthetic_frames(0, 1))
    svm.threshold = float(np.float32(np.quantile(d0, 0.999)))   # ~0.1 % of the windows positive
    del single
    casc = SlidingWindowCascade(ctx, det_kw, None, svm, feature=fdesc)
    casc.prepare(W, H, n)
    wpf = casc.windows_per_frame
    base = syn.synthetic_frames((rank * n) % 97, min(8, n))
    host_frames = torch.from_numpy(np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]).pin_memory()
    dev_frames = host_frames.to(device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); ctx.synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        l0 = ctx.launch_count()
        ms, out = [], None
        for _ in range(args.steps):
            flush.fill_(1); torch.cuda.synchronize()
            ctx.timer_start()
            out = fn()
            ms.append(ctx.timer_stop())
        barrier()
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), out, ctx.launch_count() - l0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    tot, out, launches = timed(lambda: casc.detect_single_device(dev_frames.data_ptr(), n, None, det_cap=64 * n))
    e2e_tot, out_e2e, _ = timed(lambda: casc.detect_single(host_frames.numpy(), want_distances=False, det_cap=64 * n)[0])
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        value = wpf * n * world * args.steps / (tot * 1e-3)
        e2e = wpf * n * world * args.steps / (e2e_tot * 1e-3)
        peak, peak_src = hbm_peak()
        algo = (W * H + 8 * wpf) * n
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            w_, t_, cores, sample = _single_feature_cpu(feature, 1)
            cpu = {"value": w_ / t_, "unit": "patches/s", "cores": cores, "kind": "port", "sample": sample}
        emit({"metric": "classified_patches_per_s", "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
              "config": cfg, "frames_per_s": value / wpf, "clocks": clocks,
              "e2e": {"value": e2e, "unit": "patches/s", "h2d_bytes_per_step": int(W * H * n), "d2h_bytes_per_step": int(len(out_e2e) * 96 + 8),
                      "ms_per_step": e2e_tot / args.steps},
              "gpu_launches": int(launches),
              "roofline": {"bound": "hbm", "kernel": "feature_patch_kernel + svm_kernel (whole step: the patch-filter chain and the SVM on every window)",
                           "achieved": algo / (tot / args.steps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": algo / (tot / args.steps * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                           "algorithmic_bytes_per_step": int(algo),
                           "note": "compute bound by construction: the exact float32 sequential SSD against 1024 support vectors is ~8e5 scalar operations per window"},
              "detections_per_step": int(len(out)) * world, "cpu_baseline": cpu})
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def pin_rank_to_cores(local_rank, world):
    """each rank's host threads (pipeline thread, CUDA driver threads) stay on their own share of the host cores"""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cores) >= world:
            per = len(cores) // world
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]))
            return per
        return len(cores)
    except Exception:
        return None


def measure_cascades(args, workload, feature, n, rank, world, device, ctx, dist, steps, warmup, sampler=None, gather=None, lo=0):
    """value / e2e / stage-1 profile of one cascade workload (all ranks call this; collective-free except the max-reduce of
    the timings and the final gather)."""
    import torch
    from featuredetection_b200 import capi, synthetic as syn
    from featuredetection_b200.detector import SlidingWindowCascade, DetectorSet, DETECTION_DTYPE
    names = cascade_names(workload)
    cascs = []
    for nm in names:
        det_kw, wvm, svm = syn.landmark_models(nm, args.profile)
        if args.profile == "no-exit":
            det_kw = dict(det_kw, max_positives_per_frame=400000)
        fdesc = None
        if feature != "hq64":
            fdesc = syn.feature_desc(kind=feature)
            probe = SlidingWindowCascade(ctx, det_kw, wvm, None, feature=fdesc)  # the product's own extractor builds the SVM's support vectors
            probe.prepare(W, H, 1)
            svm = feature_svm_model(syn, feature, det_kw, probe.layers(), probe.extract_features)
            del probe
        cascs.append(SlidingWindowCascade(ctx, det_kw, wvm, svm, feature=fdesc))
    dset = None
    if len(cascs) > 1:  # all detectors of the application on every frame: one detector set
        dset = DetectorSet(ctx, cascs)
        dset.prepare(W, H, n)
    else:
        cascs[0].prepare(W, H, n)
    nwin = sum(c.windows_per_frame for c in cascs)
    assert nwin == WINDOWS_PER_FRAME[workload], (nwin, workload)
    max_nwin = max(c.windows_per_frame for c in cascs)
    stage = capi.FDB_STAGE_NMS if args.profile == "realistic" else capi.FDB_STAGE_WVM
    det_cap = (64 * n if args.profile == "realistic" else max_nwin * n) * len(cascs)

    # this rank's shard of the global batch: n frames starting at global frame lo, 8 distinct frames tiled
    base = syn.synthetic_frames(lo % 97, min(8, n))
    host_frames = torch.from_numpy(np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]).pin_memory()
    dev_frames = host_frames.to(device)
    dev_dense = [torch.empty((n, c.windows_per_frame, 2), dtype=torch.int32, device=device) for c in cascs]
    dense_ptrs = [t.data_ptr() for t in dev_dense]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def flush_l2():
        flush.fill_(1)
        torch.cuda.synchronize()

    def step_resident():
        if dset is not None:
            return dset.detect_device(dev_frames.data_ptr(), n, stage=stage, dense_ptrs=dense_ptrs, det_cap=det_cap)
        return cascs[0].detect_device(dev_frames.data_ptr(), n, stage=stage, dense_ptr=dense_ptrs[0], det_cap=det_cap)

    def step_e2e():
        if dset is not None:
            return dset.detect(host_frames.numpy(), stage=stage, det_cap=det_cap)
        return cascs[0].detect(host_frames.numpy(), stage=stage, det_cap=det_cap)

    def final_gather(dets):
        """one exchange at the end of the job - the positives block of every rank (NCCL over NVLink)"""
        if gather is not None:
            gather(dets[:gather.capacity], lo)

    def timed(step_fn):
        barrier()
        launches0 = ctx.launch_count()
        ms = []
        dets = None
        for k in range(steps):
            flush_l2()
            ctx.timer_start()
            dets = step_fn()
            if k == steps - 1:
                final_gather(dets)   # inside the timed region of the last step
            ms.append(ctx.timer_stop())
        barrier()
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        per_rank = None
        if world > 1:
            mine = torch.tensor([min(ms), float(np.median(ms)), max(ms)], dtype=torch.float64, device=device)
            allr = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            per_rank = [[round(float(x), 3) for x in t.cpu().tolist()] for t in allr]
        return float(tot.item()), ms, dets, ctx.launch_count() - launches0, per_rank

    for _ in range(warmup):
        step_resident()
    if sampler is not None and rank == 0:
        sampler.start()
    total_ms, step_ms, dets, launches, per_rank = timed(step_resident)
    host_ms = dset.last_host_ms() if dset is not None else None

    # dominant kernel: CUDA events around the stage-1 kernels (serialised)
    prof = []
    for _ in range(max(min(steps, 5), 3)):
        flush_l2()
        prof.append(np.array(dset.profile_device(dev_frames.data_ptr(), n)) if dset is not None
                    else np.array(cascs[0].profile_device(dev_frames.data_ptr(), n)))
    prof = np.array(prof).mean(axis=0)

    for _ in range(2):
        step_e2e()
    e2e_total, e2e_ms, dets_e2e, _, e2e_per_rank = timed(step_e2e)
    info = dset.info() if dset is not None else {"pyramid_images": None, "pyramid_bytes": cascs[0].pyramid_bytes,
                                                  "window_launches": 1, "fast_members": 1}
    return dict(nwin=nwin, n_detectors=len(cascs), total_ms=total_ms, step_ms=step_ms, e2e_total=e2e_total, e2e_ms=e2e_ms,
                launches=int(launches), prof=prof, n_dets=int(len(dets)), d2h=int(len(dets_e2e) * DETECTION_DTYPE.itemsize + 8),
                per_rank=per_rank, e2e_per_rank=e2e_per_rank, info=info, host_ms=host_ms)


def measured_traffic(kernel, frames=1):
    """dram bytes of a kernel from the committed ncu --set full capture (profiles/traffic.json names the capture and the build)"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    t = json.load(open(p)).get(kernel)
    if not t:
        return None, None
    if "dram_bytes_per_frame" in t:
        return t["dram_bytes_per_frame"] * frames, t["source"]
    return t["dram_bytes_per_launch"], t["source"]


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU per step (default 256; 4096 faces for sdm)")
    ap.add_argument("--frames-total", type=int, default=None, help="strong scaling: this many frames per step split over the ranks (BASELINE configs[3]: 4096)")
    ap.add_argument("--workload", default="landmarks15", choices=["landmarks15", "facefrontal", "sdm", "single-psvm", "aggdet", "single-whi", "single-hog"],
                    help="landmarks15 (default, headline) = all 15 ffpDetectApp landmark detectors on every frame (BASELINE north_star / configs[3]); "
                         "facefrontal = BASELINE configs[1]; sdm = BASELINE configs[4] (supervised-descent fit, 68 landmarks, 4096 faces per GPU); "
                         "single-psvm = the `single` detector, every window through the RBF-SVM")
    ap.add_argument("--profile", default="realistic", choices=["realistic", "no-exit"])
    ap.add_argument("--feature", default=None, choices=["hq64", "hog", "whi", "lbp", "histeq", "ehog"],
                    help="feature space of the second-stage SVM (default: hog for facefrontal = BASELINE configs[1]; hq64 for landmarks15)")
    ap.add_argument("--cpu-frames", type=int, default=None, help="cpu_baseline sample: frames (each through every detector of the workload)")
    ap.add_argument("--ref-frames-per-step", type=int, default=None, help="--impl reference: frames per step (each through every detector)")
    ap.add_argument("--cpu-frames-per-core", type=int, default=32, help="sdm / single-psvm: cpu_baseline sample size per host core")
    ap.add_argument("--ref-frames-per-core", type=int, default=2, help="sdm / single-psvm, --impl reference: frames per core per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-facefrontal", action="store_true", help="landmarks15: skip the nested BASELINE configs[1] block")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.frames_total:
        args.frames = max(1, args.frames_total // world)
    if args.frames is None:
        args.frames = {"facefrontal": 256, "landmarks15": 256, "sdm": 4096, "single-psvm": 256, "aggdet": 64, "single-whi": 8, "single-hog": 8}[args.workload]
    if args.feature is None:
        args.feature = "hog" if args.workload == "facefrontal" else "hq64"
    if args.workload == "landmarks15" and args.feature != "hq64":
        raise SystemExit("bench.py: --feature applies to the facefrontal workload")
    if args.ref_frames_per_step is None:
        args.ref_frames_per_step = 32 if args.workload == "facefrontal" else 2   # >= 16 work items per core (facefrontal) / 30 heavy items (15 detectors)
    if args.cpu_frames is None:
        args.cpu_frames = 512 if args.workload == "facefrontal" else 6

    if args.workload == "single-psvm":
        if args.impl == "reference":
            run_reference_single(args, rank, world)
        else:
            run_single(args, rank, world, local_rank)
        return
    if args.workload == "aggdet":
        run_aggdet(args, rank, world, local_rank)
        return
    if args.workload in ("single-whi", "single-hog"):
        run_single_feature(args, rank, world, local_rank)
        return
    if args.workload == "sdm":
        if args.impl == "reference":
            run_reference_sdm(args, rank, world)
        else:
            run_sdm(args, rank, world, local_rank)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from featuredetection_b200 import sharding
    from featuredetection_b200.detector import Context, DETECTION_DTYPE

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    host_cores = pin_rank_to_cores(local_rank, world)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL writes its version banner / debug lines to stdout by default: keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=device)

    ctx = Context(local_rank)
    n = args.frames
    lo, _ = sharding.shard_range(n * world, rank, world)
    n_det = len(cascade_names(args.workload))
    gather = sharding.DetectionGather(max(256, 64 * n) * n_det, dist, device) if world > 1 else None
    if gather is not None:  # set-up, like loading the models: NCCL connects the ranks for a collective when it is first used
        gather(np.zeros(0, DETECTION_DTYPE), lo)
    sampler = ClockSampler(local_rank)
    m = measure_cascades(args, args.workload, args.feature, n, rank, world, device, ctx, dist, args.steps, args.warmup,
                         sampler=sampler, gather=gather, lo=lo)
    ff = None
    if args.workload == "landmarks15" and not args.no_facefrontal:
        ff = measure_cascades(args, "facefrontal", "hog", 256, rank, world, device, ctx, dist, max(args.steps, 10), 3, lo=256 * rank)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        nwin = m["nwin"]
        windows_step = nwin * n * world
        value = windows_step * args.steps / (m["total_ms"] * 1e-3)
        e2e_value = windows_step * args.steps / (m["e2e_total"] * 1e-3)
        peak, peak_src = hbm_peak()
        ms_resize, ms_down, ms_wvm, ms_deep, ms_stage1, n_launch = m["prof"]
        # dominant kernel = wvm_group_kernel (all its launches of a step together): SURVEY 8(d) bytes = frame read once + dense records
        algo_bytes = (W * H + WINDOW_BYTES * nwin) * n
        achieved = algo_bytes / (ms_wvm * 1e-3) / 1e9
        kname = "wvm_group_kernel/landmarks15" if args.workload == "landmarks15" else "wvm_group_kernel/facefrontal"
        traffic, traffic_src = measured_traffic(kname, n)
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is a rank-0, N=1 datum
            arm = CpuArm(args.workload, args.profile, args.feature)
            arm.run(arm.frames(0, 1))
            per = args.ref_frames_per_step
            wcpu, wall, split = 0, 0.0, None
            for k in range(max(1, args.cpu_frames // per)):   # the same step size as --impl reference
                w_, t_, sp = arm.run(arm.frames(100 + k * per, per))
                wcpu += w_; wall += t_
                split = sp if split is None else {q: split[q] + sp[q] for q in sp}
            cpu = {"value": wcpu / wall, "unit": "patches/s", "cores": arm.cores, "kind": arm.kind,
                   "sample": "%d frames x %d detectors of the same workload in steps of %d frames over %d processes, %.1f s wall" % (
                       max(1, args.cpu_frames // per) * per, len(arm.names), per, arm.cores, wall),
                   "pyramid": arm.describe(), "split_core_seconds": split}
            arm.close()
        line = {
            "metric": "classified_patches_per_s", "value": value, "unit": "patches/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["total_ms"] / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.frames_total else "weak", "vs_baseline": None,
            "dtype": "u8/f32/f64", "data": "synthetic",
            "config": workload_config(args, world),
            "frames_per_s": value / nwin,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": int(W * H * n),
                    "d2h_bytes_per_step": m["d2h"], "ms_per_step": m["e2e_total"] / args.steps, "frames_per_s": e2e_value / nwin},
            "gpu_launches": m["launches"],
            "roofline": {"bound": "hbm", "kernel": "window kernels: wvm_group_tc_kernel (tcgen05.mma kind::i8, packs of 3-4 detectors) + wvm_group_kernel "
                                                   "(mma.sync u8, packs of 1-2): fused HistEq64 + the first 8 WVM filters of every window of every detector "
                                                   "as exact u8 matrix products; %d launches per step, timed together" % int(n_launch),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_step": int(algo_bytes),
                         "kernel_ms_per_step": float(ms_wvm), "launches_per_step": int(n_launch),
                         "note": "instruction-issue / shared-memory bound by construction: ~4e3 warp instructions per 32 windows (shared by the "
                                 "detectors of a pack) + ~1e3 per detector vs 8 B of compulsory traffic per window (SURVEY.md 8(d)); issue-slot "
                                 "utilisation, load/store-pipe wavefronts and pipe shares in profiles/"},
            "stage1_ms": {"resize": float(ms_resize), "pyrdown": float(ms_down), "window_kernels": float(ms_wvm), "deep_kernel": float(ms_deep),
                          "total": float(ms_stage1)},
            "detector_set": m["info"],
            "host_ms_last_step": m["host_ms"],
            "detections_per_step": m["n_dets"] * world,
            "host_cores_per_rank": host_cores,
            "step_ms_per_rank_min_median_max": m["per_rank"], "e2e_step_ms_per_rank_min_median_max": m["e2e_per_rank"],
            "cpu_baseline": cpu,
        }
        if ff is not None:
            ffw = ff["nwin"] * 256 * world
            line["facefrontal"] = {
                "workload": "BASELINE configs[1]: 256-frame batch per GPU, 640x480, FaceFrontal WVM->SVM cascade, HOG SVM stage (round-1 headline)",
                "value": ffw * len(ff["step_ms"]) / (ff["total_ms"] * 1e-3), "ms_per_step": ff["total_ms"] / len(ff["step_ms"]),
                "e2e_value": ffw * len(ff["e2e_ms"]) / (ff["e2e_total"] * 1e-3), "e2e_ms_per_step": ff["e2e_total"] / len(ff["e2e_ms"]),
                "unit": "patches/s", "stage1_ms": {"resize": float(ff["prof"][0]), "pyrdown": float(ff["prof"][1]), "window_kernel": float(ff["prof"][2]),
                                                   "deep_kernel": float(ff["prof"][3]), "total": float(ff["prof"][4])}}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
