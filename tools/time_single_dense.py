"""Quick timing of the `single` psvm detector on the tensor cores (not a bench line): frames resident in HBM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from featuredetection_b200 import synthetic as syn
from featuredetection_b200.detector import Context, SlidingWindowCascade

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = Context(0)
det_kw, _, svm = syn.landmark_models("FaceFrontal")
probe = SlidingWindowCascade(ctx, det_kw, None, svm)
probe.prepare(640, 480, 1)
_, d0 = probe.detect_single(syn.synthetic_frames(0, 1))
svm.threshold = float(np.float32(np.quantile(d0, 0.999)))  # ~0.1 % of the windows positive
del probe
if len(sys.argv) > 2:
    os.environ['FDB_SVMD_DBG'] = sys.argv[2]  # ablation timing (see svm_dense.cu); results are then meaningless
c = SlidingWindowCascade(ctx, det_kw, None, svm)
c.prepare(640, 480, n)
assert c.single_dense
base = syn.synthetic_frames(0, 8)
frames = torch.from_numpy(np.concatenate([base] * ((n + 7) // 8))[:n]).cuda()
dist = torch.empty((n, c.windows_per_frame), dtype=torch.float64, device="cuda")
for _ in range(2):
    d = c.detect_single_device(frames.data_ptr(), n, dist.data_ptr(), det_cap=n * c.windows_per_frame)
ts = []
for _ in range(5):
    ctx.timer_start(); d = c.detect_single_device(frames.data_ptr(), n, dist.data_ptr(), det_cap=n * c.windows_per_frame); ts.append(ctx.timer_stop())
ms = float(np.median(ts))
w = n * c.windows_per_frame
print("single psvm dense: %d frames, %d windows, %.3f ms, %.3e windows/s, %.1f TOP/s (u8), %d positives" % (
    n, w, ms, w / ms * 1e3, 2.0 * w * svm.sv.shape[0] * 400 / ms * 1e3 / 1e12, len(d)))
