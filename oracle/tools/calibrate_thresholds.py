"""Calibrate the hierarchical thresholds of the synthetic WVM models.

Runs the CPU oracle with all thresholds at -inf on frames 0..3 (a window subsample for the large
landmark models), then picks hierarchicalThresholds[l] so that the cumulative survival after
filter l is S(l) = max(0.5^(l+1), 2e-3) ("realistic" profile: ~2 filters per window on average,
~0.2 % of the windows reach the SVM). Each threshold is placed midway between the last surviving
and the first rejected fout so that no calibration window sits exactly on a threshold.

Output: featuredetection_b200/data/thresholds_<cfg>_realistic.json (float32 bit patterns).
Test/bench infrastructure: this is the only place model *generation* touches the oracle.

usage: python oracle/tools/calibrate_thresholds.py [cfg names ...]   (default: all 15)
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from featuredetection_b200 import synthetic as syn  # noqa: E402
from oracle import fdoracle as fo  # noqa: E402

MAX_WINDOWS = 60000


def survival(l):
    return max(0.5 ** (l + 1), 2e-3)


def calibrate(name):
    idx, (nm, pw, ph, inc, mn, mx, per_level, levels, r) = syn.landmark_config(name)
    wvm = syn.make_wvm(pw, ph, per_level, levels, r, seed=100 + idx)
    patches = []
    for k in range(4):
        frame = syn.synthetic_frame(k)
        _, layers = fo.pyramid(frame, float(np.float32(inc)), float(np.float32(mn)), float(np.float32(mx)))
        for (_, scale, img) in layers:
            ex, ey = fo.lib().fdo_cvround(640 * scale), fo.lib().fdo_cvround(480 * scale)
            for y in range(0, ey - ph):
                for x in range(0, ex - pw):
                    patches.append((k, img, x, y))
    rng = np.random.default_rng(7)
    if len(patches) > MAX_WINDOWS:
        sel = rng.choice(len(patches), MAX_WINDOWS, replace=False)
        sel.sort()
        patches = [patches[i] for i in sel]
    data = np.stack([fo.hq64(img[y:y + ph, x:x + pw]).ravel() for (_, img, x, y) in patches])
    fouts = fo.Wvm(wvm).eval_all_levels(data)  # [N, n]
    N, n = fouts.shape
    alive = np.arange(N)
    thr = np.zeros(n, np.float32)
    for l in range(n):
        keep = max(1, int(round(survival(l) * N)))
        f = fouts[alive, l]
        order = np.argsort(-f, kind="stable")
        if keep >= len(alive):
            t = np.float32(f.min()) - np.float32(abs(f.min()) * 0.05 + 1e-3)
            keep = len(alive)
        else:
            a, b = np.float32(f[order[keep - 1]]), np.float32(f[order[keep]])
            t = np.float32((np.float64(a) + np.float64(b)) / 2)
            if not (t <= a and t > b):  # a == b or adjacent floats: fall back to a (ties all survive)
                t = a
        thr[l] = t
        alive = alive[f >= t]
    out = dict(cfg=name, profile="realistic", calibration_windows=int(N), survivors_last=int(len(alive)),
               thresholds_u32=[int(v) for v in thr.view(np.uint32)], thresholds=[float(v) for v in thr])
    path = syn.thresholds_path(name, "realistic")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("%s: N=%d survivors=%d -> %s" % (name, N, len(alive), path))


if __name__ == "__main__":
    names = sys.argv[1:] or [c[0] for c in syn.LANDMARK_CONFIGS]
    for nm in names:
        calibrate(nm)
