"""Multi-process (world_size 2, gloo, CPU) test of the frame-sharding logic used by bench.py at N > 1:
contiguous frame ranges per rank, replicated models, one gather of fixed-size detection blocks.
Each rank produces its detections with the CPU oracle here (the GPU path is exercised by -m gpu tests);
the gathered result must equal a single-process run over all frames."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_everything():
    from featuredetection_b200.sharding import shard_range
    for n in (0, 1, 7, 256, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    from featuredetection_b200.detector import DETECTION_DTYPE
    from featuredetection_b200.sharding import pack_detections, unpack_detections
    rng = np.random.default_rng(0)
    d = np.zeros(17, DETECTION_DTYPE)
    for f in DETECTION_DTYPE.names:
        d[f] = rng.integers(0, 1000, 17) if DETECTION_DTYPE[f].kind == "i" else rng.normal(size=17)
    blocks = [pack_detections(d, 0, 32), pack_detections(d[:5], 100, 32)]
    out = unpack_detections(blocks, DETECTION_DTYPE)
    assert len(out) == 22
    for f in ("window", "wvm_fout", "svm_distance", "probability", "center_x"):
        assert np.array_equal(out[f][:17], d[f])
    assert np.array_equal(out["frame"][17:], d["frame"][:5] + 100)
    with pytest.raises(OverflowError):
        unpack_detections([pack_detections(d, 0, 4)], DETECTION_DTYPE)


def _worker(rank, world, port, n_frames, queue):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from featuredetection_b200 import capi, synthetic as syn, sharding
    from oracle import fdoracle as fo
    dist.init_process_group("gloo", rank=rank, world_size=world)
    det_kw, wvm, svm = syn.landmark_models("FaceFrontal")
    wo, so = fo.Wvm(wvm), fo.Svm(svm)
    lo, hi = sharding.shard_range(n_frames, rank, world)
    mine = [fo.detect_frame(det_kw, wo, so, syn.synthetic_frame(k), stage=capi.FDB_STAGE_NMS, frame_index=k - lo,
                            want_dense=False)["detections"] for k in range(lo, hi)]
    dets = np.concatenate(mine) if mine else np.zeros(0, fo.DETECTION_DTYPE)
    gathered = sharding.gather_detections(dets, lo, 64, dist)
    again = sharding.DetectionGather(64, dist)(dets, lo)
    if rank == 0:
        assert len(again) == len(gathered) and np.array_equal(again["window"], gathered["window"])
    if rank == 0:
        queue.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(built):
    import torch.multiprocessing as mp
    from featuredetection_b200 import capi, synthetic as syn
    from oracle import fdoracle as fo
    n_frames, world = 5, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    det_kw, wvm, svm = syn.landmark_models("FaceFrontal")
    wo, so = fo.Wvm(wvm), fo.Svm(svm)
    serial = np.concatenate([fo.detect_frame(det_kw, wo, so, syn.synthetic_frame(k), stage=capi.FDB_STAGE_NMS, frame_index=k,
                                             want_dense=False)["detections"] for k in range(n_frames)])
    assert len(gathered) == len(serial)
    for f in ("frame", "window", "layer", "center_x", "center_y", "wvm_level", "wvm_fout", "svm_distance", "probability"):
        assert np.array_equal(gathered[f], serial[f]), f
