"""Frame-level data parallelism.

Frames are independent units (every ``update(Mat)`` rebuilds the pyramid, FeatureExtractor.hpp:32-34,
classifiers are read-only at inference), so a batch is split into contiguous frame ranges, one per
rank; models are replicated; the only exchange is a gather of fixed-size detection records at the
end (NCCL over NVLink on GPUs, gloo in the CPU tests). No data-path collective exists.
"""
import numpy as np

# columns of the gathered record (float64 keeps every field exactly: ints < 2^53)
GATHER_FIELDS = ("frame", "window", "layer", "x", "y", "center_x", "center_y", "width", "height",
                 "wvm_level", "wvm_fout", "wvm_probability", "svm_distance", "svm_probability", "probability", "positive", "reserved")


def shard_range(n_frames, rank, world):
    """Contiguous range [lo, hi) of rank `rank`: ranges differ by at most one frame."""
    base, rem = divmod(int(n_frames), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_detections(dets, frame_offset, capacity, rows=None):
    """structured detections -> [capacity + 1, len(GATHER_FIELDS)] float64 block (or its first `rows` rows, rows > number
    of detections kept); row 0 holds the count (and an overflow flag); frame indices become global."""
    block = np.zeros((capacity + 1 if rows is None else rows, len(GATHER_FIELDS)), np.float64)
    n = min(len(dets), capacity)
    block[0, 0] = len(dets)
    block[0, 1] = 1.0 if len(dets) > capacity else 0.0
    for j, f in enumerate(GATHER_FIELDS):
        col = dets[f][:n].astype(np.float64)
        if f == "frame":
            col = col + frame_offset
        block[1:n + 1, j] = col
    return block


def unpack_detections(blocks, dtype):
    """[world, capacity + 1, F] gathered blocks -> one structured array ordered by rank (= frame order)."""
    out = []
    for b in blocks:
        n = int(b[0, 0])
        if b[0, 1] != 0.0:
            raise OverflowError("a rank overflowed its gather block (%d detections)" % n)
        rec = np.zeros(n, dtype)
        for j, f in enumerate(GATHER_FIELDS):
            rec[f] = b[1:n + 1, j].astype(dtype[f])
        out.append(rec)
    return np.concatenate(out) if out else np.zeros(0, dtype)


def gather_detections(dets, frame_offset, capacity, dist, device=None, dst=0):
    """Result gather: every rank contributes one fixed-size block; rank `dst` gets them all.
    `dist` is torch.distributed (initialised); tensors live on `device` (cuda for NCCL)."""
    import torch
    block = torch.from_numpy(pack_detections(dets, frame_offset, capacity))
    if device is not None:
        block = block.to(device)
    world = dist.get_world_size()
    bucket = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(bucket, block)
    if dist.get_rank() != dst:
        return None
    return unpack_detections([b.cpu().numpy() for b in bucket], dets.dtype)


class DetectionGather:
    """Reusable gather with preallocated buffers (pinned host + device), double-buffered so that the gather of
    step k overlaps the computation of step k+1: submit() enqueues H2D + all_gather + D2H on a side stream and
    returns a ticket, collect(ticket) waits for it and unpacks on the destination rank.
    Blocks travel padded to the largest rank's count, not to the capacity (a first, 8-byte all_gather tells every rank
    that count), and only the destination rank copies the gathered blocks to the host."""

    def __init__(self, capacity, dist, device=None):
        import torch
        self.torch, self.dist, self.device, self.capacity = torch, dist, device, int(capacity)
        self.world = dist.get_world_size()
        self.fields = len(GATHER_FIELDS)
        shape = (self.capacity + 1, self.fields)
        pin = device is not None
        self.bufs = []
        for _ in range(2):
            b = {"h_in": torch.zeros(shape, dtype=torch.float64, pin_memory=pin),
                 "h_out": torch.zeros((self.world,) + shape, dtype=torch.float64, pin_memory=pin), "event": None}
            if device is not None:
                b["d_in"] = torch.zeros(shape, dtype=torch.float64, device=device)
                b["d_out"] = torch.zeros((self.world,) + shape, dtype=torch.float64, device=device)
            self.bufs.append(b)
        self.count_in = torch.zeros(1, dtype=torch.int64, device=device)
        self.count_out = torch.zeros(self.world, dtype=torch.int64, device=device)
        self.stream = torch.cuda.Stream(device) if device is not None else None
        self.turn = 0

    def submit(self, dets, frame_offset, dst=0):
        torch = self.torch
        b = self.bufs[self.turn]
        self.turn ^= 1
        n = min(len(dets), self.capacity)
        self.count_in.fill_(n)
        self.dist.all_gather_into_tensor(self.count_out, self.count_in)
        rows = int(self.count_out.max().item()) + 1  # row 0 of a block: count and overflow flag
        b["rows"], b["dtype"] = rows, dets.dtype
        used = rows * self.fields
        b["h_in"].numpy()[:rows] = pack_detections(dets, frame_offset, self.capacity, rows)
        if self.device is None:
            self.dist.all_gather_into_tensor(b["h_out"].view(-1)[:self.world * used], b["h_in"].view(-1)[:used])
        else:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                b["d_in"].view(-1)[:used].copy_(b["h_in"].view(-1)[:used], non_blocking=True)
                self.dist.all_gather_into_tensor(b["d_out"].view(-1)[:self.world * used], b["d_in"].view(-1)[:used])
                if self.dist.get_rank() == dst:
                    b["h_out"].view(-1)[:self.world * used].copy_(b["d_out"].view(-1)[:self.world * used], non_blocking=True)
                b["event"] = torch.cuda.Event()
                b["event"].record(self.stream)
        return b

    def collect(self, ticket, dst=0):
        if ticket.get("event") is not None:
            ticket["event"].synchronize()
        if self.dist.get_rank() != dst:
            return None
        rows = ticket["rows"]
        blocks = ticket["h_out"].view(-1)[:self.world * rows * self.fields].view(self.world, rows, self.fields)
        return unpack_detections(list(blocks.numpy()), ticket["dtype"])

    def __call__(self, dets, frame_offset, dst=0):
        return self.collect(self.submit(dets, frame_offset, dst), dst)
