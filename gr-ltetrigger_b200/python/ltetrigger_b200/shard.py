"""Multi-GPU sharding of the search: one process per GPU, streams dealt round-robin, no
collective on the data path.  Streams are independent flowgraphs in the reference
(one `downlink_trigger_c` per input, python/downlink_trigger_c.py:18-45), so rank r simply owns
streams r, r+N, r+2N, ... and the only exchange is the host-side merge of the (tiny) record /
detection lists at the end -- an all-gather of a few kB over whatever backend the process group
has (NCCL between GPUs, gloo in the CPU tests).

One long capture is sharded the second way the north_star names: by TIME SEGMENT with a halo.
`plan_time_segments` cuts it into equal, overlapping segments; each segment is searched as a stream of
its own (on one engine, or dealt to ranks with `owned_streams` like any other stream) and
`stitch_segments` maps the records back onto the capture's time axis, every segment contributing the
part behind its halo.  The chains of the reference carry state from window to window without bound (the
per-lag moving average, the tracking score, the CP average: lib/pss_impl.cc:111-152), so a segment is
not bit-identical to the same stretch of a sequential run; what it is bit-identical to is the
reference's search started at the segment's first sample -- the halo is the stretch that search
needs to reach tracking (track_after windows) before the part it owns begins.
"""
import numpy as np

from . import _abi as A


def owned_streams(n_streams, rank, world):
    """Global indices of the streams rank `rank` of `world` processes."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    return np.arange(rank, n_streams, world, dtype=np.int64)


def to_global(recs, owned):
    """Rewrite the engine's local stream ordinals (0..len(owned)-1) to global stream ids."""
    out = recs.copy()
    if len(out):
        out["stream"] = np.asarray(owned, np.int64)[out["stream"]].astype(np.int32)
    return out


def sort_records(recs):
    order = np.lexsort((recs["win_index"], recs["n_id_2"], recs["stream"]))
    return recs[order]


def merge_records(local_recs, group=None, dst=None):
    """All ranks contribute their WINDOW_REC arrays; returns the concatenation ordered by
    (stream, n_id_2, win_index) on every rank (dst=None) or on rank `dst` only (others get None).
    Without an initialised process group it is the identity (single GPU)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return sort_records(local_recs)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    raw = np.ascontiguousarray(local_recs).view(np.uint8).reshape(-1)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([raw.size], dtype=torch.int64, device=dev), group=group)
    counts = [int(c.item()) for c in counts]
    width = max(max(counts), 1)
    mine = torch.zeros(width, dtype=torch.uint8, device=dev)
    if raw.size:
        mine[:raw.size] = torch.from_numpy(raw.copy()).to(dev)
    parts = [torch.zeros(width, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    if dst is not None and rank != dst:
        return None
    chunks = [p[:c].cpu().numpy().view(A.WINDOW_REC) for p, c in zip(parts, counts) if c]
    merged = np.concatenate(chunks) if chunks else np.zeros(0, A.WINDOW_REC)
    return sort_records(merged)


class SegmentPlan:
    """n_segments segments of `length` input-rate samples starting at `starts[k]`; segment k owns the records
    whose half-frame or window starts in [own_from[k], own_to[k]) on the capture's search-rate axis."""

    def __init__(self, n_samples, decim, starts, length, halo):
        self.n_samples, self.decim, self.length, self.halo = n_samples, decim, length, halo
        self.starts = np.asarray(starts, np.int64)
        self.n_segments = len(self.starts)
        # a segment's last windows are never evaluated (a window needs 18365 samples from its start: the scheduler
        # rule of lib/pss_impl.cc:154-190), so the next segment takes over four half-frames before the overlap ends
        first = self.starts // decim
        self.own_from = np.where(np.arange(self.n_segments) == 0, 0, first + halo // decim - 4 * 9600)
        self.own_to = np.append(self.own_from[1:], np.iinfo(np.int64).max)


def plan_time_segments(n_samples, decim, n_segments, halo_halfframes=24):
    """Cut n_samples input-rate samples into `n_segments` equal segments that overlap by the halo
    (`halo_halfframes` x 5 ms: track_after = 16 windows to reach tracking, and as many again for the
    moving averages to settle).  Starts are multiples of 8 * decim samples (the engine's chunk
    granularity, and it keeps every segment on the decimator phase of the sequential run); the last
    segment is moved back so that it ends with the capture.  Fewer segments are planned when the
    capture is too short for each to own at least one halo's worth of signal."""
    gran = 8 * decim
    halo = halo_halfframes * 9600 * decim
    if n_segments < 1 or n_samples < 1 or halo_halfframes < 8:
        raise ValueError("bad segment request %r x %r, halo %r" % (n_segments, n_samples, halo_halfframes))
    n_segments = int(max(1, min(n_segments, (n_samples - halo) // max(halo, 1))))
    if n_segments == 1:
        return SegmentPlan(n_samples, decim, [0], n_samples // gran * gran, 0)
    step = (n_samples - halo) // n_segments // gran * gran
    length = step + halo
    starts = [k * step for k in range(n_segments)]
    starts[-1] = (n_samples - length) // gran * gran
    return SegmentPlan(n_samples, decim, starts, length, halo)


def cut_segments(iq, plan):
    """[n_samples(, 2)] capture -> [n_segments, length(, 2)] array, one row per segment."""
    return np.stack([iq[s:s + plan.length] for s in plan.starts])


def stitch_segments(recs, plan, stream=0):
    """Window records of a search over the rows of `cut_segments` (stream ordinal = segment) -> records
    on the capture's own time axis, stream id `stream`, each segment contributing what it owns, ordered
    by (n_id_2, time).  win_index is renumbered per chain."""
    out = recs.copy()
    seg = out["stream"].astype(np.int64)
    shift = plan.starts[seg] // plan.decim
    emitted = out["emit_start"] >= 0
    out["win_start"] += shift
    out["emit_start"] = np.where(emitted, out["emit_start"] + shift, -1)
    at = np.where(emitted, out["emit_start"], out["win_start"])
    out = out[(at >= plan.own_from[seg]) & (at < plan.own_to[seg])]
    at = np.where(out["emit_start"] >= 0, out["emit_start"], out["win_start"])
    out = out[np.lexsort((at, out["n_id_2"]))]
    out["stream"] = stream
    for r in range(3):
        m = out["n_id_2"] == r
        out["win_index"][m] = np.arange(int(m.sum()))
    return out


DETECTION = np.dtype([("stream", "<i4"), ("cell_id", "<i4"), ("n_id_2", "<i4"), ("n_id_1", "<i4"),
                      ("cp_normal", "<i4"), ("count", "<i4"), ("frame_start", "<i8"), ("first_seen", "<i8"),
                      ("max_psr", "<f4")])


def detections(recs):
    """Collapse window records to one row per (stream, cell_id): what `cellstore.cells()` would
    list once mib had confirmed them.  frame_start = absolute search-rate sample index of the first
    decoded subframe-0 half-frame (m1 > m0 <=> subframe 0), modulo one 10 ms frame;
    -1 if only subframe-5 half-frames were decoded."""
    cells = recs[(recs["flags"] & A.F_CELL) != 0]
    rows = []
    if len(cells):
        keys = np.stack([cells["stream"], cells["cell_id"]], 1)
        uniq, inv = np.unique(keys, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        for i, (s, c) in enumerate(uniq):
            g = cells[inv == i]
            sf0 = g[g["m1"] > g["m0"]]
            fs = int(sf0["emit_start"][0] % 19200) if len(sf0) else -1
            norm = int(((g["flags"] & A.F_CP_NORM) != 0).sum() * 2 >= len(g))
            rows.append((int(s), int(c), int(g["n_id_2"][0]), int(g["n_id_1"][0]), norm, len(g), fs,
                         int(g["emit_start"].min()), float(g["psr"].max())))
    return np.array(rows, DETECTION)


class ShardedTrigger:
    """This rank's slice of a `n_streams`-stream search: a `Trigger` over the owned streams whose
    records come back with global stream ids.  `engine_factory(n_local, **kw)` builds the engine
    (default: the CUDA `Trigger` on `device`)."""

    def __init__(self, n_streams, rank, world, device=None, engine_factory=None, **kw):
        self.n_streams, self.rank, self.world = n_streams, rank, world
        self.owned = owned_streams(n_streams, rank, world)
        if engine_factory is None:
            from .engine import Trigger
            engine_factory = lambda n, **k: Trigger(n, device=rank if device is None else device, **k)  # noqa: E731
        self.engine = engine_factory(len(self.owned), **kw) if len(self.owned) else None

    def local_view(self, iq_all):
        """Rows of a [n_streams, ...] host array this rank owns."""
        return iq_all[self.owned]

    def run(self, iq_local, **kw):
        if self.engine is None:
            return np.zeros(0, A.WINDOW_REC)
        return to_global(self.engine.run(iq_local, **kw), self.owned)

    def run_and_merge(self, iq_local, group=None, dst=None, **kw):
        return merge_records(self.run(iq_local, **kw), group=group, dst=dst)
