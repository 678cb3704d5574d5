"""Multi-GPU sharding of the search: one process per GPU, streams dealt round-robin, no
collective on the data path.  Streams are independent flowgraphs in the reference
(one `downlink_trigger_c` per input, python/downlink_trigger_c.py:18-45), so rank r simply owns
streams r, r+N, r+2N, ... and the only exchange is the host-side merge of the (tiny) record /
detection lists at the end -- an all-gather of a few kB over whatever backend the process group
has (NCCL between GPUs, gloo in the CPU tests).
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
