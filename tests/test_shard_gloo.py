"""Multi-process path on CPU: two gloo ranks shard a batch of streams the way `bench.py --gpus N`
and `shard.ShardedTrigger` do (stream i -> rank i mod N, no collective on the data path), run
their slices independently, and merge the record lists.  The CUDA engine cannot run here, so
the oracle stands in for it behind the same `.run(iq)` interface -- what is under test is the
host logic: ownership, local->global stream ids, the all-gather merge, ordering, detections."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

N_STREAMS, N_SAMPLES = 5, 19200 * 12          # odd stream count: ragged shards (3 + 2)


class OracleEngine:
    def __init__(self, n_local, decim=1, psr_threshold=4.0):
        from oracle import oracle as O
        self.O, self.decim, self.thr = O, decim, psr_threshold

    def run(self, iq):
        return self.O.trigger_run(iq, decim=self.decim, psr_threshold=self.thr, nthreads=2)


def make_batch():
    from ltetrigger_b200 import synth
    return synth.batch(N_STREAMS, N_SAMPLES, 6.0, master_seed=42)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "gr-ltetrigger_b200", "python"))
    import torch.distributed as dist
    from ltetrigger_b200 import shard
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        iq, _ = make_batch()
        st = shard.ShardedTrigger(N_STREAMS, rank, world, engine_factory=OracleEngine)
        assert list(st.owned) == list(range(rank, N_STREAMS, world))
        merged_all = st.run_and_merge(st.local_view(iq))            # every rank gets the full list
        merged_dst = st.run_and_merge(st.local_view(iq), dst=0)     # only rank 0
        assert (merged_dst is None) == (rank != 0)
        np.save(os.path.join(out_dir, "merged_%d.npy" % rank), merged_all)
        # an empty contribution must not break the merge
        empty = shard.merge_records(np.zeros(0, merged_all.dtype) if rank == 1 else merged_all[:3])
        assert len(empty) == 3
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_gloo_shard_and_merge(tmp_path, oracle):
    import torch.multiprocessing as mp
    from ltetrigger_b200 import shard
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    iq, ids = make_batch()
    want = oracle.trigger_run(iq, decim=1, psr_threshold=4.0)       # one process, all streams
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "merged_%d.npy" % rank))
        assert got.tobytes() == want.tobytes()
    det = shard.detections(want)
    assert sorted(det["stream"].tolist()) == list(range(N_STREAMS))
    for row in det:
        assert row["cell_id"] == ids[row["stream"]] and row["cp_normal"] == 1
        assert row["n_id_2"] == row["cell_id"] % 3 and row["n_id_1"] == row["cell_id"] // 3
        assert 0 <= row["frame_start"] < 19200


def test_ownership_and_global_ids():
    from ltetrigger_b200 import shard, _abi
    assert list(shard.owned_streams(10, 1, 4)) == [1, 5, 9]
    assert len(shard.owned_streams(2, 3, 4)) == 0                   # more ranks than streams
    with pytest.raises(ValueError):
        shard.owned_streams(4, 4, 4)
    recs = np.zeros(4, _abi.WINDOW_REC)
    recs["stream"] = [0, 1, 2, 1]
    out = shard.to_global(recs, shard.owned_streams(10, 1, 4))
    assert out["stream"].tolist() == [1, 5, 9, 5]
    assert len(shard.merge_records(recs)) == 4                      # no process group: identity (sorted)
    assert len(shard.detections(np.zeros(0, _abi.WINDOW_REC))) == 0
