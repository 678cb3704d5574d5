"""Drop-in mirrors of the reference's blocks for the PSS+SSS path.

    ltetrigger.pss                 include/ltetrigger/pss.h:36-88,  lib/pss_impl.cc
    ltetrigger.sss                 include/ltetrigger/sss.h:36-52,  lib/sss_impl.cc
    ltetrigger.downlink_trigger_c  python/downlink_trigger_c.py:13-73

Same constructor arguments, accessors, stream-tag keys/values and message-port names.
GNU Radio itself is not importable here, so the classes expose the block contract
(`history`, `output_multiple`, `general_work`/`work`, `consume_each`, tags) as plain
Python; `gr_emu.py` drives them the way the GR scheduler would.  All arithmetic runs in
the CUDA library (no CPU fallback).  pmt values are represented as: PMT_NIL -> None,
PMT_T/PMT_F -> True/False, pmt.from_long -> int.
"""
import ctypes as C
from collections import deque

import numpy as np

from . import _abi as A
from .engine import Trigger

SLOT_LENGTH = 960
HALF_FRAME_LENGTH = 10 * SLOT_LENGTH
SYMBOL_SZ = 128

TRACKING_LOST_TAG_KEY = "tracking_lost"     # lib/pss_impl.cc:39-40
CELL_ID_TAG_KEY = "cell_id"                 # lib/sss_impl.cc:38
CP_TYPE_TAG_KEY = "cp_type"                 # lib/sss_impl.cc:40


class tag_t:
    """gr::tag_t: absolute item offset, key, value."""
    __slots__ = ("offset", "key", "value")

    def __init__(self, offset, key, value):
        self.offset, self.key, self.value = offset, key, value

    def __repr__(self):
        return "tag_t(%d, %r, %r)" % (self.offset, self.key, self.value)


class _block:
    """The slice of gr::block both mirrors need."""

    def __init__(self, name):
        self._name = name
        self._history = 1
        self._output_multiple = 1
        self._nitems_read = 0
        self._nitems_written = 0
        self._consumed = 0
        self._out_tags = []
        self._in_tags = []

    def name(self): return self._name
    def history(self): return self._history
    def set_history(self, h): self._history = h
    def output_multiple(self): return self._output_multiple
    def set_output_multiple(self, m): self._output_multiple = m
    def nitems_read(self, port=0): return self._nitems_read
    def nitems_written(self, port=0): return self._nitems_written
    def consume_each(self, n): self._consumed = n
    def add_item_tag(self, port, offset, key, value): self._out_tags.append(tag_t(offset, key, value))

    def get_tags_in_window(self, port, rel_start, rel_end, key=None):
        a, b = self._nitems_read + rel_start, self._nitems_read + rel_end
        return [t for t in self._in_tags if a <= t.offset < b and (key is None or t.key == key)]


class pss(_block):
    """ltetrigger.pss(N_id_2, psr_threshold, track_after=16, track_every=8)

    One general_work call handles exactly one 9600-sample window, as in the reference
    (lib/pss_impl.cc:154-223): returns 0 or 9600 produced items, consumes 9600 (drop) or
    peak_pos-960+9600 (emit), tags the first item of every half-frame emitted while not
    tracking with "tracking_lost".  The forecast asks for the reference's history plus the
    largest possible consume (9599 + 18365 items) so a call never reads past its input.
    """

    def __init__(self, N_id_2, psr_threshold, track_after=16, track_every=8, device=0, max_chunk=1 << 18):
        _block.__init__(self, "pss")
        if N_id_2 not in (0, 1, 2):
            raise RuntimeError("Error initializing PSS N_id_2")          # lib/pss_impl.cc:75-76
        self._n_id_2 = N_id_2
        self._engine = Trigger(1, decim=1, psr_threshold=psr_threshold, max_chunk=max_chunk,
                               track_after=track_after, track_every=track_every, record_all=True,
                               keep_halfframes=True, device=device, root_mask=1 << N_id_2)
        # pss applies no clamp of its own (the hier block does): set the raw value
        self._engine.set_psr_threshold(psr_threshold, clamp=False)
        self._max_chunk = max_chunk
        self._pushed = 0                       # absolute count of samples handed to the engine
        self._queue = deque()                  # (record, halfframe or None) not yet released
        self.set_history(HALF_FRAME_LENGTH)    # :81
        self.set_output_multiple(HALF_FRAME_LENGTH)   # :82

    def forecast(self, noutput_items):
        return [self.history() - 1 + A.LOOKAHEAD]

    # accessors, lib/pss_impl.h:95-100
    def max_psr(self): return self._engine.stats(0, self._n_id_2).max_psr
    def mean_psr(self): return self._engine.stats(0, self._n_id_2).mean_psr
    def mean_cfo(self): return self._engine.stats(0, self._n_id_2).mean_cfo
    def psr_threshold(self): return self._engine.stats(0, self._n_id_2).psr_threshold
    def tracking_score(self): return self._engine.stats(0, self._n_id_2).tracking_score
    def set_psr_threshold(self, threshold): self._engine.set_psr_threshold(threshold, clamp=False)

    def _push(self, samples):
        step = 8
        n = len(samples) // step * step
        pos = 0
        while pos < n:
            take = min(n - pos, self._max_chunk)
            recs = self._engine.process(samples[None, pos:pos + take])
            n_emit = int(((recs["flags"] & A.F_EMIT) != 0).sum())
            hfs = self._engine.fetch_halfframes(n_emit) if n_emit else None
            k = 0
            for r in recs:
                if r["flags"] & A.F_EMIT:
                    self._queue.append((r.copy(), hfs[k].copy()))
                    k += 1
                else:
                    self._queue.append((r.copy(), None))
            pos += take
        self._pushed += n

    def general_work(self, noutput_items, ninput_items, input_items, output_items):
        """input_items[0]: complex64 view whose element history()-1 is the first new item."""
        inp = input_items[0]
        first_new = self.history() - 1
        avail_end = self._nitems_read + (ninput_items[0] - first_new)     # absolute end of readable input
        if avail_end > self._pushed:
            off = first_new + (self._pushed - self._nitems_read)
            self._push(np.ascontiguousarray(inp[off:off + (avail_end - self._pushed)], np.complex64))
        if not self._queue:
            self.consume_each(0)
            return 0
        rec, hf = self._queue.popleft()
        assert rec["win_start"] == self._nitems_read, (rec["win_start"], self._nitems_read)
        if rec["flags"] & A.F_EMIT:
            nconsume = int(rec["emit_start"] - rec["win_start"]) + HALF_FRAME_LENGTH
            output_items[0][:HALF_FRAME_LENGTH] = hf
            if rec["flags"] & A.F_TAG_LOST:
                self.add_item_tag(0, self.nitems_written(0), TRACKING_LOST_TAG_KEY, None)
            self.consume_each(nconsume)
            self.last_record = rec
            return HALF_FRAME_LENGTH
        self.consume_each(HALF_FRAME_LENGTH)
        self.last_record = rec
        return 0


class sss(_block):
    """ltetrigger.sss(N_id_2): gr::sync_block, one aligned half-frame per work call
    (lib/sss_impl.cc:83-156).  Consumes "tracking_lost", emits "cell_id" (int) and
    "cp_type" (True = normal) on item 0 of each decoded half-frame, passes samples through."""

    def __init__(self, N_id_2, device=0, frame_type=A.FRAME_FDD):
        _block.__init__(self, "sss")
        self._n_id_2 = N_id_2
        self._h = C.c_void_p()
        rc = A.lib().ltb_sss_create(device, N_id_2, C.byref(self._h))
        if rc != A.SUCCESS:
            raise RuntimeError(A.lib().ltb_last_error().decode() or "Error initializing SSS SYNC")
        if frame_type != A.FRAME_FDD:                  # TDD SSS position: not in the reference
            A.check(A.lib().ltb_sss_set_frame_type(self._h, frame_type), "ltb_sss_set_frame_type")
        self.set_output_multiple(HALF_FRAME_LENGTH)
        self.last_record = None

    def __del__(self):
        if getattr(self, "_h", None) is not None and self._h:
            A.lib().ltb_sss_destroy(self._h)
            self._h = None

    def work(self, noutput_items, input_items, output_items):
        inp = np.ascontiguousarray(input_items[0][:HALF_FRAME_LENGTH], np.complex64)
        lost = self.get_tags_in_window(0, 0, 1, TRACKING_LOST_TAG_KEY)
        rec = np.zeros(1, A.WINDOW_REC)
        rec["m0"] = rec["m1"] = rec["n_id_1"] = rec["cell_id"] = -1
        tag = np.array([1 if lost else 0], np.int32)
        A.check(A.lib().ltb_sss_work(self._h, inp.ctypes.data, tag.ctypes.data, 1, rec.ctypes.data), "ltb_sss_work")
        self.last_record = rec[0]
        if lost:
            output_items[0][:HALF_FRAME_LENGTH] = inp
            return HALF_FRAME_LENGTH
        if not (rec[0]["flags"] & A.F_CELL):
            return HALF_FRAME_LENGTH            # :119-120: no tags, output not written
        self.add_item_tag(0, self.nitems_written(0), CELL_ID_TAG_KEY, int(rec[0]["cell_id"]))
        self.add_item_tag(0, self.nitems_written(0), CP_TYPE_TAG_KEY, bool(rec[0]["flags"] & A.F_CP_NORM))
        output_items[0][:HALF_FRAME_LENGTH] = inp
        return HALF_FRAME_LENGTH


class mib(_block):
    """ltetrigger.mib(exit_on_success=False) -- host side, as in the reference
    (lib/mib_impl.cc:96-251; include/ltetrigger/mib.h).  Consumes half-frames tagged by sss,
    decodes the PBCH on the host (ltb_mib_decode, no GPU) and publishes the reference's cell
    dictionary on "track"; a "tracking_lost" tag publishes the same object on "drop".
    Message ports are lists of callbacks (`msg_connect`)."""

    PHICH_RESOURCES = ("1/6", "1/2", "1", "2")                 # lib/mib_impl.cc pack_cell

    def __init__(self, exit_on_success=False):
        _block.__init__(self, "mib")
        self._exit_on_success = exit_on_success
        self._published = False
        self._current = None
        self._ports = {"track": [], "drop": []}
        self.done = False                                      # WORK_DONE returned (exit_on_success)
        self.set_output_multiple(HALF_FRAME_LENGTH)

    def message_ports(self): return list(self._ports)
    def msg_connect(self, port, callback): self._ports[port].append(callback)

    def _pub(self, port, msg):
        for cb in self._ports[port]:
            cb(msg)

    def general_work(self, noutput_items, ninput_items, input_items, output_items):
        import time
        self.consume_each(HALF_FRAME_LENGTH)
        if self.get_tags_in_window(0, 0, 1, TRACKING_LOST_TAG_KEY):              # :107-120
            if self._published:
                self._pub("drop", self._current)
            self._published, self._current = False, None
            return 0
        if self._published:                                                        # :122-125
            return 0
        ids = self.get_tags_in_window(0, 0, 1, CELL_ID_TAG_KEY)
        cps = self.get_tags_in_window(0, 0, 1, CP_TYPE_TAG_KEY)
        if len(ids) != 1 or len(cps) != 1:                                         # :132-135
            return 0
        hf = np.ascontiguousarray(input_items[0][:HALF_FRAME_LENGTH], np.complex64)
        m = A.Mib()
        rc = A.lib().ltb_mib_decode(hf.ctypes.data, int(ids[0].value), int(bool(cps[0].value)), C.byref(m))
        if rc < 0:
            return 0                                                               # "SSS block passed us non-sense"
        if rc == 1:                                                                # SRSLTE_UE_MIB_FOUND :167-177
            self._current = {
                "cell_id": int(ids[0].value), "nof_tx_ports": int(m.nof_ports),
                "cp_len": "Normal" if cps[0].value else "Extended", "nof_prb": int(m.nof_prb),
                "phich_len": "Normal" if m.phich_length == 0 else "Extended",
                "nof_phich_resources": self.PHICH_RESOURCES[m.phich_resources],
                # the reference passes &d_sfn_offset as the `sfn` argument of srslte_pbch_mib_unpack
                # (lib/mib_impl.cc:167-172), which overwrites it with the MIB's 8-bit SFN field << 2
                # before pack_cell runs: that, not the 0..3 scrambling phase, is what the message carries
                "sfn_offset": int(m.sfn) & ~3, "tracking_start_time": int(time.time())}
            self._pub("track", self._current)
            self._published = True
            if self._exit_on_success:
                self.done = True
        return HALF_FRAME_LENGTH


class _chain_view:
    """What `downlink_trigger_c.pssK` exposes: the pss accessors of chain K of the fused engine
    (GRC probes poll e.g. pss0.tracking_score, examples/rtlsdr_ltetrigger.grc:735-752)."""

    def __init__(self, engine, k):
        self._e, self._k = engine, k

    def max_psr(self): return self._e.stats(0, self._k).max_psr
    def mean_psr(self): return self._e.stats(0, self._k).mean_psr
    def mean_cfo(self): return self._e.stats(0, self._k).mean_cfo
    def psr_threshold(self): return self._e.stats(0, self._k).psr_threshold
    def tracking_score(self): return self._e.stats(0, self._k).tracking_score
    def set_psr_threshold(self, t): self._e.set_psr_threshold(t, n_id_2=self._k, clamp=False)


MIN_PSR_THRESHOLD = 1.5  # python/downlink_trigger_c.py:10


class downlink_trigger_c:
    """Hier block: one complex input at 1.92 Msps, three pss->sss chains, message ports
    "track" and "drop" (python/downlink_trigger_c.py:18-61).

    The three pss -> sss chains run fused in one batched engine (one stream, root_mask 7); the
    three mib blocks stay on the host as in the reference (`mib0..2`, lib/mib_impl.cc) and are
    fed the emitted half-frames with their tags.  Their "track"/"drop" messages are forwarded
    to subscribers of the hier block's ports (python/downlink_trigger_c.py:47-61).  A custom
    `mib_sink(k, tags, halfframe)` may replace the mib stage (it returns (port, msg) pairs).
    """

    def __init__(self, psr_threshold, exit_on_success=False, device=0, max_chunk=1 << 18, keep_halfframes=True,
                 decim=1, input_format=A.FMT_FC32, frontend_mode=A.FRONTEND_FP32, fc32_full_scale=0.0):
        """`decim` > 1 fuses the `rational_resampler_ccc(1, decim)` the reference's apps put in
        front of the hier block (examples/cell_search_file.py:56-57) into the engine's front end;
        the input of work() is then at decim x 1.92 Msps.  `input_format` other than fc32 also fuses the
        interleaved-short / interleaved-char to complex conversion a flowgraph on an SDR's wire format starts with
        (work() then takes [n, 2] int16 / int8 arrays).  `frontend_mode = FRONTEND_TC_INT` runs that fused resampler as
        exact-integer GEMMs on the tensor cores (decim 2 ... 32, see include/ltetrigger_b200.h; fc32 input then needs
        `fc32_full_scale`, the range of the source).  Defaults: the reference's interface."""
        self.psr_threshold = self._ensure_safe_threshold(psr_threshold)
        self.exit_on_success = exit_on_success
        self._step = 8 * decim
        self._fmt = input_format
        self._engine = Trigger(1, decim=decim, psr_threshold=self.psr_threshold, max_chunk=max_chunk * decim,
                               record_all=True, keep_halfframes=keep_halfframes, device=device, input_format=input_format,
                               frontend_mode=frontend_mode, fc32_full_scale=fc32_full_scale)
        self._keep = keep_halfframes
        self.pss0, self.pss1, self.pss2 = (_chain_view(self._engine, k) for k in range(3))
        self._ports = {"track": [], "drop": []}
        self.mib_sink = None
        self.mib0, self.mib1, self.mib2 = (mib(exit_on_success) for _ in range(3))   # :33-35
        for m_ in (self.mib0, self.mib1, self.mib2):
            for port in ("track", "drop"):
                m_.msg_connect(port, lambda msg, port=port: [cb(msg) for cb in self._ports[port]])
        self._carry = np.zeros((0,) if input_format == A.FMT_FC32 else (0, 2), A.FMT_DTYPE[input_format])
        self.records = []

    def message_ports(self): return list(self._ports)
    def msg_connect(self, port, callback): self._ports[port].append(callback)

    def set_psr_threshold(self, t):
        t = self._ensure_safe_threshold(t)
        self.psr_threshold = t
        self._engine.set_psr_threshold(t, clamp=True)

    @staticmethod
    def _ensure_safe_threshold(t):
        return t if t > MIN_PSR_THRESHOLD else MIN_PSR_THRESHOLD

    def work(self, samples):
        """Consume a run of input items; returns the stream tags produced by the three sss
        blocks as (k, tag_t) pairs (offsets count items written by chain k's pss)."""
        x = np.concatenate([self._carry, np.asarray(samples, A.FMT_DTYPE[self._fmt])])
        n = len(x) // self._step * self._step
        self._carry = x[n:].copy()
        tags = []
        pos = 0
        while pos < n:
            take = min(n - pos, self._engine.max_chunk // self._step * self._step)
            recs = self._engine.process(x[None, pos:pos + take])
            hfs = None
            if self._keep:
                n_emit = int(((recs["flags"] & A.F_EMIT) != 0).sum())
                hfs = self._engine.fetch_halfframes(n_emit) if n_emit else None
            k_emit = 0
            for r in recs:
                self.records.append(r.copy())
                if not (r["flags"] & A.F_EMIT):
                    continue
                k = int(r["n_id_2"])
                off = getattr(self, "_written%d" % k, 0)
                setattr(self, "_written%d" % k, off + HALF_FRAME_LENGTH)
                these = []
                if r["flags"] & A.F_TAG_LOST:
                    these.append(tag_t(off, TRACKING_LOST_TAG_KEY, None))
                if r["flags"] & A.F_CELL:
                    these.append(tag_t(off, CELL_ID_TAG_KEY, int(r["cell_id"])))
                    these.append(tag_t(off, CP_TYPE_TAG_KEY, bool(r["flags"] & A.F_CP_NORM)))
                tags.extend((k, t) for t in these)
                if self.mib_sink is not None:
                    msgs = self.mib_sink(k, these, hfs[k_emit] if hfs is not None else None) or []
                    for port, msg in msgs:
                        for cb in self._ports[port]:
                            cb(msg)
                elif hfs is not None:
                    mb = (self.mib0, self.mib1, self.mib2)[k]
                    mb._in_tags, mb._nitems_read = these, off
                    mb.general_work(HALF_FRAME_LENGTH, [HALF_FRAME_LENGTH], [hfs[k_emit]], [None])
                k_emit += 1
            pos += take
        return tags


class cellstore:
    """ltetrigger.cellstore() -- message-only block keeping the tracked cells
    (lib/cellstore_impl.cc:46-105, include/ltetrigger/cellstore.h:59-65): "track" appends the
    cell, "drop" removes that same object; tracking() / cells() / latest_cell()."""

    def __init__(self):
        import threading
        self._cells = []
        self._mu = threading.Lock()                            # the reference guards its list too

    def message_ports(self): return ["track", "drop"]

    def track_cell(self, cell):
        with self._mu:
            self._cells.append(cell)

    def drop_cell(self, cell):
        with self._mu:
            self._cells = [c for c in self._cells if c is not cell]   # removal by identity (:103)

    def tracking(self):
        with self._mu:
            return len(self._cells) > 0

    def cells(self):
        with self._mu:
            return list(self._cells)

    def latest_cell(self):
        with self._mu:
            return self._cells[-1] if self._cells else None

    def connect(self, trigger):
        """msg_connect(trigger, "track"/"drop", self, ...) of examples/cell_search_file.py:82-88."""
        trigger.msg_connect("track", self.track_cell)
        trigger.msg_connect("drop", self.drop_cell)
        return self
