"""A minimal stand-in for the GNU Radio scheduler (GNU Radio is not importable here), enough
to drive the block mirrors in blocks.py the way the reference's flowgraphs drive the real
blocks: file_source(repeat) -> head -> pss -> sss -> sink (python/qa_downlink_trigger_c.py:
85-100 wires the hier block; this runs one of its three chains with the blocks exposed).

Contract emulated: the input buffer handed to general_work starts history()-1 items before the
first unread item (zeros before the stream start, as GR pre-fills history); general_work is
called only when forecast()'s requirement is met; consume_each / the return value advance
nitems_read / nitems_written; tags added by a block travel with the items to the next block.
"""
import numpy as np

from .blocks import HALF_FRAME_LENGTH


class ChainTrace:
    """What one pss -> sss chain did, call by call."""

    def __init__(self):
        self.pss_calls = []      # (nitems_read, noutput, nconsumed, [tags])
        self.sss_calls = []      # (nitems_read, [in tags], [out tags])
        self.pss_out = []        # emitted half-frames
        self.sss_out = []        # half-frames after sss (None where sss did not write its output)


def run_chain(samples, pss_block, sss_block, max_calls=None):
    """Feed `samples` (complex64, 1.92 Msps) through pss_block -> sss_block until the scheduler
    can no longer satisfy pss's forecast.  Returns a ChainTrace."""
    x = np.ascontiguousarray(samples, np.complex64)
    hist = pss_block.history() - 1
    buf = np.concatenate([np.zeros(hist, np.complex64), x])          # GR zero-fills the history
    need = pss_block.forecast(HALF_FRAME_LENGTH)[0]
    tr = ChainTrace()
    out = np.zeros(HALF_FRAME_LENGTH, np.complex64)
    sss_buf = np.zeros(HALF_FRAME_LENGTH, np.complex64)
    calls = 0
    while True:
        r = pss_block.nitems_read(0)
        avail = len(buf) - r                     # items readable from (r - hist) on, history included
        if avail < need or (max_calls is not None and calls >= max_calls):
            break
        pss_block._out_tags = []
        nout = pss_block.general_work(HALF_FRAME_LENGTH, [avail], [buf[r:]], [out])
        ncons = pss_block._consumed
        tags = list(pss_block._out_tags)
        tr.pss_calls.append((r, nout, ncons, tags))
        pss_block._nitems_read += ncons
        calls += 1
        if nout:
            tr.pss_out.append(out[:nout].copy())
            # hand the half-frame and its tags to sss (tag offsets are absolute in pss's output)
            sss_block._in_tags = tags
            sss_block._nitems_read = pss_block._nitems_written
            sss_block._out_tags = []
            sss_buf[:] = np.nan                  # detect "output not written" (lib/sss_impl.cc:119-120)
            n = sss_block.work(HALF_FRAME_LENGTH, [out[:nout]], [sss_buf])
            assert n == HALF_FRAME_LENGTH
            tr.sss_calls.append((sss_block._nitems_read, tags, list(sss_block._out_tags)))
            tr.sss_out.append(None if np.isnan(sss_buf[0]) else sss_buf.copy())
            sss_block._nitems_written += n
            pss_block._nitems_written += nout
        elif ncons == 0:
            break                                # nothing evaluated and nothing consumed: starved
    return tr
