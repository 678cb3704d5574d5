"""TEST INFRASTRUCTURE: the oracle's restated blocks behind the interface of `ltetrigger_b200.Trigger`, so that the host
code above the C ABI (block mirrors, hier block, CLIs) can be exercised where there is no CUDA device.  Never imported by
the product; the CPU tests monkeypatch it in where the product constructs its engine.

Semantics follow the engine's: n_streams x (roots in root_mask) chains of oracle.Pss -> oracle.Sss driven call by call
under the scheduler rule (a window runs once 18365 samples from its start have arrived); records come back ordered by
(stream, N_id_2, call); `record_all = False` keeps emitted half-frames only; `fetch_halfframes` returns the emitted
(CFO-corrected) half-frames of the last call in record order; decim > 1 and sc16 / sc8 input go through the oracle's
restated resampler and conversions."""
import numpy as np


def make(oracle):
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import _abi as A

    class OracleTrigger:
        created = []                                             # constructor arguments, for tests that check them

        def __init__(self, n_streams, decim=1, psr_threshold=4.0, max_chunk=1 << 20, input_format=A.FMT_FC32,
                     track_after=16, track_every=8, record_all=True, keep_halfframes=False, device=0, root_mask=7,
                     corr_mode=A.CORR_DIRECT, **kw):
            OracleTrigger.created.append(dict(kw, n_streams=n_streams, decim=decim, input_format=input_format,
                                              root_mask=root_mask, record_all=record_all, keep_halfframes=keep_halfframes))
            self.n_streams, self.decim, self.input_format, self.max_chunk = n_streams, decim, input_format, max_chunk
            self.record_all, self.keep = bool(record_all), bool(keep_halfframes)
            self.roots = [k for k in range(3) if root_mask >> k & 1]
            conv = oracle.CONV_DIRECT                            # block level: one window at a time, whatever corr_mode asks for
            self.pss = [[oracle.Pss(k, psr_threshold, track_after, track_every, conv_mode=conv) for k in range(3)]
                        for _ in range(n_streams)]
            self.sss = [[oracle.Sss(k) for k in range(3)] for _ in range(n_streams)]
            self.raw = [np.zeros(0, np.complex64) for _ in range(n_streams)]
            self.n_y = [0] * n_streams
            self.buf = [np.zeros(960, np.complex64) for _ in range(n_streams)]   # the zero history GNU Radio puts in front
            self.pos = [[960, 960, 960] for _ in range(n_streams)]
            self.hfs = []

        def close(self):
            pass

        def _to_search_rate(self, s, x):
            if self.input_format == A.FMT_SC16:
                x = oracle.sc16_to_fc32(np.ascontiguousarray(x, np.int16))
            elif self.input_format == A.FMT_SC8:
                x = oracle.sc8_to_fc32(np.ascontiguousarray(x, np.int8))
            x = np.asarray(x, np.complex64)
            if self.decim > 1:                                   # y[k] depends on past input only: decimate all, keep the new outputs
                self.raw[s] = np.concatenate([self.raw[s], x])
                y = oracle.decimate(self.raw[s], self.decim)
                x, self.n_y[s] = y[self.n_y[s]:], len(y)
            return x

        def process(self, iq):
            assert len(iq) == self.n_streams and iq.shape[1] % (8 * self.decim) == 0 and iq.shape[1] <= self.max_chunk
            out = []
            for s in range(self.n_streams):
                self.buf[s] = np.concatenate([self.buf[s], self._to_search_rate(s, iq[s])])
                for k in self.roots:
                    while self.pos[s][k] - 960 + oracle.LOOKAHEAD <= len(self.buf[s]) - 960:
                        nout, ncons, hf, rec = self.pss[s][k].work(self.buf[s], self.pos[s][k])
                        rec = rec.copy()
                        rec["stream"] = s
                        rec["win_start"] = self.pos[s][k] - 960
                        rec["emit_start"] = self.pos[s][k] - 960 + rec["emit_start"] if nout else -1
                        if nout:
                            _, rec = self.sss[s][k].work(hf, bool(rec["flags"] & oracle.F_TAG_LOST), rec)
                        if nout or self.record_all:
                            out.append((rec, hf if nout else None))
                        self.pos[s][k] += ncons
            recs = np.zeros(len(out), A.WINDOW_REC)
            self.hfs = []
            for i, (rec, hf) in enumerate(out):
                for f in rec.dtype.names:
                    recs[i][f] = rec[f]
                if hf is not None:
                    self.hfs.append(hf)
            return recs

        def run(self, iq, chunk=None):
            step = 8 * self.decim
            n = iq.shape[1] - iq.shape[1] % step
            chunk = min(chunk or self.max_chunk, self.max_chunk) // step * step
            parts = [self.process(iq[:, a:min(a + chunk, n)]) for a in range(0, n, chunk)]
            recs = np.concatenate(parts) if parts else np.zeros(0, A.WINDOW_REC)
            return recs[np.lexsort((recs["win_index"], recs["n_id_2"], recs["stream"]))]

        def fetch_halfframes(self, n):
            assert self.keep and n <= len(self.hfs)
            return np.stack(self.hfs[:n]) if n else np.zeros((0, 9600), np.complex64)

        def stats(self, stream, k):
            st, p = A.PssStats(), self.pss[stream][k]
            st.max_psr, st.mean_psr, st.mean_cfo = p.max_psr(), p.mean_psr(), p.mean_cfo()
            st.psr_threshold, st.tracking_score = p.psr_threshold(), p.tracking_score()
            return st

        def set_psr_threshold(self, t, stream=-1, n_id_2=-1, clamp=True):
            if clamp:
                t = max(t, lt.MIN_PSR_THRESHOLD)
            for s in range(self.n_streams):
                for k in range(3):
                    if stream in (-1, s) and n_id_2 in (-1, k):
                        self.pss[s][k].set_psr_threshold(t)

    OracleTrigger.created = []
    return OracleTrigger
