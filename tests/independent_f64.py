"""A second, independent statement of the search path in numpy float64 / complex128.

TEST INFRASTRUCTURE ONLY.  It shares no code with `oracle/` (C, float32, canonical summation orders
chosen together with the CUDA kernels) and none with the product: every formula here is written from
SURVEY.md Appendix A (the published srsLTE / GNU Radio algorithms) and from the reference's own block
code -- `lib/pss_impl.cc:94-223` (state machine, consume rule, CFO correction when tracking) and
`lib/sss_impl.cc:83-156` (reset on tag, CP detection, SSS decode, cell id).  It uses library routines
the oracle deliberately avoids (`np.convolve`, `np.fft`, `np.i0`) and double precision throughout, so
agreement of the oracle with it -- same windows, same decisions, PSR / peak values within the
north_star's 1e-4 -- says the oracle's float32 expression trees and hand-written FFTs evaluate the
same mathematics, not merely that oracle and kernels agree with each other.
"""
import numpy as np

SLOT, HALF, SYM, LAGS = 960, 9600, 128, 9726
ROOTS = (25, 29, 34)


def pss_filter(n_id_2):
    """Appendix A.1: Zadoff-Chu root -> 62 carriers around DC -> time domain -> conj / 62."""
    u = ROOTS[n_id_2]
    i = np.arange(31)
    d = np.concatenate([np.exp(-1j * np.pi * u * i * (i + 1) / 63.0),
                        np.exp(-1j * np.pi * u * (i + 32) * (i + 33) / 63.0)])
    bins = np.zeros(SYM, np.complex128)
    bins[np.arange(-31, 0) % SYM] = d[:31]
    bins[1:32] = d[31:]
    t = np.fft.ifft(bins) * SYM / np.sqrt(SYM)
    return np.conj(t) / 62.0


def decimator_taps(decim):
    """Appendix A.7: rational_resampler(1, D) default low-pass (Kaiser, beta 7), unit DC gain."""
    beta, rate = 7.0, 1.0 / decim
    tw = rate * 0.1
    fc = rate * 0.5 - tw / 2
    ntaps = int((beta / 0.1102 + 8.7) / (22.0 * tw))
    ntaps |= 1
    m = (ntaps - 1) // 2
    n = np.arange(-m, m + 1)
    w = np.i0(beta * np.sqrt(1 - (2.0 * np.arange(ntaps) / (ntaps - 1) - 1) ** 2)) / np.i0(beta)
    with np.errstate(invalid="ignore", divide="ignore"):
        taps = np.where(n == 0, 2 * fc, np.sin(2 * np.pi * fc * n) / (np.pi * n)) * w
    return taps / taps.sum()


def decimate(x, decim):
    """y[k] = sum_j taps[j] x[kD - j], zero history (lfilter(taps, 1, x)[::D])."""
    if decim == 1:
        return np.asarray(x, np.complex128)
    taps = decimator_taps(decim)
    x = np.concatenate([np.zeros(len(taps) - 1, np.complex128), np.asarray(x, np.complex128)])
    n_out = (len(x) - len(taps) + 1 + decim - 1) // decim
    x = np.concatenate([x, np.zeros(decim, np.complex128)])
    rows = np.lib.stride_tricks.as_strided(x, (n_out, len(taps)), (x.strides[0] * decim, x.strides[0]))
    return rows @ taps[::-1]


def _mseq(fb):
    x = [0, 0, 0, 0, 1]
    for i in range(26):
        x.append(sum(x[i + k] for k in fb) % 2)
    return 1 - 2 * np.array(x)


S_T, C_T, Z_T = _mseq((2, 0)), _mseq((3, 0)), _mseq((4, 2, 1, 0))


def n_id_1_table():
    tab = np.zeros((31, 31), int)        # srsLTE keeps the table in a zeroed struct: a pair that is no cell reads 0
    for n in range(168):
        qp = n // 30
        q = (n + qp * (qp + 1) // 2) // 30
        mp = n + q * (q + 1) // 2
        m0 = mp % 31
        m1 = (m0 + mp // 31 + 1) % 31
        tab[m0, m1] = n
    return tab


N_ID_1 = n_id_1_table()


class Chain:
    """One pss -> sss chain of downlink_trigger_c for one N_id_2, driven window by window."""

    def __init__(self, n_id_2, thr=4.0, track_after=16, track_every=8):
        self.n_id_2, self.thr, self.after, self.every = n_id_2, thr, track_after, track_every
        self.h = pss_filter(n_id_2)
        self.avg = np.zeros(LAGS + 3)
        self.tracking, self.score, self.timer, self.lost = False, 0, 0, False
        self.psr, self.peak, self.peak_value = 0.0, 0, 0.0
        self.cfo_hist, self.last_f, self.tab_phase = [], None, None
        self.cp_avg = [0.0, 0.0]
        i = np.arange(31)
        self.c0, self.c1 = C_T[(i + n_id_2) % 31], C_T[(i + n_id_2 + 3) % 31]

    # ---- Appendix A.2
    def find_pss(self, win):
        a = np.abs(np.convolve(win, self.h)[:LAGS]) ** 2
        self.avg[:LAGS] = 0.2 * a + 0.8 * self.avg[:LAGS]
        v = self.avg
        p = int(np.argmax(v[:LAGS]))
        ub = p + 1
        while ub < LAGS + 1 and v[ub + 1] <= v[ub]:
            ub += 1
        lb = 0
        if p > 2:
            lb = p - 1
            while lb > 1 and v[lb - 1] <= v[lb]:
                lb -= 1
        right = ub + (int(np.argmax(v[ub:LAGS])) if LAGS - ub > 0 else 0)
        left = int(np.argmax(v[:lb])) if lb > 0 else 0
        side = max(v[left], v[right])
        with np.errstate(invalid="ignore", divide="ignore"):
            psr = np.float64(v[p]) / np.float64(side)
        self.peak_value = float(v[p])
        return p, float(psr)

    def _reset_avg(self):
        self.avg[:] = 0.0

    # ---- lib/pss_impl.cc:111-152
    def _incr(self):
        if self.tracking and self.score == self.after:
            return
        self.score += 1
        if not self.tracking and self.score == self.after:
            self.tracking = True
            self._reset_avg()

    def _reset(self):
        if self.score == 0:
            return
        self.score, self.timer, self.tracking = 0, 0, False
        self._reset_avg()
        self.cfo_hist, self.last_f = [], None
        self.lost = True

    # ---- Appendix A.3
    def _cfo(self, r):
        y0, y1 = np.sum(self.h[:64] * r[:64]), np.sum(self.h[64:] * r[64:])
        return float(np.angle(np.conj(y0) * y1) / np.pi)

    def _correct(self, hf, f):
        idx = np.floor(np.mod(np.arange(HALF) * f * 4096.0, 4096.0))
        return hf * np.exp(2j * np.pi * idx / 4096.0)

    # ---- Appendix A.4
    def _detect_cp(self, hf):
        R, M = [0.0, 0.0], [0.0, 0.0]
        for k, cp in enumerate((9, 32)):
            j, r, c = SLOT - 3 * (SYM + cp), 0.0, 0.0
            for _ in range(3):
                r += np.sum(hf[j + SYM:j + SYM + cp] * np.conj(hf[j:j + cp])).real
                c += np.sum(np.abs(hf[j:j + cp]) ** 2)
                j += SYM + cp
            R[k], M[k] = r, (r / c if c else 0.0)
            self.cp_avg[k] = 0.1 * M[k] / 3 + 0.9 * self.cp_avg[k]
        if self.cp_avg[0] != self.cp_avg[1]:
            return self.cp_avg[0] > self.cp_avg[1]
        return R[0] > R[1]

    # ---- Appendix A.5 / A.6
    def _sss(self, sym):
        S = np.fft.fft(sym)
        v = np.concatenate([S[-31:], S[1:32]])
        i = np.arange(31)
        y0, y1 = v[0::2] * self.c0, v[1::2] * self.c1
        rot = (i[None, :] + i[:, None]) % 31                          # [m, i] -> (i + m) mod 31
        c0 = np.abs((y0[None, :] * S_T[rot]).sum(1)) ** 2
        m0 = int(np.argmax(c0))
        y1 = y1 * Z_T[(i + m0 % 8) % 31]
        c1 = np.abs((y1[None, :] * S_T[rot]).sum(1)) ** 2
        m1 = int(np.argmax(c1))
        lo, hi = (m0, m1) if m1 > m0 else (m1, m0)                    # A.6; the indices are unsigned there,
        nid = N_ID_1[lo, hi] if (lo < 30 and 1 <= hi <= 30) else -1   # so hi - 1 with hi = 0 is out of range
        return m0, m1, float(c0[m0]), float(c1[m1]), int(nid)

    # ---- lib/pss_impl.cc:154-223 + lib/sss_impl.cc:83-156; `x` has SLOT zeros in front of sample 0
    def window(self, x, pos):
        """One general_work call on the window starting at stream sample `pos`; returns (record, nconsume)."""
        base = pos + SLOT
        rec = dict(win_start=pos, searched=False, over=False, emit=False, tracking=False, tag_lost=False,
                   sss=False, cell_id=-1, n_id_1=-1, m0=-1, m1=-1, cp_norm=None, cfo=0.0, mean_cfo=0.0,
                   emit_start=-1)
        if not self.tracking or self.timer == 0:
            self.timer = self.every
            self.peak, self.psr = self.find_pss(x[base:base + HALF])
            rec["searched"] = True
        else:
            self.timer -= 1
        over = self.psr > self.thr
        rec["over"] = bool(over)
        if over:
            self._incr()
        else:
            self._reset()
        rec.update(peak_pos=self.peak, psr=self.psr, peak_value=self.peak_value, score=self.score)
        if not (over or self.lost):
            return rec, HALF
        start = self.peak - SLOT
        self.peak = SLOT
        hf = x[base + start:base + start + HALF]
        rec.update(emit=True, emit_start=pos + start)
        if self.tracking:
            rec["tracking"] = True
            cfo = self._cfo(hf[SLOT - SYM:SLOT])
            self.cfo_hist.append(cfo)
            mean = float(np.mean(self.cfo_hist[-200:]))    # :97-109: a 200-entry ring, mean of what it holds
            rec.update(cfo=cfo, mean_cfo=mean)
            hf = self._correct(hf, -mean / SYM)
            tag = False
        else:
            rec["tag_lost"] = tag = True
            self.lost = False
        if tag:
            self.cp_avg = [0.0, 0.0]
        else:
            norm = bool(self._detect_cp(hf))
            cp = 9 if norm else 32
            i0 = SLOT - 2 * SYM - cp
            m0, m1, v0, v1, nid = self._sss(hf[i0:i0 + SYM])
            rec.update(sss=True, cp_norm=norm, m0=m0, m1=m1, m0_val=v0, m1_val=v1, n_id_1=nid,
                       cell_id=3 * nid + self.n_id_2 if nid >= 0 else -1)
        return rec, start + HALF

    def run(self, y):
        """Every window of the stream `y` (search-rate samples) under the scheduler rule the oracle
        and the engine use: a window is evaluated once 18365 samples from its start are present."""
        x = np.concatenate([np.zeros(SLOT, np.complex128), np.asarray(y, np.complex128)])
        pos, out = 0, []
        while pos + 18365 <= len(y):
            rec, n = self.window(x, pos)
            out.append(rec)
            pos += n
        return out
