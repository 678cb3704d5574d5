"""Seeded synthetic LTE FDD downlink captures (BASELINE config C4/C5).

Normal-CP (or, with ext_cp, extended-CP) frames with the 36.211 PSS (last symbol of slots 0
and 10) and SSS (the symbol before it, SF0/SF5 variants) on the 62 centre carriers and unit-power QPSK on every other
occupied resource element, OFDM-modulated at 128*decim points (1.92*decim Msps), cut at a
random timing offset, with complex AWGN at a requested SNR and an optional carrier offset.
Models examples/snr_ltetrigger.grc (fixture x gain + Gaussian noise) of the reference.
Pure numpy: this is input generation, not part of the search path.
"""
import numpy as np

PSS_ROOTS = (25, 29, 34)


def _mseq(taps):
    x = [0, 0, 0, 0, 1] + [0] * 26
    for i in range(26):
        x[i + 5] = sum(x[i + t] for t in taps) % 2
    return 1 - 2 * np.array(x[:31])


S_TILDE, C_TILDE, Z_TILDE = _mseq((2, 0)), _mseq((3, 0)), _mseq((4, 2, 1, 0))


def pss_freq(n_id_2):
    u = PSS_ROOTS[n_id_2]
    n = np.arange(62)
    k = np.where(n < 31, n * (n + 1), (n + 1) * (n + 2))
    return np.exp(-1j * np.pi * u * k / 63.0)


def m0m1(n_id_1):
    qp = n_id_1 // 30
    q = (n_id_1 + qp * (qp + 1) // 2) // 30
    mp = n_id_1 + q * (q + 1) // 2
    m0 = mp % 31
    return m0, (m0 + mp // 31 + 1) % 31


def sss_freq(cell_id, subframe):
    n_id_1, n_id_2 = cell_id // 3, cell_id % 3
    m0, m1 = m0m1(n_id_1)
    n = np.arange(31)
    s0, s1 = S_TILDE[(n + m0) % 31], S_TILDE[(n + m1) % 31]
    c0, c1 = C_TILDE[(n + n_id_2) % 31], C_TILDE[(n + n_id_2 + 3) % 31]
    z0, z1 = Z_TILDE[(n + m0 % 8) % 31], Z_TILDE[(n + m1 % 8) % 31]
    d = np.zeros(62)
    if subframe == 0:
        d[0::2], d[1::2] = s0 * c0, s1 * c1 * z0
    else:
        d[0::2], d[1::2] = s1 * c0, s0 * c1 * z1
    return d


N_USED = {1: 72, 2: 180, 4: 300, 8: 600, 12: 900, 16: 1200}


# ---- PBCH / CRS transmitter (36.211 6.6, 6.10.1, 7.2; 36.212 5.1.1, 5.1.3.1, 5.1.4.2, 5.3.1) -------------
# Used to make synthetic cells whose MIB the host-side decoder can read, with one or two antenna
# ports (transmit diversity, 36.211 6.3.4.3).  Input generation only.
_PERM = (1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31,
         0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30)
_PRB_CODE = {6: 0, 15: 1, 25: 2, 50: 3, 75: 4, 100: 5}
_CRC_MASK = {1: 0x0000, 2: 0xFFFF, 4: 0x5555}


def gold(c_init, n):
    x1 = [1] + [0] * 30
    x2 = [(c_init >> i) & 1 for i in range(31)]
    for i in range(1600 + n):
        x1.append(x1[i + 3] ^ x1[i])
        x2.append(x2[i + 3] ^ x2[i + 2] ^ x2[i + 1] ^ x2[i])
    return np.array([x1[i + 1600] ^ x2[i + 1600] for i in range(n)], np.int64)


def _crc16(bits):
    reg = 0
    for b in bits:
        fb = ((reg >> 15) & 1) ^ int(b)
        reg = (reg << 1) & 0xFFFF
        if fb:
            reg ^= 0x1021
    return reg


def _tbcc(bits):
    """Tail-biting K = 7 code, generators 133 / 171 / 165 (octal); output index 3 i + s."""
    g = (0o133, 0o171, 0o165)
    n = len(bits)
    state = 0
    for b in bits[-6:]:                       # the register starts with the last six information bits
        state = (state >> 1) | (int(b) << 5)
    out = np.zeros(3 * n, np.int64)
    for i, b in enumerate(bits):
        reg = (int(b) << 6) | state           # newest bit in the MSB
        for s_ in range(3):
            out[3 * i + s_] = bin(reg & g[s_]).count("1") & 1
        state = reg >> 1
    return out


def _ratematch_order():
    order = []
    for s_ in range(3):
        for j in range(32):
            for r in range(2):
                y = r * 32 + _PERM[j]
                if y >= 24:                   # 24 dummy bits in front of the 40
                    order.append(3 * (y - 24) + s_)
    return np.array(order)


def mib_bits(nof_prb, phich_ext, phich_res, sfn):
    b = [(_PRB_CODE[nof_prb] >> 2) & 1, (_PRB_CODE[nof_prb] >> 1) & 1, _PRB_CODE[nof_prb] & 1, int(phich_ext),
         (phich_res >> 1) & 1, phich_res & 1]
    b += [((sfn >> 2) >> (7 - i)) & 1 for i in range(8)]
    return b + [0] * 10


def pbch_symbols(cell_id, mib24, n_ports, n_bits):
    """The QPSK symbols of one 40 ms PBCH period (n_bits = 1920 normal CP, 1728 extended)."""
    crc = _crc16(mib24) ^ _CRC_MASK[n_ports]
    a = list(mib24) + [(crc >> (15 - i)) & 1 for i in range(16)]
    coded = _tbcc(a)
    order = _ratematch_order()
    e = coded[order[np.arange(n_bits) % 120]] ^ gold(cell_id, n_bits)
    return ((1 - 2 * e[0::2]) + 1j * (1 - 2 * e[1::2])) / np.sqrt(2.0)


def crs_central(cell_id, ns, l, ext_cp):
    """CRS values r(m') for the twelve pilots of the six central resource blocks (m' = 104 .. 115)."""
    ncp = 0 if ext_cp else 1
    c = gold(1024 * (7 * (ns + 1) + l + 1) * (2 * cell_id + 1) + 2 * cell_id + ncp, 2 * 220)
    m = np.arange(104, 116)
    return ((1 - 2 * c[2 * m]) + 1j * (1 - 2 * c[2 * m + 1])) / np.sqrt(2.0)


def _central(n):
    """central-72 subcarrier index 0..71 (lowest frequency first, DC skipped) -> signed FFT bin"""
    return np.where(n < 36, n - 36, n - 35)



def lte_frame(cell_id, decim=1, rng=None, n_frames=1, ext_cp=False, tdd=False, mib=None):
    """n_frames radio frames (19200*decim samples each) at 1.92*decim Msps, mean power ~1.
    ext_cp: 6 symbols per slot with a 32*decim-sample prefix (36.211 table 6.12-1).
    tdd: frame structure type 2 -- PSS in symbol 2 of slots 2 and 12, SSS in the last symbol of
    slots 1 and 11 (36.211 6.11.1.2, 6.11.2.2); every subframe is filled like a downlink one.
    mib: dict(nof_prb, n_ports=1|2|4, phich_ext=0, phich_res=2, sfn0=0, h=(h0, h1, h2, h3)) adds the cell-specific
    reference signals of the central six resource blocks and the PBCH (FDD position: slot 1 of
    subframe 0), for one antenna port, or two / four with transmit diversity (SFBC / SFBC-FSTD) through flat
    channels h."""
    rng = rng or np.random.default_rng(cell_id)
    nfft = 128 * decim
    n_used = N_USED.get(decim, 12 * (6 * decim - decim // 2))
    half = n_used // 2
    cp0, cp = (32 * decim, 32 * decim) if ext_cp else (10 * decim, 9 * decim)
    per_slot = 6 if ext_cp else 7
    nsym = 20 * per_slot * n_frames
    bits = rng.integers(0, 2, size=(nsym, n_used, 2)) * 2 - 1
    qpsk = (bits[..., 0] + 1j * bits[..., 1]) / np.sqrt(2.0)
    grid = np.zeros((nsym, nfft), np.complex128)
    grid[:, 1:half + 1] = qpsk[:, half:]
    grid[:, nfft - half:] = qpsk[:, :half]
    pss = pss_freq(cell_id % 3)
    for f in range(n_frames):
        for slot, sf in ((0, 0), (10, 5)):
            if tdd:
                sym_pss = (f * 20 + slot + 2) * per_slot + 2
                sym_sss = (f * 20 + slot + 1) * per_slot + per_slot - 1
            else:
                sym_pss = (f * 20 + slot) * per_slot + per_slot - 1
                sym_sss = sym_pss - 1
            for sym, seq in ((sym_pss, pss), (sym_sss, sss_freq(cell_id, sf))):
                grid[sym, 1:37] = 0
                grid[sym, nfft - 36:] = 0
                grid[sym, 1:32] = seq[31:]
                grid[sym, nfft - 31:] = seq[:31]
    if mib is not None:
        grid = _add_crs_pbch(grid, cell_id, n_frames, per_slot, ext_cp, nfft, mib)
    time = np.fft.ifft(grid, axis=1) * (nfft / np.sqrt(n_used))
    out = np.empty(n_frames * 19200 * decim, np.complex128)
    pos = 0
    for s in range(nsym):
        c = cp0 if s % per_slot == 0 else cp
        out[pos:pos + c] = time[s, nfft - c:]
        out[pos + c:pos + c + nfft] = time[s]
        pos += c + nfft
    assert pos == len(out)
    return out


def _add_crs_pbch(grid, cell_id, n_frames, per_slot, ext_cp, nfft, mib):
    n_ports = mib.get("n_ports", 1)
    h = mib.get("h", (1.0, 0.6 - 0.5j, -0.3 + 0.7j, 0.5 + 0.4j))
    ports = [grid.copy()] + [np.zeros_like(grid) for _ in range(3)]   # payload, PSS and SSS leave from port 0 only
    cols = _central(np.arange(72)) % nfft
    vshift = cell_id % 6
    n_bits = 1728 if ext_cp else 1920
    per_frame = n_bits // 8                           # QPSK symbols per radio frame
    for f in range(n_frames):
        sfn = mib.get("sfn0", 0) + f
        d = pbch_symbols(cell_id, mib_bits(mib["nof_prb"], mib.get("phich_ext", 0), mib.get("phich_res", 2), sfn),
                         n_ports, n_bits)[(sfn % 4) * per_frame:(sfn % 4 + 1) * per_frame]
        for ns in range(20):
            # 36.211 6.10.1.2: ports 0 / 1 in symbols 0 and N_symb - 3 (v = 0, 3 / 3, 0); ports 2 / 3 in symbol 1
            # (v = 3 (n_s mod 2) / 3 + 3 (n_s mod 2)); a pilot RE of one port is empty on every other port
            for l, vs in ((0, (0, 3, None, None)), (per_slot - 3, (3, 0, None, None)),
                          (1, (None, None, 3 * (ns % 2), 3 + 3 * (ns % 2)))):
                sym = (f * 20 + ns) * per_slot + l
                r = crs_central(cell_id, ns, l, ext_cp)
                for p in range(4):
                    if vs[p] is None:
                        continue
                    k = 6 * np.arange(12) + (vs[p] + vshift) % 6
                    if p >= 2 and n_ports < 4:
                        continue                      # one- and two-port cells carry data in these REs
                    for q in range(4):
                        ports[q][sym, cols[k]] = 0
                    if p < n_ports:
                        ports[p][sym, cols[k]] = r
        # PBCH: slot 1, symbols 0..3, the CRS positions of four ports stay empty
        res = []
        for l in range(4):
            crs_sym = l in (0, 1) or (ext_cp and l == 3)
            for k in range(72):
                if crs_sym and k % 3 == cell_id % 3:
                    continue
                res.append(((f * 20 + 1) * per_slot + l, cols[k]))
        assert len(res) == per_frame
        for sym_, col_ in res:
            for q in range(1, 4):
                ports[q][sym_, col_] = 0
        r2 = np.sqrt(2)
        if n_ports == 1:
            for i in range(per_frame):
                ports[0][res[i]] = d[i]
        elif n_ports == 2:                            # 36.211 6.3.4.3, two ports
            for i in range(0, per_frame, 2):
                ports[0][res[i]], ports[0][res[i + 1]] = d[i] / r2, d[i + 1] / r2
                ports[1][res[i]], ports[1][res[i + 1]] = -np.conj(d[i + 1]) / r2, np.conj(d[i]) / r2
        else:                                         # four ports: SFBC on ports (0, 2), then on ports (1, 3)
            for i in range(0, per_frame, 4):
                ports[0][res[i]] = ports[0][res[i + 1]] = ports[0][res[i + 2]] = ports[0][res[i + 3]] = 0
                ports[0][res[i]], ports[0][res[i + 1]] = d[i] / r2, d[i + 1] / r2
                ports[2][res[i]], ports[2][res[i + 1]] = -np.conj(d[i + 1]) / r2, np.conj(d[i]) / r2
                ports[1][res[i + 2]], ports[1][res[i + 3]] = d[i + 2] / r2, d[i + 3] / r2
                ports[3][res[i + 2]], ports[3][res[i + 3]] = -np.conj(d[i + 3]) / r2, np.conj(d[i + 2]) / r2
    return sum(h[q] * ports[q] for q in range(n_ports))


def capture(cell_id, n_samples, snr_db=None, decim=1, seed=0, offset=None, cfo_hz=0.0, noise_only=False,
            ext_cp=False, tdd=False, mib=None):
    """One capture of n_samples at 1.92*decim Msps as complex64."""
    rng = np.random.default_rng([seed, cell_id, 0x5EED])
    frame_len = 19200 * decim
    if offset is None:
        offset = int(rng.integers(0, frame_len))
    n_frames = (offset + n_samples + frame_len - 1) // frame_len
    # a few distinct frames tiled keeps generation cheap while payload still varies
    uniq = min(n_frames, 4)
    base = lte_frame(cell_id, decim, rng, uniq, ext_cp, tdd, mib)
    reps = (n_frames + uniq - 1) // uniq
    sig = np.tile(base, reps)[offset:offset + n_samples]
    if cfo_hz:
        fs = 1.92e6 * decim
        sig = sig * np.exp(2j * np.pi * cfo_hz / fs * np.arange(n_samples))
    if noise_only:
        sig = np.zeros_like(sig)
    if snr_db is not None:
        p_sig = 1.0
        sigma = np.sqrt(p_sig / (10.0 ** (snr_db / 10.0)) / 2.0)
        sig = sig + sigma * (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples))
    return sig.astype(np.complex64)


def batch(n_streams, n_samples, snr_db, decim=1, master_seed=1234, cell_ids=None):
    """[n_streams, n_samples] complex64; cell ids dealt from a seeded permutation of 0..503."""
    perm = np.random.default_rng(master_seed).permutation(504)
    ids = np.array([perm[i % 504] for i in range(n_streams)]) if cell_ids is None else np.asarray(cell_ids)
    out = np.empty((n_streams, n_samples), np.complex64)
    for i in range(n_streams):
        out[i] = capture(int(ids[i]), n_samples, snr_db, decim, seed=master_seed ^ i)
    return out, ids


def to_sc8(x, full_scale=4.0):
    """Quantise complex64 to interleaved int8 (I, Q) with |x| = full_scale -> 127."""
    s = 127.0 / full_scale
    iq = np.empty(x.shape + (2,), np.int8)
    iq[..., 0] = np.clip(np.round(x.real * s), -128, 127)
    iq[..., 1] = np.clip(np.round(x.imag * s), -128, 127)
    return iq


def to_sc16(x, full_scale=4.0):
    """Quantise complex64 to interleaved int16 (I, Q) with |x| = full_scale -> 32767."""
    s = 32767.0 / full_scale
    iq = np.empty(x.shape + (2,), np.int16)
    iq[..., 0] = np.clip(np.round(x.real * s), -32768, 32767)
    iq[..., 1] = np.clip(np.round(x.imag * s), -32768, 32767)
    return iq
