// Host-side MIB/PBCH decode for one aligned, CFO-corrected half-frame at 1.92 Msps.
//
// The reference keeps this step on the host (lib/mib_impl.cc:96-183 calls srsLTE's
// srslte_ue_mib_decode + srslte_pbch_mib_unpack), and so does this repo: it is the consumer of
// the GPU path's output (aligned half-frames tagged with cell_id / cp_type), not part of the hot
// path, and it never touches the GPU.  srsLTE's source is absent, so this restates the 3GPP
// procedure it implements -- 36.211 6.6 (PBCH), 6.10.1 (CRS), 7.2 (Gold sequence); 36.212 5.1.1
// (CRC16), 5.1.3.1 (tail-biting convolutional code), 5.1.4.2 (rate matching), 5.3.1 (BCH) -- for
// one antenna port and for two and four ports with transmit diversity (36.211 6.3.4.3: SFBC, and SFBC-FSTD
// on the port pairs (0, 2) / (1, 3)).  Its results are pinned by the
// reference's own tests: nof_prb 6 / 25 / 50 / 100, phich_len Normal, nof_phich_resources "1",
// nof_tx_ports 1 for the four bundled test_frames (python/qa_downlink_trigger_c.py:46-65).
#include <cmath>
#include <complex>
#include <cstring>
#include <vector>

#include "../../include/ltetrigger_b200.h"

namespace {

typedef std::complex<float> cf;

// 36.211 7.2: length-31 Gold sequence
void gold(uint32_t c_init, int len, std::vector<uint8_t> &c) {
  const int Nc = 1600;
  std::vector<uint8_t> x1(Nc + len + 31), x2(Nc + len + 31);
  for (int i = 0; i < 31; ++i) { x1[i] = (i == 0); x2[i] = (c_init >> i) & 1u; }
  for (int n = 0; n < Nc + len; ++n) {
    x1[n + 31] = x1[n + 3] ^ x1[n];
    x2[n + 31] = x2[n + 3] ^ x2[n + 2] ^ x2[n + 1] ^ x2[n];
  }
  c.resize(len);
  for (int n = 0; n < len; ++n) c[n] = x1[n + Nc] ^ x2[n + Nc];
}

// forward DFT of 128 samples, only the 72 occupied subcarriers (k = -36..-1, 1..36)
void demod72(const cf *x, cf out[72]) {
  static float cs[128], sn[128];
  static bool init = false;
  if (!init) {
    for (int i = 0; i < 128; ++i) { cs[i] = (float)std::cos(2.0 * M_PI * i / 128.0); sn[i] = (float)std::sin(2.0 * M_PI * i / 128.0); }
    init = true;
  }
  for (int n = 0; n < 72; ++n) {
    const int bin = n < 36 ? 128 - 36 + n : n - 36 + 1;
    double ar = 0.0, ai = 0.0;
    for (int t = 0; t < 128; ++t) {
      const int ph = (bin * t) & 127;                      // e^{-j 2 pi bin t / 128}
      ar += (double)x[t].real() * cs[ph] + (double)x[t].imag() * sn[ph];
      ai += (double)x[t].imag() * cs[ph] - (double)x[t].real() * sn[ph];
    }
    out[n] = cf((float)ar, (float)ai);
  }
}

// 36.212 5.1.4.2.1 column permutation of the convolutional-code sub-block interleaver
const int kPerm[32] = {1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31,
                       0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30};

// order in which the 120 coded bits (stream s, index i -> 3*i + s) leave the circular buffer
void ratematch_order(int order[120]) {
  const int D = 40, C = 32, R = 2, ND = R * C - D;         // 24 dummy bits in front
  int n = 0;
  for (int s = 0; s < 3; ++s)
    for (int j = 0; j < C; ++j)
      for (int r = 0; r < R; ++r) {
        const int y = r * C + kPerm[j];                    // position in the row-wise written matrix
        if (y >= ND) order[n++] = 3 * (y - ND) + s;
      }
}

// tail-biting Viterbi, K = 7, G = (133, 171, 165) octal, 40 information bits; llr > 0 means bit 0
void viterbi_tb(const float llr[120], uint8_t bits[40]) {
  const int L = 40, REP = 3, NS = 64;
  static int out[NS][2];                                   // 3 output bits for (state, input)
  static bool init = false;
  if (!init) {
    const int g[3] = {0133, 0171, 0165};
    for (int st = 0; st < NS; ++st)
      for (int b = 0; b < 2; ++b) {
        const int reg = (b << 6) | st;                     // newest bit in the MSB
        int o = 0;
        for (int k = 0; k < 3; ++k) o |= (__builtin_parity(reg & g[k])) << k;
        out[st][b] = o;
      }
    init = true;
  }
  const int T = L * REP;
  std::vector<float> metric(NS, 0.f), next(NS);
  std::vector<uint8_t> surv((size_t)T * NS);
  for (int t = 0; t < T; ++t) {
    const float *l = llr + 3 * (t % L);
    for (int ns = 0; ns < NS; ++ns) next[ns] = -1e30f;
    for (int st = 0; st < NS; ++st)
      for (int b = 0; b < 2; ++b) {
        const int o = out[st][b];
        float m = metric[st];
        for (int k = 0; k < 3; ++k) m += ((o >> k) & 1) ? -l[k] : l[k];
        const int ns = (b << 5) | (st >> 1);
        if (m > next[ns]) { next[ns] = m; surv[(size_t)t * NS + ns] = (uint8_t)st; }
      }
    metric.swap(next);
  }
  int st = 0;
  for (int s2 = 1; s2 < NS; ++s2) if (metric[s2] > metric[st]) st = s2;
  std::vector<uint8_t> dec(T);
  for (int t = T - 1; t >= 0; --t) { dec[t] = (uint8_t)(st >> 5); st = surv[(size_t)t * NS + st]; }
  for (int i = 0; i < L; ++i) bits[i] = dec[L + i];        // the middle repetition
}

uint32_t crc16(const uint8_t *bits, int n) {               // gCRC16 = D^16 + D^12 + D^5 + 1
  uint32_t reg = 0;
  for (int i = 0; i < n; ++i) {
    const uint32_t fb = ((reg >> 15) & 1u) ^ bits[i];
    reg = (reg << 1) & 0xFFFFu;
    if (fb) reg ^= 0x1021u;
  }
  return reg;
}

}  // namespace

extern "C" int ltb_mib_decode(const ltb_cf *halfframe, int cell_id, int cp_normal, ltb_mib *out) {
  if (!halfframe || !out || cell_id < 0 || cell_id > 503) return LTB_ERROR_INVALID_INPUTS;
  const cf *x = reinterpret_cast<const cf *>(halfframe);
  const int nsym = cp_normal ? 7 : 6;
  const int ncp = cp_normal ? 1 : 0;
  auto sym_start = [&](int slot, int l) {                  // first sample after the CP
    return cp_normal ? slot * 960 + 10 + 137 * l : slot * 960 + 32 + 160 * l;
  };
  // ---- OFDM demodulation of slot 1, and of symbol 1 of slot 0 (second comb of ports 2 / 3) ----------
  cf grid[7][72], grid01[72];
  for (int l = 0; l < nsym; ++l) demod72(x + sym_start(1, l), grid[l]);
  demod72(x + sym_start(0, 1), grid01);
  // ---- channel estimates from the CRS of slot 1 (36.211 6.10.1) ------------------------------------
  // port 0: symbols l = 0 (v = 0) and l = nsym - 3 (v = 3); port 1: the mirrored comb (v = 3, 0);
  // subcarriers k = 6 m + (v + v_shift) % 6
  // ports 2 / 3: symbol l = 1 of every slot, v = 3 (n_s mod 2) / 3 + 3 (n_s mod 2): slots 0 and 1 give the two combs
  const int vshift = cell_id % 6;
  cf hk[4][72];
  for (int port = 0; port < 4; ++port) {
    std::vector<int> pos;
    std::vector<cf> val;
    const int ls[2] = {port < 2 ? 0 : 1, port < 2 ? nsym - 3 : 1};
    const int nss[2] = {port < 2 ? 1 : 0, 1};
    const int vs[2] = {port == 0 ? 0 : port == 1 ? 3 : port == 2 ? 0 : 3, port == 0 ? 3 : port == 1 ? 0 : port == 2 ? 3 : 6};
    for (int a = 0; a < 2; ++a) {
      std::vector<uint8_t> c;
      const int ns = nss[a], l = ls[a];
      const cf *g = (ns == 1) ? grid[l] : grid01;
      const uint32_t c_init = 1024u * (7u * (ns + 1) + l + 1) * (2u * cell_id + 1) + 2u * cell_id + ncp;
      gold(c_init, 2 * 220, c);
      for (int m = 0; m < 12; ++m) {
        const int k = 6 * m + (vs[a] + vshift) % 6;
        const int mp = m + 110 - 6;                        // central 6 RB of any bandwidth
        const cf r((1 - 2 * (int)c[2 * mp]) * (float)M_SQRT1_2, (1 - 2 * (int)c[2 * mp + 1]) * (float)M_SQRT1_2);
        pos.push_back(k);
        val.push_back(g[k] * std::conj(r));
      }
    }
    // merge the two staggered combs (static channel over the slot), sort by subcarrier, interpolate
    for (size_t i = 0; i < pos.size(); ++i)
      for (size_t j = i + 1; j < pos.size(); ++j)
        if (pos[j] < pos[i]) { std::swap(pos[i], pos[j]); std::swap(val[i], val[j]); }
    for (int k = 0; k < 72; ++k) {
      size_t j = 0;
      while (j + 1 < pos.size() && pos[j + 1] <= k) ++j;
      const size_t j2 = (j + 1 < pos.size()) ? j + 1 : j;
      cf h = val[j];
      if (j2 != j) {
        const float t = (float)(k - pos[j]) / (float)(pos[j2] - pos[j]);
        h = val[j] + (val[j2] - val[j]) * t;               // also extrapolates below the first pilot
      }
      hk[port][k] = h;
    }
  }
  // ---- PBCH resource elements: symbols 0..3, all CRS positions of ports 0..3 reserved ----------
  std::vector<cf> rx, h0, h1, h2, h3;
  for (int l = 0; l < 4; ++l)
    for (int k = 0; k < 72; ++k) {
      const bool crs_sym = (l == 0 || l == 1 || (!cp_normal && l == 3));
      if (crs_sym && (k % 3) == (cell_id % 3)) continue;
      rx.push_back(grid[l][k]);
      h0.push_back(hk[0][k]);
      h1.push_back(hk[1][k]);
      h2.push_back(hk[2][k]);
      h3.push_back(hk[3][k]);
    }
  const int nre = (int)rx.size();                          // 240 (normal) / 216 (extended)
  const int E = 2 * nre;
  std::vector<uint8_t> c;
  gold((uint32_t)cell_id, 4 * E, c);
  int order[120];
  ratematch_order(order);
  // srslte_pbch_decode tries 1, 2 and 4 antenna ports and accepts a CRC that matches that number's mask
  for (int nant = 1; nant <= 4; nant *= 2) {
    std::vector<float> llr((size_t)E);
    if (nant == 1) {
      for (int i = 0; i < nre; ++i) {
        const cf z = rx[i] * std::conj(h0[i]);             // matched filter (scale does not matter)
        llr[2 * i] = z.real(); llr[2 * i + 1] = z.imag();
      }
    } else {
      // 36.211 6.3.4.3: r(2i) = ha d(2i) - hb conj(d(2i+1)),  r(2i+1) = ha d(2i+1) + hb conj(d(2i)); two ports:
      // (ha, hb) = (h0, h1) on every pair; four ports: (h0, h2) on the pairs 4i, 4i+1 and (h1, h3) on 4i+2, 4i+3
      for (int i = 0; i + 1 < nre; i += 2) {
        const bool second = nant == 4 && (i & 2);
        const std::vector<cf> &ha = nant == 2 ? h0 : second ? h1 : h0, &hb = nant == 2 ? h1 : second ? h3 : h2;
        const cf a0 = 0.5f * (ha[i] + ha[i + 1]), a1 = 0.5f * (hb[i] + hb[i + 1]);
        const cf d0 = std::conj(a0) * rx[i] + a1 * std::conj(rx[i + 1]);
        const cf d1 = std::conj(a0) * rx[i + 1] - a1 * std::conj(rx[i]);
        llr[2 * i] = d0.real(); llr[2 * i + 1] = d0.imag();
        llr[2 * i + 2] = d1.real(); llr[2 * i + 3] = d1.imag();
      }
    }
    // ---- scrambling phase hypotheses (frame number mod 4), de-ratematch, decode, CRC --------------
    for (int off = 0; off < 4; ++off) {
      float soft[120];
      std::memset(soft, 0, sizeof soft);
      for (int k = 0; k < E; ++k) {
        const int kk = off * E + k;                        // position inside the 4-frame codeword
        const float v = c[kk] ? -llr[k] : llr[k];
        soft[order[kk % 120]] += v;
      }
      uint8_t bits[40];
      viterbi_tb(soft, bits);
      bool any = false;
      for (int i = 0; i < 24; ++i) any |= bits[i] != 0;
      if (!any) continue;                                  // all-zero payload passes the CRC trivially
      uint32_t rxcrc = 0;
      for (int i = 0; i < 16; ++i) rxcrc = (rxcrc << 1) | bits[24 + i];
      const uint32_t diff = crc16(bits, 24) ^ rxcrc;
      int ports = 0;
      if (nant == 1 && diff == 0x0000u) ports = 1;
      else if (nant == 2 && diff == 0xFFFFu) ports = 2;
      else if (nant == 4 && diff == 0x5555u) ports = 4;
      if (!ports) continue;
      static const int prb[8] = {6, 15, 25, 50, 75, 100, 0, 0};
      const int bw = (bits[0] << 2) | (bits[1] << 1) | bits[2];
      if (!prb[bw]) continue;
      int sfn = 0;
      for (int i = 0; i < 8; ++i) sfn = (sfn << 1) | bits[6 + i];
      out->nof_prb = prb[bw];
      out->nof_ports = ports;
      out->phich_length = bits[3];
      out->phich_resources = (bits[4] << 1) | bits[5];
      out->sfn = (sfn << 2) | off;
      out->sfn_offset = off;
      return 1;                                            // SRSLTE_UE_MIB_FOUND
    }
  }
  return 0;                                                // SRSLTE_UE_MIB_NOTFOUND
}
