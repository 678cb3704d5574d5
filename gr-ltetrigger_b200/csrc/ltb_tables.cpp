// Constant tables.  All values are computed in double precision on the host and rounded
// to float once; the device kernels only ever see these floats.
//
// Reference call sites these stand in for:
//   srslte_pss_init / srslte_pss_set_N_id_2     lib/pss_impl.cc:72-75
//   srslte_cfo_init (cexptab)                   lib/pss_impl.cc:78
//   srslte_sync_init / srslte_sss_set_N_id_2    lib/sss_impl.cc:63-70
//   gr_filter.rational_resampler_ccc(1, D)      examples/cell_search_file.py:56-57
#include "ltb_tables.h"

#include <cmath>
#include <cstring>

namespace ltb {

namespace {
const double kPi = 3.14159265358979323846;

inline double unit_cos(int j) { return std::cos(2.0 * kPi * (double)(j & 127) / 128.0); }
inline double unit_sin(int j) { return std::sin(2.0 * kPi * (double)(j & 127) / 128.0); }

// 36.211 6.11.1.1 Zadoff-Chu root sequence; the angle is reduced with integer arithmetic
// (u*k mod 126) so d[i] == d[61-i] holds exactly and root 34 is the conjugate of root 29.
void zadoff_chu(int root, double *dre, double *dim) {
  for (int i = 0; i < 62; ++i) {
    const long k = (i < 31) ? (long)i * (i + 1) : (long)(i + 1) * (i + 2);
    const long r = ((long)root * k) % 126;
    const double ang = -kPi * (double)r / 63.0;
    dre[i] = std::cos(ang);
    dim[i] = std::sin(ang);
  }
}

// modified Bessel I0 by its power series (gr::fft::window::kaiser's Izero)
double bessel_i0(double x) {
  double sum = 1, u = 1;
  const double halfx = x / 2.0;
  int n = 1;
  do {
    double t = halfx / (double)n;
    n += 1;
    t *= t;
    u *= t;
    sum += u;
  } while (u >= 1e-21 * sum);
  return sum;
}

void m_sequence(const int *taps, int ntaps, int *out) {
  int x[31];
  std::memset(x, 0, sizeof x);
  x[4] = 1;
  for (int i = 0; i < 26; ++i) {
    int acc = 0;
    for (int t = 0; t < ntaps; ++t) acc += x[i + taps[t]];
    x[i + 5] = acc % 2;
  }
  for (int i = 0; i < 31; ++i) out[i] = 1 - 2 * x[i];
}
}  // namespace

void make_pss_taps(int n_id_2, PssTaps &out) {
  double dre[62], dim[62];
  zadoff_chu(n_id_2 == 0 ? 25 : 29, dre, dim);
  const double scale = 1.0 / std::sqrt(128.0) / 62.0;
  for (int n = 0; n <= 64; ++n) {
    double tr = 0.0, ti = 0.0;
    for (int i = 0; i < 62; ++i) {
      const int bin = (i < 31) ? i - 31 : i - 30;   // -31..-1, +1..+31 (DC empty)
      const int j = ((bin * n) % 128 + 128) % 128;
      const double c = unit_cos(j), s = unit_sin(j);
      tr += dre[i] * c - dim[i] * s;
      ti += dre[i] * s + dim[i] * c;
    }
    out.re[n] = (float)(tr * scale);
    out.im[n] = (float)(-ti * scale);
    if (n_id_2 == 2) out.im[n] = -out.im[n];
  }
  for (int n = 65; n < 128; ++n) {
    out.re[n] = out.re[128 - n];
    out.im[n] = out.im[128 - n];
  }
}

std::vector<float> make_decim_taps(int decim) {
  std::vector<float> taps;
  if (decim <= 1) return taps;
  const double beta = 7.0, fractional_bw = 0.4, halfband = 0.5;
  const double rate = 1.0 / (double)decim;
  const double trans_width = rate * (halfband - fractional_bw);
  const double mid = rate * halfband - trans_width / 2.0;
  const double atten = beta / 0.1102 + 8.7;
  int ntaps = (int)(atten * 1.0 / (22.0 * trans_width));
  if ((ntaps & 1) == 0) ntaps++;
  std::vector<float> w(ntaps);
  const double ibeta = 1.0 / bessel_i0(beta), inm1 = 1.0 / (double)(ntaps - 1);
  for (int i = 0; i < ntaps; ++i) {
    const double t = 2 * i * inm1 - 1;
    w[i] = (float)(bessel_i0(beta * std::sqrt(1.0 - t * t)) * ibeta);
  }
  taps.resize(ntaps);
  const int M = (ntaps - 1) / 2;
  const double fwT0 = 2 * kPi * mid / 1.0;
  for (int n = -M; n <= M; ++n) {
    if (n == 0) taps[n + M] = (float)(fwT0 / kPi * w[n + M]);
    else        taps[n + M] = (float)(std::sin(n * fwT0) / (n * kPi) * w[n + M]);
  }
  double fmax = taps[M];
  for (int n = 1; n <= M; ++n) fmax += 2 * taps[n + M];
  const double gain = 1.0 / fmax;
  for (int i = 0; i < ntaps; ++i) taps[i] = (float)(taps[i] * gain);
  return taps;
}

std::vector<float> make_decim_branch_taps(int decim) {
  const int Q = 33;
  const std::vector<float> taps = make_decim_taps(decim);
  if (decim < 2 || (int)taps.size() > Q * decim) return std::vector<float>();
  std::vector<float> out((size_t)decim * Q, 0.0f);
  for (int v = 0; v < decim; ++v)
    for (int q = 0; q < Q; ++q) {
      const int j = q * decim + v;
      if (j < (int)taps.size()) out[(size_t)v * Q + q] = taps[j];
    }
  return out;
}

static int tc_ilog2(int d) { int l = 0; while (d > 1) { d >>= 1; ++l; } return l; }
int tc_tap_shift_for(int decim) { return 23 + tc_ilog2(decim); }

std::vector<int32_t> make_tc_taps(int decim, long long *sum_t) {
  const std::vector<float> taps = make_decim_taps(decim);
  std::vector<int32_t> T(taps.size());
  long long sum = 0;
  const double scale = std::ldexp(1.0, tc_tap_shift_for(decim));
  for (size_t j = 0; j < taps.size(); ++j) {
    T[j] = (int32_t)std::llrint((double)taps[j] * scale);
    sum += T[j];
  }
  if (sum_t) *sum_t = sum;
  return T;
}

std::vector<int8_t> make_tc_btab(int fmt, int decim) {
  constexpr int kRows = 208, kTile = kRows * 128;
  const std::vector<int32_t> T = make_tc_taps(decim, nullptr);
  std::vector<int8_t> tab((size_t)kTile, 0);
  auto sw_off = [](int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); };
  auto digit = [](int t, int v, bool *ok) {            // balanced base-256 digits, v = 0..2
    const int d0 = ((t + 128) & 255) - 128, t1 = (t - d0) >> 8;
    const int d1 = ((t1 + 128) & 255) - 128, t2 = (t1 - d1) >> 8;
    if (t2 < -128 || t2 > 127) *ok = false;
    return v == 0 ? d0 : v == 1 ? d1 : v == 2 ? t2 : 0;
  };
  auto tap = [&](int j) { return (j >= 0 && j < (int)T.size()) ? T[j] : 0; };
  auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
  bool ok = true;
  // a k-step is 32 bytes of one component: spk samples of bps bytes each (fc32: three digit bytes and one byte that
  // meets zeros).  Row 4 d + v of the table of phase ph holds, at byte bps' p + byte, the tap digit that byte of
  // sample p multiplies into weight 256^v (fc32: 256^(v+1), its weight-1 product is not formed): T[decim d - ph - p].
  const int bpc = fmt == 0 ? 4 : fmt == 1 ? 2 : 1, spk = 32 / bpc, g = gcd(spk, decim), nph = decim / g;
  if (nph > 4) return std::vector<int8_t>();
  for (int n = 0; n < kRows; ++n) {
    const int d = n / 4, v = n % 4;
    for (int i = 0; i < nph; ++i)
      for (int p = 0; p < spk; ++p) {
        const int t = tap(decim * d - i * g - p);
        if (!t) continue;
        for (int byte = 0; byte < (fmt == 0 ? 3 : bpc); ++byte) {
          const int dg = fmt == 0 ? v + 1 - byte : v - byte;                    // tap digit index
          if (dg < 0 || dg > 2) continue;
          const int kb = 32 * i + bpc * p + byte;
          tab[sw_off(n, kb >> 4) + (kb & 15)] = (int8_t)digit(t, dg, &ok);
        }
      }
  }
  if (!ok) tab.clear();                                 // a tap that needs a fourth digit: never for these taps
  return tab;
}

void make_sss_tables(int n_id_2, SssTables &out) {
  int c_tilde[31];
  const int s_taps[2] = {2, 0}, c_taps[2] = {3, 0}, z_taps[4] = {4, 2, 1, 0};
  m_sequence(s_taps, 2, out.s_tilde);
  m_sequence(c_taps, 2, c_tilde);
  m_sequence(z_taps, 4, out.z_tilde);
  for (int i = 0; i < 31; ++i) {
    out.c0[i] = c_tilde[(i + n_id_2) % 31];
    out.c1[i] = c_tilde[(i + n_id_2 + 3) % 31];
  }
  std::memset(out.n_id_1, 0, sizeof out.n_id_1);
  for (int nid = 0; nid < 168; ++nid) {
    const int qp = nid / 30;
    const int q = (nid + qp * (qp + 1) / 2) / 30;
    const int mp = nid + q * (q + 1) / 2;
    const int m0 = mp % 31;
    const int m1 = (m0 + mp / 31 + 1) % 31;
    out.n_id_1[m0 * 30 + (m1 - 1)] = nid;
  }
}

void make_fft1024_twiddles(float *re, float *im) {
  for (int i = 0; i < 1024; ++i) {
    const double a = 2.0 * kPi * (double)i / 1024.0;
    re[i] = (float)std::cos(a);
    im[i] = (float)(-std::sin(a));
  }
  re[0] = 1.f;    im[0] = 0.f;   re[256] = 0.f; im[256] = -1.f;
  re[512] = -1.f; im[512] = 0.f; re[768] = 0.f; im[768] = 1.f;
}

void make_os_filter(int n_id_2, float *H_re, float *H_im) {
  PssTaps h;
  make_pss_taps(n_id_2, h);
  for (int f = 0; f < 1024; ++f) {
    double ar = 0.0, ai = 0.0;
    for (int m = 0; m < 128; ++m) {
      const double a = 2.0 * kPi * (double)((f * m) & 1023) / 1024.0;
      const double c = std::cos(a), sn = -std::sin(a);
      ar += (double)h.re[m] * c - (double)h.im[m] * sn;
      ai += (double)h.re[m] * sn + (double)h.im[m] * c;
    }
    H_re[f] = (float)(ar / 1024.0);
    H_im[f] = (float)(ai / 1024.0);
  }
}

void make_cexp_table(float *re, float *im) {
  for (int i = 0; i < 4096; ++i) {
    const double a = 2.0 * kPi * (double)i / 4096.0;
    re[i] = (float)std::cos(a);
    im[i] = (float)std::sin(a);
  }
  re[4096] = re[0];
  im[4096] = im[0];
}

void make_fft128_twiddles(float *re, float *im) {
  for (int k = 0; k < 64; ++k) {
    re[k] = (float)unit_cos(k);
    im[k] = (float)(-unit_sin(k));
  }
}

}  // namespace ltb
