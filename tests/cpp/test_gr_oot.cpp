// Drives the GNU Radio wrappers (gr-ltetrigger_b200/gr_oot/lib) through the stand-in gr::block API
// of tests/cpp/gr_stub the way the scheduler drives the reference's blocks:
// file_source(repeat) -> head -> pss(k) -> sss(k).  Same output lines as test_blocks.cpp, so
// tests/test_blocks_cpp.py checks both against the oracle with one routine.
//   usage: test_gr_oot <fc32 file at 1.92 Msps> <seconds> <N_id_2> <psr_threshold> [hier]
// The scheduler stand-in offers the block at most what GNU Radio's buffer would hold: the upstream buffer is
// sized max(8192, 2 * (history + output_multiple)) = 38400 items for this block (flat_flowgraph::allocate_buffer),
// one of which stays empty, so never more than 38399 items are visible to a general_work call.
// With "hier" (and LTB_SHARE_ENGINE=1 in the environment) the three chains of downlink_trigger_c are built --
// pss(0), pss(1), pss(2), then the three sss blocks, as python/downlink_trigger_c.py:27-45 does -- and driven
// round robin on one engine; the printed lines are those of chain N_id_2.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include <ltetrigger/pss.h>
#include <ltetrigger/sss.h>

using gr::ltetrigger::pss;
using gr::ltetrigger::sss;

int main(int argc, char **argv) {
  if (argc < 5) { std::fprintf(stderr, "usage: %s file seconds N_id_2 threshold\n", argv[0]); return 2; }
  std::ifstream f(argv[1], std::ios::binary);
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const size_t frame = raw.size() / sizeof(gr_complex);
  if (!frame) { std::fprintf(stderr, "empty input\n"); return 2; }
  const gr_complex *src = reinterpret_cast<const gr_complex *>(raw.data());
  const size_t n = (size_t)(std::atof(argv[2]) * 1.92e6) / 8 * 8;
  const int k = std::atoi(argv[3]);
  const float thr = (float)std::atof(argv[4]);
  const int half_frame_length = 9600;

  try { pss::make(5, thr); std::printf("E no throw\n"); return 1; } catch (const std::runtime_error &e) { std::printf("E %s\n", e.what()); }

  const bool hier = argc > 5 && std::string(argv[5]) == "hier";
  pss::sptr pall[3];
  sss::sptr sall[3];
  if (hier) {
    for (int j = 0; j < 3; ++j) pall[j] = pss::make(j, thr);
    for (int j = 0; j < 3; ++j) sall[j] = sss::make(j);
  } else {
    pall[k] = pss::make(k, thr);
    sall[k] = sss::make(k);
  }
  pss::sptr p = pall[k];
  sss::sptr s = sall[k];
  if (p->history() != 9600 || p->output_multiple() != 9600 || s->output_multiple() != 9600 ||
      s->tag_propagation_policy() != gr::block::TPP_ALL_TO_ALL) { std::printf("E block contract\n"); return 1; }
  const size_t hist = p->history() - 1;
  std::vector<gr_complex> buf(hist + n);                    // GR zero-fills the history
  for (size_t i = 0; i < n; ++i) buf[hist + i] = src[i % frame];
  gr_vector_int need(1, 0);
  p->forecast(half_frame_length, need);
  std::vector<gr_complex> out(half_frame_length), out2(half_frame_length);
  const long gr_buffer_items = 2 * (9600 + 9600) > 8192 ? 2 * (9600 + 9600) : 8192;
  for (;;) {
    if (hier) {
      // the other two chains advance as far as chain k has (round robin), through the same kind of calls
      for (int j = 0; j < 3; ++j) {
        if (j == k) continue;
        while (pall[j]->nitems_read(0) <= p->nitems_read(0)) {
          const uint64_t rj = pall[j]->nitems_read(0);
          long av = (long)buf.size() - (long)rj;
          if (av > gr_buffer_items - 1) av = gr_buffer_items - 1;
          if (av < need[0]) break;
          gr_vector_int ninj(1, (int)av);
          gr_vector_const_void_star inj(1, &buf[rj]);
          gr_vector_void_star oj(1, out2.data());
          pall[j]->test_out_tags.clear();
          const int no = pall[j]->general_work(half_frame_length, ninj, inj, oj);
          const int nc = pall[j]->test_consumed();
          if (no) {
            sall[j]->test_in_tags = pall[j]->test_out_tags;
            sall[j]->test_out_tags.clear();
            gr_vector_const_void_star in3(1, out2.data());
            std::vector<gr_complex> out3(half_frame_length);
            gr_vector_void_star o3(1, out3.data());
            const int n3 = sall[j]->work(half_frame_length, in3, o3);
            sall[j]->test_advance(n3, n3);
          }
          pall[j]->test_advance(nc, no);
          if (!no && !nc) break;
        }
      }
    }
    const uint64_t r = p->nitems_read(0);
    long avail = (long)buf.size() - (long)r;
    if (avail > gr_buffer_items - 1) avail = gr_buffer_items - 1;
    if (avail < need[0]) break;
    gr_vector_int nin(1, (int)avail);
    gr_vector_const_void_star in(1, &buf[r]);
    gr_vector_void_star o(1, out.data());
    p->test_out_tags.clear();
    const int nout = p->general_work(half_frame_length, nin, in, o);
    const int ncons = p->test_consumed();
    bool lost = false;
    for (size_t i = 0; i < p->test_out_tags.size(); ++i)
      lost |= pmt::symbol_to_string(p->test_out_tags[i].key) == "tracking_lost" && pmt::eq(p->test_out_tags[i].value, pmt::PMT_NIL) &&
              p->test_out_tags[i].offset == p->nitems_written(0);
    std::printf("P %llu %d %d %d\n", (unsigned long long)r, nout, ncons, (int)lost);
    if (nout) {
      s->test_in_tags = p->test_out_tags;                   // tags travel with the items
      s->test_out_tags.clear();
      gr_vector_const_void_star in2(1, out.data());
      gr_vector_void_star o2(1, out2.data());
      const int n2 = s->work(half_frame_length, in2, o2);
      long cell = -1, cp = -1;
      for (size_t i = 0; i < s->test_out_tags.size(); ++i) {
        const gr::tag_t &t = s->test_out_tags[i];
        if (pmt::symbol_to_string(t.key) == "cell_id") cell = pmt::to_long(t.value);
        if (pmt::symbol_to_string(t.key) == "cp_type") cp = pmt::eq(t.value, pmt::PMT_T) ? 1 : 0;
      }
      unsigned long long sum = 0;
      const uint32_t *w = reinterpret_cast<const uint32_t *>(out.data());
      for (int i = 0; i < 2 * half_frame_length; ++i) sum = sum * 1000003ull + w[i];
      std::printf("S %llu %ld %ld %llu\n", (unsigned long long)p->nitems_written(0), cell, cp, sum);
      s->test_advance(n2, n2);
    }
    p->test_advance(ncons, nout);
    if (!nout && !ncons) break;
  }
  std::printf("A %.9g %.9g %.9g %.9g %.9g\n", p->max_psr(), p->mean_psr(), p->mean_cfo(), p->psr_threshold(), p->tracking_score());
  return 0;
}
