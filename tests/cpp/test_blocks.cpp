// Drives the C++ block adapters (include/ltetrigger_b200_blocks.hpp) the way the GNU Radio
// scheduler drives the reference's blocks: file_source(repeat) -> head -> pss(k) -> sss(k).
// Prints one line per call; tests/test_gpu_blocks_cpp.py compares them with the oracle.
//   usage: test_blocks <fc32 file at 1.92 Msps> <seconds> <N_id_2> <psr_threshold>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "ltetrigger_b200_blocks.hpp"

using namespace ltetrigger_b200;

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

  // construction errors behave like the reference's (std::runtime_error)
  try { pss::make(5, thr); std::printf("E no throw\n"); return 1; } catch (const std::runtime_error &e) { std::printf("E %s\n", e.what()); }

  pss::sptr p = pss::make(k, thr);
  sss::sptr s = sss::make(k);
  const size_t hist = p->history() - 1;
  std::vector<gr_complex> buf(hist + n);                    // GR zero-fills the history
  for (size_t i = 0; i < n; ++i) buf[hist + i] = src[i % frame];   // file_source(repeat) -> head
  std::vector<int> need;
  p->forecast(half_frame_length, need);
  std::vector<gr_complex> out(half_frame_length), out2(half_frame_length);
  for (;;) {
    const uint64_t r = p->nitems_read(0);
    const long avail = (long)buf.size() - (long)r;
    if (avail < need[0]) break;
    std::vector<int> nin(1, (int)avail);
    std::vector<const void *> in(1, &buf[r]);
    std::vector<void *> o(1, out.data());
    p->output_tags().clear();
    const int nout = p->general_work(half_frame_length, nin, in, o);
    const int ncons = p->consumed();
    const bool lost = !p->output_tags().empty();
    std::printf("P %llu %d %d %d\n", (unsigned long long)r, nout, ncons, (int)lost);
    if (nout) {
      s->input_tags() = p->output_tags();
      s->output_tags().clear();
      std::vector<const void *> in2(1, out.data());
      std::vector<void *> o2(1, out2.data());
      s->work(half_frame_length, in2, o2);
      long cell = -1, cp = -1;
      for (const tag_t &t : s->output_tags()) {
        if (t.key == cell_id_tag_key) cell = t.value;
        if (t.key == cp_type_tag_key) cp = t.value;
      }
      unsigned long long sum = 0;                           // checksum of the emitted samples' bit patterns
      const uint32_t *w = reinterpret_cast<const uint32_t *>(out.data());
      for (int i = 0; i < 2 * half_frame_length; ++i) sum = sum * 1000003ull + w[i];
      std::printf("S %llu %ld %ld %llu\n", (unsigned long long)p->nitems_written(0), cell, cp, sum);
      s->advance(half_frame_length, half_frame_length);
    }
    p->advance(ncons, nout);
    if (!nout && !ncons) break;
  }
  std::printf("A %.9g %.9g %.9g %.9g %.9g\n", p->max_psr(), p->mean_psr(), p->mean_cfo(), p->psr_threshold(), p->tracking_score());
  return 0;
}
