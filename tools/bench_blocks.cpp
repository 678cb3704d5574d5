// Throughput of the drop-in block path: the reference's hier block topology
//   source -> 3 x ( pss(N_id_2 = k) -> sss(N_id_2 = k) )          python/downlink_trigger_c.py:27-45
// driven the way the GNU Radio scheduler drives it (tests/cpp/test_blocks.cpp), through the C++ adapters of
// include/ltetrigger_b200_blocks.hpp, with either one engine per pss block ("separate": three H2D copies of
// the same stream) or one engine_group shared by the three chains ("shared"), and a look-ahead of L windows.
//   usage: bench_blocks <fc32 file at 1.92 Msps> <seconds> <threshold> <separate|shared> <lookahead windows>
// Prints one JSON line: wall-clock samples/s of the input stream, time to the first cell_id tag, call counts.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <vector>

#include "ltetrigger_b200_blocks.hpp"

using namespace ltetrigger_b200;

int main(int argc, char **argv) {
  if (argc < 6) { std::fprintf(stderr, "usage: %s file seconds threshold separate|shared lookahead\n", argv[0]); return 2; }
  std::ifstream f(argv[1], std::ios::binary);
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const size_t frame = raw.size() / sizeof(gr_complex);
  if (!frame) { std::fprintf(stderr, "empty input\n"); return 2; }
  const gr_complex *src = reinterpret_cast<const gr_complex *>(raw.data());
  const size_t n = (size_t)(std::atof(argv[2]) * 1.92e6) / 8 * 8;
  const float thr = (float)std::atof(argv[3]);
  const bool shared = std::strcmp(argv[4], "shared") == 0;
  const int look = std::atoi(argv[5]);

  pss::sptr p[3];
  sss::sptr s[3];
  engine_group::sptr g;
  if (shared) g = engine_group::make(thr);
  for (int k = 0; k < 3; ++k) {
    p[k] = shared ? pss::make(k, g) : pss::make(k, thr);
    s[k] = shared ? sss::make(k, g) : sss::make(k);
    p[k]->set_lookahead_windows(look);
  }
  const size_t hist = p[0]->history() - 1;
  std::vector<gr_complex> buf(hist + n);
  for (size_t i = 0; i < n; ++i) buf[hist + i] = src[i % frame];
  std::vector<int> need;
  p[0]->forecast(half_frame_length, need);
  std::vector<gr_complex> out(half_frame_length), out2(half_frame_length);
  long calls = 0, emitted = 0, cells = 0, first_cell = -1;
  double first_cell_ms = -1;
  const auto t0 = std::chrono::steady_clock::now();
  bool progress = true;
  while (progress) {                                        // round-robin over the three chains, like the scheduler
    progress = false;
    for (int k = 0; k < 3; ++k) {
      const uint64_t r = p[k]->nitems_read(0);
      const long avail = (long)buf.size() - (long)r;
      if (avail < need[0]) continue;
      std::vector<int> nin(1, (int)avail);
      std::vector<const void *> in(1, &buf[r]);
      std::vector<void *> o(1, out.data());
      p[k]->output_tags().clear();
      const int nout = p[k]->general_work(half_frame_length, nin, in, o);
      const int ncons = p[k]->consumed();
      calls++;
      if (nout) {
        emitted++;
        s[k]->input_tags() = p[k]->output_tags();
        s[k]->output_tags().clear();
        std::vector<const void *> in2(1, out.data());
        std::vector<void *> o2(1, out2.data());
        s[k]->work(half_frame_length, in2, o2);
        for (const tag_t &t : s[k]->output_tags())
          if (t.key == cell_id_tag_key) {
            cells++;
            if (first_cell < 0) {
              first_cell = t.value;
              first_cell_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            }
          }
        s[k]->advance(half_frame_length, half_frame_length);
      }
      p[k]->advance(ncons, nout);
      if (nout || ncons) progress = true;
    }
  }
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("{\"mode\": \"%s\", \"lookahead_windows\": %d, \"seconds_of_signal\": %.3f, \"wall_s\": %.4f, \"samples_per_s\": %.1f, "
              "\"general_work_calls\": %ld, \"halfframes_emitted\": %ld, \"cell_tags\": %ld, \"first_cell_id\": %ld, \"first_cell_wall_ms\": %.2f}\n",
              shared ? "shared" : "separate", look, n / 1.92e6, wall, n / wall, calls, emitted, cells, first_cell, first_cell_ms);
  return 0;
}
