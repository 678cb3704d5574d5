// ltetrigger_b200_blocks.hpp -- C++ host side above the C ABI (include/ltetrigger_b200.h).
//
// Mirrors the reference's compiled block interface for the PSS+SSS path without GNU Radio
// types (GNU Radio is absent where this is built; the thin gr::block subclass a maintainer
// would wrap around these classes is shown in INTEGRATION.md):
//
//   gr::ltetrigger::pss   include/ltetrigger/pss.h:36-88   lib/pss_impl.{h,cc}
//   gr::ltetrigger::sss   include/ltetrigger/sss.h:36-52   lib/sss_impl.{h,cc}
//
// Same factory names and arguments (`make`), accessors, history / output_multiple, work
// signatures (raw item pointers and counts), consume/return semantics, stream-tag keys and
// values, and the same error behaviour: construction failures throw std::runtime_error with the
// reference's messages (lib/pss_impl.cc:72-79, lib/sss_impl.cc:63-70).  All arithmetic runs in
// libltetrigger_b200.so on the GPU; there is no CPU path.
#ifndef LTETRIGGER_B200_BLOCKS_HPP
#define LTETRIGGER_B200_BLOCKS_HPP

#include <complex>
#include <cstdint>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ltetrigger_b200.h"

namespace ltetrigger_b200 {

typedef std::complex<float> gr_complex;

static const int slot_length = LTB_SLOT_LEN;             // lib/pss_impl.h:52-55
static const int half_frame_length = LTB_HALF_FRAME;
static const int symbol_sz = LTB_SYMBOL_SZ;

// stream tags: pmt values reduced to what this path uses (PMT_NIL, PMT_T / PMT_F, pmt::from_long)
struct tag_t {
  enum kind_t { NIL, BOOL, LONG };
  uint64_t offset;
  std::string key;
  kind_t kind;
  long value;
};
static const char *const tracking_lost_tag_key = "tracking_lost";   // lib/pss_impl.cc:39-40
static const char *const cell_id_tag_key = "cell_id";               // lib/sss_impl.cc:38
static const char *const cp_type_tag_key = "cp_type";               // lib/sss_impl.cc:40

// the slice of gr::block the two blocks use
class block {
 public:
  explicit block(const std::string &name) : d_name(name) {}
  virtual ~block() {}
  const std::string &name() const { return d_name; }
  unsigned history() const { return d_history; }
  int output_multiple() const { return d_output_multiple; }
  uint64_t nitems_read(unsigned) const { return d_nitems_read; }
  uint64_t nitems_written(unsigned) const { return d_nitems_written; }
  // scheduler side: tags arriving on the input, tags produced, consumed count of the last call
  std::vector<tag_t> &input_tags() { return d_in_tags; }
  std::vector<tag_t> &output_tags() { return d_out_tags; }
  int consumed() const { return d_consumed; }
  void advance(int nconsumed, int nproduced) { d_nitems_read += nconsumed; d_nitems_written += nproduced; }

 protected:
  void set_history(unsigned h) { d_history = h; }
  void set_output_multiple(int m) { d_output_multiple = m; }
  void consume_each(int n) { d_consumed = n; }
  void add_item_tag(unsigned, uint64_t offset, const std::string &key, tag_t::kind_t kind, long value) {
    d_out_tags.push_back(tag_t{offset, key, kind, value});
  }
  void get_tags_in_window(std::vector<tag_t> &v, unsigned, uint64_t rel_start, uint64_t rel_end, const std::string &key) {
    for (const tag_t &t : d_in_tags)
      if (t.offset >= d_nitems_read + rel_start && t.offset < d_nitems_read + rel_end && t.key == key) v.push_back(t);
  }

 private:
  std::string d_name;
  unsigned d_history = 1;
  int d_output_multiple = 1;
  uint64_t d_nitems_read = 0, d_nitems_written = 0;
  int d_consumed = 0;
  std::vector<tag_t> d_in_tags, d_out_tags;
};

// ---- one engine for the three chains of a hier block ----------------------------------------------
// downlink_trigger_c connects ONE input stream to three pss blocks (python/downlink_trigger_c.py:27-45).
// Built on their own, the three adapters would create three engines and copy the same samples to the GPU
// three times.  An engine_group owns one ltb_trigger with all three roots; the pss (and sss) adapters made
// with it hand their scheduler calls to the group, which feeds each new sample once (whichever block the
// scheduler happens to run first brings it) and queues the window records per root.
class engine_group {
 public:
  typedef std::shared_ptr<engine_group> sptr;
  static sptr make(float psr_threshold, int track_after = 16, int track_every = 8, int device = 0) {
    return sptr(new engine_group(psr_threshold, track_after, track_every, device));
  }
  ~engine_group() { if (d_ltb) ltb_trigger_destroy(d_ltb); }

  struct item { ltb_window_rec rec; std::vector<gr_complex> halfframe; };

  // Called from pss(N_id_2 = root)::general_work.  `in` holds the stream from absolute item `abs0` to
  // `avail_end`; the engine is fed up to want_end.  Returns false if no call of this root is ready.
  bool next(int root, const gr_complex *in, uint64_t abs0, uint64_t avail_end, uint64_t want_end, item &out) {
    std::lock_guard<std::mutex> lk(d_mu);
    if (want_end > avail_end) want_end = avail_end;
    while (d_ready[root].empty() && want_end > d_pushed + 7) {
      if (d_pushed < abs0) throw std::runtime_error("engine_group: the blocks do not read the same stream");
      int64_t n = (int64_t)((want_end - d_pushed) / 8 * 8);
      if (n > d_max_chunk) n = d_max_chunk;
      push(in + (d_pushed - abs0), n);
    }
    if (d_ready[root].empty()) return false;
    out = std::move(d_ready[root].front());
    d_ready[root].pop_front();
    if (out.rec.flags & LTB_F_EMIT) d_sss[root].push_back(out.rec);      // what sss::work will say about it
    return true;
  }
  // the SSS result of the next half-frame pss(root) emitted (computed on the GPU in the same call)
  bool next_sss(int root, ltb_window_rec &rec) {
    std::lock_guard<std::mutex> lk(d_mu);
    if (d_sss[root].empty()) return false;
    rec = d_sss[root].front();
    d_sss[root].pop_front();
    return true;
  }
  void set_psr_threshold(int root, float thr) {
    std::lock_guard<std::mutex> lk(d_mu);
    ltb_trigger_set_psr_threshold(d_ltb, 0, root, thr, 0);
  }
  ltb_pss_stats stats(int root) {
    std::lock_guard<std::mutex> lk(d_mu);
    ltb_pss_stats s = ltb_pss_stats();
    ltb_trigger_get_stats(d_ltb, 0, root, &s);
    return s;
  }

 private:
  engine_group(float psr_threshold, int track_after, int track_every, int device) {
    ltb_trigger_config cfg = ltb_trigger_config();
    cfg.struct_size = sizeof cfg;
    cfg.device = device;
    cfg.n_streams = 1;
    cfg.input_format = LTB_FMT_FC32;
    cfg.decim = 1;
    cfg.root_mask = 7;
    cfg.max_chunk = d_max_chunk;
    cfg.psr_threshold = psr_threshold;
    cfg.track_after = track_after;
    cfg.track_every = track_every;
    cfg.record_all = 1;
    cfg.keep_halfframes = 1;
    if (ltb_trigger_create(&cfg, &d_ltb)) throw std::runtime_error(std::string("Error initializing PSS: ") + ltb_last_error());
    ltb_trigger_set_psr_threshold(d_ltb, 0, -1, psr_threshold, 0);      // the blocks themselves do not clamp
  }
  void push(const gr_complex *x, int64_t n) {
    std::vector<ltb_window_rec> recs((size_t)(3 * (n / 8640 + 8)));
    int n_recs = 0;
    if (ltb_trigger_process_host(d_ltb, x, (int64_t)(n * (int64_t)sizeof(gr_complex)), n, recs.data(), (int)recs.size(), &n_recs))
      throw std::runtime_error(std::string("pss: ") + ltb_last_error());
    std::vector<ltb_cf> hf((size_t)(n_recs > 0 ? n_recs : 1) * half_frame_length);
    int n_hf = 0;
    if (n_recs > 0 && ltb_trigger_fetch_halfframes(d_ltb, hf.data(), n_recs, &n_hf))
      throw std::runtime_error(std::string("pss: ") + ltb_last_error());
    for (int i = 0, k = 0; i < n_recs; ++i) {                            // records come ordered (root, win_index)
      item it;
      it.rec = recs[i];
      if (recs[i].flags & LTB_F_EMIT) {
        const gr_complex *p = reinterpret_cast<const gr_complex *>(&hf[(size_t)k++ * half_frame_length]);
        it.halfframe.assign(p, p + half_frame_length);
      }
      d_ready[recs[i].n_id_2].push_back(std::move(it));
    }
    d_pushed += (uint64_t)n;
  }

  ltb_trigger *d_ltb = nullptr;
  std::mutex d_mu;
  int64_t d_max_chunk = 1 << 18;
  uint64_t d_pushed = 0;
  std::deque<item> d_ready[3];
  std::deque<ltb_window_rec> d_sss[3];
};

// ---- ltetrigger::pss ---------------------------------------------------------------------
class pss : public block {
 public:
  typedef std::shared_ptr<pss> sptr;
  // include/ltetrigger/pss.h:66-69
  static sptr make(int N_id_2, float psr_threshold, int track_after = 16, int track_every = 8, int device = 0) {
    return sptr(new pss(N_id_2, psr_threshold, track_after, track_every, device, engine_group::sptr()));
  }
  // the same block as one of the three chains of a shared engine (see engine_group); its threshold and tracking
  // parameters are the group's until set_psr_threshold changes this root's
  static sptr make(int N_id_2, const engine_group::sptr &group) {
    return sptr(new pss(N_id_2, 0.f, 0, 0, 0, group));
  }
  ~pss() { if (d_ltb) ltb_trigger_destroy(d_ltb); }

  // accessors, lib/pss_impl.h:95-100
  float max_psr() const { return stats().max_psr; }
  float mean_psr() const { return stats().mean_psr; }
  float mean_cfo() const { return stats().mean_cfo; }
  // GNU Radio calls setters and getters from the GUI / Python thread while the scheduler thread is inside
  // general_work; the C ABI wants one host thread per engine at a time, so every use of d_ltb takes d_mu
  void set_psr_threshold(float threshold) {
    if (d_group) { d_group->set_psr_threshold(d_N_id_2, threshold); return; }
    std::lock_guard<std::mutex> lk(d_mu);
    ltb_trigger_set_psr_threshold(d_ltb, 0, d_N_id_2, threshold, 0);
  }
  // How far the engine may run ahead of the scheduler, in general_work calls.  1 (default): the engine
  // receives exactly what the next call needs (nitems_read + 18365 items), so set_psr_threshold applies to
  // the next call and the accessors describe the call that just returned, as in the reference.  Larger
  // values hand the engine up to that many windows per push (file playback at full rate): the accessors
  // then describe up to n - 1 calls in the future and a new threshold applies that much later.
  void set_lookahead_windows(int n) { d_lookahead_windows = n < 1 ? 1 : n; }
  float psr_threshold() const { return stats().psr_threshold; }
  float tracking_score() const { return stats().tracking_score; }

  // items the scheduler must provide before calling general_work: the reference's history plus the
  // largest consume of one call (its assert at lib/pss_impl.cc:191 states the same requirement)
  void forecast(int, std::vector<int> &ninput_items_required) {
    ninput_items_required.assign(1, (int)history() - 1 + LTB_LOOKAHEAD);
  }

  // lib/pss_impl.cc:154-223.  input_items[0] holds history()-1 old items followed by the new ones;
  // returns 0 or 9600 produced items, consume count via consumed().
  int general_work(int, std::vector<int> &ninput_items, std::vector<const void *> &input_items,
                   std::vector<void *> &output_items) {
    const gr_complex *in = static_cast<const gr_complex *>(input_items[0]) + (history() - 1);
    gr_complex *out = static_cast<gr_complex *>(output_items[0]);
    const uint64_t avail_end = nitems_read(0) + (uint64_t)(ninput_items[0] - (int)(history() - 1));
    // hand the engine what the next d_lookahead_windows calls need (multiples of 8), no more: the first
    // call needs nitems_read + LTB_LOOKAHEAD, every further one at most another 18365 - 960 items
    uint64_t want_end = nitems_read(0) + (uint64_t)LTB_LOOKAHEAD + 7 +
                        (uint64_t)(d_lookahead_windows - 1) * (uint64_t)(LTB_LOOKAHEAD - LTB_SLOT_LEN);
    if (want_end > avail_end) want_end = avail_end;
    ltb_window_rec rec;
    std::vector<gr_complex> hf;
    if (d_group) {
      engine_group::item it;
      if (!d_group->next(d_N_id_2, in, nitems_read(0), avail_end, want_end, it)) { consume_each(0); return 0; }
      rec = it.rec;
      hf = std::move(it.halfframe);
    } else {
      while (d_ready.empty() && want_end > d_pushed + 7) {
        int64_t n = (int64_t)((want_end - d_pushed) / 8 * 8);
        if (n > d_max_chunk) n = d_max_chunk;
        push(in + (d_pushed - nitems_read(0)), n);
      }
      if (d_ready.empty()) { consume_each(0); return 0; }
      rec = d_ready.front().first;
      hf = std::move(d_ready.front().second);
      d_ready.pop_front();
    }
    if ((uint64_t)rec.win_start != nitems_read(0)) throw std::runtime_error("pss: scheduler and engine out of step");
    d_last = rec;
    if (rec.flags & LTB_F_EMIT) {
      std::copy(hf.begin(), hf.end(), out);                                         // :193 (+ :204 when tracking)
      if (rec.flags & LTB_F_TAG_LOST)
        add_item_tag(0, nitems_written(0), tracking_lost_tag_key, tag_t::NIL, 0);   // :212
      consume_each((int)(rec.emit_start - rec.win_start) + half_frame_length);      // :195
      return half_frame_length;
    }
    consume_each(half_frame_length);                                                // :217
    return 0;
  }
  const ltb_window_rec &last_record() const { return d_last; }

 private:
  pss(int N_id_2, float psr_threshold, int track_after, int track_every, int device, const engine_group::sptr &group)
      : block("pss"), d_N_id_2(N_id_2), d_group(group) {
    if (N_id_2 < 0 || N_id_2 > 2) throw std::runtime_error("Error initializing PSS N_id_2");
    set_history(half_frame_length);           // lib/pss_impl.cc:81
    set_output_multiple(half_frame_length);   // :82
    if (d_group) return;
    ltb_trigger_config cfg = ltb_trigger_config();
    cfg.struct_size = sizeof cfg;
    cfg.device = device;
    cfg.n_streams = 1;
    cfg.input_format = LTB_FMT_FC32;
    cfg.decim = 1;
    cfg.root_mask = 1 << N_id_2;
    cfg.max_chunk = d_max_chunk;
    cfg.psr_threshold = psr_threshold;
    cfg.track_after = track_after;
    cfg.track_every = track_every;
    cfg.record_all = 1;
    cfg.keep_halfframes = 1;
    if (ltb_trigger_create(&cfg, &d_ltb)) throw std::runtime_error(std::string("Error initializing PSS: ") + ltb_last_error());
    ltb_trigger_set_psr_threshold(d_ltb, 0, N_id_2, psr_threshold, 0);   // the block itself does not clamp
  }
  ltb_pss_stats stats() const {
    if (d_group) return d_group->stats(d_N_id_2);
    std::lock_guard<std::mutex> lk(d_mu);
    ltb_pss_stats s = ltb_pss_stats();
    ltb_trigger_get_stats(d_ltb, 0, d_N_id_2, &s);
    return s;
  }
  void push(const gr_complex *x, int64_t n) {
    std::lock_guard<std::mutex> lk(d_mu);
    std::vector<ltb_window_rec> recs((size_t)(n / 8640 + 8));
    int n_recs = 0;
    if (ltb_trigger_process_host(d_ltb, x, (int64_t)(n * (int64_t)sizeof(gr_complex)), n, recs.data(), (int)recs.size(), &n_recs))
      throw std::runtime_error(std::string("pss: ") + ltb_last_error());
    std::vector<ltb_cf> hf((size_t)(n_recs > 0 ? n_recs : 1) * half_frame_length);
    int n_hf = 0;
    if (n_recs > 0 && ltb_trigger_fetch_halfframes(d_ltb, hf.data(), n_recs, &n_hf))
      throw std::runtime_error(std::string("pss: ") + ltb_last_error());
    for (int i = 0, k = 0; i < n_recs; ++i) {
      std::vector<gr_complex> v;
      if (recs[i].flags & LTB_F_EMIT) {
        const gr_complex *p = reinterpret_cast<const gr_complex *>(&hf[(size_t)k++ * half_frame_length]);
        v.assign(p, p + half_frame_length);
      }
      d_ready.push_back(std::make_pair(recs[i], std::move(v)));
    }
    d_pushed += (uint64_t)n;
  }

  int d_N_id_2;
  engine_group::sptr d_group;                              // null: this block owns its engine
  ltb_trigger *d_ltb = nullptr;
  mutable std::mutex d_mu;                                 // serialises every ltb_trigger_* call on d_ltb
  int d_lookahead_windows = 1;
  int64_t d_max_chunk = 1 << 18;
  uint64_t d_pushed = 0;                                   // absolute count of items handed to the engine
  std::deque<std::pair<ltb_window_rec, std::vector<gr_complex> > > d_ready;   // calls already evaluated
  ltb_window_rec d_last = ltb_window_rec();
};

// ---- ltetrigger::sss ---------------------------------------------------------------------
class sss : public block {
 public:
  typedef std::shared_ptr<sss> sptr;
  static sptr make(int N_id_2, int device = 0) { return sptr(new sss(N_id_2, device, engine_group::sptr())); }   // include/ltetrigger/sss.h:51
  // downstream of pss::make(N_id_2, group): the engine has already decoded the SSS of every half-frame that pss
  // emitted, so work() only attaches the tags.  The input MUST be that pss block's output, in order.
  static sptr make(int N_id_2, const engine_group::sptr &group) { return sptr(new sss(N_id_2, 0, group)); }
  ~sss() { if (d_sss) ltb_sss_destroy(d_sss); }

  // lib/sss_impl.cc:83-156: one aligned half-frame per call, returns 9600
  int work(int, std::vector<const void *> &input_items, std::vector<void *> &output_items) {
    const gr_complex *in = static_cast<const gr_complex *>(input_items[0]);
    gr_complex *out = static_cast<gr_complex *>(output_items[0]);
    std::vector<tag_t> tags;
    get_tags_in_window(tags, 0, 0, 1, tracking_lost_tag_key);                       // :91
    int32_t lost = !tags.empty();
    ltb_window_rec rec = ltb_window_rec();
    rec.m0 = rec.m1 = rec.n_id_1 = rec.cell_id = -1;
    if (d_group) {
      if (!d_group->next_sss(d_N_id_2, rec)) throw std::runtime_error("sss: no half-frame pending from the shared engine");
      if (((rec.flags & LTB_F_TAG_LOST) != 0) != (lost != 0)) throw std::runtime_error("sss: input is not the paired pss block's output");
    } else if (ltb_sss_work(d_sss, reinterpret_cast<const ltb_cf *>(in), &lost, 1, &rec)) {
      throw std::runtime_error(std::string("sss: ") + ltb_last_error());
    }
    d_last = rec;
    if (!lost && !(rec.flags & LTB_F_CELL)) return half_frame_length;               // :119-120, out not written
    if (!lost) {
      add_item_tag(0, nitems_written(0), cell_id_tag_key, tag_t::LONG, rec.cell_id);                        // :141-145
      add_item_tag(0, nitems_written(0), cp_type_tag_key, tag_t::BOOL, (rec.flags & LTB_F_CP_NORM) ? 1 : 0);  // :146-150
    }
    std::copy(in, in + half_frame_length, out);                                     // :98 / :152
    return half_frame_length;
  }
  const ltb_window_rec &last_record() const { return d_last; }

 private:
  sss(int N_id_2, int device, const engine_group::sptr &group) : block("sss"), d_N_id_2(N_id_2), d_group(group) {
    if (N_id_2 < 0 || N_id_2 > 2) throw std::runtime_error("Error initializing SSS N_id_2");
    if (!d_group && ltb_sss_create(device, N_id_2, &d_sss)) throw std::runtime_error("Error initializing SSS SYNC");
    set_output_multiple(half_frame_length);   // lib/sss_impl.cc:72
  }
  int d_N_id_2;
  engine_group::sptr d_group;
  ltb_sss *d_sss = nullptr;
  ltb_window_rec d_last = ltb_window_rec();
};

}  // namespace ltetrigger_b200
#endif  // LTETRIGGER_B200_BLOCKS_HPP
