// See pss_b200_impl.h.  Build this file instead of lib/pss_impl.cc (gr_oot/README.md).
#include "pss_b200_impl.h"

#include <gnuradio/io_signature.h>

#include <cstdlib>
#include <mutex>
#include <vector>

namespace gr {
namespace ltetrigger {

namespace {
struct open_group {
  ltetrigger_b200::engine_group::sptr g;
  float thr; int ta, te;
  bool pss_taken[3], sss_taken[3];
};
std::mutex g_reg_mu;
std::vector<open_group> g_groups;
bool sharing_enabled() {
  const char *e = std::getenv("LTB_SHARE_ENGINE");
  return e && e[0] == '1';
}
}  // namespace

ltetrigger_b200::engine_group::sptr b200_engine_registry::join_pss(int N_id_2, float thr, int ta, int te) {
  if (!sharing_enabled() || N_id_2 < 0 || N_id_2 > 2) return ltetrigger_b200::engine_group::sptr();
  std::lock_guard<std::mutex> lk(g_reg_mu);
  if (!g_groups.empty()) {
    open_group &o = g_groups.back();
    if (o.thr == thr && o.ta == ta && o.te == te && !o.pss_taken[N_id_2]) { o.pss_taken[N_id_2] = true; return o.g; }
  }
  open_group o;
  o.g = ltetrigger_b200::engine_group::make(thr, ta, te);
  o.thr = thr; o.ta = ta; o.te = te;
  for (int k = 0; k < 3; ++k) o.pss_taken[k] = o.sss_taken[k] = false;
  o.pss_taken[N_id_2] = true;
  g_groups.push_back(o);
  if (g_groups.size() > 64) g_groups.erase(g_groups.begin());          // the groups themselves live in their blocks
  return g_groups.back().g;
}

ltetrigger_b200::engine_group::sptr b200_engine_registry::join_sss(int N_id_2) {
  if (!sharing_enabled() || N_id_2 < 0 || N_id_2 > 2) return ltetrigger_b200::engine_group::sptr();
  std::lock_guard<std::mutex> lk(g_reg_mu);
  for (size_t i = g_groups.size(); i-- > 0;) {
    open_group &o = g_groups[i];
    if (o.pss_taken[N_id_2] && !o.sss_taken[N_id_2]) { o.sss_taken[N_id_2] = true; return o.g; }
  }
  return ltetrigger_b200::engine_group::sptr();
}

const pmt::pmt_t pss_b200_impl::tracking_lost_tag_key = pmt::intern(ltetrigger_b200::tracking_lost_tag_key);

pss::sptr pss::make(int N_id_2, float psr_threshold, int track_after, int track_every) {
  return gnuradio::get_initial_sptr(new pss_b200_impl(N_id_2, psr_threshold, track_after, track_every));
}

pss_b200_impl::pss_b200_impl(int N_id_2, float psr_threshold, int track_after, int track_every)
    : gr::block("pss", gr::io_signature::make(1, 1, sizeof(gr_complex)), gr::io_signature::make(1, 1, sizeof(gr_complex))),
      // throws std::runtime_error with the reference's messages (lib/pss_impl.cc:72-79)
      d_core(make_core(N_id_2, psr_threshold, track_after, track_every)) {
  set_history(d_core->history());                  // lib/pss_impl.cc:81
  set_output_multiple(d_core->output_multiple());  // :82
}

ltetrigger_b200::pss::sptr pss_b200_impl::make_core(int N_id_2, float psr_threshold, int track_after, int track_every) {
  ltetrigger_b200::engine_group::sptr g = b200_engine_registry::join_pss(N_id_2, psr_threshold, track_after, track_every);
  return g ? ltetrigger_b200::pss::make(N_id_2, g) : ltetrigger_b200::pss::make(N_id_2, psr_threshold, track_after, track_every);
}

pss_b200_impl::~pss_b200_impl() {}

// The reference leaves forecast at its default and asserts the input is long enough
// (lib/pss_impl.cc:191, compiled out in Release); asking for it up front is the same requirement.
// 9599 + 18365 = 27964 items.  GNU Radio sizes the upstream buffer to at least
// 2 * (history + output_multiple) = 38400 items for this block (flat_flowgraph::allocate_buffer), so the
// request can always be met; tests/cpp/test_gr_oot.cpp drives the block under exactly that limit.
void pss_b200_impl::forecast(int noutput_items, gr_vector_int &ninput_items_required) {
  std::vector<int> need;
  d_core->forecast(noutput_items, need);
  for (size_t i = 0; i < ninput_items_required.size(); ++i) ninput_items_required[i] = need[0];
}

int pss_b200_impl::general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
                                gr_vector_void_star &output_items) {
  // keep the adapter's item counters in step with the scheduler's (they only differ after a
  // flowgraph restart, which GNU Radio models as a new block)
  if (d_core->nitems_read(0) != nitems_read(0)) throw std::runtime_error("pss: item counters out of step");
  std::vector<int> nin(ninput_items.begin(), ninput_items.end());
  std::vector<const void *> in(input_items.begin(), input_items.end());
  std::vector<void *> out(output_items.begin(), output_items.end());
  d_core->output_tags().clear();
  const int produced = d_core->general_work(noutput_items, nin, in, out);
  for (size_t i = 0; i < d_core->output_tags().size(); ++i)          // "tracking_lost", PMT_NIL (:210-213)
    add_item_tag(0, d_core->output_tags()[i].offset, tracking_lost_tag_key, pmt::PMT_NIL);
  consume_each(d_core->consumed());
  d_core->advance(d_core->consumed(), produced);
  return produced;
}

}  // namespace ltetrigger
}  // namespace gr
