// GNU Radio block ltetrigger.pss backed by libltetrigger_b200.so.
//
// Stands where lib/pss_impl.h of the reference stands: it derives from the reference's own public
// class gr::ltetrigger::pss (include/ltetrigger/pss.h:36-88, unchanged), so Python, GRC and
// downlink_trigger_c see the same block.  The scheduler contract (history 9600, output multiple
// 9600, one general_work call = one search window, consume_each / return value, the
// "tracking_lost" tag on item 0) is that of lib/pss_impl.cc:42-223; the arithmetic runs on the GPU
// through the C++ adapter in include/ltetrigger_b200_blocks.hpp.
#ifndef INCLUDED_LTETRIGGER_PSS_B200_IMPL_H
#define INCLUDED_LTETRIGGER_PSS_B200_IMPL_H

#include <ltetrigger/pss.h>

#include "ltetrigger_b200_blocks.hpp"

namespace gr {
namespace ltetrigger {

// Opt-in sharing of one GPU engine by the three chains of a hier block (LTB_SHARE_ENGINE=1 in the environment
// while the blocks are constructed, e.g. set by python/downlink_trigger_c.py around its three pss / sss
// constructors).  The reference's Python hier block makes pss(0), pss(1), pss(2) with the same parameters on
// the same input (python/downlink_trigger_c.py:27-45) and has no handle to pass between them, so the wrappers
// group by construction order: a pss block joins the open group if its parameters match and its N_id_2 is not
// taken yet, otherwise it opens a new one; an sss block joins the most recent group whose N_id_2 has no sss.
// Without the variable every block owns its engine (three host-to-device copies of the same stream).
struct b200_engine_registry {
  static ltetrigger_b200::engine_group::sptr join_pss(int N_id_2, float psr_threshold, int track_after, int track_every);
  static ltetrigger_b200::engine_group::sptr join_sss(int N_id_2);
};

class pss_b200_impl : public pss {
 public:
  pss_b200_impl(int N_id_2, float psr_threshold, int track_after, int track_every);
  ~pss_b200_impl();

  float max_psr() const { return d_core->max_psr(); }
  float mean_psr() const { return d_core->mean_psr(); }
  float mean_cfo() const { return d_core->mean_cfo(); }
  void set_psr_threshold(float threshold) { d_core->set_psr_threshold(threshold); }
  float psr_threshold() const { return d_core->psr_threshold(); }
  float tracking_score() const { return d_core->tracking_score(); }

  void forecast(int noutput_items, gr_vector_int &ninput_items_required);
  int general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items,
                   gr_vector_void_star &output_items);

 private:
  static ltetrigger_b200::pss::sptr make_core(int N_id_2, float psr_threshold, int track_after, int track_every);
  static const pmt::pmt_t tracking_lost_tag_key;
  ltetrigger_b200::pss::sptr d_core;
};

}  // namespace ltetrigger
}  // namespace gr
#endif
