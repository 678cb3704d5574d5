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
  static const pmt::pmt_t tracking_lost_tag_key;
  ltetrigger_b200::pss::sptr d_core;
};

}  // namespace ltetrigger
}  // namespace gr
#endif
