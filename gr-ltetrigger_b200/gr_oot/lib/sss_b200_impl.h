// GNU Radio block ltetrigger.sss backed by libltetrigger_b200.so; stands where lib/sss_impl.h of the
// reference stands and derives from its public class gr::ltetrigger::sss
// (include/ltetrigger/sss.h:36-52).  Contract of lib/sss_impl.cc:45-156: sync block, output
// multiple 9600, TPP_ALL_TO_ALL, consumes "tracking_lost", emits "cell_id" and "cp_type".
#ifndef INCLUDED_LTETRIGGER_SSS_B200_IMPL_H
#define INCLUDED_LTETRIGGER_SSS_B200_IMPL_H

#include <ltetrigger/sss.h>

#include "ltetrigger_b200_blocks.hpp"

namespace gr {
namespace ltetrigger {

class sss_b200_impl : public sss {
 public:
  explicit sss_b200_impl(int N_id_2);
  ~sss_b200_impl();
  int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);

 private:
  static ltetrigger_b200::sss::sptr make_core(int N_id_2);
  static const pmt::pmt_t cell_id_tag_key, cp_type_tag_key, tracking_lost_tag_key;
  ltetrigger_b200::sss::sptr d_core;
  std::vector<gr::tag_t> d_tags;
};

}  // namespace ltetrigger
}  // namespace gr
#endif
