// See sss_b200_impl.h.  Build this file instead of lib/sss_impl.cc (gr_oot/README.md).
#include "sss_b200_impl.h"

#include <gnuradio/io_signature.h>

#include "pss_b200_impl.h"

namespace gr {
namespace ltetrigger {

const pmt::pmt_t sss_b200_impl::cell_id_tag_key = pmt::intern(ltetrigger_b200::cell_id_tag_key);
const pmt::pmt_t sss_b200_impl::cp_type_tag_key = pmt::intern(ltetrigger_b200::cp_type_tag_key);
const pmt::pmt_t sss_b200_impl::tracking_lost_tag_key = pmt::intern(ltetrigger_b200::tracking_lost_tag_key);

sss::sptr sss::make(int N_id_2) { return gnuradio::get_initial_sptr(new sss_b200_impl(N_id_2)); }

sss_b200_impl::sss_b200_impl(int N_id_2)
    : gr::sync_block("sss", gr::io_signature::make(1, 1, sizeof(gr_complex)), gr::io_signature::make(1, 1, sizeof(gr_complex))),
      d_core(make_core(N_id_2)) {                    // throws "Error initializing SSS SYNC" (lib/sss_impl.cc:63-70)
  set_tag_propagation_policy(TPP_ALL_TO_ALL);        // :61
  set_output_multiple(d_core->output_multiple());    // :72
}

ltetrigger_b200::sss::sptr sss_b200_impl::make_core(int N_id_2) {
  ltetrigger_b200::engine_group::sptr g = b200_engine_registry::join_sss(N_id_2);   // LTB_SHARE_ENGINE=1: see pss_b200_impl.h
  return g ? ltetrigger_b200::sss::make(N_id_2, g) : ltetrigger_b200::sss::make(N_id_2);
}

sss_b200_impl::~sss_b200_impl() {}

int sss_b200_impl::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) {
  // one aligned half-frame per call, like the reference (it returns half_frame_length whatever
  // noutput_items is, lib/sss_impl.cc:98,120,155)
  d_tags.clear();
  get_tags_in_window(d_tags, 0, 0, 1, tracking_lost_tag_key);        // :91
  d_core->input_tags().clear();
  if (!d_tags.empty())
    d_core->input_tags().push_back(ltetrigger_b200::tag_t{d_core->nitems_read(0), ltetrigger_b200::tracking_lost_tag_key,
                                                          ltetrigger_b200::tag_t::NIL, 0});
  d_core->output_tags().clear();
  std::vector<const void *> in(input_items.begin(), input_items.end());
  std::vector<void *> out(output_items.begin(), output_items.end());
  const int produced = d_core->work(noutput_items, in, out);
  for (size_t i = 0; i < d_core->output_tags().size(); ++i) {
    const ltetrigger_b200::tag_t &t = d_core->output_tags()[i];
    if (t.key == ltetrigger_b200::cell_id_tag_key)
      add_item_tag(0, nitems_written(0), cell_id_tag_key, pmt::from_long(t.value));          // :141-142
    else
      add_item_tag(0, nitems_written(0), cp_type_tag_key, t.value ? pmt::PMT_T : pmt::PMT_F);  // :144-150
  }
  d_core->advance(produced, produced);
  return produced;
}

}  // namespace ltetrigger
}  // namespace gr
