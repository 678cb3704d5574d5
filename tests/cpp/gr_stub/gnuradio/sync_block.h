#pragma once
#include <gnuradio/block.h>
namespace gr {
class sync_block : public block {
 public:
  virtual int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) = 0;
 protected:
  sync_block() {}
  sync_block(const std::string &name, io_signature::sptr in, io_signature::sptr out) : block(name, in, out) {}
};
}
