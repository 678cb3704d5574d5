// gr::block as the two blocks see it, plus public test_* hooks for the scheduler loop of the test
#pragma once
#include <gnuradio/io_signature.h>
#include <gnuradio/tags.h>

#include <complex>
#include <stdexcept>
#include <string>
#include <vector>

typedef std::complex<float> gr_complex;
typedef std::vector<int> gr_vector_int;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;

namespace gr {
class block {
 public:
  enum tag_propagation_policy_t { TPP_DONT = 0, TPP_ALL_TO_ALL = 1, TPP_ONE_TO_ONE = 2 };
  virtual ~block() {}
  const std::string &name() const { return d_name; }
  unsigned history() const { return d_history; }
  void set_history(unsigned h) { d_history = h; }
  int output_multiple() const { return d_output_multiple; }
  void set_output_multiple(int m) { d_output_multiple = m; }
  tag_propagation_policy_t tag_propagation_policy() const { return d_tpp; }
  void set_tag_propagation_policy(tag_propagation_policy_t p) { d_tpp = p; }
  uint64_t nitems_read(unsigned) { return d_read; }
  uint64_t nitems_written(unsigned) { return d_written; }
  void consume_each(int n) { d_consumed = n; }
  virtual void forecast(int noutput_items, gr_vector_int &ninput_items_required) {
    for (size_t i = 0; i < ninput_items_required.size(); ++i) ninput_items_required[i] = noutput_items + (int)history() - 1;
  }
  virtual int general_work(int, gr_vector_int &, gr_vector_const_void_star &, gr_vector_void_star &) {
    throw std::runtime_error("general_work not implemented");
  }
  void add_item_tag(unsigned, uint64_t offset, const pmt::pmt_t &key, const pmt::pmt_t &value,
                    const pmt::pmt_t &srcid = pmt::PMT_F) {
    tag_t t; t.offset = offset; t.key = key; t.value = value; t.srcid = srcid;
    test_out_tags.push_back(t);
  }
  void get_tags_in_window(std::vector<tag_t> &v, unsigned, uint64_t rel_start, uint64_t rel_end, const pmt::pmt_t &key) {
    for (size_t i = 0; i < test_in_tags.size(); ++i) {
      const tag_t &t = test_in_tags[i];
      if (t.offset >= d_read + rel_start && t.offset < d_read + rel_end && pmt::eq(t.key, key)) v.push_back(t);
    }
  }
  // ---- scheduler side (test only) ----
  std::vector<tag_t> test_in_tags, test_out_tags;
  int test_consumed() const { return d_consumed; }
  void test_advance(int nconsumed, int nproduced) { d_read += nconsumed; d_written += nproduced; }
 protected:
  block() {}
  block(const std::string &name, io_signature::sptr in, io_signature::sptr out) : d_name(name), d_in(in), d_out(out) {}
 private:
  std::string d_name;
  io_signature::sptr d_in, d_out;
  unsigned d_history = 1;
  int d_output_multiple = 1;
  tag_propagation_policy_t d_tpp = TPP_ALL_TO_ALL;
  uint64_t d_read = 0, d_written = 0;
  int d_consumed = 0;
};
}  // namespace gr

namespace gnuradio {
template <class T> boost::shared_ptr<T> get_initial_sptr(T *p) { return boost::shared_ptr<T>(p); }
}
