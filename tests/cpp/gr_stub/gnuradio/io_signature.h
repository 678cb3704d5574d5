#pragma once
#include <boost/shared_ptr.hpp>
namespace gr {
class io_signature {
 public:
  typedef boost::shared_ptr<io_signature> sptr;
  static sptr make(int min_streams, int max_streams, int sizeof_item) {
    sptr s(new io_signature()); s->d_min = min_streams; s->d_max = max_streams; s->d_size = sizeof_item; return s;
  }
  int min_streams() const { return d_min; }
  int max_streams() const { return d_max; }
  int sizeof_stream_item(int) const { return d_size; }
 private:
  int d_min, d_max, d_size;
};
}
