// declaration-only stand-in for the reference's include/ltetrigger/sss.h
#pragma once
#include <gnuradio/sync_block.h>
#include <ltetrigger/api.h>
namespace gr { namespace ltetrigger {
class LTETRIGGER_API sss : virtual public gr::sync_block {
 public:
  typedef boost::shared_ptr<sss> sptr;
  static sptr make(int N_id_2);
};
} }
