// declaration-only stand-in for the reference's include/ltetrigger/pss.h (same class, same virtuals)
#pragma once
#include <gnuradio/block.h>
#include <ltetrigger/api.h>
namespace gr { namespace ltetrigger {
class LTETRIGGER_API pss : virtual public gr::block {
 public:
  typedef boost::shared_ptr<pss> sptr;
  static sptr make(int N_id_2, float psr_threshold, int track_after = 16, int track_every = 8);
  virtual float max_psr() const = 0;
  virtual float mean_psr() const = 0;
  virtual float mean_cfo() const = 0;
  virtual void set_psr_threshold(float threshold) = 0;
  virtual float psr_threshold() const = 0;
  virtual float tracking_score() const = 0;
};
} }
