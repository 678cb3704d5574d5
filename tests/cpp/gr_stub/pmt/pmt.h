// minimal pmt: symbols, nil, booleans, longs
#pragma once
#include <map>
#include <memory>
#include <string>
namespace pmt {
struct pmt_base { int kind; std::string sym; long val; };   // kind: 0 nil, 1 bool, 2 long, 3 symbol
typedef std::shared_ptr<pmt_base> pmt_t;
inline pmt_t make_(int k, const std::string &s, long v) { pmt_t p(new pmt_base()); p->kind = k; p->sym = s; p->val = v; return p; }
inline pmt_t intern(const std::string &s) {
  static std::map<std::string, pmt_t> table;
  pmt_t &p = table[s];
  if (!p) p = make_(3, s, 0);
  return p;
}
inline pmt_t string_to_symbol(const std::string &s) { return intern(s); }
inline std::string symbol_to_string(const pmt_t &p) { return p->sym; }
inline pmt_t from_long(long v) { return make_(2, "", v); }
inline long to_long(const pmt_t &p) { return p->val; }
inline bool eq(const pmt_t &a, const pmt_t &b) { return a.get() == b.get(); }
// one object per process, whatever the translation unit (GNU Radio's are library globals)
inline pmt_t get_PMT_NIL() { static const pmt_t p = make_(0, "", 0); return p; }
inline pmt_t get_PMT_T() { static const pmt_t p = make_(1, "", 1); return p; }
inline pmt_t get_PMT_F() { static const pmt_t p = make_(1, "", 0); return p; }
#define PMT_NIL get_PMT_NIL()
#define PMT_T get_PMT_T()
#define PMT_F get_PMT_F()
inline bool is_true(const pmt_t &p) { return !(p->kind == 1 && p->val == 0); }
}  // namespace pmt
