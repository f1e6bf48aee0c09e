// TEST INFRASTRUCTURE -- test double of boost::split / boost::is_any_of as the reference's
// global_costmap uses them (token_compress_off: adjacent separators give empty tokens).
#pragma once
#include <string>
#include <vector>
namespace boost {
struct is_any_of { std::string set; explicit is_any_of(const std::string& s) : set(s) {} };
template <class Seq>
inline Seq& split(Seq& out, const std::string& in, const is_any_of& pred) {
  out.clear();
  std::string cur;
  for (char c : in) {
    if (pred.set.find(c) != std::string::npos) { out.push_back(cur); cur.clear(); }
    else cur.push_back(c);
  }
  out.push_back(cur);
  return out;
}
}  // namespace boost
