// Stand-in for yaml-cpp (absent here): just enough for VIOFilterSettings.h / common.h to compile.
// The harness fills VIOFilter::Settings field by field and never parses YAML through this.
#pragma once
#include <string>
namespace YAML {
class Node {
  public:
    explicit operator bool() const { return false; }
    template <typename K> Node operator[](const K&) const { return Node(); }
    template <typename T> T as() const { return T(); }
};
inline Node LoadFile(const std::string&) { return Node(); }
}  // namespace YAML
