// oracle/shim/yaml-cpp/yaml.h — TEST INFRASTRUCTURE ONLY.  The reference's Settings.h (pulled in by FeatureMatcher.h)
// declares a YAML::Node member; nothing on the matcher path reads a configuration file.  Declared so the header parses.
#pragma once
#include <cstdlib>
#include <string>
namespace YAML {
struct Node {
    Node operator[](const std::string&) const { return Node(); }
    template <class T> T as() const { std::abort(); }
};
inline Node LoadFile(const std::string&) { std::abort(); }
}  // namespace YAML
