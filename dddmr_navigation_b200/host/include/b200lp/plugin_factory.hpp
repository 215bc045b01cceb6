// b200lp::PluginFactory — type-string -> instance, the job pluginlib::ClassLoader::createSharedInstance does in the
// reference (trajectory_generators_ros.cpp:75, mpc_critics_ros.cpp:79). The type strings are the reference's
// (trajectory_generators.xml, mpc_critics.xml), so its YAML files load the B200 adapters unchanged. With ROS 2 the
// same classes are exported through PLUGINLIB_EXPORT_CLASS instead.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>

namespace b200lp {
template <class Base>
class PluginFactory {
 public:
  using Maker = std::function<std::shared_ptr<Base>()>;
  static PluginFactory& instance() {
    static PluginFactory f;
    return f;
  }
  void add(const std::string& type, Maker m) { makers_[type] = std::move(m); }
  std::shared_ptr<Base> createSharedInstance(const std::string& type) const {
    auto it = makers_.find(type);
    if (it == makers_.end()) throw std::runtime_error("b200lp: no plugin registered for type '" + type + "'");
    return it->second();
  }
  bool isClassAvailable(const std::string& type) const { return makers_.count(type) != 0; }

 private:
  std::map<std::string, Maker> makers_;
};
}  // namespace b200lp
