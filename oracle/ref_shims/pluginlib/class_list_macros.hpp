// pluginlib stand-in (TEST INFRASTRUCTURE, see ../lpref_eigen.hpp): PLUGINLIB_EXPORT_CLASS registers a factory under the
// class's type string in a process-wide registry, which is how oracle/lpref_driver.cpp instantiates the reference's plugins
// without including their (guard-less) headers twice.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>

namespace lpref {
template <class Base>
std::map<std::string, std::function<std::shared_ptr<Base>()>>& registry() {
  static std::map<std::string, std::function<std::shared_ptr<Base>()>> r;
  return r;
}
template <class Base>
std::shared_ptr<Base> create(const std::string& type) {
  auto it = registry<Base>().find(type);
  if (it == registry<Base>().end()) throw std::runtime_error("no plugin registered as " + type);
  return it->second();
}
}  // namespace lpref

#define LPREF_CAT2(a, b) a##b
#define LPREF_CAT(a, b) LPREF_CAT2(a, b)
#define PLUGINLIB_EXPORT_CLASS(class_type, base_class_type)                                                         \
  namespace {                                                                                                       \
  struct LPREF_CAT(LprefRegister, __LINE__) {                                                                       \
    LPREF_CAT(LprefRegister, __LINE__)() {                                                                          \
      lpref::registry<base_class_type>()[#class_type] = [] { return std::shared_ptr<base_class_type>(new class_type()); }; \
    }                                                                                                               \
  } LPREF_CAT(lpref_register_instance, __LINE__);                                                                   \
  }
