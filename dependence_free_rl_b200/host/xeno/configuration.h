// xeno/configuration.h -- command-line flags with the reference's flagstore interface
// (xeno/configuration.h:17-118: define_flag / set_flag / get_flag / parse_from_args), written for the
// host mirror: the trainer mains use it to turn the reference's compile-time problem definition
// (bin_packing.h:12 `constexpr num_bins`, :24 capacity, :73-78 item shapes) into launch parameters.
//   --name=value | --name value | --boolean_flag | -n value | -b | --  (everything after it is positional)
#ifndef XENO_CONFIGURATION_
#define XENO_CONFIGURATION_

#include <any>
#include <map>
#include <span>
#include <string>
#include <string_view>
#include <typeinfo>
#include <vector>

#include <xeno/exception.h>

namespace xeno {

class flagstore {
public:
  template <typename T> void define_flag(std::string_view name, char short_name, const T &default_value) {
    const std::size_t slot = values_.size();
    if (!name.empty())
      by_name_[std::string(name)] = slot;
    if (short_name != 0)
      by_letter_[short_name] = slot;
    values_.emplace_back(default_value);
  }

  template <typename V> void set_flag(std::string_view name, V &&v) { values_[slot_of(name)] = std::forward<V>(v); }
  template <typename V> void set_flag(char name, V &&v) { values_[slot_of(name)] = std::forward<V>(v); }

  template <typename T> const T &get_flag(std::string_view name) { return typed<T>(values_[slot_of(name)], std::string(name)); }
  template <typename T> const T &get_flag(char name) { return typed<T>(values_[slot_of(name)], std::string(1, name)); }

  // Consumes the options, returns the positional arguments in order.
  std::vector<std::string_view> parse_from_args(std::span<std::string_view> argv) {
    std::vector<std::string_view> positional;
    bool only_positional = false;
    for (std::size_t i = 0; i < argv.size(); ++i) {
      const std::string_view arg = argv[i];
      if (only_positional || arg.size() < 2 || arg[0] != '-') {
        if (arg != "-")
          positional.push_back(arg);
        continue;
      }
      if (arg == "--") {
        only_positional = true;
        continue;
      }
      if (arg[1] == '-') {  // --name[=value] | --name value
        const std::string_view body = arg.substr(2);
        const std::size_t eq = body.find('=');
        const std::string_view name = body.substr(0, eq);
        std::any &slot = values_[slot_of(name)];
        if (eq != std::string_view::npos)
          assign(slot, name, body.substr(eq + 1));
        else if (slot.type() == typeid(bool))
          assign(slot, name, "");
        else if (i + 1 < argv.size())
          assign(slot, name, argv[++i]);
        else
          throw xeno::error("flag " + std::string(name) + " needs a value");
      } else {  // -n value | -b
        if (arg.size() != 2)
          throw xeno::error("short flags take one letter: " + std::string(arg));
        std::any &slot = values_[slot_of(arg[1])];
        if (slot.type() == typeid(bool))
          assign(slot, arg.substr(1), "");
        else if (i + 1 < argv.size())
          assign(slot, arg.substr(1), argv[++i]);
        else
          throw xeno::error("flag " + std::string(arg) + " needs a value");
      }
    }
    return positional;
  }
  std::vector<std::string_view> parse_from_args(int argc, char **argv) {  // skips argv[0]
    std::vector<std::string_view> v(argv + (argc > 0 ? 1 : 0), argv + argc);
    return parse_from_args(std::span<std::string_view>(v));
  }

private:
  template <typename T> static const T &typed(const std::any &a, const std::string &name) {
    if (a.type() != typeid(T))
      throw xeno::error("flag " + name + " has type " + a.type().name());
    return *std::any_cast<T>(&a);
  }
  static void assign(std::any &slot, std::string_view name, std::string_view text) {
    const std::string s(text);
    try {
      if (slot.type() == typeid(std::string))
        slot = s;
      else if (slot.type() == typeid(int))
        slot = std::stoi(s);
      else if (slot.type() == typeid(long))
        slot = std::stol(s);
      else if (slot.type() == typeid(std::size_t))
        slot = (std::size_t)std::stoull(s);
      else if (slot.type() == typeid(float))
        slot = std::stof(s);
      else if (slot.type() == typeid(double))
        slot = std::stod(s);
      else if (slot.type() == typeid(bool)) {
        if (s.empty() || s == "true")
          slot = true;
        else if (s == "false")
          slot = false;
        else
          throw xeno::error("flag " + std::string(name) + " expects boolean value");
      } else
        throw xeno::error("flag " + std::string(name) + " has a type the command line cannot set");
    } catch (const std::logic_error &) {  // stoi / stod: invalid_argument, out_of_range
      throw xeno::error("flag " + std::string(name) + ": cannot parse '" + s + "'");
    }
  }
  std::size_t slot_of(std::string_view name) const {
    auto it = by_name_.find(std::string(name));
    if (it == by_name_.end())
      throw xeno::error("undefined flag " + std::string(name));
    return it->second;
  }
  std::size_t slot_of(char name) const {
    auto it = by_letter_.find(name);
    if (it == by_letter_.end())
      throw xeno::error(std::string("undefined flag short name ") + name);
    return it->second;
  }

  std::vector<std::any> values_;
  std::map<std::string, std::size_t> by_name_;
  std::map<char, std::size_t> by_letter_;
};

} // namespace xeno

#endif // XENO_CONFIGURATION_
