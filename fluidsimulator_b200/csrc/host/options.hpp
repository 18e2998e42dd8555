// options.hpp — `--name value` / `--name=value` / flags / `--` positionals / `-h`, the
// behaviour of the reference's cli/ module (reference cli/include/fluid/cli.h:10-36,
// cli/src/cli.cpp:34-115), including its error strings.
#pragma once

#include <map>
#include <string>
#include <vector>

namespace fluid {
namespace b200 {

struct Option {
  std::string name;
  bool takes_value = false;
  std::string help;
};

struct ParsedOptions {
  bool ok = true;
  std::string error;
  std::map<std::string, std::string> values;
  std::vector<std::string> positionals;
  bool has(const std::string& name) const { return values.count(name) != 0; }
  std::string value(const std::string& name, const std::string& fallback = "") const {
    auto it = values.find(name);
    return it == values.end() ? fallback : it->second;
  }
};

ParsedOptions parse_options(int argc, char** argv, const std::vector<Option>& options);
std::string usage_text(const char* argv0, const std::vector<Option>& options);

}  // namespace b200
}  // namespace fluid
