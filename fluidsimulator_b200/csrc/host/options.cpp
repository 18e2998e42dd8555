#include "options.hpp"

namespace fluid {
namespace b200 {

ParsedOptions parse_options(int argc, char** argv, const std::vector<Option>& options) {
  ParsedOptions parsed;
  auto lookup = [&](const std::string& name) -> const Option* {
    for (const Option& o : options)
      if (o.name == name) return &o;
    return nullptr;
  };
  auto reject = [&](const std::string& why) {
    parsed.ok = false;
    parsed.error = why;
    return parsed;
  };
  for (int i = 1; i < argc; ++i) {
    const std::string arg = argv[i];
    if (arg == "--") {  // everything after is positional
      parsed.positionals.insert(parsed.positionals.end(), argv + i + 1, argv + argc);
      break;
    }
    if (arg == "-h" || arg == "--help") {
      parsed.values.emplace("help", "1");
      continue;
    }
    if (arg.compare(0, 2, "--") != 0) {
      parsed.positionals.push_back(arg);
      continue;
    }
    std::string name = arg.substr(2), value;
    const std::size_t eq = name.find('=');
    const bool inline_value = eq != std::string::npos;
    if (inline_value) {
      value = name.substr(eq + 1);
      name.erase(eq);
    }
    const Option* opt = lookup(name);
    if (!opt) return reject("Unknown option: --" + name);
    if (!opt->takes_value) {
      if (inline_value && !value.empty()) return reject("Option does not take a value: --" + name);
      parsed.values[name] = "1";
      continue;
    }
    if (!inline_value) {
      if (i + 1 >= argc) return reject("Missing value for option: --" + name);
      value = argv[++i];
    }
    parsed.values[name] = value;
  }
  return parsed;
}

std::string usage_text(const char* argv0, const std::vector<Option>& options) {
  std::string out = std::string("Usage: ") + argv0 + " [options]\n\nOptions:\n  -h, --help\n";
  for (const Option& o : options) {
    out += "  --" + o.name;
    if (o.takes_value) out += " <value>";
    if (!o.help.empty()) out += "\n      " + o.help;
    out += '\n';
  }
  return out;
}

}  // namespace b200
}  // namespace fluid
