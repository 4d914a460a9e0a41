#include "ConfigMap.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <vector>

namespace euler_b200 {

namespace {
constexpr size_t kMaxLine = 200; // inih MAX_LINE: fgets reads at most 199 chars per "line"

std::string rstrip(std::string s)
{
  while (!s.empty() && std::isspace(static_cast<unsigned char>(s.back())))
    s.pop_back();
  return s;
}
size_t lskip(const std::string& s, size_t pos = 0)
{
  while (pos < s.size() && std::isspace(static_cast<unsigned char>(s[pos])))
    ++pos;
  return pos;
}
// first occurrence of c, or of a ';' preceded by white space; npos if neither
size_t find_char_or_comment(const std::string& s, size_t pos, char c)
{
  bool was_ws = false;
  for (; pos < s.size(); ++pos) {
    if (c != '\0' && s[pos] == c)
      return pos;
    if (was_ws && s[pos] == ';')
      return pos;
    was_ws = std::isspace(static_cast<unsigned char>(s[pos])) != 0;
  }
  return std::string::npos;
}
} // namespace

std::string ConfigMap::makeKey(const std::string& section, const std::string& name)
{
  std::string key = section + "." + name;
  for (char& ch : key)
    ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
  return key;
}

void ConfigMap::parseLine(std::string line, std::string& section, std::string& prev_name, int lineno)
{
  line = rstrip(line);
  const size_t st = lskip(line);
  const std::string body = line.substr(st);
  if (!prev_name.empty() && !body.empty() && st > 0) {
    // continuation of the previous name's value (replaces it, as the reference's handler does)
    values_[makeKey(section, prev_name)] = body;
    return;
  }
  if (body.empty() || body[0] == ';' || body[0] == '#')
    return;
  if (body[0] == '[') {
    const size_t end = find_char_or_comment(body, 1, ']');
    if (end != std::string::npos && body[end] == ']') {
      section = body.substr(1, end - 1).substr(0, 49); // MAX_SECTION 50
      prev_name.clear();
    } else if (!error_) {
      error_ = lineno;
    }
    return;
  }
  const size_t eq = find_char_or_comment(body, 0, '=');
  if (eq != std::string::npos && body[eq] == '=') {
    std::string name = rstrip(body.substr(0, eq));
    std::string value = body.substr(eq + 1);
    value = value.substr(lskip(value));
    const size_t cm = find_char_or_comment(value, 0, '\0');
    if (cm != std::string::npos)
      value = value.substr(0, cm);
    value = rstrip(value);
    prev_name = name.substr(0, 49); // MAX_NAME 50
    values_[makeKey(section, name)] = value;
  } else if (!error_) {
    error_ = lineno;
  }
}

ConfigMap::ConfigMap(const std::string& filename)
{
  FILE* f = std::fopen(filename.c_str(), "r");
  if (!f) {
    error_ = -1;
    return;
  }
  char buf[kMaxLine];
  std::string section, prev_name;
  int lineno = 0;
  while (std::fgets(buf, sizeof(buf), f) != nullptr)
    parseLine(buf, section, prev_name, ++lineno);
  std::fclose(f);
}

ConfigMap ConfigMap::fromBuffer(const char* buffer, size_t size)
{
  ConfigMap m;
  std::istringstream ss(std::string(buffer, size));
  std::string section, prev_name, line;
  int lineno = 0;
  while (std::getline(ss, line))
    m.parseLine(line, section, prev_name, ++lineno);
  return m;
}

std::string ConfigMap::getString(const std::string& section, const std::string& name, const std::string& default_value) const
{
  auto it = values_.find(makeKey(section, name));
  return it != values_.end() ? it->second : default_value;
}
void ConfigMap::setString(const std::string& section, const std::string& name, const std::string& value)
{
  values_[makeKey(section, name)] = value;
}
long ConfigMap::getInteger(const std::string& section, const std::string& name, long default_value) const
{
  const std::string v = getString(section, name, "");
  char* end = nullptr;
  const long n = std::strtol(v.c_str(), &end, 0);
  return end > v.c_str() ? n : default_value;
}
void ConfigMap::setInteger(const std::string& section, const std::string& name, long value)
{
  setString(section, name, std::to_string(value));
}
float ConfigMap::getFloat(const std::string& section, const std::string& name, float default_value) const
{
  const std::string v = getString(section, name, "");
  char* end = nullptr;
  const float x = std::strtof(v.c_str(), &end);
  return end > v.c_str() ? x : default_value;
}
void ConfigMap::setFloat(const std::string& section, const std::string& name, float value)
{
  std::stringstream ss;
  ss << value;
  setString(section, name, ss.str());
}
bool ConfigMap::getBool(const std::string& section, const std::string& name, bool default_value) const
{
  bool val = default_value;
  const std::string v = getString(section, name, "");
  if (v == "1" || v == "yes" || v == "true" || v == "on")
    val = true;
  if (v == "0" || v == "no" || v == "false" || v == "off")
    val = false;
  if (v.empty())
    val = default_value;
  return val;
}
void ConfigMap::setBool(const std::string& section, const std::string& name, bool value)
{
  setString(section, name, value ? "true" : "false");
}

void ConfigMap::applyOverrides(const std::string& overrides)
{
  std::istringstream ss(overrides);
  std::string line;
  while (std::getline(ss, line)) {
    const size_t eq = line.find('=');
    const size_t dot = line.find('.');
    if (eq == std::string::npos || dot == std::string::npos || dot > eq)
      continue;
    setString(line.substr(0, dot), line.substr(dot + 1, eq - dot - 1), line.substr(eq + 1));
  }
}

std::string ConfigMap::toIni() const
{
  std::map<std::string, std::vector<std::pair<std::string, std::string>>> by_section;
  for (const auto& kv : values_) {
    const size_t dot = kv.first.find('.');
    by_section[kv.first.substr(0, dot)].emplace_back(kv.first.substr(dot + 1), kv.second);
  }
  std::string out;
  for (const auto& s : by_section) {
    out += "[" + s.first + "]\n";
    for (const auto& nv : s.second)
      out += nv.first + "=" + nv.second + "\n";
    out += "\n";
  }
  return out;
}

ConfigMap broadcast_parameters(const std::string& filename)
{
  ConfigMap m(filename);
  if (m.ParseError() < 0)
    std::printf("Error : can't read input parameter file \"%s\"\nPlease check file actually exists !\n", filename.c_str());
  return m;
}

} // namespace euler_b200
