// ini-file dictionary with the reference's exact look-up semantics.
//
// Mirrors euler_kokkos::ConfigMap / INIReader (src/utils/config/ConfigMap.{h,cpp},
// src/utils/config/inih/INIReader.{h,cpp}, inih/ini.cpp):
//  * keys are "section.name", lower-cased (INIReader.cpp:100-107) => look-ups are case-insensitive;
//  * getFloat parses with strtof and RETURNS A float (ConfigMap.cpp:38-46): every real-valued
//    parameter is a float widened to double (gamma0=1.666 -> 1.66600000858306884765625).  This is
//    load-bearing for 1e-12 parity with the reference;
//  * getInteger uses strtol(base 0) (INIReader.cpp:58-67); getBool accepts 1/yes/true/on and
//    0/no/false/off (ConfigMap.cpp:64-82);
//  * line syntax of inih: leading/trailing blanks stripped, ';' or '#' at line start is a comment,
//    " ;" starts an inline comment, "[section]", "name = value", a non-blank line that starts with
//    white space continues (replaces) the previous name's value, lines longer than 199 chars are
//    split as fgets would (ini.cpp:14,66-140).
#pragma once
#include <map>
#include <string>

namespace euler_b200 {

class ConfigMap {
public:
  ConfigMap() = default;
  explicit ConfigMap(const std::string& filename);
  static ConfigMap fromBuffer(const char* buffer, size_t size);

  //! 0 on success, -1 if the file could not be opened, else the first line in error
  int ParseError() const { return error_; }

  std::string getString(const std::string& section, const std::string& name, const std::string& default_value) const;
  void setString(const std::string& section, const std::string& name, const std::string& value);
  long getInteger(const std::string& section, const std::string& name, long default_value) const;
  void setInteger(const std::string& section, const std::string& name, long value);
  float getFloat(const std::string& section, const std::string& name, float default_value) const;
  void setFloat(const std::string& section, const std::string& name, float value);
  bool getBool(const std::string& section, const std::string& name, bool default_value) const;
  void setBool(const std::string& section, const std::string& name, bool value);

  //! apply "section.name=value" overrides separated by newlines (used by the bench / tests)
  void applyOverrides(const std::string& overrides);
  //! serialise back to ini text (sections sorted)
  std::string toIni() const;
  const std::map<std::string, std::string>& values() const { return values_; }

private:
  void parseLine(std::string line, std::string& section, std::string& prev_name, int lineno);
  static std::string makeKey(const std::string& section, const std::string& name);
  std::map<std::string, std::string> values_;
  int error_ = 0;
};

//! single-process stand-in of the reference's MPI broadcast_parameters (ConfigMap.cpp:100-173)
ConfigMap broadcast_parameters(const std::string& filename);

} // namespace euler_b200
