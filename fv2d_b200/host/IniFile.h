// IniFile — .ini parser for the host driver.
//
// Behavioural restatement (not a copy) of the parser the reference uses,
// jtilly/inih `INIReader` (reference external/inih/INIReader.h), because run parity
// depends on its quirks:
//   * lines are read with fgets into a 200-byte buffer, so longer lines are split
//     (INIReader.h:91-92, 169-184);
//   * optional UTF-8 BOM on line 1 (:188-194);
//   * `;` / `#` start a comment line (:197-200); an inline `;` comment needs a
//     preceding whitespace character (:143-150);
//   * a non-blank line with leading whitespace continues the previous value (:202-216);
//   * `[section]` (49 chars kept, :219-231) and `name = value` / `name : value` (:233-252);
//   * keys are `section=name` lower-cased (:442-448); a repeated key appends
//     "\n" + value (:450-459); section names are recorded in their raw case;
//   * GetFloat parses with strtof and returns *float* (:420-427); GetInteger uses
//     strtol(...,0) (:401-409); GetBoolean accepts true/yes/on/1, false/no/off/0
//     case-insensitively (:429-440).
#pragma once

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>

namespace fv2d
{

class IniFile
{
public:
  IniFile() = default;
  explicit IniFile(const std::string &filename) { error_ = parse(filename); }

  int ParseError() const { return error_; }
  const std::set<std::string> &Sections() const { return sections_; }
  const std::map<std::string, std::string> &Values() const { return values_; }

  static std::string MakeKey(const std::string &section, const std::string &name)
  {
    std::string key = section + "=" + name;
    std::transform(key.begin(), key.end(), key.begin(), [](unsigned char c) { return char(std::tolower(c)); });
    return key;
  }

  std::string Get(const std::string &section, const std::string &name, const std::string &default_value) const
  {
    auto it = values_.find(MakeKey(section, name));
    return it == values_.end() ? default_value : it->second;
  }

  long GetInteger(const std::string &section, const std::string &name, long default_value) const
  {
    const std::string s = Get(section, name, "");
    char *end           = nullptr;
    long n              = std::strtol(s.c_str(), &end, 0);
    return end > s.c_str() ? n : default_value;
  }

  // Single precision on purpose: every real parameter of a run is float-valued (Q1).
  float GetFloat(const std::string &section, const std::string &name, float default_value) const
  {
    const std::string s = Get(section, name, "");
    char *end           = nullptr;
    float x             = std::strtof(s.c_str(), &end);
    return end > s.c_str() ? x : default_value;
  }

  bool GetBoolean(const std::string &section, const std::string &name, bool default_value) const
  {
    std::string s = Get(section, name, "");
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return char(std::tolower(c)); });
    if (s == "true" || s == "yes" || s == "on" || s == "1")
      return true;
    if (s == "false" || s == "no" || s == "off" || s == "0")
      return false;
    return default_value;
  }

  // Additions the reference makes in its IniReader subclass (SimInfo.h:100-120).
  bool HasSection(const std::string &section) const
  {
    const std::string key = MakeKey(section, "");
    auto pos              = values_.lower_bound(key);
    if (pos == values_.end())
      return false;
    return pos->first.compare(0, key.length(), key) == 0;
  }
  bool HasValue(const std::string &section, const std::string &name) const
  {
    return values_.count(MakeKey(section, name)) != 0;
  }

private:
  static constexpr int kMaxLine    = 200;
  static constexpr int kMaxSection = 50;
  static constexpr int kMaxName    = 50;

  int error_ = 0;
  std::map<std::string, std::string> values_;
  std::set<std::string> sections_;

  static char *rstrip(char *s)
  {
    char *p = s + std::strlen(s);
    while (p > s && std::isspace((unsigned char)(*--p)))
      *p = '\0';
    return s;
  }
  static char *lskip(char *s)
  {
    while (*s && std::isspace((unsigned char)(*s)))
      s++;
    return s;
  }
  // first char in `chars`, or start of an inline comment (";" preceded by whitespace), or end
  static char *find_stop(char *s, const char *chars)
  {
    bool was_space = false;
    while (*s && (!chars || !std::strchr(chars, *s)) && !(was_space && *s == ';'))
    {
      was_space = std::isspace((unsigned char)(*s)) != 0;
      s++;
    }
    return s;
  }

  void store(const char *section, const char *name, const char *value)
  {
    std::string &slot = values_[MakeKey(section, name)];
    if (!slot.empty())
      slot += "\n";
    slot += value;
    sections_.insert(section);
  }

  int parse(const std::string &filename)
  {
    FILE *file = std::fopen(filename.c_str(), "r");
    if (!file)
      return -1;
    char line[kMaxLine];
    char section[kMaxSection] = "";
    char prev_name[kMaxName]  = "";
    int lineno = 0, error = 0;

    while (std::fgets(line, kMaxLine, file) != nullptr)
    {
      lineno++;
      char *start = line;
      if (lineno == 1 && (unsigned char)start[0] == 0xEF && (unsigned char)start[1] == 0xBB &&
          (unsigned char)start[2] == 0xBF)
        start += 3;
      start = lskip(rstrip(start));

      if (*start == ';' || *start == '#')
      {
        // comment line
      }
      else if (*prev_name && *start && start > line)
      {
        // continuation of the previous value
        char *end = find_stop(start, nullptr);
        if (*end)
          *end = '\0';
        rstrip(start);
        store(section, prev_name, start);
      }
      else if (*start == '[')
      {
        char *end = find_stop(start + 1, "]");
        if (*end == ']')
        {
          *end = '\0';
          std::strncpy(section, start + 1, sizeof(section));
          section[sizeof(section) - 1] = '\0';
          *prev_name                   = '\0';
        }
        else if (!error)
          error = lineno;
      }
      else if (*start)
      {
        char *end = find_stop(start, "=:");
        if (*end == '=' || *end == ':')
        {
          *end        = '\0';
          char *name  = rstrip(start);
          char *value = lskip(end + 1);
          end         = find_stop(value, nullptr);
          if (*end)
            *end = '\0';
          rstrip(value);
          std::strncpy(prev_name, name, sizeof(prev_name));
          prev_name[sizeof(prev_name) - 1] = '\0';
          store(section, name, value);
        }
        else if (!error)
          error = lineno;
      }
    }
    std::fclose(file);
    return error;
  }
};

} // namespace fv2d
