// INI reader with the reference's Ini_reader interface (source/Ini_reader.h:17-77:
// set_section / read<T>(key, T*) / read<T>(section, key, T*), section_not_found and
// key_not_found thrown for missing entries) on top of a small parser that follows
// the mINI semantics the reference relies on (source/ini.h): section and key names
// are case-insensitive, ';' starts a comment line, text is trimmed, lines without
// '=' are ignored, a repeated key keeps the last value.
#ifndef SY2D_HOST_INI_READER_H_
#define SY2D_HOST_INI_READER_H_

#include <algorithm>
#include <cctype>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

class Ini_reader {
 public:
  struct section_not_found : std::runtime_error {
    std::string section;
    explicit section_not_found(const std::string& s = std::string()) : std::runtime_error("ini: section not found: " + s), section(s) {}
  };
  struct key_not_found : std::runtime_error {
    std::string key;
    explicit key_not_found(const std::string& k = std::string()) : std::runtime_error("ini: key not found: " + k), key(k) {}
  };

  explicit Ini_reader(const std::string& filename) {
    std::ifstream in(filename);
    std::string line, section;
    while (std::getline(in, line)) {
      line = trim(line);
      if (line.empty() || line[0] == ';') continue;
      if (line[0] == '[') {
        const auto close = line.find(']');
        if (close != std::string::npos) section = lower(trim(line.substr(1, close - 1)));
        data_[section];
        continue;
      }
      const auto eq = line.find('=');
      if (eq == std::string::npos) continue;
      data_[section][lower(trim(line.substr(0, eq)))] = trim(line.substr(eq + 1));
    }
  }

  bool has(const std::string& section) const { return data_.count(lower(section)) != 0; }
  bool has(const std::string& section, const std::string& key) const {
    auto s = data_.find(lower(section));
    return s != data_.end() && s->second.count(lower(key)) != 0;
  }

  template <typename T>
  void read(const std::string& section, const std::string& key, T* valuep) {
    if (!has(section)) throw section_not_found(section);
    if (!has(section, key)) throw key_not_found(key);
    string_as_T<T>(data_[lower(section)][lower(key)], *valuep);
  }
  void set_section(const std::string& section) { section_ = section; }
  template <typename T>
  void read(const std::string& key, T* valuep) { read(section_, key, valuep); }

 private:
  std::map<std::string, std::map<std::string, std::string>> data_;
  std::string section_;

  static std::string trim(const std::string& s) {
    const auto b = s.find_first_not_of(" \t\r\n");
    if (b == std::string::npos) return std::string();
    return s.substr(b, s.find_last_not_of(" \t\r\n") - b + 1);
  }
  static std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return static_cast<char>(std::tolower(c)); });
    return s;
  }
  template <class T>
  static void string_as_T(const std::string& s, T& t) {  // stream extraction, as the reference does
    std::istringstream ist(s);
    ist >> t;
  }
};

template <>
inline void Ini_reader::string_as_T<std::string>(const std::string& s, std::string& t) { t = s; }
template <>
inline void Ini_reader::string_as_T<bool>(const std::string& s, bool& b) {
  std::string u = s;
  std::transform(u.begin(), u.end(), u.begin(), [](unsigned char c) { return static_cast<char>(std::toupper(c)); });
  b = !(u == "FALSE" || u == "F" || u == "NO" || u == "N" || u == "0" || u == "NONE");
}

#endif
