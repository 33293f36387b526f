// settings_file.h -- the reference's text configuration (src/settings.{h,cu}, settings.txt):
// whitespace-separated `key value` pairs, `#` starts a comment that runs to end of line,
// keys looked up through printf-style formats such as "conv[%d].value.wet".
// Same class names and accessors; a missing numeric key throws (settings.cu:72-84).
#pragma once
#include <cstdint>
#include <map>
#include <string>

class Setting {
public:
    std::string key, value;
    bool isTrue() const { return value == "yes" || value == "true"; }
    bool isFalse() const { return !isTrue(); }
    uint8_t u8() const { return (uint8_t)(std::stoi(value) & 0xFF); }
    uint16_t u16() const { return (uint16_t)(std::stoi(value) & 0xFFFF); }
    uint32_t u32() const { return (uint32_t)std::stoi(value); }
    float f32() const { return std::stof(value); }
    const std::string &str() const { return value; }
};

class Settings : public std::map<std::string, Setting> {
public:
    void open(const std::string &path);
    void parse(const std::string &text);
    bool has(const std::string &fmt, ...) const;
    bool isTrue(const std::string &fmt, ...);
    bool isFalse(const std::string &fmt, ...);
    uint8_t u8(const std::string &fmt, ...);
    uint16_t u16(const std::string &fmt, ...);
    uint32_t u32(const std::string &fmt, ...);
    float f32(const std::string &fmt, ...);
    const std::string &str(const std::string &fmt, ...);
};
