#include "settings_file.h"

#include <cstdarg>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

#include "logger.h"

static std::string format_key(const std::string &fmt, va_list ap)
{
    va_list ap2;
    va_copy(ap2, ap);
    const int n = vsnprintf(nullptr, 0, fmt.c_str(), ap2);
    va_end(ap2);
    std::vector<char> buf((size_t)(n > 0 ? n : 0) + 1);
    vsnprintf(buf.data(), buf.size(), fmt.c_str(), ap);
    return std::string(buf.data());
}

#define CA_KEY(fmt, key) va_list _ap; va_start(_ap, fmt); const std::string key = format_key(fmt, _ap); va_end(_ap)

void Settings::parse(const std::string &text)
{
    // token based like settings.cu:7-22: a key token starting with '#' discards the rest of its
    // line; otherwise the next whitespace-delimited token (even on a later line) is the value.
    std::istringstream is(text);
    std::string key;
    while (is >> key) {
        if (key[0] == '#') { std::string rest; std::getline(is, rest); continue; }
        std::string value;
        is >> value;
        (*this)[key] = Setting{key, value};
        Log::info("Settings", "%-24s %s", key.c_str(), value.c_str());
    }
}

void Settings::open(const std::string &path)
{
    std::ifstream is(path, std::ifstream::binary);
    if (!is) throw std::runtime_error("cannot open settings file " + path);
    std::stringstream ss;
    ss << is.rdbuf();
    parse(ss.str());
}

bool Settings::has(const std::string &fmt, ...) const { CA_KEY(fmt, key); return find(key) != end(); }

template <class F>
static auto typed(Settings &s, const std::string &key, F f) -> decltype(f(s[key]))
{
    try { return f(s[key]); }
    catch (std::exception &) { Log::error("Settings", "Error for key %s", key.c_str()); throw; }
}

bool Settings::isTrue(const std::string &fmt, ...) { CA_KEY(fmt, key); return typed(*this, key, [](Setting &v) { return v.isTrue(); }); }
bool Settings::isFalse(const std::string &fmt, ...) { CA_KEY(fmt, key); return typed(*this, key, [](Setting &v) { return v.isFalse(); }); }
uint8_t Settings::u8(const std::string &fmt, ...) { CA_KEY(fmt, key); return typed(*this, key, [](Setting &v) { return v.u8(); }); }
uint16_t Settings::u16(const std::string &fmt, ...) { CA_KEY(fmt, key); return typed(*this, key, [](Setting &v) { return v.u16(); }); }
uint32_t Settings::u32(const std::string &fmt, ...) { CA_KEY(fmt, key); return typed(*this, key, [](Setting &v) { return v.u32(); }); }
float Settings::f32(const std::string &fmt, ...) { CA_KEY(fmt, key); return typed(*this, key, [](Setting &v) { return v.f32(); }); }
const std::string &Settings::str(const std::string &fmt, ...) { CA_KEY(fmt, key); return (*this)[key].value; }
