// logger.h -- minimal stand-in for the reference's Log class (src/log.h:28-47): same static
// entry points (info / warn / error / newline), plain timestamp-free lines on stderr.
// Quiet unless CA_LOG=1 so the real-time path and the test output stay clean.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>

class Log {
public:
    static bool enabled() { static const bool on = [] { const char *e = getenv("CA_LOG"); return e && *e && *e != '0'; }(); return on; }
    static void info(const std::string &id, const char *fmt, ...) noexcept { va_list ap; va_start(ap, fmt); emit('I', id.c_str(), fmt, ap, false); va_end(ap); }
    static void warn(const std::string &id, const char *fmt, ...) noexcept { va_list ap; va_start(ap, fmt); emit('W', id.c_str(), fmt, ap, true); va_end(ap); }
    static void error(const std::string &id, const char *fmt, ...) noexcept { va_list ap; va_start(ap, fmt); emit('E', id.c_str(), fmt, ap, true); va_end(ap); }
    static void newline() noexcept { if (enabled()) fputc('\n', stderr); }
    static void newline(const char *fmt, ...) noexcept { va_list ap; va_start(ap, fmt); emit(' ', "", fmt, ap, false); va_end(ap); }

private:
    static void emit(char lvl, const char *id, const char *fmt, va_list ap, bool always) noexcept
    {
        if (!always && !enabled()) return;
        char line[512];
        vsnprintf(line, sizeof(line), fmt, ap);
        fprintf(stderr, "%c [%s] %s\n", lvl, id, line);
    }
};
