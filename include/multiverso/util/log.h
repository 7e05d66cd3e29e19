// multiverso-b200 host runtime :: logging (counterpart of the reference's
// include/multiverso/util/log.h:9-142 -- same levels, CHECK macros and line format,
// plus a rank prefix; re-implemented on std::mutex + vsnprintf).
#ifndef MULTIVERSO_UTIL_LOG_H_
#define MULTIVERSO_UTIL_LOG_H_
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>

namespace multiverso {

enum class LogLevel : int { Debug = 0, Info = 1, Error = 2, Fatal = 3 };

class Logger {
 public:
  Logger() = default;
  ~Logger();
  int ResetLogFile(const std::string& filename);           // "" closes the file
  void ResetLogLevel(LogLevel level) { level_ = level; }
  void ResetKillFatal(bool kill) { kill_fatal_ = kill; }
  void SetRank(int rank) { rank_ = rank; }
  void Write(LogLevel level, const char* fmt, va_list args);
  void Debug(const char* fmt, ...);
  void Info(const char* fmt, ...);
  void Error(const char* fmt, ...);
  void Fatal(const char* fmt, ...);

 private:
  std::mutex mu_;
  FILE* file_ = nullptr;
  LogLevel level_ = LogLevel::Info;
  bool kill_fatal_ = true;
  int rank_ = -1;
};

// Process-global facade with the reference's static interface.
class Log {
 public:
  static Logger& Get();
  static int ResetLogFile(const std::string& f) { return Get().ResetLogFile(f); }
  static void ResetLogLevel(LogLevel l) { Get().ResetLogLevel(l); }
  static void ResetKillFatal(bool k);
  static void Debug(const char* fmt, ...);
  static void Info(const char* fmt, ...);
  static void Error(const char* fmt, ...);
  static void Fatal(const char* fmt, ...);
};

#define CHECK(cond)                                                                    \
  do {                                                                                 \
    if (!(cond))                                                                       \
      ::multiverso::Log::Fatal("Check failed: %s at %s:%d\n", #cond, __FILE__, __LINE__); \
  } while (0)
#define CHECK_NOTNULL(ptr)                                                             \
  do {                                                                                 \
    if ((ptr) == nullptr)                                                              \
      ::multiverso::Log::Fatal("%s must not be NULL at %s:%d\n", #ptr, __FILE__, __LINE__); \
  } while (0)

}  // namespace multiverso
#endif
