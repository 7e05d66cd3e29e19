// Logger implementation (see include/multiverso/util/log.h).
#include "multiverso/util/log.h"
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <stdexcept>
#include "multiverso/util/configure.h"

namespace multiverso {

bool g_mv_kill_fatal = true;

MV_DEFINE_bool(logtostderr, false, "log to stderr instead of stdout");

Logger::~Logger() {
  if (file_) fclose(file_);
}

int Logger::ResetLogFile(const std::string& filename) {
  std::lock_guard<std::mutex> lk(mu_);
  if (file_) {
    fclose(file_);
    file_ = nullptr;
  }
  if (filename.empty()) return 0;
  file_ = fopen(filename.c_str(), "a");
  return file_ ? 0 : -1;
}

void Logger::Write(LogLevel level, const char* fmt, va_list args) {
  if (static_cast<int>(level) < static_cast<int>(level_)) return;
  static const char* kNames[] = {"DEBUG", "INFO", "ERROR", "FATAL"};
  char body[4096];
  vsnprintf(body, sizeof body, fmt, args);
  char stamp[32];
  time_t now = time(nullptr);
  struct tm tmv;
  localtime_r(&now, &tmv);
  strftime(stamp, sizeof stamp, "%Y-%m-%d %H:%M:%S", &tmv);
  char rank[24] = "";
  if (rank_ >= 0) snprintf(rank, sizeof rank, " [rank %d]", rank_);
  std::lock_guard<std::mutex> lk(mu_);
  FILE* out = MV_CONFIG(logtostderr) ? stderr : stdout;
  size_t len = strlen(body);
  const char* nl = (len && body[len - 1] == '\n') ? "" : "\n";
  fprintf(out, "[%s] [%s]%s %s%s", kNames[static_cast<int>(level)], stamp, rank, body, nl);
  fflush(out);
  if (file_) {
    fprintf(file_, "[%s] [%s]%s %s%s", kNames[static_cast<int>(level)], stamp, rank, body, nl);
    fflush(file_);
  }
}

#define MV_LOG_FORWARD(level)      \
  va_list args;                    \
  va_start(args, fmt);             \
  Write(level, fmt, args);         \
  va_end(args)

void Logger::Debug(const char* fmt, ...) { MV_LOG_FORWARD(LogLevel::Debug); }
void Logger::Info(const char* fmt, ...) { MV_LOG_FORWARD(LogLevel::Info); }
void Logger::Error(const char* fmt, ...) { MV_LOG_FORWARD(LogLevel::Error); }
void Logger::Fatal(const char* fmt, ...) {
  MV_LOG_FORWARD(LogLevel::Fatal);
  if (kill_fatal_) exit(1);
}

void Log::ResetKillFatal(bool k) {
  g_mv_kill_fatal = k;
  Get().ResetKillFatal(k);
}

Logger& Log::Get() {
  static Logger logger;
  return logger;
}

#define MV_LOG_STATIC(level)         \
  va_list args;                      \
  va_start(args, fmt);               \
  Get().Write(level, fmt, args);     \
  va_end(args)

void Log::Debug(const char* fmt, ...) { MV_LOG_STATIC(LogLevel::Debug); }
void Log::Info(const char* fmt, ...) { MV_LOG_STATIC(LogLevel::Info); }
void Log::Error(const char* fmt, ...) { MV_LOG_STATIC(LogLevel::Error); }
void Log::Fatal(const char* fmt, ...) {
  MV_LOG_STATIC(LogLevel::Fatal);
  // ResetKillFatal(false) (tests): throw instead of exiting
  if (g_mv_kill_fatal) exit(1);
  throw std::runtime_error("multiverso fatal error");
}

}  // namespace multiverso
