/* Stand-in for libhdfs used by tests/test_host_runtime.py::test_hdfs_stream: implements the
 * seven entry points HDFSStream binds (hdfs.h signatures) on top of a local directory
 * ($FAKE_HDFS_ROOT), so the dlopen path of csrc/host/io/hdfs_stream.cpp is exercised end to
 * end without Hadoop. Reads are deliberately short (<= 64 KiB) like a real DFS client's. */
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { char root[512]; char namenode[128]; int port; } FakeFs;

void* hdfsConnect(const char* namenode, uint16_t port) {
  const char* root = getenv("FAKE_HDFS_ROOT");
  if (!root) return NULL;
  FakeFs* fs = calloc(1, sizeof *fs);
  snprintf(fs->root, sizeof fs->root, "%s", root);
  snprintf(fs->namenode, sizeof fs->namenode, "%s", namenode ? namenode : "");
  fs->port = port;
  char marker[640];
  snprintf(marker, sizeof marker, "%s/.connected", root);
  FILE* m = fopen(marker, "w");
  if (m) { fprintf(m, "%s %d\n", fs->namenode, fs->port); fclose(m); }
  return fs;
}
int hdfsDisconnect(void* fs) { free(fs); return 0; }
void* hdfsOpenFile(void* fs, const char* path, int flags, int buffer_size, short replication, int32_t block_size) {
  (void)buffer_size; (void)replication; (void)block_size;
  char full[1024];
  snprintf(full, sizeof full, "%s%s", ((FakeFs*)fs)->root, path);
  const char* mode = (flags & O_APPEND) ? "ab" : ((flags & O_WRONLY) ? "wb" : "rb");
  return fopen(full, mode);
}
int hdfsCloseFile(void* fs, void* file) { (void)fs; return fclose((FILE*)file); }
int32_t hdfsRead(void* fs, void* file, void* buffer, int32_t length) {
  (void)fs;
  if (length > 65536) length = 65536;
  return (int32_t)fread(buffer, 1, (size_t)length, (FILE*)file);
}
int32_t hdfsWrite(void* fs, void* file, const void* buffer, int32_t length) {
  (void)fs;
  return (int32_t)fwrite(buffer, 1, (size_t)length, (FILE*)file);
}
int hdfsFlush(void* fs, void* file) { (void)fs; return fflush((FILE*)file); }
