// GIL-free bulk file IO for the KV-cache storage tier (vllm_b200/kv_offload): whole-buffer pread/pwrite with
// short-transfer loops, optional O_DIRECT, fsync-on-close, and a tiny thread pool so several blocks move at once.
// Role of the reference's csrc/fs_io.cpp.
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

extern "C" {

// Returns bytes written, or -errno.
int64_t vbn_write_file(const char* path, const void* buf, int64_t nbytes, int direct, int do_fsync) {
  int flags = O_WRONLY | O_CREAT | O_TRUNC;
#ifdef O_DIRECT
  if (direct && (reinterpret_cast<uintptr_t>(buf) % 4096 == 0) && (nbytes % 4096 == 0)) flags |= O_DIRECT;
#endif
  std::string tmp = std::string(path) + ".tmp";
  int fd = ::open(tmp.c_str(), flags, 0644);
  if (fd < 0) return -errno;
  int64_t done = 0;
  const char* p = static_cast<const char*>(buf);
  while (done < nbytes) {
    ssize_t n = ::pwrite(fd, p + done, (size_t)(nbytes - done), (off_t)done);
    if (n < 0) { if (errno == EINTR) continue; int e = errno; ::close(fd); ::unlink(tmp.c_str()); return -e; }
    done += n;
  }
  if (do_fsync) ::fsync(fd);
  ::close(fd);
  if (::rename(tmp.c_str(), path) != 0) { int e = errno; ::unlink(tmp.c_str()); return -e; }   // atomic publish
  return done;
}

// Returns bytes read (may be < nbytes at EOF), or -errno.
int64_t vbn_read_file(const char* path, void* buf, int64_t nbytes, int64_t offset, int direct) {
  int flags = O_RDONLY;
#ifdef O_DIRECT
  if (direct && (reinterpret_cast<uintptr_t>(buf) % 4096 == 0) && (nbytes % 4096 == 0) && (offset % 4096 == 0)) flags |= O_DIRECT;
#endif
  int fd = ::open(path, flags);
  if (fd < 0) return -errno;
  int64_t done = 0;
  char* p = static_cast<char*>(buf);
  while (done < nbytes) {
    ssize_t n = ::pread(fd, p + done, (size_t)(nbytes - done), (off_t)(offset + done));
    if (n < 0) { if (errno == EINTR) continue; int e = errno; ::close(fd); return -e; }
    if (n == 0) break;
    done += n;
  }
  ::close(fd);
  return done;
}

int64_t vbn_file_size(const char* path) {
  struct stat st;
  if (::stat(path, &st) != 0) return -errno;
  return (int64_t)st.st_size;
}

// Batched variants: n independent files, moved by up to `threads` threads.  results[i] = bytes or -errno.
void vbn_write_files(int n, const char* const* paths, const void* const* bufs, const int64_t* nbytes, int threads,
                     int direct, int64_t* results) {
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([=]() { for (int i = t; i < n; i += threads) results[i] = vbn_write_file(paths[i], bufs[i], nbytes[i], direct, 0); });
  for (auto& th : pool) th.join();
}

void vbn_read_files(int n, const char* const* paths, void* const* bufs, const int64_t* nbytes, int threads, int direct,
                    int64_t* results) {
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([=]() { for (int i = t; i < n; i += threads) results[i] = vbn_read_file(paths[i], bufs[i], nbytes[i], 0, direct); });
  for (auto& th : pool) th.join();
}

}  // extern "C"
