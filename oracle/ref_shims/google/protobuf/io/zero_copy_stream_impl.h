// google/protobuf/io/zero_copy_stream_impl.h — SHIM (test infrastructure; not protobuf):
// FileInputStream over a file descriptor (solve.cc:426-427).
#ifndef LFR_SHIM_PB_ZERO_COPY_STREAM_IMPL_H_
#define LFR_SHIM_PB_ZERO_COPY_STREAM_IMPL_H_
#include <unistd.h>

#include <vector>

#include "google/protobuf/io/zero_copy_stream.h"
namespace google {
namespace protobuf {
namespace io {
class FileInputStream : public ZeroCopyInputStream {
 public:
  explicit FileInputStream(int fd) : fd_(fd), loaded_(false) {}
  virtual bool ReadAll(const unsigned char** data, size_t* size) {
    if (!loaded_) {
      loaded_ = true;
      if (fd_ < 0) return false;
      unsigned char chunk[1 << 16];
      for (;;) {
        const ssize_t n = ::read(fd_, chunk, sizeof chunk);
        if (n < 0) return false;
        if (n == 0) break;
        buf_.insert(buf_.end(), chunk, chunk + n);
      }
    }
    *data = buf_.data();
    *size = buf_.size();
    return fd_ >= 0;
  }

 private:
  int fd_;
  bool loaded_;
  std::vector<unsigned char> buf_;
};
}  // namespace io
}  // namespace protobuf
}  // namespace google
#endif
