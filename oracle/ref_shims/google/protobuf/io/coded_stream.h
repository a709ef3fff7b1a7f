// google/protobuf/io/coded_stream.h — SHIM (test infrastructure; not protobuf): solve.cc:428-430
#ifndef LFR_SHIM_PB_CODED_STREAM_H_
#define LFR_SHIM_PB_CODED_STREAM_H_
#include "google/protobuf/io/zero_copy_stream.h"
namespace google {
namespace protobuf {
namespace io {
class CodedInputStream {
 public:
  explicit CodedInputStream(ZeroCopyInputStream* input) : input_(input), limit_(64 << 20) {}
  void SetTotalBytesLimit(int total_bytes_limit, int /*warning_threshold*/) { limit_ = total_bytes_limit; }
  ZeroCopyInputStream* input() const { return input_; }
  long long limit() const { return limit_; }

 private:
  ZeroCopyInputStream* input_;
  long long limit_;
};
}  // namespace io
}  // namespace protobuf
}  // namespace google
#endif
