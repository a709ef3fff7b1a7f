// google/protobuf/io/zero_copy_stream.h — SHIM (test infrastructure; not protobuf): solve.cc:427
#ifndef LFR_SHIM_PB_ZERO_COPY_STREAM_H_
#define LFR_SHIM_PB_ZERO_COPY_STREAM_H_
#include <cstddef>
namespace google {
namespace protobuf {
namespace io {
class ZeroCopyInputStream {
 public:
  virtual ~ZeroCopyInputStream() {}
  // whole remaining content; the shim's only consumer is CodedInputStream
  virtual bool ReadAll(const unsigned char** data, size_t* size) = 0;
};
}  // namespace io
}  // namespace protobuf
}  // namespace google
#endif
