// Persistent host worker pool + the 2-bit (pk2) packer of the .bk streamer (pack.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>

namespace tpg {

class HostPool {
 public:
  static HostPool& instance();
  // Splits [0, n) into at most max_threads() contiguous parts of >= grain items and runs
  // body(begin, end) on the pool (the caller takes a share); returns when all parts are done.
  void run(int64_t n, int64_t grain, const std::function<void(int64_t, int64_t)>& body);
  // The same split, but on the workers only: begin() returns at once (the caller goes on to
  // enqueue GPU work), end() waits and returns the milliseconds from begin() to the last part.
  // One job at a time, like run(); every begin() must be matched by an end().
  void begin(int64_t n, int64_t grain, std::function<void(int64_t, int64_t)> body);
  double end();
  static int max_threads();  // min(hardware threads, 32), or TPG_HOST_THREADS

 private:
  HostPool();
  struct Impl;
  Impl* impl_;
};

// memcpy with non-temporal stores (large result pieces: written once, no read-for-ownership).
void copy_streaming(void* dst, const void* src, size_t n);

// pk2 layout: bytes per packed column for a staged row window of `span` individuals.
int64_t pk2_stride(int64_t span);
// Packs one column (span raw bytes) into stride bytes at dst (code 3 = missing / padding).
void pk2_pack_column(const uint8_t* src, int64_t span, uint8_t* dst, int64_t stride);

}  // namespace tpg
