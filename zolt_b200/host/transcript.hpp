// Host Fiat-Shamir transcript: Zolt's Keccak `Transcript(F)` (reference src/transcripts/mod.zig:49-221).
// The north star keeps the transcript on the host; this mirror exists so the C++ host side and the
// benchmark can run the real round loop (absorb evaluations -> squeeze challenge -> bind on the GPU).
//
// Behaviour reproduced exactly (it is NOT standard SHA-3): bytes are XORed into state[position];
// the permutation runs when position reaches the 136-byte rate; no padding; `position` is not reset by
// a squeeze; appendScalar absorbs the label and then the raw Montgomery limbs (little-endian);
// challengeScalar absorbs the label, permutes, and maps state[0..32) through Fr::fromBytes.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string_view>

#include "fr_host.hpp"

namespace zolt::host {

class Transcript {
 public:
  static constexpr size_t kRate = 136;

  explicit Transcript(std::string_view domain) { appendBytes(domain.data(), domain.size()); }

  void appendBytes(const void* data, size_t n) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    for (size_t i = 0; i < n; i++) {
      lanes_[position_ >> 3] ^= (uint64_t)p[i] << (8 * (position_ & 7));
      if (++position_ >= kRate) {
        permute();
        position_ = 0;
      }
    }
  }
  void appendMessage(std::string_view label, const void* data, size_t n) {
    appendBytes(label.data(), label.size());
    appendBytes(data, n);
  }
  void appendScalar(std::string_view label, const Fr& s) {
    appendBytes(label.data(), label.size());
    appendBytes(s.limbs.data(), 32);  // little-endian host: limb bytes are already LE
  }
  Fr challengeScalar(std::string_view label) {
    appendBytes(label.data(), label.size());
    permute();
    uint8_t buf[32];
    std::memcpy(buf, lanes_, 32);
    return Fr::fromBytes(buf);
  }
  size_t position() const { return position_; }
  const uint64_t* state() const { return lanes_; }

 private:
  uint64_t lanes_[25] = {0};
  size_t position_ = 0;

  static uint64_t rol(uint64_t x, unsigned n) { return n ? (x << n) | (x >> (64 - n)) : x; }

  // Keccak-f[1600], lane formulation (theta / rho+pi / chi / iota)
  void permute() {
    static constexpr uint64_t kRC[24] = {
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
        0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
        0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
        0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
        0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    // rotation offsets r[x][y] indexed by lane x + 5y
    static constexpr unsigned kRot[25] = {0,  1,  62, 28, 27, 36, 44, 6,  55, 20, 3,  10, 43,
                                          25, 39, 41, 45, 15, 21, 8,  18, 2,  61, 56, 14};
    uint64_t* a = lanes_;
    for (int round = 0; round < 24; round++) {
      uint64_t c[5], d[5], b[25];
      for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
      for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
      for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
      for (int x = 0; x < 5; x++)
        for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(a[x + 5 * y], kRot[x + 5 * y]);
      for (int y = 0; y < 5; y++)
        for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
      a[0] ^= kRC[round];
    }
  }
};

}  // namespace zolt::host
