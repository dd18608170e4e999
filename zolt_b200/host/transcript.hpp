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
    size_t i = 0;
    while (i < n) {
      if ((position_ & 7) == 0 && n - i >= 8) {  // a whole lane at once (the state is little-endian lanes, like the host)
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        lanes_[position_ >> 3] ^= w;
        position_ += 8;
        i += 8;
      } else {
        lanes_[position_ >> 3] ^= (uint64_t)p[i] << (8 * (position_ & 7));
        position_++;
        i++;
      }
      if (position_ >= kRate) {
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

  static uint64_t rol(uint64_t x, unsigned n) { return (x << n) | (x >> (64 - n)); }  // 0 < n < 64

  // Keccak-f[1600], lane formulation (theta / rho+pi / chi / iota)
  void permute() {
    static constexpr uint64_t kRC[24] = {
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
        0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
        0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
        0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
        0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    // One round = theta, rho + pi, chi with every lane index and rotation offset written out (the rolled form with
    // `% 5` index arithmetic cost 3 us per permutation; a round loop over straight-line code is ~6x faster, and the
    // permutation sits on the critical path of every sumcheck round, between two kernel launches).
    uint64_t* a = lanes_;
    for (int round = 0; round < 24; round++) {
      const uint64_t c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20], c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21], c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22],
                     c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23], c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
      const uint64_t d0 = c4 ^ rol(c1, 1), d1 = c0 ^ rol(c2, 1), d2 = c1 ^ rol(c3, 1), d3 = c2 ^ rol(c4, 1), d4 = c3 ^ rol(c0, 1);
      const uint64_t b0 = a[0] ^ d0;
      const uint64_t b1 = rol(a[6] ^ d1, 44);
      const uint64_t b2 = rol(a[12] ^ d2, 43);
      const uint64_t b3 = rol(a[18] ^ d3, 21);
      const uint64_t b4 = rol(a[24] ^ d4, 14);
      const uint64_t b5 = rol(a[3] ^ d3, 28);
      const uint64_t b6 = rol(a[9] ^ d4, 20);
      const uint64_t b7 = rol(a[10] ^ d0, 3);
      const uint64_t b8 = rol(a[16] ^ d1, 45);
      const uint64_t b9 = rol(a[22] ^ d2, 61);
      const uint64_t b10 = rol(a[1] ^ d1, 1);
      const uint64_t b11 = rol(a[7] ^ d2, 6);
      const uint64_t b12 = rol(a[13] ^ d3, 25);
      const uint64_t b13 = rol(a[19] ^ d4, 8);
      const uint64_t b14 = rol(a[20] ^ d0, 18);
      const uint64_t b15 = rol(a[4] ^ d4, 27);
      const uint64_t b16 = rol(a[5] ^ d0, 36);
      const uint64_t b17 = rol(a[11] ^ d1, 10);
      const uint64_t b18 = rol(a[17] ^ d2, 15);
      const uint64_t b19 = rol(a[23] ^ d3, 56);
      const uint64_t b20 = rol(a[2] ^ d2, 62);
      const uint64_t b21 = rol(a[8] ^ d3, 55);
      const uint64_t b22 = rol(a[14] ^ d4, 39);
      const uint64_t b23 = rol(a[15] ^ d0, 41);
      const uint64_t b24 = rol(a[21] ^ d1, 2);
      a[0] = b0 ^ (~b1 & b2);
      a[1] = b1 ^ (~b2 & b3);
      a[2] = b2 ^ (~b3 & b4);
      a[3] = b3 ^ (~b4 & b0);
      a[4] = b4 ^ (~b0 & b1);
      a[5] = b5 ^ (~b6 & b7);
      a[6] = b6 ^ (~b7 & b8);
      a[7] = b7 ^ (~b8 & b9);
      a[8] = b8 ^ (~b9 & b5);
      a[9] = b9 ^ (~b5 & b6);
      a[10] = b10 ^ (~b11 & b12);
      a[11] = b11 ^ (~b12 & b13);
      a[12] = b12 ^ (~b13 & b14);
      a[13] = b13 ^ (~b14 & b10);
      a[14] = b14 ^ (~b10 & b11);
      a[15] = b15 ^ (~b16 & b17);
      a[16] = b16 ^ (~b17 & b18);
      a[17] = b17 ^ (~b18 & b19);
      a[18] = b18 ^ (~b19 & b15);
      a[19] = b19 ^ (~b15 & b16);
      a[20] = b20 ^ (~b21 & b22);
      a[21] = b21 ^ (~b22 & b23);
      a[22] = b22 ^ (~b23 & b24);
      a[23] = b23 ^ (~b24 & b20);
      a[24] = b24 ^ (~b20 & b21);
      a[0] ^= kRC[round];
    }
  }
};

}  // namespace zolt::host
