// numtheory.h -- host-side number theory for plan construction.
//
// Mirrors what the Haskell side of the reference computes before it calls into
// C (none of it exists in lol-cpp itself):
//   prime powers of m in increasing prime order   lol/Crypto/Lol/FactoredDefs.hs:92-94, 360-361
//   omega = (smallest generator of Z_q^*)^((q-1)/m)   lol/Crypto/Lol/Types/Unsafe/ZqBasic.hs:144-165
//   mhat^-1 mod q                                  ZqBasic.hs:167-171, FactoredDefs.hs:374-376
//   b^-1 mod a by extended Euclid                  lol-cpp/.../CPP/zq.cpp:20-54
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

namespace lolb {

typedef unsigned __int128 u128;

inline uint64_t mulmod64(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)((u128)a * b % q); }

inline uint64_t powmod64(uint64_t b, uint64_t e, uint64_t q)
{
  uint64_t r = 1 % q;
  b %= q;
  while (e) {
    if (e & 1) r = mulmod64(r, b, q);
    b = mulmod64(b, b, q);
    e >>= 1;
  }
  return r;
}

inline int64_t ipow64(int64_t b, int e)
{
  int64_t r = 1;
  while (e-- > 0) r *= b;
  return r;
}

// distinct prime factors with multiplicity, ascending
inline std::vector<std::pair<uint64_t, int>> factorize(uint64_t v)
{
  std::vector<std::pair<uint64_t, int>> out;
  for (uint64_t p = 2; p * p <= v; p += (p == 2 ? 1 : 2)) {
    if (v % p == 0) {
      int e = 0;
      while (v % p == 0) { v /= p; e++; }
      out.push_back({p, e});
    }
  }
  if (v > 1) out.push_back({v, 1});
  return out;
}

inline bool is_prime(uint64_t v)
{
  if (v < 2) return false;
  for (uint64_t p = 2; p * p <= v; p += (p == 2 ? 1 : 2))
    if (v % p == 0) return false;
  return true;
}

// inverse of b modulo a in [0,a); 0 when gcd(a,b) != 1
inline int64_t mod_inverse(int64_t a, int64_t b)
{
  int64_t r0 = a, r1 = ((b % a) + a) % a, t0 = 0, t1 = 1;
  while (r1 != 0) {
    int64_t qt = r0 / r1;
    int64_t rr = r0 - qt * r1; r0 = r1; r1 = rr;
    int64_t tt = t0 - qt * t1; t0 = t1; t1 = tt;
  }
  if (r0 != 1) return 0;
  return ((t0 % a) + a) % a;
}

// smallest x in [0,q) generating Z_q^* (q prime); 0 on failure
inline uint64_t smallest_generator(uint64_t q)
{
  if (!is_prime(q)) return 0;
  if (q == 2) return 1;
  const uint64_t order = q - 1;
  std::vector<uint64_t> exps;
  for (auto& f : factorize(order)) exps.push_back(order / f.first);
  for (uint64_t x = 2; x < q; x++) {
    bool gen = true;
    for (uint64_t e : exps)
      if (powmod64(x, e, q) == 1) { gen = false; break; }
    if (gen) return x;
  }
  return 0;
}

// principal m-th root of unity mod q as ZqBasic.hs derives it; 0 when Z_q has no CRT of index m
inline uint64_t principal_root(uint64_t m, uint64_t q)
{
  if (!is_prime(q) || (q - 1) % m != 0) return 0;
  uint64_t g = smallest_generator(q);
  if (g == 0) return 0;
  return powmod64(g, (q - 1) / m, q);
}

// base-p digit reversal on `digits` digits (crt.cpp:21-33)
inline int64_t digit_rev(int p, int digits, int64_t j)
{
  int64_t out = 0;
  for (int d = 0; d < digits; d++) { out = out * p + j % p; j /= p; }
  return out;
}

}  // namespace lolb
