// Host-side, NumPy-compatible random streams (no device code).
//
// The reference draws its minibatch permutations with numpy: np.random.default_rng(seed) (ppo.py:72) and
// Generator.shuffle (ppo.py:276); SAC draws replay indices with Generator.integers (sac/pytorch/replay_buffer.py:33-34).
// NumPy is a third-party dependency of the reference (numpy>=2.2.6, pyproject.toml:24); its published algorithms are
// restated here so that index streams are bit-exact:
//   SeedSequence(seed).generate_state(4, uint64)  -> PCG64 seeding   (numpy/random/bit_generator.pyx, O'Neill's seed_seq_fe)
//   PCG64: 128-bit LCG (PCG_DEFAULT_MULTIPLIER_128) + XSL-RR 64-bit output, state advanced BEFORE output
//   next_uint32: low half of a 64-bit draw first, high half buffered across calls
//   shuffle: Fisher-Yates from the top with masked-rejection bounded draws (random_interval)
//   integers: Lemire multiply-shift with rejection on the 32-bit stream (buffered_bounded_lemire_uint32)
// Pinned against numpy 2.3.5 in tests/test_host_logic.py (test_pcg64_*).
#include <stdint.h>

#include <vector>

#include "../../include/rlx_b200.h"

namespace rlx {
void set_error(const char* fmt, ...);
}
#define RLX_CHECK_ARG(cond, msg)                                  \
  do {                                                            \
    if (!(cond)) {                                                \
      rlx::set_error("%s: invalid argument: %s", __func__, msg);  \
      return RLX_ERR_INVALID_ARG;                                 \
    }                                                             \
  } while (0)

namespace {

typedef unsigned __int128 u128;

const u128 kMult = ((u128)2549297995355413924ULL << 64) | (u128)4865540595714422341ULL;

inline u128 get_state(const rlx_pcg64* st) { return ((u128)st->s[0] << 64) | st->s[1]; }
inline u128 get_inc(const rlx_pcg64* st) { return ((u128)st->s[2] << 64) | st->s[3]; }
inline void put_state(rlx_pcg64* st, u128 v) {
  st->s[0] = (uint64_t)(v >> 64);
  st->s[1] = (uint64_t)v;
}

inline uint64_t next64(rlx_pcg64* st) {
  const u128 s = get_state(st) * kMult + get_inc(st);
  put_state(st, s);
  const uint64_t hi = (uint64_t)(s >> 64), lo = (uint64_t)s;
  const uint64_t x = hi ^ lo;
  const unsigned rot = (unsigned)(s >> 122);
  return (x >> rot) | (x << ((64 - rot) & 63));
}

inline uint32_t next32(rlx_pcg64* st) {
  if (st->s[4]) {
    st->s[4] = 0;
    return (uint32_t)st->s[5];
  }
  const uint64_t v = next64(st);
  st->s[4] = 1;
  st->s[5] = v >> 32;
  return (uint32_t)v;
}

// ---- SeedSequence (pool of 4 uint32 words)
const uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u, INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu;
const uint32_t MIX_MULT_L = 0xca01f9ddu, MIX_MULT_R = 0x4973f715u;

inline uint32_t hashmix(uint32_t value, uint32_t& hash_const) {
  value ^= hash_const;
  hash_const *= MULT_A;
  value *= hash_const;
  value ^= value >> 16;
  return value;
}
inline uint32_t mix(uint32_t x, uint32_t y) {
  uint32_t r = MIX_MULT_L * x - MIX_MULT_R * y;
  r ^= r >> 16;
  return r;
}

}  // namespace

extern "C" int rlx_pcg64_seed(uint64_t seed, rlx_pcg64* st) {
  RLX_CHECK_ARG(st != nullptr, "state is null");
  // entropy -> little-endian uint32 words (0 -> [0])
  uint32_t ent[2];
  int n_ent = 1;
  ent[0] = (uint32_t)seed;
  ent[1] = (uint32_t)(seed >> 32);
  if (ent[1] != 0) n_ent = 2;
  uint32_t pool[4];
  uint32_t hc = INIT_A;
  for (int i = 0; i < 4; ++i) pool[i] = hashmix(i < n_ent ? ent[i] : 0u, hc);
  for (int i_src = 0; i_src < 4; ++i_src)
    for (int i_dst = 0; i_dst < 4; ++i_dst)
      if (i_src != i_dst) pool[i_dst] = mix(pool[i_dst], hashmix(pool[i_src], hc));
  // generate_state(4, uint64) == 8 uint32 words, pairs little-endian
  uint32_t w[8];
  uint32_t hb = INIT_B;
  for (int i = 0; i < 8; ++i) {
    uint32_t v = pool[i & 3];
    v ^= hb;
    hb *= MULT_B;
    v *= hb;
    v ^= v >> 16;
    w[i] = v;
  }
  uint64_t val[4];
  for (int i = 0; i < 4; ++i) val[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
  // pcg64_set_seed: initstate = (val0 << 64 | val1), initseq = (val2 << 64 | val3)
  const u128 initstate = ((u128)val[0] << 64) | val[1];
  const u128 initseq = ((u128)val[2] << 64) | val[3];
  const u128 inc = (initseq << 1) | 1;
  st->s[2] = (uint64_t)(inc >> 64);
  st->s[3] = (uint64_t)inc;
  u128 s = 0;
  s = s * kMult + inc;
  s += initstate;
  s = s * kMult + inc;
  put_state(st, s);
  st->s[4] = 0;
  st->s[5] = 0;
  return RLX_OK;
}

extern "C" uint64_t rlx_pcg64_next64(rlx_pcg64* st) { return next64(st); }
extern "C" uint32_t rlx_pcg64_next32(rlx_pcg64* st) { return next32(st); }

extern "C" int rlx_pcg64_shuffle_i64(rlx_pcg64* st, int64_t* a, int64_t n) {
  RLX_CHECK_ARG(st != nullptr && (a != nullptr || n == 0) && n >= 0, "bad arguments");
  rlx_pcg64 loc = *st;  // keep the hot state in registers
  for (int64_t i = n - 1; i >= 1; --i) {
    uint64_t mask = (uint64_t)i;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    uint64_t j;
    if ((uint64_t)i <= 0xffffffffULL) {
      while ((j = (next32(&loc) & mask)) > (uint64_t)i) {}
    } else {
      while ((j = (next64(&loc) & mask)) > (uint64_t)i) {}
    }
    const int64_t t = a[i];
    a[i] = a[j];
    a[j] = t;
  }
  *st = loc;
  return RLX_OK;
}

extern "C" int rlx_pcg64_integers_i64(rlx_pcg64* st, int64_t high, int64_t* out, int64_t n) {
  RLX_CHECK_ARG(st != nullptr && (out != nullptr || n == 0) && n >= 0, "bad arguments");
  RLX_CHECK_ARG(high >= 1, "high must be >= 1");
  const uint64_t rng = (uint64_t)high - 1;
  if (rng == 0) {
    for (int64_t i = 0; i < n; ++i) out[i] = 0;
    return RLX_OK;
  }
  rlx_pcg64 loc = *st;
  if (rng == 0xFFFFFFFFULL) {
    for (int64_t i = 0; i < n; ++i) out[i] = (int64_t)next32(&loc);
  } else if (rng < 0xFFFFFFFFULL) {
    const uint32_t rng_excl = (uint32_t)rng + 1u;
    for (int64_t i = 0; i < n; ++i) {
      uint64_t m = (uint64_t)next32(&loc) * rng_excl;
      uint32_t leftover = (uint32_t)m;
      if (leftover < rng_excl) {
        const uint32_t threshold = (0xFFFFFFFFu - (uint32_t)rng) % rng_excl;
        while (leftover < threshold) {
          m = (uint64_t)next32(&loc) * rng_excl;
          leftover = (uint32_t)m;
        }
      }
      out[i] = (int64_t)(m >> 32);
    }
  } else {
    // 64-bit Lemire (random_bounded_uint64_fill, rng > 2^32 - 1)
    const uint64_t rng_excl = rng + 1;
    for (int64_t i = 0; i < n; ++i) {
      u128 m = (u128)next64(&loc) * rng_excl;
      uint64_t leftover = (uint64_t)m;
      if (leftover < rng_excl) {
        const uint64_t threshold = (0xFFFFFFFFFFFFFFFFULL - rng) % rng_excl;
        while (leftover < threshold) {
          m = (u128)next64(&loc) * rng_excl;
          leftover = (uint64_t)m;
        }
      }
      out[i] = (int64_t)(m >> 64);
    }
  }
  *st = loc;
  return RLX_OK;
}

namespace {
// random_bounded_uint64(bitgen, 0, rng, 0, use_masked=false) of numpy/random/src/distributions/distributions.c: a value in [0, rng]
inline uint64_t bounded_lemire(rlx_pcg64* loc, uint64_t rng) {
  if (rng == 0) return 0;
  if (rng == 0xFFFFFFFFULL) return next32(loc);
  if (rng < 0xFFFFFFFFULL) {
    const uint32_t rng_excl = (uint32_t)rng + 1u;
    uint64_t m = (uint64_t)next32(loc) * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
      const uint32_t threshold = (0xFFFFFFFFu - (uint32_t)rng) % rng_excl;
      while (leftover < threshold) {
        m = (uint64_t)next32(loc) * rng_excl;
        leftover = (uint32_t)m;
      }
    }
    return m >> 32;
  }
  if (rng == 0xFFFFFFFFFFFFFFFFULL) return next64(loc);
  const uint64_t rng_excl = rng + 1;
  u128 m = (u128)next64(loc) * rng_excl;
  uint64_t leftover = (uint64_t)m;
  if (leftover < rng_excl) {
    const uint64_t threshold = (0xFFFFFFFFFFFFFFFFULL - rng) % rng_excl;
    while (leftover < threshold) {
      m = (u128)next64(loc) * rng_excl;
      leftover = (uint64_t)m;
    }
  }
  return (uint64_t)(m >> 64);
}
// _shuffle_int of numpy/random/_generator.pyx: Fisher-Yates on positions [first, n)
inline void shuffle_int(rlx_pcg64* loc, int64_t n, int64_t first, int64_t* data) {
  for (int64_t i = n - 1; i >= first; --i) {
    const int64_t j = (int64_t)bounded_lemire(loc, (uint64_t)i);
    const int64_t t = data[i];
    data[i] = data[j];
    data[j] = t;
  }
}
}  // namespace

// Generator.choice(pop_size, size=size, replace=False)  (numpy/random/_generator.pyx, the `p is None`, shuffle=True branch):
// a tail shuffle of arange(pop_size) when the sample is a large part of a large population, Floyd's algorithm with an open-addressing
// hash set (followed by a shuffle of the sample) otherwise.  ref use: espo.py:256.
extern "C" int rlx_pcg64_choice_i64(rlx_pcg64* st, int64_t pop_size, int64_t size, int64_t* out) {
  RLX_CHECK_ARG(st != nullptr && pop_size >= 0 && size >= 0 && (out != nullptr || size == 0), "bad arguments");
  RLX_CHECK_ARG(size <= pop_size, "cannot take a larger sample than population when replace is False");
  if (size == 0) return RLX_OK;
  rlx_pcg64 loc = *st;
  const int64_t cutoff = 50;
  if (pop_size > 10000 && size > pop_size / cutoff) {
    std::vector<int64_t> idx((size_t)pop_size);
    for (int64_t i = 0; i < pop_size; ++i) idx[(size_t)i] = i;
    const int64_t first = (pop_size - size > 1) ? pop_size - size : 1;
    shuffle_int(&loc, pop_size, first, idx.data());
    for (int64_t i = 0; i < size; ++i) out[i] = idx[(size_t)(pop_size - size + i)];
  } else {
    uint64_t mask = (uint64_t)(1.2 * (double)size);
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    std::vector<uint64_t> hash((size_t)mask + 1, ~0ULL);
    for (int64_t j = pop_size - size; j < pop_size; ++j) {
      const uint64_t val = bounded_lemire(&loc, (uint64_t)j);
      uint64_t slot = val & mask;
      while (hash[slot] != ~0ULL && hash[slot] != val) slot = (slot + 1) & mask;
      if (hash[slot] == ~0ULL) {  // val not drawn yet
        hash[slot] = val;
        out[j - pop_size + size] = (int64_t)val;
      } else {                    // already in the sample: take j itself
        slot = (uint64_t)j & mask;
        while (hash[slot] != ~0ULL) slot = (slot + 1) & mask;
        hash[slot] = (uint64_t)j;
        out[j - pop_size + size] = j;
      }
    }
    shuffle_int(&loc, size, 1, out);
  }
  *st = loc;
  return RLX_OK;
}
