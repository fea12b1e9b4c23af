/*
 * mm_hash.h -- bit-exact k-mer hashing shared by the CUDA kernels (and, compiled with a host
 * compiler, by the CPU unit tests of this header).
 *
 * Restates, for a fixed k-mer length K held in 64-bit register words:
 *   MurmurHash3_x64_128(key, K, seed=42), low 64 bits   reference src/common/murmur3.h:236-303,
 *                                                        called from commonFunc.hpp:138-147 (getHash)
 *   upper-casing + "anything not ACGT is N"             commonFunc.hpp:75-107
 *   per-base complement                                 commonFunc.hpp:50-73
 * The reference hashes the ASCII bytes of the forward k-mer and of its reverse complement
 * separately (commonFunc.hpp:225-237); there is no rolling hash, so both full evaluations are done
 * here too. The two k-mers are kept as sliding byte windows in registers: the forward window shifts
 * right by one byte per base, the reverse-complement window shifts left.
 */
#ifndef MM_HASH_H
#define MM_HASH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define MM_HD __host__ __device__ __forceinline__
#else
#define MM_HD inline
#endif

#define MM_SEED 42ULL /* commonFunc.hpp:37 */

MM_HD uint64_t mm_rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

/* Building blocks of the hash. On the device they are spelled in 32-bit halves: a 64x64->64 multiply by a constant is
 * one mul.wide + two mad.lo (nvcc's own expansion of `x * c` spends two extra adds per multiply), a rotate is two
 * funnel shifts, `h*5 + c` is shift-and-add. The per-base loop of the sketch kernel is issue-bound, so the
 * instruction count is the throughput. Same results as the plain C expressions (tests/test_host_cpu.py checks the
 * host spelling against the reference's getHash, the GPU sketch tests check the device spelling). */
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint64_t mm_pack64(uint32_t lo, uint32_t hi)
{
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void mm_unpack64(uint64_t x, uint32_t &lo, uint32_t &hi)
{
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
}
/* x * C mod 2^64; NBYTES = number of low bytes of x that can be non-zero */
template <uint64_t C, int NBYTES>
__device__ __forceinline__ uint64_t mm_mulc(uint64_t x)
{
  uint32_t xl, xh, pl, ph;
  uint64_t p;
  mm_unpack64(x, xl, xh);
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(xl), "n"((uint32_t)(C & 0xFFFFFFFFULL)));
  mm_unpack64(p, pl, ph);
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(ph) : "r"(xl), "n"((uint32_t)(C >> 32)));
  if (NBYTES > 4) asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(ph) : "r"(xh), "n"((uint32_t)(C & 0xFFFFFFFFULL)));
  return mm_pack64(pl, ph);
}
template <int R>
__device__ __forceinline__ uint64_t mm_rotl(uint64_t x)
{
  uint32_t lo, hi;
  mm_unpack64(x, lo, hi);
  if (R < 32) return mm_pack64(__funnelshift_l(hi, lo, R), __funnelshift_l(lo, hi, R));
  return mm_pack64(__funnelshift_l(lo, hi, R - 32), __funnelshift_l(hi, lo, R - 32));
}
/* x * 5 + A (A < 2^32) as (x << 2) + x + A: two LEA + two adds on the ALU pipe. The multiply form costs an IMAD.WIDE, and
 * the 32x32->64 multiply is the one instruction of this loop that does not overlap with ALU work (measured:
 * mashmap_b200/csrc/bench/mm_issue_peak.cu), so every one that can be avoided is worth more than its count. */
template <uint32_t A>
__device__ __forceinline__ uint64_t mm_mul5_add(uint64_t x)
{
  return (x << 2) + x + (uint64_t)A;
}
__device__ __forceinline__ uint64_t mm_xorshr33(uint64_t k)
{ /* k ^ (k >> 33): only the low half changes */
  uint32_t lo, hi;
  mm_unpack64(k, lo, hi);
  return mm_pack64(lo ^ (hi >> 1), hi);
}
#else
MM_HD uint64_t mm_pack64(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }
MM_HD void mm_unpack64(uint64_t x, uint32_t &lo, uint32_t &hi) { lo = (uint32_t)x; hi = (uint32_t)(x >> 32); }
template <uint64_t C, int NBYTES>
MM_HD uint64_t mm_mulc(uint64_t x) { return x * C; }
template <int R>
MM_HD uint64_t mm_rotl(uint64_t x) { return mm_rotl64(x, R); }
template <uint32_t A>
MM_HD uint64_t mm_mul5_add(uint64_t x) { return x * 5 + A; }
MM_HD uint64_t mm_xorshr33(uint64_t k) { return k ^ (k >> 33); }
#endif

MM_HD uint64_t mm_fmix64(uint64_t k)
{ /* murmur3.h fmix64 */
  k = mm_xorshr33(k);
  k = mm_mulc<0xff51afd7ed558ccdULL, 8>(k);
  k = mm_xorshr33(k);
  k = mm_mulc<0xc4ceb9fe1a85ec53ULL, 8>(k);
  k = mm_xorshr33(k);
  return k;
}

template <int K>
struct mm_kmer_words {
  static constexpr int NW = (K + 7) / 8;                 /* 64-bit words holding K bytes        */
  static constexpr int TOP_BYTES = K - 8 * (NW - 1);     /* bytes used in the top word (1..8)   */
  static constexpr uint64_t TOP_MASK =
      TOP_BYTES == 8 ? ~0ULL : ((1ULL << (8 * (TOP_BYTES & 7))) - 1ULL);
};

/* low 64 bits of MurmurHash3_x64_128 over the K bytes packed little-endian in w[] (unused high
 * bytes of the top word must be zero), seed 42. murmur3.h:236-303 */
template <int K>
MM_HD uint64_t mm_murmur3_k(const uint64_t *w)
{
  constexpr int NB = K / 16;
  constexpr int TL = K & 15;
  constexpr uint64_t c1 = 0x87c37b91114253d5ULL;
  constexpr uint64_t c2 = 0x4cf5ad432745937fULL;
  uint64_t h1 = MM_SEED, h2 = MM_SEED;
#pragma unroll
  for (int b = 0; b < NB; b++) {
    uint64_t k1 = w[2 * b], k2 = w[2 * b + 1];
    k1 = mm_mulc<c1, 8>(k1); k1 = mm_rotl<31>(k1); k1 = mm_mulc<c2, 8>(k1); h1 ^= k1;
    if (b == 0) { /* h2 is still the seed: (x + seed) * 5 + c = x * 5 + (c + 5 * seed), two instructions fewer */
      h1 = mm_rotl<27>(h1); h1 = mm_mul5_add<0x52dce729u + 5u * (uint32_t)MM_SEED>(h1);
    } else {
      h1 = mm_rotl<27>(h1); h1 += h2; h1 = mm_mul5_add<0x52dce729u>(h1);
    }
    k2 = mm_mulc<c2, 8>(k2); k2 = mm_rotl<33>(k2); k2 = mm_mulc<c1, 8>(k2); h2 ^= k2;
    h2 = mm_rotl<31>(h2); h2 += h1; h2 = mm_mul5_add<0x38495ab5u>(h2);
  }
  if (TL > 8) {
    uint64_t k2 = w[2 * NB + 1];
    k2 = mm_mulc<c2, (TL > 8 ? TL - 8 : 8)>(k2); k2 = mm_rotl<33>(k2); k2 = mm_mulc<c1, 8>(k2); h2 ^= k2;
  }
  if (TL > 0) {
    uint64_t k1 = w[2 * NB];
    k1 = mm_mulc<c1, (TL > 8 ? 8 : (TL > 0 ? TL : 8))>(k1); k1 = mm_rotl<31>(k1); k1 = mm_mulc<c2, 8>(k1); h1 ^= k1;
  }
  h1 ^= (uint64_t)K; h2 ^= (uint64_t)K;
  h1 += h2; h2 += h1;
  h1 = mm_fmix64(h1); h2 = mm_fmix64(h2);
  h1 += h2;
  return h1;
}

/* Base classification. Returns the 2-bit code (A=0 C=1 T=2 G=3: bits 1-2 of the upper-cased
 * letter) and sets is_n for anything that is not ACGT after upper-casing (commonFunc.hpp:97-107;
 * bytes >= 127 index outside the reference's table -- treated as N). */
MM_HD uint32_t mm_base_code(uint32_t byte, bool &is_n)
{
  uint32_t c = byte & 0xDFu; /* a-z -> A-Z; only 0x41/0x61 map to 'A', etc. */
  is_n = !(c == 0x41u || c == 0x43u || c == 0x47u || c == 0x54u);
  return (c >> 1) & 3u;
}
/* ASCII of a base code and of its complement: A<->T (0<->2), C<->G (1<->3). */
MM_HD uint32_t mm_code_ascii(uint32_t code) { return (0x47544341u >> (8 * code)) & 0xFFu; }
MM_HD uint32_t mm_code_comp_ascii(uint32_t code) { return (0x47544341u >> (8 * (code ^ 2u))) & 0xFFu; }

/* Forward / reverse-complement k-mer windows. After K pushes, f[] holds seq[i..i+K) and r[] holds
 * the reverse complement of it, both packed little-endian exactly like the byte strings the
 * reference passes to getHash (commonFunc.hpp:225,229 / :357,363). */
template <int K>
struct mm_kmer_window {
  static constexpr int NW = mm_kmer_words<K>::NW;
  uint64_t f[NW];
  uint64_t r[NW];

  MM_HD void reset()
  {
#pragma unroll
    for (int i = 0; i < NW; i++) { f[i] = 0; r[i] = 0; }
  }
  MM_HD void push(uint32_t code)
  {
    const uint64_t fa = mm_code_ascii(code);
    const uint64_t ra = mm_code_comp_ascii(code);
#pragma unroll
    for (int i = 0; i < NW - 1; i++) f[i] = (f[i] >> 8) | (f[i + 1] << 56);
    f[NW - 1] = (f[NW - 1] >> 8) | (fa << (8 * ((K - 1) & 7)));
#pragma unroll
    for (int i = NW - 1; i > 0; i--) r[i] = (r[i] << 8) | (r[i - 1] >> 56);
    r[0] = (r[0] << 8) | ra;
    r[NW - 1] &= mm_kmer_words<K>::TOP_MASK;
  }
  MM_HD uint64_t hash_fwd() const { return mm_murmur3_k<K>(f); }
  MM_HD uint64_t hash_rev() const { return mm_murmur3_k<K>(r); }
};

/* Interval point packed for sorting: (seqId, pos, side) ascending with CLOSE before OPEN, i.e.
 * IntervalPoint::operator< (base_types.hpp:75-78). */
MM_HD uint64_t mm_pack_point(int32_t seqId, int32_t pos, int open)
{
  return ((uint64_t)(uint32_t)seqId << 33) | ((uint64_t)(uint32_t)pos << 1) | (uint64_t)(open ? 1 : 0);
}
MM_HD int32_t mm_point_seq(uint64_t p) { return (int32_t)(p >> 33); }
MM_HD int32_t mm_point_pos(uint64_t p) { return (int32_t)((p >> 1) & 0xFFFFFFFFu); }
MM_HD int mm_point_open(uint64_t p) { return (int)(p & 1); }

#endif /* MM_HASH_H */
