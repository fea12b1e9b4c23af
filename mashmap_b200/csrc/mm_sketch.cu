/*
 * mm_sketch.cu -- K1: bottom-s MinHash sketch of every query segment.
 *
 * Replaces CommonFunc::sketchSequence (reference src/map/include/commonFunc.hpp:182-288) as
 * called from Map::getSeedHits (computeMap.hpp:817-843), i.e. rows a1-a3 of SURVEY 8(a):
 *   normalise bases (commonFunc.hpp:97-107), hash every k-mer and its reverse complement with
 *   MurmurHash3_x64_128 (commonFunc.hpp:225-237), keep k-mers without N whose two hashes differ
 *   (:234), canonical = min (:237), strand vote = +1 if fwd < rev else -1 (:240), and output the
 *   s smallest DISTINCT canonical hashes with first position, last position and the sign of the
 *   vote sum (:242-286), ascending by hash.
 * sketchSequence is a pure set function (the heap top only decreases once full, so
 * every occurrence of a surviving hash is seen), so any selection that yields that set is
 * bit-exact. This kernel does it without a heap:
 *
 *   one CTA per segment (persistent grid), segment bytes staged HBM -> shared memory with a 1-D
 *   TMA bulk copy (cp.async.bulk + mbarrier, double buffered across segments);
 *   each thread slides forward / reverse-complement k-mer windows in registers over a contiguous
 *   run of positions and evaluates both Murmur3 hashes (INT-ALU bound: ~10 64-bit multiplies each);
 *   canonical hashes <= T (T ~ c*s/n * 2^64) are inserted into a shared-memory open-addressing
 *   table keyed by hash (atomicCAS), accumulating min position / max position / vote sum
 *   (atomicMin / atomicMax / atomicAdd) -- this de-duplicates before any sorting;
 *   if fewer than s distinct hashes survived although larger ones exist, or the table overflowed,
 *   T is raised / lowered / bisected and the pass is redone (rare; always terminates because
 *   distinct-count(T) grows by at most one per unit of T);
 *   the <= C survivors are ordered with a 256-bucket counting sort on the leading bits plus
 *   in-bucket ranking, and the first s are written out.
 */
#include "mm_internal.h"

namespace {

#ifndef MM_SK_THREADS
#define MM_SK_THREADS 128
#endif
constexpr int SK_THREADS = MM_SK_THREADS;
constexpr int SK_BUCKETS = 256;
constexpr uint64_t SK_EMPTY = ~0ULL;

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
/* 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP) */
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "MM_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra MM_DONE;\n"
      "bra MM_WAIT;\n"
      "MM_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

struct sk_ctrl {
  int distinct;
  int overflow;
  int above;       /* some valid canonical hash was > T */
  int has_max;     /* the hash value 0xFFFF...F (the table's empty marker) occurred */
  int max_first, max_last, max_votes;
  int _pad;
};

constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int SK_LIST_PER_WARP = 128; /* survivors buffered per warp before they are inserted into the table */

struct sk_smem_layout {
  uint32_t stage_bytes;  /* per staging buffer */
  uint32_t off_bar, off_keys, off_first, off_last, off_votes, off_order, off_bcnt, off_bstart, off_bfill,
      off_ctrl, off_list_h, off_list_m, total;
};

__host__ __device__ inline sk_smem_layout sk_layout(int seg_length, int C)
{
  sk_smem_layout L;
  L.stage_bytes = (uint32_t)(((seg_length + 15) & ~15) + 32);
  uint32_t o = 2 * L.stage_bytes;
  L.off_bar = o; o += 16;
  L.off_keys = o; o += 8u * C;
  L.off_first = o; o += 4u * C;
  L.off_last = o; o += 4u * C;
  L.off_votes = o; o += 4u * C;
  L.off_ctrl = o; o += (uint32_t)sizeof(sk_ctrl);
  o = (o + 15) & ~15u;
  /* the per-warp survivor lists are dead once the table is built; the ordering scratch (order[] + bucket counters)
   * lives in the same bytes -- 5 KB less per CTA is one more resident CTA per SM */
  const uint32_t lists = (8u + 4u) * SK_WARPS * SK_LIST_PER_WARP;
  const uint32_t ordering = ((2u * C + 15) & ~15u) + 3u * 4u * SK_BUCKETS;
  L.off_list_h = o;
  L.off_list_m = o + 8u * SK_WARPS * SK_LIST_PER_WARP;
  L.off_order = o;
  L.off_bcnt = o + ((2u * C + 15) & ~15u);
  L.off_bstart = L.off_bcnt + 4u * SK_BUCKETS;
  L.off_bfill = L.off_bstart + 4u * SK_BUCKETS;
  o += lists > ordering ? lists : ordering;
  L.total = (o + 15) & ~15u;
  return L;
}

/* insert one occurrence of canonical hash h at position pos with strand vote sv (+1/-1) */
__device__ __forceinline__ void sk_insert(unsigned long long *keys, int *first, int *last, int *votes,
                                          uint32_t mask, int limit, sk_ctrl *ctrl, uint64_t h, int pos, int sv)
{
  if (h == SK_EMPTY) { /* cannot be a table key; keep it in a dedicated entry (it sorts last) */
    ctrl->has_max = 1;
    atomicMin(&ctrl->max_first, pos);
    atomicMax(&ctrl->max_last, pos);
    atomicAdd(&ctrl->max_votes, sv);
    return;
  }
  uint32_t slot = ((uint32_t)h ^ (uint32_t)(h >> 32)) & mask;
  for (uint32_t probe = 0; probe <= mask; probe++) {
    if (*(volatile int *)&ctrl->overflow) return;
    unsigned long long cur = *(volatile unsigned long long *)&keys[slot];
    if (cur == SK_EMPTY) {
      cur = atomicCAS(&keys[slot], (unsigned long long)SK_EMPTY, (unsigned long long)h);
      if (cur == SK_EMPTY) {
        int d = atomicAdd(&ctrl->distinct, 1) + 1;
        if (d > limit) ctrl->overflow = 1;
        cur = h;
      }
    }
    if (cur == h) {
      atomicMin(&first[slot], pos);
      atomicMax(&last[slot], pos);
      atomicAdd(&votes[slot], sv);
      return;
    }
    slot = (slot + 1) & mask;
  }
  ctrl->overflow = 1;
}

template <int K>
__global__ void __launch_bounds__(SK_THREADS)
k_sketch(const uint8_t *__restrict__ bases, const mm_segment *__restrict__ segs, uint32_t n_segs, int S,
         int seg_length, int C, uint64_t *__restrict__ sk_hash, int2 *__restrict__ sk_pos,
         int8_t *__restrict__ sk_strand, mm_segment_result *__restrict__ seg_res)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const sk_smem_layout L = sk_layout(seg_length, C);
  uint64_t *bars = (uint64_t *)(smem + L.off_bar);
  unsigned long long *keys = (unsigned long long *)(smem + L.off_keys);
  int *first = (int *)(smem + L.off_first);
  int *last = (int *)(smem + L.off_last);
  int *votes = (int *)(smem + L.off_votes);
  uint16_t *order = (uint16_t *)(smem + L.off_order);
  uint32_t *bcnt = (uint32_t *)(smem + L.off_bcnt);
  uint32_t *bstart = (uint32_t *)(smem + L.off_bstart);
  uint32_t *bfill = (uint32_t *)(smem + L.off_bfill);
  sk_ctrl *ctrl = (sk_ctrl *)(smem + L.off_ctrl);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t *list_h = (uint64_t *)(smem + L.off_list_h) + (size_t)wid * SK_LIST_PER_WARP;
  uint32_t *list_m = (uint32_t *)(smem + L.off_list_m) + (size_t)wid * SK_LIST_PER_WARP;

  const int tid = threadIdx.x;
  const uint32_t mask = (uint32_t)C - 1;
  const int limit = C / 2;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_proxy_async();
  }
  __syncthreads();

  auto issue = [&](uint32_t seg, int stage) {
    const uint64_t off = segs[seg].offset;
    const int len = segs[seg].length;
    const uint64_t g0 = off & ~15ULL;
    const uint32_t bytes = (uint32_t)(((off + (uint64_t)len + 15ULL) & ~15ULL) - g0);
    mbar_expect_tx(&bars[stage], bytes);
    bulk_g2s(smem + (size_t)stage * L.stage_bytes, bases + g0, bytes, &bars[stage]);
  };

  uint32_t it = 0;
  if (tid == 0 && blockIdx.x < n_segs) issue(blockIdx.x, 0);

  for (uint32_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x, it++) {
    const int stage = it & 1;
    const uint32_t next = seg + gridDim.x;
    if (tid == 0 && next < n_segs) {
      fence_proxy_async(); /* generic-proxy reads of that buffer (previous iteration) before the async write */
      issue(next, stage ^ 1);
    }
    const uint64_t off = segs[seg].offset;
    const int len = segs[seg].length;
    const uint8_t *s = smem + (size_t)stage * L.stage_bytes + (off & 15ULL);
    mbar_wait(&bars[stage], (it >> 1) & 1);

    const int n = len - K + 1; /* number of k-mer positions (commonFunc.hpp:217) */
    const int P = n > 0 ? (n + SK_THREADS - 1) / SK_THREADS : 0;
    const int p0 = tid * P;
    const int p1 = min(n, p0 + P);

    /* initial threshold: expect c*S distinct survivors, c = 1.1 + 6/sqrt(S). The canonical hash is the MIN of two
     * uniform hashes, so P(canonical <= t) = 1 - (1-t)^2: solve that for the wanted fraction f = c*S/n. */
    uint64_t T = SK_EMPTY;
    if (n > 0) {
      const double c = 1.1 + 6.0 / sqrt((double)S);
      const double f = c * (double)S / (double)n;
      if (f < 1.0) T = (uint64_t)((1.0 - sqrt(1.0 - f)) * 18446744073709551616.0);
    }
    uint64_t lo = 0, hi = 0;
    bool have_lo = false, have_hi = false;

    while (true) {
      for (int i = tid; i < C; i += SK_THREADS) {
        keys[i] = SK_EMPTY;
        first[i] = 0x7fffffff;
        last[i] = -1;
        votes[i] = 0;
      }
      if (tid == 0) {
        ctrl->distinct = 0; ctrl->overflow = 0; ctrl->above = 0; ctrl->has_max = 0;
        ctrl->max_first = 0x7fffffff; ctrl->max_last = -1; ctrl->max_votes = 0;
      }
      __syncthreads();

      /* Hashing pass. Survivors (canonical hash <= T) are appended to this warp's list -- ballot + popc, no atomics --
       * and inserted into the table afterwards by all threads, so that the rare insert path (7 % of the positions, but
       * some lane of almost every warp iteration) does not serialise the hashing loop. */
      uint32_t wcount = 0; /* warp-uniform */
      {
        mm_kmer_window<K> w;
        w.reset();
        int run = 0; /* consecutive non-N bases ending at the current byte */
        bool above = false;
        const bool has_work = p0 < p1;
        if (has_work) {
#pragma unroll 1
          for (int j = 0; j < K - 1; j++) {
            bool isn;
            const uint32_t code = mm_base_code(s[p0 + j], isn);
            run = isn ? 0 : run + 1;
            w.push(code);
          }
        }
#pragma unroll 1
        for (int jj = 0; jj < P; jj++) {
          const int i = p0 + jj;
          bool surv = false;
          uint64_t h = 0;
          uint32_t meta = 0;
          if (i < p1) {
            bool isn;
            const uint32_t code = mm_base_code(s[i + K - 1], isn);
            run = isn ? 0 : run + 1;
            w.push(code);
            const uint64_t hf = w.hash_fwd();
            const uint64_t hb = w.hash_rev();
            if (run >= K && hf != hb) { /* commonFunc.hpp:234 */
              h = hf < hb ? hf : hb;
              meta = ((uint32_t)i << 1) | (hf < hb ? 1u : 0u);
              if (h <= T) surv = true; else above = true;
            }
          }
          const uint32_t sm = __ballot_sync(0xffffffffu, surv);
          if (sm) {
            const uint32_t idx = wcount + __popc(sm & ((1u << lane) - 1u));
            if (surv) {
              if (idx < (uint32_t)SK_LIST_PER_WARP) { list_h[idx] = h; list_m[idx] = meta; }
              else sk_insert(keys, first, last, votes, mask, limit, ctrl, h, (int)(meta >> 1), (meta & 1u) ? 1 : -1);
            }
            wcount += __popc(sm);
          }
        }
        if (above) ctrl->above = 1;
      }
      __syncwarp();
      {
        const uint32_t nl = min(wcount, (uint32_t)SK_LIST_PER_WARP);
        for (uint32_t q = lane; q < nl; q += 32) {
          const uint32_t meta = list_m[q];
          sk_insert(keys, first, last, votes, mask, limit, ctrl, list_h[q], (int)(meta >> 1), (meta & 1u) ? 1 : -1);
        }
      }
      __syncthreads();
      const int d = ctrl->distinct + ctrl->has_max;
      const int ovf = ctrl->overflow;
      const int abv = ctrl->above;
      if (ovf) { /* too many survivors: lower T */
        hi = T; have_hi = true;
        T = have_lo ? lo + (hi - lo) / 2 : T / 2;
      } else if (d < S && abv) { /* too few: raise T */
        lo = T; have_lo = true;
        if (have_hi) T = lo + (hi - lo) / 2;
        else T = (T > (SK_EMPTY >> 2)) ? SK_EMPTY : (T << 2) | 3ULL;
      } else {
        break;
      }
      __syncthreads();
    }

    /* ---- order the survivors: counting sort on the leading bits, rank inside the bucket ---- */
    const int dt = ctrl->distinct; /* entries in the table (excludes the has_max entry) */
    for (int i = tid; i < SK_BUCKETS; i += SK_THREADS) { bcnt[i] = 0; bfill[i] = 0; } /* aliases the (consumed) lists */
    __syncthreads();
    const int sh = max(0, (64 - __clzll((long long)T)) - 8);
    for (int i = tid; i < C; i += SK_THREADS) {
      const uint64_t k = keys[i];
      if (k != SK_EMPTY) atomicAdd(&bcnt[(uint32_t)(k >> sh)], 1u);
    }
    __syncthreads();
    if (tid < 32) { /* exclusive prefix over 256 buckets by one warp */
      uint32_t loc[SK_BUCKETS / 32];
      uint32_t sum = 0;
#pragma unroll
      for (int j = 0; j < SK_BUCKETS / 32; j++) { loc[j] = sum; sum += bcnt[tid * (SK_BUCKETS / 32) + j]; }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += v;
      }
      const uint32_t excl = incl - sum;
#pragma unroll
      for (int j = 0; j < SK_BUCKETS / 32; j++) bstart[tid * (SK_BUCKETS / 32) + j] = excl + loc[j];
    }
    __syncthreads();
    for (int i = tid; i < C; i += SK_THREADS) {
      const uint64_t k = keys[i];
      if (k != SK_EMPTY) {
        const uint32_t b = (uint32_t)(k >> sh);
        const uint32_t p = bstart[b] + atomicAdd(&bfill[b], 1u);
        order[p] = (uint16_t)i;
      }
    }
    __syncthreads();
    const size_t obase = (size_t)seg * (size_t)S;
    for (int p = tid; p < dt; p += SK_THREADS) {
      const uint32_t slot = order[p];
      const uint64_t k = keys[slot];
      const uint32_t b = (uint32_t)(k >> sh);
      const uint32_t bs = bstart[b], be = bs + bcnt[b];
      uint32_t rank = bs;
      for (uint32_t q = bs; q < be; q++) rank += (keys[order[q]] < k) ? 1u : 0u;
      if ((int)rank < S) {
        sk_hash[obase + rank] = k;
        sk_pos[obase + rank] = make_int2(first[slot], last[slot]);
        const int v = votes[slot];
        sk_strand[obase + rank] = (int8_t)(v > 0 ? 1 : (v == 0 ? 0 : -1)); /* commonFunc.hpp:282 */
      }
    }
    if (tid == 0) {
      int count = dt;
      if (ctrl->has_max) {
        if (dt < S) {
          sk_hash[obase + dt] = SK_EMPTY;
          sk_pos[obase + dt] = make_int2(ctrl->max_first, ctrl->max_last);
          const int v = ctrl->max_votes;
          sk_strand[obase + dt] = (int8_t)(v > 0 ? 1 : (v == 0 ? 0 : -1));
        }
        count = dt + 1;
      }
      if (count > S) count = S;
      mm_segment_result r;
      r.sketch_max_hash = 0; /* filled by the L1 kernel from sk_hash[count-1] */
      r.sketch_raw_count = count;
      r.sketch_size = count;
      r.n_points = 0; r.minimum_hits = 0; r.best_intersection = 0;
      r.first_candidate = 0; r.n_candidates = 0; r._pad = 0;
      seg_res[seg] = r;
    }
    __syncthreads(); /* all reads of the staging buffer and of the table are done */
  }
}

template <int K>
cudaError_t launch_k(const mm_params &p, const mm_dev_batch &b, cudaStream_t st, int sm_count, int C, size_t smem)
{
  cudaError_t e = cudaFuncSetAttribute(k_sketch<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sketch<K>, SK_THREADS, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  uint32_t grid = (uint32_t)sm_count * (uint32_t)occ; /* persistent: a whole number of CTAs per SM */
  if (grid > b.n_segs) grid = b.n_segs;
  if (grid == 0) return cudaSuccess;
  k_sketch<K><<<grid, SK_THREADS, smem, st>>>(b.bases, b.segs, b.n_segs, p.sketch_size, p.seg_length, C, b.sk_hash,
                                              b.sk_pos, b.sk_strand, b.seg_res);
  return cudaGetLastError();
}

} // namespace

#define MM_FOR_EACH_K(X) X(11) X(13) X(15) X(16) X(17) X(19) X(21) X(23) X(25) X(27) X(29) X(31) X(32)

int mm_sketch_kmer_supported(int k)
{
  switch (k) {
#define X(KK) case KK:
    MM_FOR_EACH_K(X)
#undef X
    return 1;
    default: return 0;
  }
}

size_t mm_sketch_smem_bytes(int seg_length, int sketch_size, int *table_cap)
{
  /* survivors ~ c*S with c = 1.1 + 6/sqrt(S); the table may be half full at most */
  const double c = 1.1 + 6.0 / sqrt((double)sketch_size);
  const double want = 2.0 * (c * sketch_size + 8.0 * sqrt(c * sketch_size) + 16.0);
  int C = 512;
  while (C < want) C <<= 1;
  if (C > 32768) return 0; /* order[] holds 16-bit slots */
  if (table_cap) *table_cap = C;
  const sk_smem_layout L = sk_layout(seg_length, C);
  if (L.total > 227u * 1024u) return 0;
  return L.total;
}

cudaError_t mm_launch_sketch(const mm_params &p, const mm_dev_batch &b, cudaStream_t st, int sm_count)
{
  int C = 0;
  const size_t smem = mm_sketch_smem_bytes(p.seg_length, p.sketch_size, &C);
  if (smem == 0) return cudaErrorInvalidValue;
  switch (p.kmer_size) {
#define X(KK) case KK: return launch_k<KK>(p, b, st, sm_count, C, smem);
    MM_FOR_EACH_K(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}
