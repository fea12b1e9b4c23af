/*
 * mm_l2_stream.cu -- K3 (fast path): L2 windowed-MinHash scan as a merge of two sorted streams.
 *
 * Same contract as mm_l2.cu (Map::computeL2MappedRegions, reference computeMap.hpp:1275-1451, with the
 * SlideMapper of slidingMap.hpp:27-212); what changes is how the work is laid out on the GPU.
 *
 * The reference keeps the live reference minmers in a min-heap on wpos_end (computeMap.hpp:1296-1300) to find the
 * ones to evict before each insertion (:1344-1358). Because the index is static, that heap is replaced by a second
 * copy of each contig's entries sorted by wpos_end (the "death order", built once at upload): the entries evicted
 * while the scan moves from one position to the next are a contiguous run of that copy. A candidate's scan is then
 * a two-pointer merge of
 *   the insert stream  = index entries with rangeStart - L - 1 <= wpos <= rangeEnd, in wpos order, and
 *   the delete stream  = death-order entries with rangeStart < wpos_end <= rangeEnd, in wpos_end order,
 * with "delete while wpos_end <= wpos of the next insert" (the eviction rule, <=). Entries of the delete stream were
 * all inserted before they are met (wpos < wpos_end, interval length <= L), and the state after a batch of
 * evictions does not depend on their order (the pivot invariant is restored by every single operation).
 *
 * Three kernels:
 *   k_l2_ranges  one thread per candidate: the four binary searches that delimit its two streams, and the
 *                number of operation records it needs (host-free exclusive scan follows);
 *   k_l2_prep    one warp per candidate: lanes take consecutive stream entries (coalesced SoA reads), binary-search
 *                the hash in the query sketch held in shared memory (slidingMap.hpp:128-131) and write one 8-byte
 *                record {position, slot | match | vote} per operation; no-op deletes and set-up entries that are
 *                never inserted are compacted away with warp ballots;
 *   k_l2_scan    ONE LANE per candidate, 32 candidates per warp: the sequential rank/pivot state machine
 *                (insert_minmer / delete_minmer, slidingMap.hpp:125-211) and the region tracking
 *                (computeMap.hpp:1373-1450) run in registers; the per-query-hash counters are one packed 32-bit word
 *                per slot in shared memory, laid out [slot][lane] so the 32 candidates never bank-conflict.
 * Candidates that produce more loci than the fixed slots reserved per candidate are flagged and redone by the
 * general warp-per-candidate kernel of mm_l2.cu.
 */
#include <cub/cub.cuh>

#include "mm_internal.h"

namespace {

constexpr int L2S_WARPS = 4;
constexpr int L2S_THREADS = L2S_WARPS * 32;
constexpr int L2C_WARPS = 2;              /* k_l2_scan: warps per CTA (each warp = 32 candidates) */
constexpr int L2C_THREADS = L2C_WARPS * 32;
/* per-slot state word (16 bits): num_before_inc bits 0..10, active bit 11, strand_vote bits 12..15 (signed) */
constexpr uint32_t W_NBI_MASK = 0x7FFu;
constexpr uint32_t W_ACT = 1u << 11;

__device__ __forceinline__ int w_sv(uint32_t w) { return ((int)(w << 16)) >> 28; }
__device__ __forceinline__ uint32_t w_set_sv(uint32_t w, int sv) { return (w & 0x0FFFu) | (((uint32_t)sv & 0xFu) << 12); }

/* first index in [lo,hi) with a[i] >= v */
__device__ __forceinline__ uint64_t lower_bound_i32(const int32_t *a, uint64_t lo, uint64_t hi, int32_t v)
{
  while (lo < hi) {
    const uint64_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void k_l2_ranges(const mm_params prm, const mm_dev_index ix, const mm_dev_batch b, uint32_t n_cands)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cands) return;
  const mm_l1_candidate cd = b.cands[c];
  const uint64_t cs = ix.contig_start[cd.seqId], ce = ix.contig_start[cd.seqId + 1];
  mm_l2_range r;
  /* firstOpenIt (computeMap.hpp:1290-1293) .. last entry with wpos <= rangeEnd (:1340) */
  r.it0 = lower_bound_i32(ix.idx_wpos, cs, ce, cd.rangeStartPos - prm.seg_length - 1);
  const uint64_t it1 = lower_bound_i32(ix.idx_wpos, r.it0, ce, cd.rangeEndPos + 1);
  /* evictions: wpos_end > rangeStart (anything smaller is never live) and <= rangeEnd (the last insert position) */
  r.d0 = lower_bound_i32(ix.idx2_wend, cs, ce, cd.rangeStartPos + 1);
  const uint64_t d1 = lower_bound_i32(ix.idx2_wend, r.d0, ce, cd.rangeEndPos + 1);
  r.nI = (uint32_t)(it1 - r.it0);
  r.nD = (uint32_t)(d1 - r.d0);
  r.next_wpos = 0;
  if (r.nI > 0) r.next_wpos = (it1 < ce) ? ix.idx_wpos[it1] : ix.idx_wpos[it1 - 1]; /* std::next(windowIt,...) (:1387-1390) */
  r._pad = 0;
  b.l2_ranges[c] = r;
  b.l2_rec_off[c] = (uint64_t)r.nI + (uint64_t)r.nD; /* counts; scanned in place afterwards */
}

__global__ void __launch_bounds__(L2S_THREADS)
k_l2_prep(const mm_params prm, const mm_dev_index ix, const mm_dev_batch b, uint32_t n_cands)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int S = prm.sketch_size;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t *qhash = (uint64_t *)smem_raw + (size_t)wid * (S + 2);
  int8_t *qstr = (int8_t *)((uint64_t *)smem_raw + (size_t)L2S_WARPS * (S + 2)) + (size_t)wid * (S + 2);
  const uint32_t FULL = 0xffffffffu;

  for (uint32_t c = blockIdx.x * L2S_WARPS + wid; c < n_cands; c += gridDim.x * L2S_WARPS) {
    const mm_l1_candidate cd = b.cands[c];
    mm_l2_range r = b.l2_ranges[c];
    const uint64_t off = b.l2_rec_off[c];
    if (off + r.nI + r.nD > b.l2_recs_cap) continue; /* host sized the buffer from the scan; cannot happen */
    const int n = b.seg_res[cd.segment].sketch_size;
    const size_t sbase = (size_t)cd.segment * (size_t)S;
    __syncwarp();
    for (int j = lane; j < n; j += 32) {
      qhash[j + 1] = b.sk_hash[sbase + j];
      qstr[j + 1] = b.sk_strand[sbase + j];
    }
    __syncwarp();
    uint2 *out = b.l2_recs + off;
    /* ---- insert stream ---- */
    uint32_t n_out = 0;
    for (uint32_t t0 = 0; t0 < r.nI; t0 += 32) {
      const uint32_t t = t0 + lane;
      bool keep = false;
      uint2 rec = make_uint2(0, 0);
      if (t < r.nI) {
        const uint64_t e = r.it0 + t;
        const uint64_t h = ix.idx_hash[e];
        const int wpos = ix.idx_wpos[e];
        int a = 1, z = n + 1; /* lower_bound over q_1..q_n (slidingMap.hpp:128-131) */
        while (a < z) {
          const int mid = (a + z) >> 1;
          if (qhash[mid] < h) a = mid + 1; else z = mid;
        }
        const bool match = a <= n && qhash[a] == h;
        uint32_t info = (uint32_t)a;
        if (match) {
          info |= MM_L2_MATCH;
          const int vote = (int)qstr[a] * (int)ix.idx_strand[e]; /* q_strand * mi.strand (slidingMap.hpp:141) */
          info |= (uint32_t)(vote & 3) << 17;                    /* 2-bit two's complement: -1, 0, +1 */
        }
        if (wpos < cd.rangeStartPos) { /* set-up entry (computeMap.hpp:1323-1338) */
          keep = ix.idx_wend[e] > cd.rangeStartPos && a <= n;
        } else {
          keep = true; /* every main entry is an evaluation point, even when it changes nothing */
        }
        rec = make_uint2((uint32_t)wpos, info);
      }
      const uint32_t km = __ballot_sync(FULL, keep);
      if (keep) out[n_out + __popc(km & ((1u << lane) - 1u))] = rec;
      n_out += __popc(km);
    }
    const uint32_t nI2 = n_out;
    /* ---- delete stream ---- */
    uint2 *dout = out + nI2;
    uint32_t n_del = 0;
    for (uint32_t t0 = 0; t0 < r.nD; t0 += 32) {
      const uint32_t t = t0 + lane;
      bool keep = false;
      uint2 rec = make_uint2(0, 0);
      if (t < r.nD) {
        const uint64_t e = r.d0 + t;
        const uint64_t h = ix.idx2_hash[e];
        int a = 1, z = n + 1;
        while (a < z) {
          const int mid = (a + z) >> 1;
          if (qhash[mid] < h) a = mid + 1; else z = mid;
        }
        keep = a <= n; /* hashes above every query hash never touch the state (slidingMap.hpp:133-136,179-182) */
        const bool match = keep && qhash[a] == h;
        rec = make_uint2((uint32_t)ix.idx2_wend[e], (uint32_t)a | (match ? MM_L2_MATCH : 0u));
      }
      const uint32_t km = __ballot_sync(FULL, keep);
      if (keep) dout[n_del + __popc(km & ((1u << lane) - 1u))] = rec;
      n_del += __popc(km);
    }
    if (lane == 0) {
      r.nI = nI2;
      r.nD = n_del;
      b.l2_ranges[c] = r;
    }
  }
}

/* order of the scan: candidates sorted by their (post-compaction) operation count, longest first, so that the 32 lanes
 * of a warp run scans of similar length */
__global__ void k_l2_order_keys(const mm_dev_batch b, uint32_t n_cands, uint32_t *keys, uint32_t *vals)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cands) return;
  const mm_l2_range r = b.l2_ranges[c];
  keys[c] = min(r.nI + r.nD, 0xFFFFu);
  vals[c] = c;
}

/* One lane's sequential reader of 8-byte op records: the current 4 records and the next 4 live in registers (32-byte
 * aligned chunks, two 16-byte loads each), the line after those is prefetched into L2. The scan consumes a record every
 * few hundred cycles per lane and L1 is almost entirely given to shared memory, so a plain `rec = p[i]` per step would
 * expose the full DRAM/L2 latency on every step. Reads run up to 8 records past the stream's end (the buffer has slack). */
struct rec_stream {
  const uint4 *next; /* chunk after `n0,n1` */
  uint4 c0, c1, n0, n1;
  uint32_t k;        /* record inside the current chunk, 0..3 */
  __device__ __forceinline__ void open(const uint2 *base, uint64_t first)
  {
    const uint4 *p = (const uint4 *)(base + (first & ~3ULL));
    k = (uint32_t)(first & 3ULL);
    c0 = p[0]; c1 = p[1]; n0 = p[2]; n1 = p[3];
    next = p + 4;
  }
  __device__ __forceinline__ uint2 at(uint32_t kk) const /* kk in 0..4: record kk of the current chunk / first of the next */
  {
    uint2 r;
    r.x = kk == 0 ? c0.x : kk == 1 ? c0.z : kk == 2 ? c1.x : kk == 3 ? c1.z : n0.x;
    r.y = kk == 0 ? c0.y : kk == 1 ? c0.w : kk == 2 ? c1.y : kk == 3 ? c1.w : n0.y;
    return r;
  }
  __device__ __forceinline__ uint2 cur() const { return at(k); }
  __device__ __forceinline__ uint32_t peek_x() const { return at(k + 1).x; }
  __device__ __forceinline__ void advance()
  {
    if (++k == 4) {
      k = 0;
      c0 = n0; c1 = n1;
      n0 = next[0]; n1 = next[1];
      asm volatile("prefetch.global.L2 [%0];" ::"l"(next + 8));
      next += 2;
    }
  }
};

struct lane_locus {
  int start, end, mean, shared, strand;
};

__global__ void __launch_bounds__(L2C_THREADS)
k_l2_scan(const mm_params prm, const mm_dev_index ix, const mm_dev_batch b, uint32_t n_cands)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int S = prm.sketch_size;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint16_t *words = (uint16_t *)smem_raw + (size_t)wid * (size_t)(S + 2) * 32; /* [slot][lane] */
  const uint32_t FULL = 0xffffffffu;
  const int LPC = (int)b.l2_loci_per_cand;
  const int segL = prm.seg_length;

  for (uint32_t cbase = (blockIdx.x * L2C_WARPS + wid) * 32; cbase < n_cands; cbase += gridDim.x * L2C_WARPS * 32) {
    const bool valid = cbase + lane < n_cands;
    const uint32_t c = valid ? (b.l2_perm ? b.l2_perm[cbase + lane] : cbase + lane) : 0u;
    mm_l1_candidate cd;
    mm_l2_range r;
    r.nI = 0; r.nD = 0; r.it0 = 0; r.d0 = 0; r.next_wpos = 0;
    cd.seqId = 0; cd.rangeStartPos = 0; cd.rangeEndPos = 0; cd.segment = 0;
    int n = 0;
    if (valid) {
      cd = b.cands[c];
      r = b.l2_ranges[c];
      n = b.seg_res[cd.segment].sketch_size;
    }
    /* SlideMapper::init (slidingMap.hpp:104-121): every query slot counts itself once */
    int nmax = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, o));
    __syncwarp();
    for (int j = 0; j <= nmax + 1; j++) {
      uint32_t w = 1u;
      if (j == 0) w = 0u;
      if (j > n) w = W_NBI_MASK; /* beyond the sketch: rank can never fit */
      words[j * 32 + lane] = (uint16_t)w;
    }
    __syncwarp();
    int pivot = n, pivRank = n, shared = 0, votes = 0;
    int best = 1; /* bestSketchSize (computeMap.hpp:1317) */
    bool in_cand = false, has_back = false;
    lane_locus cur = {0, 0, 0, 0, 0}, back = {0, 0, 0, 0, 0};
    int n_loci = 0;
    mm_l2_locus *lout = b.loci + (size_t)c * (size_t)LPC;
    auto store = [&](int k, const lane_locus &l) {
      if (k < LPC) {
        mm_l2_locus o;
        o.seqId = cd.seqId; o.meanOptimalPos = l.mean; o.optimalStart = l.start; o.optimalEnd = l.end;
        o.sharedSketchSize = l.shared; o.strand = l.strand;
        lout[k] = o;
      }
    };
    auto push_or_merge = [&](const lane_locus &l) { /* computeMap.hpp:1417-1426, :1440-1449 */
      if (!has_back) { back = l; has_back = true; }
      else if (back.end + segL < l.start) { store(n_loci, back); n_loci++; back = l; }
      else { back.end = l.end; back.mean = (back.start + back.end) / 2; }
    };

    uint32_t i = 0, d = 0;
    rec_stream is, ds;
    is.open(b.l2_recs, valid ? b.l2_rec_off[c] : 0ULL);
    ds.open(b.l2_recs, valid ? b.l2_rec_off[c] + r.nI : 0ULL);
    while (__any_sync(FULL, i < r.nI)) {
      if (i < r.nI) {
        const uint2 irec = is.cur(), drec = ds.cur();
        const int ipos = (int)irec.x;
        const bool is_main = ipos >= cd.rangeStartPos;
        const bool do_del = is_main && d < r.nD && (int)drec.x <= ipos; /* evict while wpos_end <= wpos (:1344) */
        const uint32_t info = do_del ? drec.y : irec.y;
        const int slot = (int)(info & MM_L2_SLOT_MASK);
        const bool match = (info & MM_L2_MATCH) != 0;
        const int prev_votes = votes; /* computeMap.hpp:1342 (only read after an insert) */
        if (slot <= n) {
          /* insert_minmer / delete_minmer (slidingMap.hpp:125-165, :171-211) as one straight-line update: the four
           * cases (insert/delete x hash in the query sketch or not) are selected with predicates, so the 32 candidates of
           * the warp do not diverge. q is the slot whose membership in the pivot prefix may change: the pivot itself on
           * an insert (it is popped when the rank overflows), the slot after it on a delete (it is pulled in when it
           * fits). words[n+1] holds an unreachable count, which stands for the reference's `pivot != end` test. */
          const uint32_t w = words[slot * 32 + lane];
          const bool le = slot <= pivot;
          const int sgn = do_del ? -1 : 1;
          const int svw = w_sv(w);
          int vote = (int)((info >> 17) & 3u);
          vote = vote == 3 ? -1 : vote;
          const int sv2 = do_del ? 0 : svw + vote;
          const uint32_t w_match = w_set_sv(do_del ? (w & ~W_ACT) : (w | W_ACT), sv2);
          const uint32_t w_plain = w + (uint32_t)sgn;
          const int pr = pivRank + (le ? sgn : 0);
          const int q = pivot + (do_del ? 1 : 0);
          uint32_t x = words[q * 32 + lane];
          if (q == slot) x = w_plain;
          const int xn = (int)(x & W_NBI_MASK), xa = (x & W_ACT) ? 1 : 0, xs = w_sv(x);
          const int mv = match ? 0 : (do_del ? ((pr + xn <= n) ? 1 : 0) : ((pr > n) ? -1 : 0));
          shared += match ? (le ? sgn : 0) : mv * xa;
          votes += match ? (le ? (do_del ? -svw : sv2) : 0) : mv * xs;
          pivRank = match ? pivRank : pr + mv * xn;
          pivot += mv;
          words[slot * 32 + lane] = (uint16_t)(match ? w_match : w_plain);
        }
        if (do_del) {
          d++;
          ds.advance();
        } else {
          const int npos = (i + 1 < r.nI) ? (int)is.peek_x() : r.next_wpos;
          if (is_main) { /* region tracking (computeMap.hpp:1373-1430) */
            if (shared > best) {
              n_loci = 0; has_back = false; /* l2_vec_out.clear() */
              in_cand = true;
              best = shared;
              cur.shared = shared; cur.start = ipos; cur.end = npos;
            } else if (shared == best) {
              if (!in_cand) { cur.shared = shared; cur.start = ipos; }
              in_cand = true;
              cur.end = npos;
            } else {
              if (in_cand) {
                cur.end = npos;
                cur.mean = (cur.start + cur.end) / 2;
                cur.strand = prev_votes >= 0 ? 1 : -1;
                push_or_merge(cur);
                cur.start = cur.end = cur.mean = cur.shared = cur.strand = 0;
              }
              in_cand = false;
            }
          }
          i++;
          is.advance();
        }
      }
    }
    if (valid) {
      if (in_cand) { /* computeMap.hpp:1435-1450 */
        cur.mean = (cur.start + cur.end) / 2;
        cur.strand = votes >= 0 ? 1 : -1;
        push_or_merge(cur);
      }
      if (has_back) { store(n_loci, back); n_loci++; }
      if (n_loci > LPC) {
        atomicAdd(b.counters + 7, 1u); /* redo by the general kernel */
        b.cands[c].first_locus = 0;
        b.cands[c].n_loci = 0xFFFFFFFFu;
      } else {
        b.cands[c].first_locus = c * (uint32_t)LPC;
        b.cands[c].n_loci = (uint32_t)n_loci;
      }
    }
  }
}

__global__ void k_fill_death_keys(const int32_t *idx_wend, const uint64_t *contig_start, int32_t n_contigs, uint64_t n,
                                  uint64_t *keys, uint32_t *vals)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = n_contigs; /* contig of entry i: last c with contig_start[c] <= i */
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (contig_start[mid] <= i) lo = mid; else hi = mid;
  }
  keys[i] = ((uint64_t)(uint32_t)lo << 32) | (uint64_t)(uint32_t)idx_wend[i];
  vals[i] = (uint32_t)i;
}
__global__ void k_gather_death(const uint64_t *idx_hash, const uint64_t *keys_sorted, const uint32_t *vals_sorted, uint64_t n,
                               uint64_t *idx2_hash, int32_t *idx2_wend)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  idx2_hash[i] = idx_hash[vals_sorted[i]];
  idx2_wend[i] = (int32_t)(uint32_t)keys_sorted[i];
}

__global__ void k_split_minmers(const mm_minmer *aos, uint64_t n, uint64_t *hash, int32_t *wpos, int32_t *wend, int8_t *strand)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const mm_minmer m = aos[i];
  hash[i] = m.hash; wpos[i] = m.wpos; wend[i] = m.wpos_end; strand[i] = (int8_t)m.strand;
}
__global__ void k_pack_points(const mm_ipoint *aos, uint64_t n, int32_t n_contigs, uint64_t *packed, uint32_t *err)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const mm_ipoint p = aos[i];
  if (p.seqId < 0 || p.seqId >= n_contigs || p.pos < 0) atomicOr(err, 1u);
  packed[i] = mm_pack_point(p.seqId, p.pos, p.side > 0);
}
/* keys are distinct: a slot is claimed by CAS on its value word, the key is written afterwards (no reader yet) */
__global__ void k_build_table(const uint64_t *keys, const uint64_t *offs, const uint8_t *is_freq, uint64_t n_keys, mm_tab_slot *tab,
                              int tab_log2, uint32_t *err)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys) return;
  uint64_t cnt = offs[i + 1] - offs[i];
  if (is_freq[i] && cnt > MM_VAL_CNT_MASK) cnt = MM_VAL_CNT_MASK; /* a frequent seed's list is never gathered: only the flag is read */
  if (cnt == 0 || cnt > MM_VAL_CNT_MASK || offs[i] >= (1ULL << (64 - MM_VAL_OFF_SHIFT))) { atomicOr(err, 2u); return; }
  const uint64_t val = (offs[i] << MM_VAL_OFF_SHIFT) | (cnt << 1) | (is_freq[i] ? 1ULL : 0ULL);
  const uint32_t mask = (1u << tab_log2) - 1u;
  uint32_t slot = mm_tab_slot_of(keys[i], tab_log2);
  for (uint32_t probe = 0; probe <= mask; probe++) {
    const unsigned long long old = atomicCAS((unsigned long long *)&tab[slot].val, 0ULL, (unsigned long long)val);
    if (old == 0ULL) { tab[slot].key = keys[i]; return; }
    slot = (slot + 1) & mask;
  }
  atomicOr(err, 2u);
}
/* duplicates would occupy two slots: detect them after the build */
__global__ void k_check_table_dups(const mm_tab_slot *tab, int tab_log2, uint32_t *err)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t slots = 1ULL << tab_log2;
  if (i >= slots || tab[i].val == 0) return;
  const uint32_t mask = (uint32_t)(slots - 1);
  uint32_t j = ((uint32_t)i + 1) & mask;
  while (tab[j].val != 0) { /* the probe run that follows */
    if (tab[j].key == tab[i].key) { atomicOr(err, 4u); return; }
    j = (j + 1) & mask;
    if (j == (uint32_t)i) return;
  }
}

} // namespace

cudaError_t mm_upload_split_minmers(const mm_minmer *aos, uint64_t n, uint64_t *hash, int32_t *wpos, int32_t *wend, int8_t *strand,
                                    cudaStream_t st)
{
  if (n == 0) return cudaSuccess;
  k_split_minmers<<<(uint32_t)((n + 255) / 256), 256, 0, st>>>(aos, n, hash, wpos, wend, strand);
  return cudaGetLastError();
}
cudaError_t mm_upload_pack_points(const mm_ipoint *aos, uint64_t n, int32_t n_contigs, uint64_t *packed, uint32_t *err, cudaStream_t st)
{
  if (n == 0) return cudaSuccess;
  k_pack_points<<<(uint32_t)((n + 255) / 256), 256, 0, st>>>(aos, n, n_contigs, packed, err);
  return cudaGetLastError();
}
cudaError_t mm_upload_build_table(const uint64_t *keys, const uint64_t *offs, const uint8_t *is_freq, uint64_t n_keys, mm_tab_slot *tab,
                                  int tab_log2, uint32_t *err, cudaStream_t st)
{
  if (n_keys == 0) return cudaSuccess;
  k_build_table<<<(uint32_t)((n_keys + 255) / 256), 256, 0, st>>>(keys, offs, is_freq, n_keys, tab, tab_log2, err);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const uint64_t slots = 1ULL << tab_log2;
  k_check_table_dups<<<(uint32_t)((slots + 255) / 256), 256, 0, st>>>(tab, tab_log2, err);
  return cudaGetLastError();
}

/* per contig, entries sorted by wpos_end (stable): one device radix sort on (seqId, wpos_end) */
cudaError_t mm_build_death_order(const uint64_t *idx_hash, const int32_t *idx_wend, const uint64_t *contig_start,
                                 int32_t n_contigs, uint64_t n, uint64_t *idx2_hash, int32_t *idx2_wend, cudaStream_t st)
{
  if (n == 0) return cudaSuccess;
  if (n >= (1ULL << 32)) return cudaErrorInvalidValue;
  uint64_t *keys = nullptr, *keys2 = nullptr;
  uint32_t *vals = nullptr, *vals2 = nullptr;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  cudaError_t e;
  if ((e = cudaMalloc(&keys, n * 8)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&keys2, n * 8)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&vals, n * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&vals2, n * 4)) != cudaSuccess) return e;
  const uint32_t grid = (uint32_t)((n + 255) / 256);
  k_fill_death_keys<<<grid, 256, 0, st>>>(idx_wend, contig_start, n_contigs, n, keys, vals);
  int end_bit = 32;
  while ((1LL << (end_bit - 32)) < (long long)n_contigs + 1) end_bit++;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, vals2, (int64_t)n, 0, end_bit, st);
  if ((e = cudaMalloc(&tmp, tmp_bytes)) != cudaSuccess) return e;
  e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, (int64_t)n, 0, end_bit, st);
  if (e != cudaSuccess) return e;
  k_gather_death<<<grid, 256, 0, st>>>(idx_hash, keys2, vals2, n, idx2_hash, idx2_wend);
  e = cudaStreamSynchronize(st);
  cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(vals2); cudaFree(tmp);
  return e != cudaSuccess ? e : cudaGetLastError();
}

size_t mm_l2_scan_tmp_bytes(uint32_t n_cands)
{
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (uint64_t *)nullptr, (uint64_t *)nullptr, (int)n_cands + 1);
  return bytes;
}

cudaError_t mm_launch_l2_ranges(const mm_params &p, const mm_dev_index &ix, const mm_dev_batch &b, uint32_t n_cands,
                                void *scan_tmp, size_t scan_tmp_bytes, cudaStream_t st)
{
  if (n_cands == 0) return cudaSuccess;
  k_l2_ranges<<<(n_cands + 127) / 128, 128, 0, st>>>(p, ix, b, n_cands);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  /* exclusive scan over n_cands+1 elements: the last one (written 0 by the caller) becomes the total */
  return cub::DeviceScan::ExclusiveSum(scan_tmp, scan_tmp_bytes, b.l2_rec_off, b.l2_rec_off, (int)n_cands + 1, st);
}

/* temp bytes needed by mm_launch_l2_order for n candidates (4 u32 arrays + the library's sort storage) */
size_t mm_l2_order_bytes(uint32_t n_cands)
{
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                            (uint32_t *)nullptr, (int)n_cands, 0, 16);
  return (size_t)n_cands * 16 + 1024 + tmp;
}

/* work: mm_l2_order_bytes(n_cands) bytes; returns the permutation (device pointer inside work) in *perm */
cudaError_t mm_launch_l2_order(const mm_dev_batch &b, uint32_t n_cands, void *work, size_t work_bytes, uint32_t **perm, cudaStream_t st)
{
  uint32_t *k0 = (uint32_t *)work, *v0 = k0 + n_cands, *k1 = v0 + n_cands, *v1 = k1 + n_cands;
  void *tmp = (void *)(((uintptr_t)(v1 + n_cands) + 255) & ~(uintptr_t)255);
  size_t tmp_bytes = work_bytes - (size_t)((char *)tmp - (char *)work);
  k_l2_order_keys<<<(n_cands + 255) / 256, 256, 0, st>>>(b, n_cands, k0, v0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  *perm = v1;
  return cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, k0, k1, v0, v1, (int)n_cands, 0, 16, st);
}

static size_t l2_prep_smem(const mm_params &p) { return (size_t)L2S_WARPS * (size_t)(p.sketch_size + 2) * 9 + 16; }
static size_t l2_scan_smem(const mm_params &p) { return (size_t)L2C_WARPS * (size_t)(p.sketch_size + 2) * 32 * 2; }

cudaError_t mm_launch_l2_prep(const mm_params &p, const mm_dev_index &ix, const mm_dev_batch &b, uint32_t n_cands, cudaStream_t st,
                              int sm_count)
{
  if (n_cands == 0) return cudaSuccess;
  const size_t smem = l2_prep_smem(p);
  cudaError_t e = cudaFuncSetAttribute(k_l2_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_l2_prep, L2S_THREADS, smem);
  if (e != cudaSuccess) return e;
  uint32_t grid = (uint32_t)sm_count * (uint32_t)max(occ, 1);
  grid = min(grid, (n_cands + L2S_WARPS - 1) / L2S_WARPS);
  k_l2_prep<<<grid, L2S_THREADS, smem, st>>>(p, ix, b, n_cands);
  return cudaGetLastError();
}

cudaError_t mm_launch_l2_scan(const mm_params &p, const mm_dev_index &ix, const mm_dev_batch &b, uint32_t n_cands, cudaStream_t st,
                              int sm_count)
{
  if (n_cands == 0) return cudaSuccess;
  const size_t smem = l2_scan_smem(p);
  if (smem > 227 * 1024 || p.sketch_size > 1000) return cudaErrorInvalidValue; /* 11-bit counters: caller falls back */
  cudaError_t e = cudaFuncSetAttribute(k_l2_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_l2_scan, L2C_THREADS, smem);
  if (e != cudaSuccess) return e;
  uint32_t grid = (uint32_t)sm_count * (uint32_t)max(occ, 1);
  grid = min(grid, (n_cands + L2C_THREADS - 1) / L2C_THREADS);
  k_l2_scan<<<grid, L2C_THREADS, smem, st>>>(p, ix, b, n_cands);
  return cudaGetLastError();
}
