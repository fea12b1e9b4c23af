/*
 * mm_index_build.cu -- the reference index built on the GPU.
 *
 * Replaces, for a reference that is already in device memory as text:
 *   CommonFunc::addMinmers        reference src/map/include/commonFunc.hpp:301-570   (sliding-window minmer intervals)
 *   Sketch::index                 winSketch.hpp:379-404                              (hash -> interval points, fusion rule)
 *   Sketch::computeFreqHist / computeFreqSeedSet / dropFreqSeedSet   :410-453, :488-504 (frequent-seed filter)
 * and leaves the device arrays that mm_index_upload would have produced from host arrays.
 *
 * The window scan of addMinmers is a sequential state machine whose every record boundary is an L2 evaluation point, so it
 * is not re-derived: mm_winmachine.h restates it once (tested record for record against the reference on the CPU) and
 * this file runs that machine in parallel over CHUNKS of every contig, one GPU thread per chunk:
 *   k_window_scan   chunk [a, b) starts WARM positions early from an empty machine (records suppressed until a). At a it
 *                   takes a digest of its state, marks the records that are open as "started earlier", scans to b, takes
 *                   another digest and exports which hashes are open (with the start of their record).
 *   host            chunk j is accepted iff chunk j-1 is, digest_start(j) == digest_end(j-1) and the machine never took an
 *                   expired heap entry (wm_machine::drained: the only way history older than the window can matter).
 *   k_window_fix    rejected chunks (N runs, low complexity; none on ordinary sequence) are re-scanned by ONE thread per
 *                   run of them that first rebuilds the exact state at the run's start from the accepted chunk before it.
 *   k_patch_starts  a record that was open at its chunk's start gets its wpos from the previous chunk's export.
 * Then the post-processing of :522-568 with scans and radix sorts (malformed records, strand collapse, chunking to <= w,
 * order by (seqId, wpos, wpos_end) -- STABLE in emission order where the reference's std::sort leaves exact ties in
 * libstdc++'s order, see DESIGN.md --, adjacent de-duplication), Sketch::index as a sort by hash + adjacent-record rule,
 * the frequency histogram on the device and its threshold on the host (a few hundred numbers).
 */
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "mm_index_build.h"
#include "mm_winmachine.h"

namespace {

struct wb_chunk {
  int32_t contig;
  int32_t a, b;    /* k-mer positions [a, b) */
  int32_t npos;    /* positions of the contig (len - k + 1) */
};
struct wb_chunk_out {
  uint64_t d_start, d_end;
  uint32_t n_rec;
  uint32_t flags;  /* 1 machine failure, 2 record buffer full, 4 expired heap entry taken (history-dependent) */
  uint32_t n_open;
  uint32_t _pad;
};
struct wb_open {
  uint64_t hash;
  int32_t wpos;
  uint32_t inherited; /* the record was already open at the chunk's start: wpos is the warm-up machine's, to be resolved */
};
struct wb_chain {
  uint32_t first, n;     /* rejected chunks [first, first + n) */
  uint64_t out_offset;   /* first record slot of the chain in the fix buffer; chunk q gets fix_cap slots at out_offset + (q - first) * fix_cap */
};

struct wb_slab_layout {
  size_t off_ring, off_heap, off_nodes, off_mem, off_mh, off_mslot, off_sfree, bytes;
  int32_t ring_cap, heap_cap, node_cap, mem_cap;
};
wb_slab_layout slab_layout(int w, int s)
{
  wb_slab_layout L;
  L.ring_cap = wm_ring_cap(w); L.heap_cap = wm_heap_cap(w); L.node_cap = wm_node_cap(w); L.mem_cap = wm_mem_cap(s);
  size_t o = 0;
  L.off_ring = o; o += (size_t)L.ring_cap * sizeof(wm_kmer);
  L.off_heap = o; o += (size_t)L.heap_cap * sizeof(wm_kmer);
  L.off_nodes = o; o += (size_t)L.node_cap * sizeof(wm_node);
  o = (o + 15) & ~(size_t)15;
  L.off_mem = o; o += (size_t)L.mem_cap * sizeof(wm_member);
  L.off_mh = o; o += (size_t)L.mem_cap * 8;
  L.off_mslot = o; o += (size_t)L.mem_cap * 2;
  L.off_sfree = o; o += (size_t)L.mem_cap * 2;
  L.bytes = (o + 255) & ~(size_t)255;
  return L;
}
__device__ __forceinline__ void attach(wm_machine &m, unsigned char *slab, const wb_slab_layout &L)
{
  m.ring = (wm_kmer *)(slab + L.off_ring); m.ring_cap = L.ring_cap;
  m.heap = (wm_kmer *)(slab + L.off_heap); m.heap_cap = L.heap_cap;
  m.nodes = (wm_node *)(slab + L.off_nodes); m.node_cap = L.node_cap;
  m.slots = (wm_member *)(slab + L.off_mem); m.mem_cap = L.mem_cap;
  m.mh = (uint64_t *)(slab + L.off_mh); m.mslot = (uint16_t *)(slab + L.off_mslot); m.sfree = (uint16_t *)(slab + L.off_sfree);
}
__device__ __forceinline__ void export_open(const wm_machine &m, wb_open *ex, wb_chunk_out &o)
{
  o.n_open = (uint32_t)m.mem_n;
  for (int32_t j = 0; j < m.mem_n; j++) { const wm_member &e = wm_at(m, j); ex[j].hash = e.hash; ex[j].wpos = e.wpos; ex[j].inherited = e.inherited; }
}
__device__ __forceinline__ int32_t find_open(const wb_open *ex, uint32_t n, uint64_t h)
{
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (ex[mid].hash < h) lo = mid + 1; else hi = mid;
  }
  return (lo < n && ex[lo].hash == h) ? (int32_t)lo : -1;
}

/* pass 1: every chunk on its own, from a warm-up */
template <int K>
__global__ void __launch_bounds__(128)
k_window_scan(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ contig_off, const wb_chunk *__restrict__ chunks,
              uint32_t n_chunks, int w, int s, int warm, unsigned char *slabs, wb_slab_layout L, wm_record *rec_buf, uint32_t rec_cap,
              wb_chunk_out *outs, wb_open *exports, uint32_t export_stride)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, T = gridDim.x * blockDim.x;
  unsigned char *slab = slabs + (size_t)t * L.bytes;
  for (uint32_t c = t; c < n_chunks; c += T) {
    const wb_chunk ch = chunks[c];
    const uint8_t *base = seq + contig_off[ch.contig];
    wm_machine m;
    attach(m, slab, L);
    m.out = rec_buf + (size_t)c * rec_cap; m.out_cap = rec_cap;
    wm_init(m, K, w, s);
    wm_kmer_bytes<K> win;
    wb_chunk_out o;
    o.d_start = 0; o.flags = 0; o._pad = 0;
    if (ch.a > 0) {
      const int32_t from = ch.a - warm > 0 ? ch.a - warm : 0;
      m.emit_from = ch.a;
      wm_scan<K>(m, win, base, from, ch.a, true);
      o.d_start = wm_digest(m, ch.a - 1 + K - w);
      if (m.drained) o.flags |= 4u;
      m.drained = 0;
      for (int32_t j = 0; j < m.mem_n; j++) wm_at(m, j).inherited = 1; /* their records started before a */
      wm_scan<K>(m, win, base, ch.a, ch.b, false);
    } else {
      wm_scan<K>(m, win, base, 0, ch.b, true);
    }
    if (ch.a > 0 && m.drained) o.flags |= 4u; /* a chunk that starts at 0 is exact whatever its heap did */
    o.d_end = wm_digest(m, ch.b - 1 + K - w);
    export_open(m, exports + (size_t)c * export_stride, o);
    if (ch.b == ch.npos) wm_flush(m, ch.npos);
    if (m.out_n >= m.out_cap) o.flags |= 2u;
    else if (m.fail) o.flags |= 1u;
    o.n_rec = (uint32_t)m.out_n;
    outs[c] = o;
  }
}

/* pass 2: a run of rejected chunks, scanned by one thread from the exact state at the run's start */
template <int K>
__global__ void __launch_bounds__(64)
k_window_fix(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ contig_off, const wb_chunk *__restrict__ chunks,
             const wb_chain *__restrict__ chains, uint32_t n_chains, int w, int s, int warm, unsigned char *slabs, wb_slab_layout L,
             wm_record *fix_buf, uint32_t fix_cap, wb_chunk_out *outs, wb_open *exports, uint32_t export_stride)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_chains) return;
  const wb_chain cn = chains[t];
  unsigned char *slab = slabs + (size_t)t * L.bytes;
  wm_machine m;
  attach(m, slab, L);
  wm_record dummy;
  m.out = &dummy; m.out_cap = 0;
  wm_init(m, K, w, s);
  wm_kmer_bytes<K> win;
  const wb_chunk first = chunks[cn.first];
  const uint8_t *base = seq + contig_off[first.contig];
  bool fresh = true;
  if (first.a > 0) { /* rebuild the exact state at first.a: the accepted chunk before it, silently */
    const wb_chunk pv = chunks[cn.first - 1];
    m.emit_from = 0x7fffffff;
    if (pv.a > 0) {
      const int32_t from = pv.a - warm > 0 ? pv.a - warm : 0;
      wm_scan<K>(m, win, base, from, pv.a, true);
      const wb_open *ex = exports + (size_t)(cn.first - 2) * export_stride;
      const uint32_t nx = outs[cn.first - 2].n_open;
      for (int32_t j = 0; j < m.mem_n; j++) {
        const int32_t at = find_open(ex, nx, m.mh[j]);
        if (at >= 0) wm_at(m, j).wpos = ex[at].wpos;
      }
      wm_scan<K>(m, win, base, pv.a, pv.b, false);
    } else {
      wm_scan<K>(m, win, base, 0, pv.b, true);
    }
    fresh = false;
  }
  m.emit_from = 0;
  for (uint32_t q = cn.first; q < cn.first + cn.n; q++) {
    const wb_chunk ch = chunks[q];
    m.out = fix_buf + cn.out_offset + (size_t)(q - cn.first) * fix_cap; m.out_cap = fix_cap; m.out_n = 0;
    m.fail = 0;
    wm_scan<K>(m, win, base, ch.a, ch.b, fresh);
    fresh = false;
    wb_chunk_out o = outs[q];
    o.flags = 8u; /* fixed: exact by construction */
    o.d_end = wm_digest(m, ch.b - 1 + K - w);
    export_open(m, exports + (size_t)q * export_stride, o);
    if (ch.b == ch.npos) wm_flush(m, ch.npos);
    if (m.out_n >= m.out_cap) o.flags |= 2u;
    else if (m.fail) o.flags |= 1u;
    o.n_rec = (uint32_t)m.out_n;
    outs[q] = o;
  }
}

/* A record can stay open over several chunks: an exported entry that was itself inherited takes its start from the previous
 * chunk's (already resolved) export. Sequential along the chunks of a contig, one block per contig, cheap (s entries per chunk). */
__global__ void k_resolve_exports(const uint32_t *__restrict__ contig_first, const uint32_t *__restrict__ contig_n, uint32_t n_used,
                                  const wb_chunk_out *__restrict__ outs, wb_open *exports, uint32_t export_stride, uint32_t *err)
{
  const uint32_t c = blockIdx.x;
  if (c >= n_used) return;
  const uint32_t first = contig_first[c], n = contig_n[c];
  for (uint32_t q = first + 1; q < first + n; q++) {
    const wb_open *pv = exports + (size_t)(q - 1) * export_stride;
    wb_open *cur = exports + (size_t)q * export_stride;
    const uint32_t np = outs[q - 1].n_open;
    for (uint32_t e = threadIdx.x; e < outs[q].n_open; e += blockDim.x) {
      if (!cur[e].inherited) continue;
      const int32_t at = find_open(pv, np, cur[e].hash);
      if (at < 0) { atomicOr(err, 2u); continue; }
      cur[e].wpos = pv[at].wpos;
      cur[e].inherited = 0;
    }
    __syncthreads();
  }
}

/* records of accepted (not re-scanned) chunks that were open at the chunk's start: wpos from the previous chunk's export */
__global__ void k_patch_starts(const wb_chunk *__restrict__ chunks, const wb_chunk_out *__restrict__ outs, uint32_t n_chunks,
                               wm_record *rec_buf, uint32_t rec_cap, const wb_open *__restrict__ exports, uint32_t export_stride,
                               uint32_t *err)
{
  const uint32_t c = blockIdx.x;
  if (c >= n_chunks) return;
  if (chunks[c].a == 0 || (outs[c].flags & 8u)) return;
  const wb_open *ex = exports + (size_t)(c - 1) * export_stride;
  const uint32_t nx = outs[c - 1].n_open;
  wm_record *r = rec_buf + (size_t)c * rec_cap;
  for (uint32_t i = threadIdx.x; i < outs[c].n_rec; i += blockDim.x) {
    if (!r[i].inherited) continue;
    const int32_t at = find_open(ex, nx, r[i].hash);
    if (at < 0) { atomicOr(err, 1u); continue; }
    r[i].wpos = ex[at].wpos;
  }
}

/* ---- post-processing of addMinmers (:522-568) ---------------------------------------------------------------------- */

/* chunk buffers -> one raw array in emission order; per record: kept as it is (1) / number of pieces it is cut into */
__global__ void k_gather_raw(const wb_chunk *__restrict__ chunks, const wb_chunk_out *__restrict__ outs, uint32_t n_chunks,
                             const uint64_t *__restrict__ raw_off, const wm_record *__restrict__ rec_buf, uint32_t rec_cap,
                             const wm_record *__restrict__ fix_buf, const uint64_t *__restrict__ fix_off, int w,
                             uint64_t *r_hash, int32_t *r_wpos, int32_t *r_wend, int32_t *r_seq, int8_t *r_strand,
                             uint32_t *keep, uint32_t *pieces)
{
  const uint32_t c = blockIdx.x;
  if (c >= n_chunks) return;
  const wb_chunk_out o = outs[c];
  const wm_record *src = (o.flags & 8u) ? fix_buf + fix_off[c] : rec_buf + (size_t)c * rec_cap;
  const uint64_t at = raw_off[c];
  const int32_t seqId = chunks[c].contig;
  for (uint32_t i = threadIdx.x; i < o.n_rec; i += blockDim.x) {
    const wm_record r = src[i];
    const uint64_t d = at + i;
    r_hash[d] = r.hash; r_wpos[d] = r.wpos; r_wend[d] = r.wpos_end; r_seq[d] = seqId;
    r_strand[d] = (int8_t)(r.votes < 0 ? -1 : 1); /* :534 */
    const bool bad = r.wpos < 0 || r.wpos_end < 0 || r.wpos == r.wpos_end; /* :523-528 */
    const int64_t len = (int64_t)r.wpos_end - (int64_t)r.wpos;
    uint32_t k = 0, p = 0;
    if (!bad) {
      if (r.wpos_end > r.wpos + w) p = (uint32_t)ceilf((float)(r.wpos_end - r.wpos) / (float)w); /* :536-537 */
      else k = 1;
      (void)len;
    }
    keep[d] = k; pieces[d] = p;
  }
}
/* kept records first (emission order), then all pieces (parent's emission order, piece index): the vector the reference sorts */
__global__ void k_scatter_records(uint64_t n_raw, const uint64_t *__restrict__ r_hash, const int32_t *__restrict__ r_wpos,
                                  const int32_t *__restrict__ r_wend, const int32_t *__restrict__ r_seq, const int8_t *__restrict__ r_strand,
                                  const uint32_t *__restrict__ keep, const uint32_t *__restrict__ pieces,
                                  const uint64_t *__restrict__ keep_off, const uint64_t *__restrict__ piece_off, uint64_t n_keep, int w,
                                  uint64_t *o_hash, int32_t *o_wpos, int32_t *o_wend, int32_t *o_seq, int8_t *o_strand)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_raw) return;
  if (keep[i]) {
    const uint64_t d = keep_off[i];
    o_hash[d] = r_hash[i]; o_wpos[d] = r_wpos[i]; o_wend[d] = r_wend[i]; o_seq[d] = r_seq[i]; o_strand[d] = r_strand[i];
  }
  const uint32_t p = pieces[i];
  if (p) {
    const uint64_t d0 = n_keep + piece_off[i];
    const int32_t a = r_wpos[i], e = r_wend[i];
    for (uint32_t c = 0; c < p; c++) { /* :538-553 */
      const uint64_t d = d0 + c;
      o_hash[d] = r_hash[i]; o_seq[d] = r_seq[i]; o_strand[d] = r_strand[i];
      o_wpos[d] = a + (int32_t)c * w;
      const int32_t hi = a + (int32_t)c * w + w;
      o_wend[d] = hi < e ? hi : e;
    }
  }
}
__global__ void k_iota_keys32(uint64_t n, const int32_t *__restrict__ src, uint32_t *keys, uint32_t *vals)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = (uint32_t)src[i]; vals[i] = (uint32_t)i; }
}
__global__ void k_keys_seq_wpos(uint64_t n, const uint32_t *__restrict__ perm, const int32_t *__restrict__ seq, const int32_t *__restrict__ wpos,
                                uint64_t *keys)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const uint32_t j = perm[i]; keys[i] = ((uint64_t)(uint32_t)seq[j] << 32) | (uint64_t)(uint32_t)wpos[j]; }
}
/* gather in sorted order and flag the records std::unique keeps (:563-568: same wpos and hash as the one before, per contig) */
__global__ void k_gather_sorted(uint64_t n, const uint32_t *__restrict__ perm, const uint64_t *__restrict__ i_hash,
                                const int32_t *__restrict__ i_wpos, const int32_t *__restrict__ i_wend, const int32_t *__restrict__ i_seq,
                                const int8_t *__restrict__ i_strand, uint64_t *o_hash, int32_t *o_wpos, int32_t *o_wend, int32_t *o_seq,
                                int8_t *o_strand, uint32_t *uniq)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t j = perm[i];
  o_hash[i] = i_hash[j]; o_wpos[i] = i_wpos[j]; o_wend[i] = i_wend[j]; o_seq[i] = i_seq[j]; o_strand[i] = i_strand[j];
  uint32_t u = 1;
  if (i > 0) {
    const uint32_t p = perm[i - 1];
    if (i_seq[p] == i_seq[j] && i_wpos[p] == i_wpos[j] && i_hash[p] == i_hash[j]) u = 0;
  }
  uniq[i] = u;
}
__global__ void k_compact5(uint64_t n, const uint32_t *__restrict__ flag, const uint64_t *__restrict__ off, const uint64_t *__restrict__ i_hash,
                           const int32_t *__restrict__ i_wpos, const int32_t *__restrict__ i_wend, const int32_t *__restrict__ i_seq,
                           const int8_t *__restrict__ i_strand, uint64_t *o_hash, int32_t *o_wpos, int32_t *o_wend, int32_t *o_seq,
                           int8_t *o_strand)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  const uint64_t d = off[i];
  o_hash[d] = i_hash[i]; o_wpos[d] = i_wpos[i]; o_wend[d] = i_wend[i]; o_seq[d] = i_seq[i]; o_strand[d] = i_strand[i];
}

/* ---- Sketch::index (:379-404) ------------------------------------------------------------------------------------------
 * In hash-sorted order (stable: index order inside a hash) a record opens a new interval unless the previous record of the
 * same hash ends exactly where it starts (then the CLOSE point moves to its end; seqId is not compared, as in the reference). */
__global__ void k_iota_keys64(uint64_t n, const uint64_t *__restrict__ src, uint64_t *keys, uint32_t *vals)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = src[i]; vals[i] = (uint32_t)i; }
}
__global__ void k_lookup_flags(uint64_t n, const uint64_t *__restrict__ hs, const uint32_t *__restrict__ perm, const int32_t *__restrict__ wpos,
                               const int32_t *__restrict__ wend, uint32_t *key_start, uint32_t *run_start)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool ks = i == 0 || hs[i] != hs[i - 1];
  key_start[i] = ks ? 1u : 0u;
  run_start[i] = (ks || wend[perm[i - (i ? 1 : 0)]] != wpos[perm[i]]) ? 1u : 0u;
}
__global__ void k_lookup_emit(uint64_t n, const uint64_t *__restrict__ hs, const uint32_t *__restrict__ perm, const int32_t *__restrict__ wpos,
                              const int32_t *__restrict__ wend, const int32_t *__restrict__ seq, const uint32_t *__restrict__ key_start,
                              const uint32_t *__restrict__ run_start, const uint64_t *__restrict__ key_idx, const uint64_t *__restrict__ run_idx,
                              uint64_t *keys, uint64_t *offs, uint64_t *pts, uint64_t *rec_key)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t j = perm[i];
  /* exclusive scans of the start flags: at a start the index of the new run / key, elsewhere that index + 1 */
  const uint64_t r = run_start[i] ? run_idx[i] : run_idx[i] - 1;
  const uint64_t k = key_start[i] ? key_idx[i] : key_idx[i] - 1;
  rec_key[i] = k;
  if (key_start[i]) { keys[k] = hs[i]; offs[k] = 2 * r; }
  /* the run's OPEN point carries the seqId of its first record, and so does its CLOSE point (only its pos is moved, :397) */
  if (run_start[i]) pts[2 * r] = mm_pack_point(seq[j], wpos[j], 1);
  const bool last_of_run = i + 1 == n || run_start[i + 1];
  if (last_of_run) {
    /* first record of this run: walk back (runs are almost always one or two records long) */
    uint64_t f = i;
    while (!run_start[f]) f--;
    pts[2 * r + 1] = mm_pack_point(seq[perm[f]], wend[j], 0);
  }
}
__global__ void k_key_counts(uint64_t n_keys, const uint64_t *__restrict__ offs, uint64_t n_points, uint32_t *cnt)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_keys) cnt[i] = (uint32_t)((i + 1 < n_keys ? offs[i + 1] : n_points) - offs[i]);
}
__global__ void k_histogram(uint64_t n_keys, const uint32_t *__restrict__ cnt, unsigned long long *hist, uint32_t hist_n)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_keys) atomicAdd(&hist[cnt[i] < hist_n ? cnt[i] : hist_n - 1], 1ULL);
}
__global__ void k_mark_freq(uint64_t n_keys, const uint32_t *__restrict__ cnt, uint32_t threshold, uint8_t *is_freq)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_keys) is_freq[i] = cnt[i] >= threshold ? 1 : 0;
}
/* keep[index position] = the record's hash is not a frequent seed (dropFreqSeedSet :497-504) */
__global__ void k_keep_not_freq(uint64_t n, const uint32_t *__restrict__ perm, const uint64_t *__restrict__ rec_key, const uint8_t *__restrict__ is_freq,
                                uint32_t *keep)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[perm[i]] = is_freq[rec_key[i]] ? 0u : 1u;
}

struct Dev { /* frees what it allocated when it goes out of scope */
  std::vector<void *> p;
  ~Dev() { for (void *x : p) cudaFree(x); }
  template <typename T> cudaError_t alloc(T *&ptr, uint64_t n)
  {
    ptr = nullptr;
    cudaError_t e = cudaMalloc((void **)&ptr, std::max<uint64_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) p.push_back(ptr);
    return e;
  }
  void release(void *x) { p.erase(std::remove(p.begin(), p.end(), x), p.end()); }
  void free_now(void *x) { if (x) { cudaFree(x); release(x); } }
};

#define CE(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return e_ == cudaErrorMemoryAllocation ? MM_ENOMEM : MM_ECUDA; } \
  } while (0)

inline uint32_t blocks(uint64_t n, uint32_t per = 256) { return (uint32_t)((n + per - 1) / per); }

cudaError_t exclusive_sum_u32_to_u64(const uint32_t *in, uint64_t *out, uint64_t n, cudaStream_t st, Dev &dv)
{
  void *tmp = nullptr;
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, cub::TransformInputIterator<uint64_t, cub::CastOp<uint64_t>, const uint32_t *>(in, cub::CastOp<uint64_t>()), out, (int64_t)n, st);
  cudaError_t e = cudaMalloc(&tmp, bytes + 16);
  if (e != cudaSuccess) return e;
  e = cub::DeviceScan::ExclusiveSum(tmp, bytes, cub::TransformInputIterator<uint64_t, cub::CastOp<uint64_t>, const uint32_t *>(in, cub::CastOp<uint64_t>()), out, (int64_t)n, st);
  cudaStreamSynchronize(st);
  cudaFree(tmp);
  (void)dv;
  return e;
}
template <typename KeyT>
cudaError_t sort_pairs(KeyT *k_in, KeyT *k_out, uint32_t *v_in, uint32_t *v_out, uint64_t n, int end_bit, cudaStream_t st)
{
  void *tmp = nullptr;
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k_in, k_out, v_in, v_out, (int64_t)n, 0, end_bit, st);
  cudaError_t e = cudaMalloc(&tmp, bytes + 16);
  if (e != cudaSuccess) return e;
  e = cub::DeviceRadixSort::SortPairs(tmp, bytes, k_in, k_out, v_in, v_out, (int64_t)n, 0, end_bit, st);
  cudaStreamSynchronize(st);
  cudaFree(tmp);
  return e;
}

template <int K>
void launch_scan(const uint8_t *seq, const uint64_t *off, const wb_chunk *chunks, uint32_t n_chunks, int w, int s, int warm, unsigned char *slabs,
                 const wb_slab_layout &L, wm_record *rec, uint32_t rec_cap, wb_chunk_out *outs, wb_open *ex, uint32_t stride, uint32_t grid,
                 cudaStream_t st)
{
  k_window_scan<K><<<grid, 128, 0, st>>>(seq, off, chunks, n_chunks, w, s, warm, slabs, L, rec, rec_cap, outs, ex, stride);
}
template <int K>
void launch_fix(const uint8_t *seq, const uint64_t *off, const wb_chunk *chunks, const wb_chain *chains, uint32_t n_chains, int w, int s, int warm,
                unsigned char *slabs, const wb_slab_layout &L, wm_record *fix, uint32_t fix_cap, wb_chunk_out *outs, wb_open *ex, uint32_t stride,
                cudaStream_t st)
{
  k_window_fix<K><<<(n_chains + 63) / 64, 64, 0, st>>>(seq, off, chunks, chains, n_chains, w, s, warm, slabs, L, fix, fix_cap, outs, ex, stride);
}

} // namespace

#define MM_FOR_EACH_K(X) \
  X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) \
  X(28) X(29) X(30) X(31) X(32)

void mm_built_index_free(mm_built_index *b)
{
  if (!b) return;
  cudaFree(b->hash); cudaFree(b->wpos); cudaFree(b->wend); cudaFree(b->seq); cudaFree(b->strand);
  cudaFree(b->keys); cudaFree(b->offs); cudaFree(b->is_freq); cudaFree(b->pts);
  *b = mm_built_index{};
}

int mm_build_index_device(const mm_params &p, const uint8_t *d_seq, const uint64_t *h_contig_off, int32_t n_contigs,
                          float kmer_pct_threshold, cudaStream_t st, int sm_count, mm_built_index *out, std::string &err)
{
  *out = mm_built_index{};
  const int K = p.kmer_size, w = p.seg_length, s = p.sketch_size;
  if (!mm_sketch_kmer_supported(K)) { err = "k-mer size not compiled in"; return MM_EINVAL; }
  Dev dv;
  cudaEvent_t ev[4];
  for (auto &e : ev) cudaEventCreate(&e);
  cudaEventRecord(ev[0], st);

  /* ---- chunks ---- */
  const int warm = w + 2 * K + 64;
  int chunk_len = 16384;
  if (const char *e = getenv("MM_INDEX_CHUNK")) chunk_len = std::max(1024, atoi(e)); /* tests: small chunks */
  std::vector<wb_chunk> chunks;
  for (int32_t c = 0; c < n_contigs; c++) {
    const uint64_t len = h_contig_off[c + 1] - h_contig_off[c];
    if (len >= (1ULL << 31)) { err = "a contig is longer than 2^31 bases"; return MM_EINVAL; }
    const int32_t npos = (int32_t)len - K + 1;
    if (npos <= 0) continue;
    for (int32_t a = 0; a < npos; a += chunk_len) chunks.push_back(wb_chunk{c, a, std::min(npos, a + chunk_len), npos});
  }
  const uint32_t n_chunks = (uint32_t)chunks.size();
  out->n_chunks = n_chunks;
  uint64_t *d_off = nullptr;
  CE(dv.alloc(d_off, (uint64_t)n_contigs + 1));
  CE(cudaMemcpyAsync(d_off, h_contig_off, ((size_t)n_contigs + 1) * 8, cudaMemcpyHostToDevice, st));

  uint64_t n_raw = 0;
  uint64_t *r_hash = nullptr; int32_t *r_wpos = nullptr, *r_wend = nullptr, *r_seq = nullptr; int8_t *r_strand = nullptr;
  uint32_t *r_keep = nullptr, *r_pieces = nullptr;
  if (n_chunks) {
    wb_chunk *d_chunks = nullptr;
    CE(dv.alloc(d_chunks, n_chunks));
    CE(cudaMemcpyAsync(d_chunks, chunks.data(), (size_t)n_chunks * sizeof(wb_chunk), cudaMemcpyHostToDevice, st));
    const wb_slab_layout L = slab_layout(w, s);
    int tpsm = 768; /* machines per SM: the scan is latency-bound (dependent accesses to a per-thread slab), more threads hide more */
    if (const char *e = getenv("MM_INDEX_TPSM")) tpsm = std::max(128, atoi(e) / 128 * 128);
    uint32_t threads = (uint32_t)sm_count * (uint32_t)tpsm;
    if (threads > n_chunks) threads = (n_chunks + 127) / 128 * 128;
    const uint32_t grid = threads / 128;
    unsigned char *slabs = nullptr;
    CE(dv.alloc(slabs, (uint64_t)threads * L.bytes));
    const uint32_t rec_cap = (uint32_t)(chunk_len / 4 + s + 64);
    wm_record *rec = nullptr;
    CE(dv.alloc(rec, (uint64_t)n_chunks * rec_cap));
    wb_chunk_out *d_outs = nullptr;
    CE(dv.alloc(d_outs, n_chunks));
    const uint32_t stride = (uint32_t)wm_mem_cap(s);
    wb_open *d_ex = nullptr;
    CE(dv.alloc(d_ex, (uint64_t)n_chunks * stride));
    switch (K) {
#define X(KK) case KK: launch_scan<KK>(d_seq, d_off, d_chunks, n_chunks, w, s, warm, slabs, L, rec, rec_cap, d_outs, d_ex, stride, grid, st); break;
      MM_FOR_EACH_K(X)
#undef X
    }
    CE(cudaGetLastError());
    std::vector<wb_chunk_out> outs(n_chunks);
    CE(cudaMemcpyAsync(outs.data(), d_outs, (size_t)n_chunks * sizeof(wb_chunk_out), cudaMemcpyDeviceToHost, st));
    CE(cudaStreamSynchronize(st));

    /* chunks of each contig (for the sequential resolution of inherited record starts) */
    std::vector<uint32_t> cf, cnn;
    for (uint32_t c = 0; c < n_chunks; c++) {
      if (c == 0 || chunks[c].contig != chunks[c - 1].contig) { cf.push_back(c); cnn.push_back(0); }
      cnn.back()++;
    }
    uint32_t *d_cf = nullptr, *d_cn = nullptr, *d_err = nullptr;
    CE(dv.alloc(d_cf, cf.size())); CE(dv.alloc(d_cn, cf.size())); CE(dv.alloc(d_err, 1));
    CE(cudaMemcpyAsync(d_cf, cf.data(), cf.size() * 4, cudaMemcpyHostToDevice, st));
    CE(cudaMemcpyAsync(d_cn, cnn.data(), cf.size() * 4, cudaMemcpyHostToDevice, st));
    CE(cudaMemsetAsync(d_err, 0, 4, st));
    auto resolve = [&]() {
      k_resolve_exports<<<(uint32_t)cf.size(), 128, 0, st>>>(d_cf, d_cn, (uint32_t)cf.size(), d_outs, d_ex, stride, d_err);
      return cudaGetLastError();
    };

    /* ---- acceptance chain, fix-up rounds ----
     * A chunk is good if it starts a contig and ran clean, or if its predecessor is good, it ran clean (no failure, no
     * expired heap entry taken) and its state digest at its start equals its predecessor's at its end. Runs of chunks that
     * are not good form chains; a chain always starts right after a chunk that was accepted on its pass-1 (warm-up) run, is
     * re-scanned by one thread from that chunk's exact end state, and is extended and re-scanned as a whole if the chunk
     * that follows it does not match the chain's new end state. */
    std::vector<uint8_t> ok(n_chunks, 0);
    std::vector<int32_t> in_chain(n_chunks, -1);
    struct HostChain { uint32_t first, n; bool dirty; };
    std::vector<HostChain> hchains;
    wm_record *fix = nullptr;
    std::vector<uint64_t> fix_off(n_chunks, 0);
    uint64_t fix_used = 0, fix_capacity = 0;
    const uint32_t fix_cap = (uint32_t)(3 * chunk_len + s + 64); /* at most three records per position */
    for (int round = 0; round < 256; round++) {
      for (uint32_t c = 0; c < n_chunks; c++) {
        const bool first = chunks[c].a == 0;
        bool good;
        if (in_chain[c] >= 0) {
          if (outs[c].flags & 3u) { err = "window machine capacity exceeded while re-scanning a chunk"; return MM_ECAPACITY; }
          good = !hchains[(size_t)in_chain[c]].dirty;
        } else if (first) {
          good = !(outs[c].flags & 3u);
        } else {
          good = ok[c - 1] && !(outs[c].flags & 7u) && outs[c].d_start == outs[c - 1].d_end;
        }
        ok[c] = good ? 1 : 0;
        if (!good && in_chain[c] < 0) {
          if (!first && in_chain[c - 1] >= 0) { /* the chain before it grows by this chunk and is re-scanned as a whole */
            HostChain &hc = hchains[(size_t)in_chain[c - 1]];
            hc.n++; hc.dirty = true;
            in_chain[c] = in_chain[c - 1];
          } else {
            in_chain[c] = (int32_t)hchains.size();
            hchains.push_back(HostChain{c, 1, true});
          }
        }
      }
      std::vector<wb_chain> chains;
      for (auto &hc : hchains)
        if (hc.dirty) chains.push_back(wb_chain{hc.first, hc.n, 0});
      if (chains.empty()) break;
      out->fix_rounds = (uint32_t)round + 1;
      uint64_t need = 0;
      for (auto &cn : chains) { cn.out_offset = fix_used + need; need += (uint64_t)cn.n * fix_cap; }
      if ((fix_used + need) * sizeof(wm_record) > (48ULL << 30)) { err = "too many chunks need an exact re-scan (N-rich / low-complexity reference): use the host builder"; return MM_ECAPACITY; }
      if (fix_used + need > fix_capacity) { /* grow, keeping what earlier rounds wrote */
        wm_record *bigger = nullptr;
        const uint64_t cap2 = (fix_used + need) + (fix_used + need) / 2;
        CE(dv.alloc(bigger, cap2));
        if (fix_used) CE(cudaMemcpyAsync(bigger, fix, fix_used * sizeof(wm_record), cudaMemcpyDeviceToDevice, st));
        CE(cudaStreamSynchronize(st));
        dv.free_now(fix);
        fix = bigger; fix_capacity = cap2;
      }
      for (auto &cn : chains)
        for (uint32_t q = 0; q < cn.n; q++) fix_off[cn.first + q] = cn.out_offset + (uint64_t)q * fix_cap;
      fix_used += need;
      CE(resolve()); /* the re-scan takes record starts from the exports of the chunks before the chain */
      wb_chain *d_chains = nullptr;
      CE(dv.alloc(d_chains, chains.size()));
      CE(cudaMemcpyAsync(d_chains, chains.data(), chains.size() * sizeof(wb_chain), cudaMemcpyHostToDevice, st));
      unsigned char *fslabs = slabs;
      if (chains.size() > threads) { CE(dv.alloc(fslabs, (uint64_t)chains.size() * L.bytes)); }
      switch (K) {
#define X(KK) case KK: launch_fix<KK>(d_seq, d_off, d_chunks, d_chains, (uint32_t)chains.size(), w, s, warm, fslabs, L, fix, fix_cap, d_outs, d_ex, stride, st); break;
        MM_FOR_EACH_K(X)
#undef X
      }
      CE(cudaGetLastError());
      CE(cudaMemcpyAsync(outs.data(), d_outs, (size_t)n_chunks * sizeof(wb_chunk_out), cudaMemcpyDeviceToHost, st));
      CE(cudaStreamSynchronize(st));
      if (fslabs != slabs) dv.free_now(fslabs);
      dv.free_now(d_chains);
      for (auto &hc : hchains) hc.dirty = false;
    }
    out->n_fixed_chunks = 0;
    for (auto &hc : hchains) out->n_fixed_chunks += hc.n;
    for (uint32_t c = 0; c < n_chunks; c++)
      if (!ok[c]) { err = "chunk stitching did not converge"; return MM_ECUDA; }
    dv.free_now(slabs);

    CE(resolve());
    k_patch_starts<<<n_chunks, 128, 0, st>>>(d_chunks, d_outs, n_chunks, rec, rec_cap, d_ex, stride, d_err);
    CE(cudaGetLastError());
    uint32_t h_err = 0;
    CE(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st));
    CE(cudaStreamSynchronize(st));
    if (h_err) { err = "a record open at a chunk start is missing from the previous chunk's export (code " + std::to_string(h_err) + ")"; return MM_ECUDA; }
    dv.free_now(d_ex);

    /* ---- raw records in emission order ---- */
    std::vector<uint64_t> raw_off(n_chunks + 1, 0);
    for (uint32_t c = 0; c < n_chunks; c++) raw_off[c + 1] = raw_off[c] + outs[c].n_rec;
    n_raw = raw_off[n_chunks];
    uint64_t *d_raw_off = nullptr, *d_fix_off = nullptr;
    CE(dv.alloc(d_raw_off, (uint64_t)n_chunks + 1));
    CE(dv.alloc(d_fix_off, n_chunks));
    CE(cudaMemcpyAsync(d_raw_off, raw_off.data(), ((size_t)n_chunks + 1) * 8, cudaMemcpyHostToDevice, st));
    CE(cudaMemcpyAsync(d_fix_off, fix_off.data(), (size_t)n_chunks * 8, cudaMemcpyHostToDevice, st));
    CE(dv.alloc(r_hash, n_raw)); CE(dv.alloc(r_wpos, n_raw)); CE(dv.alloc(r_wend, n_raw)); CE(dv.alloc(r_seq, n_raw)); CE(dv.alloc(r_strand, n_raw));
    CE(dv.alloc(r_keep, n_raw)); CE(dv.alloc(r_pieces, n_raw));
    k_gather_raw<<<n_chunks, 256, 0, st>>>(d_chunks, d_outs, n_chunks, d_raw_off, rec, rec_cap, fix, d_fix_off, w, r_hash, r_wpos, r_wend, r_seq,
                                           r_strand, r_keep, r_pieces);
    CE(cudaGetLastError());
    CE(cudaStreamSynchronize(st));
    dv.free_now(rec); dv.free_now(fix); dv.free_now(d_outs); dv.free_now(d_chunks); dv.free_now(d_raw_off); dv.free_now(d_fix_off);
  }
  cudaEventRecord(ev[1], st);

  /* ---- kept records + pieces -> the vector the reference sorts ---- */
  uint64_t n_all = 0;
  uint64_t *a_hash = nullptr; int32_t *a_wpos = nullptr, *a_wend = nullptr, *a_seq = nullptr; int8_t *a_strand = nullptr;
  if (n_raw) {
    uint64_t *keep_off = nullptr, *piece_off = nullptr;
    CE(dv.alloc(keep_off, n_raw + 1)); CE(dv.alloc(piece_off, n_raw + 1));
    CE(exclusive_sum_u32_to_u64(r_keep, keep_off, n_raw, st, dv));
    CE(exclusive_sum_u32_to_u64(r_pieces, piece_off, n_raw, st, dv));
    uint64_t last_k = 0, last_p = 0;
    uint32_t lk = 0, lp = 0;
    CE(cudaMemcpy(&last_k, keep_off + n_raw - 1, 8, cudaMemcpyDeviceToHost)); CE(cudaMemcpy(&lk, r_keep + n_raw - 1, 4, cudaMemcpyDeviceToHost));
    CE(cudaMemcpy(&last_p, piece_off + n_raw - 1, 8, cudaMemcpyDeviceToHost)); CE(cudaMemcpy(&lp, r_pieces + n_raw - 1, 4, cudaMemcpyDeviceToHost));
    const uint64_t n_keep = last_k + lk, n_pieces = last_p + lp;
    n_all = n_keep + n_pieces;
    if (n_all >= (1ULL << 32)) { err = "more than 2^32 minmer records"; return MM_EINVAL; }
    CE(dv.alloc(a_hash, n_all)); CE(dv.alloc(a_wpos, n_all)); CE(dv.alloc(a_wend, n_all)); CE(dv.alloc(a_seq, n_all)); CE(dv.alloc(a_strand, n_all));
    k_scatter_records<<<blocks(n_raw), 256, 0, st>>>(n_raw, r_hash, r_wpos, r_wend, r_seq, r_strand, r_keep, r_pieces, keep_off, piece_off, n_keep, w,
                                                     a_hash, a_wpos, a_wend, a_seq, a_strand);
    CE(cudaGetLastError());
    CE(cudaStreamSynchronize(st));
    dv.free_now(keep_off); dv.free_now(piece_off);
    dv.free_now(r_hash); dv.free_now(r_wpos); dv.free_now(r_wend); dv.free_now(r_seq); dv.free_now(r_strand); dv.free_now(r_keep); dv.free_now(r_pieces);
  }

  /* ---- order by (seqId, wpos, wpos_end), stable; adjacent de-duplication ---- */
  uint64_t n_mi = 0;
  uint64_t *m_hash = nullptr; int32_t *m_wpos = nullptr, *m_wend = nullptr, *m_seq = nullptr; int8_t *m_strand = nullptr;
  if (n_all) {
    uint32_t *k32a = nullptr, *k32b = nullptr, *va = nullptr, *vb = nullptr;
    CE(dv.alloc(k32a, n_all)); CE(dv.alloc(k32b, n_all)); CE(dv.alloc(va, n_all)); CE(dv.alloc(vb, n_all));
    k_iota_keys32<<<blocks(n_all), 256, 0, st>>>(n_all, a_wend, k32a, va);
    CE(sort_pairs<uint32_t>(k32a, k32b, va, vb, n_all, 32, st));   /* by wpos_end */
    dv.free_now(k32a); dv.free_now(k32b);
    uint64_t *k64a = nullptr, *k64b = nullptr;
    CE(dv.alloc(k64a, n_all)); CE(dv.alloc(k64b, n_all));
    k_keys_seq_wpos<<<blocks(n_all), 256, 0, st>>>(n_all, vb, a_seq, a_wpos, k64a);
    int end_bit = 32;
    while ((1LL << (end_bit - 32)) < (long long)n_contigs + 1) end_bit++;
    CE(sort_pairs<uint64_t>(k64a, k64b, vb, va, n_all, end_bit, st)); /* then by (seqId, wpos): LSD, stable */
    dv.free_now(k64a); dv.free_now(k64b); dv.free_now(vb);
    uint64_t *s_hash = nullptr; int32_t *s_wpos = nullptr, *s_wend = nullptr, *s_seq = nullptr; int8_t *s_strand = nullptr; uint32_t *uniq = nullptr;
    CE(dv.alloc(s_hash, n_all)); CE(dv.alloc(s_wpos, n_all)); CE(dv.alloc(s_wend, n_all)); CE(dv.alloc(s_seq, n_all)); CE(dv.alloc(s_strand, n_all));
    CE(dv.alloc(uniq, n_all));
    k_gather_sorted<<<blocks(n_all), 256, 0, st>>>(n_all, va, a_hash, a_wpos, a_wend, a_seq, a_strand, s_hash, s_wpos, s_wend, s_seq, s_strand, uniq);
    CE(cudaGetLastError());
    CE(cudaStreamSynchronize(st));
    dv.free_now(va);
    dv.free_now(a_hash); dv.free_now(a_wpos); dv.free_now(a_wend); dv.free_now(a_seq); dv.free_now(a_strand);
    uint64_t *uoff = nullptr;
    CE(dv.alloc(uoff, n_all + 1));
    CE(exclusive_sum_u32_to_u64(uniq, uoff, n_all, st, dv));
    uint64_t lo = 0; uint32_t lu = 0;
    CE(cudaMemcpy(&lo, uoff + n_all - 1, 8, cudaMemcpyDeviceToHost)); CE(cudaMemcpy(&lu, uniq + n_all - 1, 4, cudaMemcpyDeviceToHost));
    n_mi = lo + lu;
    CE(dv.alloc(m_hash, n_mi)); CE(dv.alloc(m_wpos, n_mi)); CE(dv.alloc(m_wend, n_mi)); CE(dv.alloc(m_seq, n_mi)); CE(dv.alloc(m_strand, n_mi));
    k_compact5<<<blocks(n_all), 256, 0, st>>>(n_all, uniq, uoff, s_hash, s_wpos, s_wend, s_seq, s_strand, m_hash, m_wpos, m_wend, m_seq, m_strand);
    CE(cudaGetLastError());
    CE(cudaStreamSynchronize(st));
    dv.free_now(uoff); dv.free_now(uniq);
    dv.free_now(s_hash); dv.free_now(s_wpos); dv.free_now(s_wend); dv.free_now(s_seq); dv.free_now(s_strand);
  }
  out->n_minmers_before_filter = n_mi;
  cudaEventRecord(ev[2], st);

  /* ---- Sketch::index + frequency filter ---- */
  uint64_t n_keys = 0, n_points = 0;
  uint64_t *keys = nullptr, *offs = nullptr, *pts = nullptr; uint8_t *is_freq = nullptr;
  int32_t threshold = 0x7fffffff;
  uint64_t n_final = 0;
  uint64_t *f_hash = nullptr; int32_t *f_wpos = nullptr, *f_wend = nullptr, *f_seq = nullptr; int8_t *f_strand = nullptr;
  if (n_mi) {
    uint64_t *hk = nullptr, *hs = nullptr; uint32_t *va = nullptr, *perm = nullptr;
    CE(dv.alloc(hk, n_mi)); CE(dv.alloc(hs, n_mi)); CE(dv.alloc(va, n_mi)); CE(dv.alloc(perm, n_mi));
    k_iota_keys64<<<blocks(n_mi), 256, 0, st>>>(n_mi, m_hash, hk, va);
    CE(sort_pairs<uint64_t>(hk, hs, va, perm, n_mi, 64, st));
    dv.free_now(hk); dv.free_now(va);
    uint32_t *key_start = nullptr, *run_start = nullptr; uint64_t *key_idx = nullptr, *run_idx = nullptr, *rec_key = nullptr;
    CE(dv.alloc(key_start, n_mi)); CE(dv.alloc(run_start, n_mi)); CE(dv.alloc(key_idx, n_mi + 1)); CE(dv.alloc(run_idx, n_mi + 1)); CE(dv.alloc(rec_key, n_mi));
    k_lookup_flags<<<blocks(n_mi), 256, 0, st>>>(n_mi, hs, perm, m_wpos, m_wend, key_start, run_start);
    CE(exclusive_sum_u32_to_u64(key_start, key_idx, n_mi, st, dv));
    CE(exclusive_sum_u32_to_u64(run_start, run_idx, n_mi, st, dv));
    uint64_t lk = 0, lr = 0; uint32_t fk = 0, fr = 0;
    CE(cudaMemcpy(&lk, key_idx + n_mi - 1, 8, cudaMemcpyDeviceToHost)); CE(cudaMemcpy(&fk, key_start + n_mi - 1, 4, cudaMemcpyDeviceToHost));
    CE(cudaMemcpy(&lr, run_idx + n_mi - 1, 8, cudaMemcpyDeviceToHost)); CE(cudaMemcpy(&fr, run_start + n_mi - 1, 4, cudaMemcpyDeviceToHost));
    n_keys = lk + fk;
    n_points = 2 * (lr + fr);
    /* exclusive scans give, at a start flag, the index of the new key / run; at other records index + 1 of the current one */
    CE(dv.alloc(keys, n_keys)); CE(dv.alloc(offs, n_keys + 1)); CE(dv.alloc(pts, n_points + 1)); CE(dv.alloc(is_freq, n_keys));
    k_lookup_emit<<<blocks(n_mi), 256, 0, st>>>(n_mi, hs, perm, m_wpos, m_wend, m_seq, key_start, run_start, key_idx, run_idx, keys, offs, pts, rec_key);
    CE(cudaGetLastError());
    CE(cudaMemcpyAsync(offs + n_keys, &n_points, 8, cudaMemcpyHostToDevice, st));
    CE(cudaStreamSynchronize(st));
    dv.free_now(key_start); dv.free_now(run_start); dv.free_now(key_idx); dv.free_now(run_idx); dv.free_now(hs);
    /* histogram of interval points per key (winSketch.hpp:415-417) */
    uint32_t *cnt = nullptr;
    CE(dv.alloc(cnt, n_keys));
    k_key_counts<<<blocks(n_keys), 256, 0, st>>>(n_keys, offs, n_points, cnt);
    uint32_t max_cnt = 0;
    {
      uint32_t *d_max = nullptr;
      CE(dv.alloc(d_max, 1));
      void *tmp = nullptr; size_t bytes = 0;
      cub::DeviceReduce::Max(nullptr, bytes, cnt, d_max, (int64_t)n_keys, st);
      CE(cudaMalloc(&tmp, bytes + 16));
      cub::DeviceReduce::Max(tmp, bytes, cnt, d_max, (int64_t)n_keys, st);
      CE(cudaMemcpyAsync(&max_cnt, d_max, 4, cudaMemcpyDeviceToHost, st));
      CE(cudaStreamSynchronize(st));
      cudaFree(tmp);
      dv.free_now(d_max);
    }
    const uint32_t hist_n = max_cnt + 2;
    unsigned long long *d_hist = nullptr;
    CE(dv.alloc(d_hist, hist_n));
    CE(cudaMemsetAsync(d_hist, 0, (size_t)hist_n * 8, st));
    k_histogram<<<blocks(n_keys), 256, 0, st>>>(n_keys, cnt, d_hist, hist_n);
    std::vector<unsigned long long> hist(hist_n);
    CE(cudaMemcpyAsync(hist.data(), d_hist, (size_t)hist_n * 8, cudaMemcpyDeviceToHost, st));
    CE(cudaStreamSynchronize(st));
    dv.free_now(d_hist);
    { /* computeFreqHist :431-441, the same arithmetic (int64 * float / 100 -> int64; walk from the most frequent) */
      const int64_t totalUniqueMinmers = (int64_t)n_keys;
      const int64_t minmerToIgnore = totalUniqueMinmers * kmer_pct_threshold / 100;
      int64_t sum = 0;
      for (int64_t f = (int64_t)hist_n - 1; f >= 0; f--) {
        if (!hist[(size_t)f]) continue;
        sum += (int64_t)hist[(size_t)f];
        if (sum < minmerToIgnore) threshold = (int32_t)f;
        else if (sum == minmerToIgnore) { threshold = (int32_t)f; break; }
        else break;
      }
      out->hist_min_count = 0; out->hist_max_count = max_cnt;
      for (uint32_t f = 0; f < hist_n; f++) if (hist[f]) { out->hist_min_count = f; out->hist_min_keys = hist[f]; break; }
      out->hist_max_keys = hist[max_cnt];
    }
    k_mark_freq<<<blocks(n_keys), 256, 0, st>>>(n_keys, cnt, (uint32_t)threshold, is_freq);
    dv.free_now(cnt);
    /* dropFreqSeedSet: the frequent hashes leave minmerIndex (not the lookup) */
    uint32_t *keep = nullptr; uint64_t *koff = nullptr;
    CE(dv.alloc(keep, n_mi)); CE(dv.alloc(koff, n_mi + 1));
    k_keep_not_freq<<<blocks(n_mi), 256, 0, st>>>(n_mi, perm, rec_key, is_freq, keep);
    CE(exclusive_sum_u32_to_u64(keep, koff, n_mi, st, dv));
    uint64_t lo = 0; uint32_t lu = 0;
    CE(cudaMemcpy(&lo, koff + n_mi - 1, 8, cudaMemcpyDeviceToHost)); CE(cudaMemcpy(&lu, keep + n_mi - 1, 4, cudaMemcpyDeviceToHost));
    n_final = lo + lu;
    dv.free_now(perm); dv.free_now(rec_key);
    CE(dv.alloc(f_hash, n_final)); CE(dv.alloc(f_wpos, n_final)); CE(dv.alloc(f_wend, n_final)); CE(dv.alloc(f_seq, n_final)); CE(dv.alloc(f_strand, n_final));
    k_compact5<<<blocks(n_mi), 256, 0, st>>>(n_mi, keep, koff, m_hash, m_wpos, m_wend, m_seq, m_strand, f_hash, f_wpos, f_wend, f_seq, f_strand);
    CE(cudaGetLastError());
    CE(cudaStreamSynchronize(st));
    dv.free_now(keep); dv.free_now(koff);
    dv.free_now(m_hash); dv.free_now(m_wpos); dv.free_now(m_wend); dv.free_now(m_seq); dv.free_now(m_strand);
  }
  cudaEventRecord(ev[3], st);
  CE(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&out->ms_scan, ev[0], ev[1]);
  cudaEventElapsedTime(&out->ms_post, ev[1], ev[2]);
  cudaEventElapsedTime(&out->ms_lookup, ev[2], ev[3]);
  for (auto &e : ev) cudaEventDestroy(e);

  out->n_minmers = n_final; out->n_keys = n_keys; out->n_points = n_points; out->freq_threshold = threshold;
  out->hash = f_hash; out->wpos = f_wpos; out->wend = f_wend; out->seq = f_seq; out->strand = f_strand;
  out->keys = keys; out->offs = offs; out->is_freq = is_freq; out->pts = pts;
  for (void *x : {(void *)f_hash, (void *)f_wpos, (void *)f_wend, (void *)f_seq, (void *)f_strand, (void *)keys, (void *)offs, (void *)is_freq, (void *)pts})
    dv.release(x); /* now owned by *out (mm_built_index_free) */
  return MM_OK;
}
