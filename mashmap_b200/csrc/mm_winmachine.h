/*
 * mm_winmachine.h -- the sliding-window minmer machine of the reference index, as a fixed-capacity state machine that one
 * GPU thread (mm_index_build.cu) or the host (tests, the CPU side of skch::Sketch) can run.
 *
 * What it computes: CommonFunc::addMinmers (reference src/map/include/commonFunc.hpp:301-570), the per-position part
 * (:337-506): for every window start p of a contig, the s smallest distinct canonical k-mer hashes of the window
 * [p, p + w - k]; one record (hash, wpos, wpos_end, strand-vote sum) is emitted whenever a hash leaves that set (:386-391,
 * :469-472) or its strand-vote sum passes through zero (:395-402, :429-436). Every record's wpos is a point where the L2
 * stage evaluates a window (computeMap.hpp:1340-1376), so the records have to be the reference's, including what its
 * bookkeeping does in corner cases: an evicted member's occurrence AT the window start is not handed
 * back to the waiting heap (:478 `>`), the refill loop reads the top of a heap it may just have emptied (:495), at most
 * one member is evicted per position (:466 `if`), expired heap entries are only purged past 2w entries (:344).
 *
 * How: not the reference's containers (std::map of std::deque, std::vector heap, std::deque) but four flat arrays in a
 * slab the caller provides -- a ring of the window's valid k-mers, the members sorted by hash with their occurrence lists
 * threaded through a node pool, a binary min-heap on (hash, position) of the waiting k-mers -- so the state has a fixed
 * footprint, can be digested and compared (chunks of a contig are processed in parallel and stitched, see
 * mm_index_build.cu), and needs no allocation. The order in which equal-priority work is done never depends on the
 * containers' internals: (hash, position) is a total order, so any correct heap pops the same sequence.
 */
#ifndef MM_WINMACHINE_H
#define MM_WINMACHINE_H

#include <stdint.h>

#include "mm_hash.h"

#if defined(__CUDACC__)
#define WM_HD __host__ __device__ __forceinline__
#else
#define WM_HD inline
#endif

struct wm_kmer {  /* a valid k-mer: canonical hash, position, strand vote (+1 forward hash smaller, -1 reverse) */
  uint64_t hash;
  int32_t pos;
  int32_t strand;
};
struct wm_member {  /* a member of the window's sketch (reference: MinmerInfo + deque<KmerInfo>, commonFunc.hpp:318) */
  uint64_t hash;
  int32_t wpos;    /* start of the record that is open for this hash; -1 = none */
  int32_t votes;   /* running strand-vote sum (MinmerInfo::strand during the scan) */
  int32_t head, tail, count; /* occurrence list in the node pool, oldest first */
  uint32_t inherited;        /* the open record started before this machine's official start (stitching) */
};
struct wm_node {
  int32_t pos, strand, next;
};
struct wm_record {  /* one emitted record, before the post-processing of :522-568 */
  uint64_t hash;
  int32_t wpos, wpos_end;
  int32_t votes;      /* vote sum at emission (:534 turns it into FWD / REV) */
  uint32_t inherited; /* wpos is the machine's warm-up value: to be replaced by the predecessor chunk's */
};

struct wm_machine {
  /* parameters */
  int32_t k, w, s;
  /* storage (caller-provided) */
  wm_kmer *ring; int32_t ring_cap, ring_head, ring_n;
  /* the members: hashes ascending in mh[], mslot[i] = payload slot of the i-th smallest (the payloads stay where they are:
   * a member entering or leaving the sketch moves 10 bytes per larger member, not a whole record) */
  uint64_t *mh; uint16_t *mslot; wm_member *slots; uint16_t *sfree; int32_t sfree_n;
  int32_t mem_n, mem_cap;
  wm_node *nodes; int32_t node_cap, node_free;
  wm_kmer *heap; int32_t heap_n, heap_cap;
  wm_record *out; uint64_t out_n, out_cap;
  int32_t ambig;       /* bases until the window is free of N again (:413-416, :447-450) */
  int32_t emit_from;   /* records are kept only once the scan has reached this position (warm-up of a chunk) */
  /* diagnostics */
  int32_t fail;        /* 1 = a capacity was exceeded or the reference would have dereferenced end(): results unusable */
  int32_t drained;     /* a refill took an expired / already popped heap entry: from here on the state may depend on history
                          older than the window (see wm_step) */
};

WM_HD bool wm_heap_after(const wm_kmer &a, const wm_kmer &b)
{ /* KIHeap_cmp (:325-326): a sorts after b */
  return a.hash > b.hash || (a.hash == b.hash && a.pos > b.pos);
}
WM_HD void wm_heap_push(wm_machine &m, const wm_kmer &x)
{
  if (m.heap_n >= m.heap_cap) { m.fail = 1; return; }
  int32_t i = m.heap_n++;
  while (i > 0) {
    const int32_t p = (i - 1) >> 1;
    if (!wm_heap_after(m.heap[p], x)) break;
    m.heap[i] = m.heap[p];
    i = p;
  }
  m.heap[i] = x;
}
/* removes the top. Like std::pop_heap + pop_back on a vector, the removed element stays readable at index 0 when the heap
 * becomes empty (the reference reads it there, :495) */
WM_HD void wm_heap_pop(wm_machine &m)
{
  const int32_t n = --m.heap_n;
  if (n <= 0) return; /* single element: it stays in slot 0 */
  const wm_kmer x = m.heap[n];
  int32_t i = 0;
  while (true) {
    int32_t c = 2 * i + 1;
    if (c >= n) break;
    if (c + 1 < n && wm_heap_after(m.heap[c], m.heap[c + 1])) c++;
    if (!wm_heap_after(x, m.heap[c])) break;
    m.heap[i] = m.heap[c];
    i = c;
  }
  m.heap[i] = x;
}
WM_HD void wm_heap_purge(wm_machine &m, int32_t wid)
{ /* :344-354: drop expired entries, rebuild */
  int32_t n = 0;
  for (int32_t i = 0; i < m.heap_n; i++)
    if (!(m.heap[i].pos < wid)) m.heap[n++] = m.heap[i];
  m.heap_n = n;
  for (int32_t start = n / 2 - 1; start >= 0; start--) { /* sift down from the last parent */
    const wm_kmer x = m.heap[start];
    int32_t i = start;
    while (true) {
      int32_t c = 2 * i + 1;
      if (c >= n) break;
      if (c + 1 < n && wm_heap_after(m.heap[c], m.heap[c + 1])) c++;
      if (!wm_heap_after(x, m.heap[c])) break;
      m.heap[i] = m.heap[c];
      i = c;
    }
    m.heap[i] = x;
  }
}

/* the i-th smallest member */
WM_HD wm_member &wm_at(const wm_machine &m, int32_t i) { return m.slots[m.mslot[i]]; }
/* index of the first member with hash >= h */
WM_HD int32_t wm_lower_bound(const wm_machine &m, uint64_t h)
{
  int32_t lo = 0, hi = m.mem_n;
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (m.mh[mid] < h) lo = mid + 1; else hi = mid;
  }
  return lo;
}
WM_HD int32_t wm_find(const wm_machine &m, uint64_t h)
{
  const int32_t i = wm_lower_bound(m, h);
  return (i < m.mem_n && m.mh[i] == h) ? i : -1;
}
WM_HD void wm_list_push_back(wm_machine &m, wm_member &e, int32_t pos, int32_t strand)
{
  const int32_t n = m.node_free;
  if (n < 0) { m.fail = 1; return; }
  m.node_free = m.nodes[n].next;
  m.nodes[n].pos = pos; m.nodes[n].strand = strand; m.nodes[n].next = -1;
  if (e.count == 0) e.head = n; else m.nodes[e.tail].next = n;
  e.tail = n;
  e.count++;
}
WM_HD void wm_list_pop_front(wm_machine &m, wm_member &e)
{
  if (e.count == 0) return; /* std::deque::pop_front on an empty deque is undefined; the reference never gets here */
  const int32_t n = e.head;
  e.head = m.nodes[n].next;
  m.nodes[n].next = m.node_free;
  m.node_free = n;
  e.count--;
}
WM_HD void wm_list_free(wm_machine &m, wm_member &e)
{
  while (e.count > 0) wm_list_pop_front(m, e);
}
/* The sorted member arrays live in a per-machine slab (global memory on the device). An element-by-element shift is a
 * chain of load -> store -> load on one cache line (the store invalidates the line the next load needs): moved 8 at a
 * time, all loads of a batch first, the shift costs one memory round trip per 8 members instead of one per member (this
 * loop was 64 % of the window-scan kernel's stall samples). */
WM_HD void wm_members_shift_down(wm_machine &m, int32_t idx) /* [idx+1, mem_n) -> [idx, mem_n-1) */
{
  int32_t i = idx;
  while (i + 8 < m.mem_n) {
    uint64_t h[8]; uint16_t sl[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { h[u] = m.mh[i + 1 + u]; sl[u] = m.mslot[i + 1 + u]; }
#pragma unroll
    for (int u = 0; u < 8; u++) { m.mh[i + u] = h[u]; m.mslot[i + u] = sl[u]; }
    i += 8;
  }
  for (; i + 1 < m.mem_n; i++) { m.mh[i] = m.mh[i + 1]; m.mslot[i] = m.mslot[i + 1]; }
}
WM_HD void wm_members_shift_up(wm_machine &m, int32_t at) /* [at, mem_n) -> [at+1, mem_n] */
{
  int32_t i = m.mem_n;
  while (i - at >= 8) {
    uint64_t h[8]; uint16_t sl[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { h[u] = m.mh[i - 1 - u]; sl[u] = m.mslot[i - 1 - u]; }
#pragma unroll
    for (int u = 0; u < 8; u++) { m.mh[i - u] = h[u]; m.mslot[i - u] = sl[u]; }
    i -= 8;
  }
  for (; i > at; i--) { m.mh[i] = m.mh[i - 1]; m.mslot[i] = m.mslot[i - 1]; }
}
WM_HD void wm_erase_member(wm_machine &m, int32_t idx)
{
  wm_list_free(m, wm_at(m, idx));
  m.sfree[m.sfree_n++] = m.mslot[idx];
  wm_members_shift_down(m, idx);
  m.mem_n--;
}
/* a new member at its place in hash order; returns its index (or -1) */
WM_HD int32_t wm_insert_member(wm_machine &m, uint64_t h)
{
  if (m.mem_n >= m.mem_cap || m.sfree_n <= 0) { m.fail = 1; return -1; }
  const int32_t at = wm_lower_bound(m, h);
  wm_members_shift_up(m, at);
  m.mem_n++;
  m.mh[at] = h;
  m.mslot[at] = m.sfree[--m.sfree_n];
  wm_member &e = wm_at(m, at);
  e.hash = h; e.wpos = -1; e.votes = 0; e.head = -1; e.tail = -1; e.count = 0; e.inherited = 0;
  return at;
}
WM_HD void wm_emit(wm_machine &m, wm_member &e, int32_t wpos_end, int32_t scan_pos)
{
  if (scan_pos >= m.emit_from) {
    if (m.out_n >= m.out_cap) { m.fail = 1; }
    else {
      wm_record &r = m.out[m.out_n++];
      r.hash = e.hash; r.wpos = e.wpos; r.wpos_end = wpos_end; r.votes = e.votes; r.inherited = e.inherited;
    }
  }
  e.inherited = 0;
}

WM_HD void wm_init(wm_machine &m, int32_t k, int32_t w, int32_t s)
{
  m.k = k; m.w = w; m.s = s;
  m.ring_head = 0; m.ring_n = 0; m.mem_n = 0; m.heap_n = 0; m.out_n = 0; m.ambig = 0; m.emit_from = 0;
  m.fail = 0; m.drained = 0;
  for (int32_t i = 0; i < m.mem_cap; i++) m.sfree[i] = (uint16_t)(m.mem_cap - 1 - i);
  m.sfree_n = m.mem_cap;
  for (int32_t i = 0; i < m.node_cap; i++) m.nodes[i].next = i + 1 < m.node_cap ? i + 1 : -1;
  m.node_free = m.node_cap > 0 ? 0 : -1;
  if (m.heap_cap > 0) { m.heap[0].hash = 0; m.heap[0].pos = 0; m.heap[0].strand = 0; }
}

/* capacities a slab must provide for (w, s) */
WM_HD int32_t wm_ring_cap(int32_t w) { return w + 8; }
WM_HD int32_t wm_node_cap(int32_t w) { return w + w / 2 + 64; } /* occurrences of members: at most the window, plus stale ones */
WM_HD int32_t wm_heap_cap(int32_t w) { return 3 * w + 64; }      /* purged past 2w (:344); a position adds at most a few entries */
WM_HD int32_t wm_mem_cap(int32_t s) { return s + 4; }

/*
 * One position of the scan (:337-506). i = k-mer position; hash_fwd / hash_bwd = MurmurHash3 of the k-mer and of its
 * reverse complement; last_base_is_n = seq[i + k - 1] == 'N' after normalisation.
 */
WM_HD void wm_step(wm_machine &m, int32_t i, uint64_t hash_fwd, uint64_t hash_bwd, bool last_base_is_n)
{
  const int32_t wid = i + m.k - m.w; /* currentWindowId (:340) */
  if (m.heap_n > 2 * m.w) wm_heap_purge(m, wid);

  const uint64_t cur = hash_fwd < hash_bwd ? hash_fwd : hash_bwd;
  const int32_t cur_strand = hash_fwd < hash_bwd ? 1 : -1;

  /* the k-mer that left the window (:376-410) */
  if (m.ring_n > 0 && m.ring[m.ring_head].pos < wid) {
    const uint64_t lh = m.ring[m.ring_head].hash;
    const int32_t ls = m.ring[m.ring_head].strand;
    if (m.mem_n > 0 && lh <= m.mh[m.mem_n - 1]) {
      const int32_t idx = wm_find(m, lh);
      if (idx < 0) {
        m.fail = 1; /* the reference dereferences end() here */
      } else {
        wm_member &e = wm_at(m, idx);
        if (e.count == 1) {
          wm_emit(m, e, wid, i);
          wm_erase_member(m, idx);
        } else {
          if (e.votes - ls == 0 || e.votes == 0) {
            wm_emit(m, e, wid, i);
            e.wpos = wid;
          }
          e.votes -= ls;
          wm_list_pop_front(m, e);
        }
      }
    }
    m.ring_head = m.ring_head + 1 == m.ring_cap ? 0 : m.ring_head + 1;
    m.ring_n--;
  }

  if (last_base_is_n) m.ambig = m.k;
  if (hash_bwd != hash_fwd && m.ambig == 0) { /* the arriving k-mer (:417-445) */
    if (m.ring_n >= m.ring_cap) m.fail = 1;
    else {
      int32_t at = m.ring_head + m.ring_n;
      if (at >= m.ring_cap) at -= m.ring_cap;
      m.ring[at].hash = cur; m.ring[at].pos = i; m.ring[at].strand = cur_strand;
      m.ring_n++;
    }
    /* a hash above the largest member cannot be one (19 of 20 arrivals on ordinary sequence): no search */
    const int32_t idx = (m.mem_n > 0 && cur <= m.mh[m.mem_n - 1]) ? wm_find(m, cur) : -1;
    if (idx >= 0) {
      wm_member &e = wm_at(m, idx);
      wm_list_push_back(m, e, i, cur_strand);
      if (e.votes + cur_strand == 0 || e.votes == 0) {
        wm_emit(m, e, wid, i);
        e.wpos = wid;
      }
      e.votes += cur_strand;
    } else {
      wm_kmer x; x.hash = cur; x.pos = i; x.strand = cur_strand;
      wm_heap_push(m, x);
    }
  }
  if (m.ambig > 0) m.ambig--;

  if (wid >= 0) { /* refill from the waiting heap (:455-505) */
    while (m.heap_n > 0 && m.heap[0].pos < wid) wm_heap_pop(m);
    if (m.mem_n > 0 && m.heap_n > 0 && m.mem_n == m.s && m.heap[0].hash < m.mh[m.mem_n - 1]) {
      wm_member &big = wm_at(m, m.mem_n - 1);
      wm_emit(m, big, wid, i);
      for (int32_t n = big.head, c = 0; c < big.count; c++, n = m.nodes[n].next) {
        if (m.nodes[n].pos > wid) { /* `>`: an occurrence at the window start is dropped (:478) */
          wm_kmer x; x.hash = big.hash; x.pos = m.nodes[n].pos; x.strand = m.nodes[n].strand;
          wm_heap_push(m, x);
        }
      }
      wm_erase_member(m, m.mem_n - 1);
    }
    while (m.heap_n > 0 && m.mem_n < m.s) {
      if (m.heap[0].pos < wid) wm_heap_pop(m);
      /* the reference takes the top now without looking at it again: if the heap is empty it reads the element that was
       * just popped (:495), if the top is another expired entry it becomes a member. Only then does the machine's future
       * depend on expired entries, i.e. on history older than the window: remember it (chunk stitching) */
      if (m.heap_n == 0 || m.heap[0].pos < wid) m.drained = 1;
      const wm_kmer nk = m.heap[0];
      int32_t idx = wm_find(m, nk.hash);
      if (idx < 0) idx = wm_insert_member(m, nk.hash);
      if (idx < 0) break;
      { /* sortedWindow[h].first = MinmerInfo{h, wid, -1, seq, 0}: resets an existing member's record, keeps its list */
        wm_member &e = wm_at(m, idx);
        e.wpos = wid; e.votes = 0; e.inherited = 0;
      }
      while (m.heap_n > 0 && m.heap[0].hash == nk.hash) {
        wm_member &e = wm_at(m, idx);
        wm_list_push_back(m, e, m.heap[0].pos, m.heap[0].strand);
        e.votes += m.heap[0].strand;
        wm_heap_pop(m);
      }
    }
  }
}

/* the records still open at the end of the contig (:508-520), in hash order */
WM_HD void wm_flush(wm_machine &m, int32_t n_positions)
{
  for (int32_t i = 0; i < m.mem_n && i < m.s; i++) {
    wm_member &e = wm_at(m, i);
    if (e.wpos != -1) wm_emit(m, e, n_positions, n_positions);
  }
}

/* order-independent digest of the part of the state that decides everything the machine does from here on (the wpos of
 * the open records excepted: those are inherited from the predecessor chunk): window ring, members with votes and
 * occurrence lists, the live part of the waiting heap, the N counter. */
WM_HD uint64_t wm_mix(uint64_t x)
{
  x ^= x >> 31; x *= 0x7fb5d329728ea185ULL; x ^= x >> 27; x *= 0x81dadef4bc2dd44dULL; x ^= x >> 33;
  return x;
}
WM_HD uint64_t wm_digest(const wm_machine &m, int32_t wid)
{
  uint64_t d = wm_mix((uint64_t)(uint32_t)m.ambig + 0x1234567ULL) + wm_mix((uint64_t)m.ring_n << 20 | (uint64_t)m.mem_n);
  for (int32_t j = 0; j < m.ring_n; j++) {
    int32_t at = m.ring_head + j;
    if (at >= m.ring_cap) at -= m.ring_cap;
    d += wm_mix(m.ring[at].hash ^ wm_mix(((uint64_t)(uint32_t)m.ring[at].pos << 2) | (uint64_t)(m.ring[at].strand & 3)));
  }
  for (int32_t j = 0; j < m.mem_n; j++) {
    const wm_member &e = wm_at(m, j);
    uint64_t x = wm_mix(e.hash + 0x9e3779b97f4a7c15ULL) ^ wm_mix(((uint64_t)(uint32_t)e.votes << 32) | (uint32_t)e.count) ^ (e.wpos == -1 ? 77 : 0);
    uint64_t seq = 0;
    for (int32_t n = e.head, c = 0; c < e.count; c++, n = m.nodes[n].next)
      seq = wm_mix(seq + (((uint64_t)(uint32_t)m.nodes[n].pos << 2) | (uint64_t)(m.nodes[n].strand & 3)) + 1);
    d += wm_mix(x ^ seq);
  }
  for (int32_t j = 0; j < m.heap_n; j++)
    if (!(m.heap[j].pos < wid))
      d += wm_mix(m.heap[j].hash * 3 + wm_mix(((uint64_t)(uint32_t)m.heap[j].pos << 2) | (uint64_t)(m.heap[j].strand & 3)) + 5);
  return d;
}

/* ---- scanning a stretch of a contig: bases -> both hashes per position -> wm_step ------------------------------------ */

WM_HD uint32_t wm_norm(uint32_t c)
{ /* makeUpperCaseAndValidDNA (commonFunc.hpp:97-107) of one byte */
  if (c > 96 && c < 123) c -= 32;
  return (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : (uint32_t)'N';
}
WM_HD uint32_t wm_comp(uint32_t c)
{ /* reverseComplement's per-base map (commonFunc.hpp:50-73): anything else (N) stays */
  return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
}

/* the k-mer and its reverse complement as the byte strings the reference hashes (an N is hashed as the letter N: a k-mer at
 * the very start of a contig is not protected by the N counter, commonFunc.hpp:413 looks at the LAST base only) */
template <int K>
struct wm_kmer_bytes {
  static constexpr int NW = mm_kmer_words<K>::NW;
  uint64_t f[NW], r[NW];
  WM_HD void reset()
  {
    for (int i = 0; i < NW; i++) { f[i] = 0; r[i] = 0; }
  }
  WM_HD void push(uint32_t base /* normalised */)
  {
    const uint64_t fa = base, ra = wm_comp(base);
    for (int i = 0; i < NW - 1; i++) f[i] = (f[i] >> 8) | (f[i + 1] << 56);
    f[NW - 1] = (f[NW - 1] >> 8) | (fa << (8 * ((K - 1) & 7)));
    for (int i = NW - 1; i > 0; i--) r[i] = (r[i] << 8) | (r[i - 1] >> 56);
    r[0] = (r[0] << 8) | ra;
    r[NW - 1] &= mm_kmer_words<K>::TOP_MASK;
  }
};

/* Scans k-mer positions [from, to) of a contig of contig_len bases with machine m. `fresh` = the machine starts here
 * (empty state at `from`): the byte windows are primed with bases from .. from+K-2 and the N counter is set to what the
 * reference's would be at `from` (it saw those bases as "last base" of earlier positions, unless from == 0). Otherwise the
 * windows `win` continue from a previous call that ended at `from`. */
template <int K>
WM_HD void wm_scan(wm_machine &m, wm_kmer_bytes<K> &win, const uint8_t *seq, int32_t from, int32_t to, bool fresh)
{
  if (fresh) {
    win.reset();
    int32_t ambig = 0;
    for (int32_t j = 0; j < K - 1; j++) {
      const uint32_t b = wm_norm(seq[from + j]);
      win.push(b);
      if (b == 'N' && from + j >= K - 1) ambig = j + 1; /* that base was the last base of position from + j - (K - 1) >= 0 */
    }
    m.ambig = ambig;
  }
  for (int32_t i = from; i < to; i++) {
    const uint32_t b = wm_norm(seq[i + K - 1]);
    win.push(b);
    const uint64_t hf = mm_murmur3_k<K>(win.f), hb = mm_murmur3_k<K>(win.r);
    wm_step(m, i, hf, hb, b == 'N');
  }
}

#endif /* MM_WINMACHINE_H */
