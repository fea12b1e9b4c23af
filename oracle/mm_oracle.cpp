/* ORACLE / TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product.
 *
 * CPU restatement of the reference's mapping hot path (marbl/MashMap v3.1.3), plain sequential C++ that
 * follows the reference routine by routine; every function cites the reference file:line it restates.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Pinning: the reference ships no golden vectors for this path, so this restatement is pinned
 * against the reference itself: tests/test_oracle.py compares every stage (hash, sketch, interval points,
 * L1 candidates, L2 loci, fragment mappings, read mappings) with oracle/_ref/libmm_ref.so -- the unmodified
 * reference compiled from /root/reference by oracle/Makefile -- and against the committed fixtures in
 * tests/golden/ that were generated from it (tests/golden/make_golden.py).
 * The three GSL functions come from gsl_shim/ ("parity unpinned" at that boundary, see gsl_cdf.h).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <numeric>
#include <set>
#include <string>
#include <tuple>
#include <thread>
#include <atomic>
#include <unordered_map>
#include <vector>

#include "gsl/gsl_cdf.h"
#include "mm_oracle_types.h"

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

typedef uint64_t hash_t;
enum { FWD = 1, AMBIG = 0, REV = -1 };
enum { OPEN = 1, CLOSE = -1 };
const float CONFIDENCE = 0.95f;       /* map_parameters.hpp:96 */
const double SS_TABLE_MAX = 1000.0;   /* map_parameters.hpp:94 */

/* ---- murmur3.h:236-303 (MurmurHash3_x64_128), commonFunc.hpp:138-147 (getHash, seed 42, low word) ---- */
inline uint64_t rotl64(uint64_t x, int8_t r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t fmix64(uint64_t k)
{
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}
hash_t getHash(const char *seq, int len)
{
  const uint8_t *data = (const uint8_t *)seq;
  const int nblocks = len / 16;
  uint64_t h1 = 42, h2 = 42;
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  for (int i = 0; i < nblocks; i++) {
    uint64_t k1, k2;
    memcpy(&k1, data + 16 * i, 8); memcpy(&k2, data + 16 * i + 8, 8);
    k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  }
  const uint8_t *tail = data + nblocks * 16;
  uint64_t k1 = 0, k2 = 0;
  switch (len & 15) {
    case 15: k2 ^= (uint64_t)(tail[14]) << 48; /* fallthrough */
    case 14: k2 ^= (uint64_t)(tail[13]) << 40; /* fallthrough */
    case 13: k2 ^= (uint64_t)(tail[12]) << 32; /* fallthrough */
    case 12: k2 ^= (uint64_t)(tail[11]) << 24; /* fallthrough */
    case 11: k2 ^= (uint64_t)(tail[10]) << 16; /* fallthrough */
    case 10: k2 ^= (uint64_t)(tail[9]) << 8;   /* fallthrough */
    case 9:  k2 ^= (uint64_t)(tail[8]) << 0;
             k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; /* fallthrough */
    case 8:  k1 ^= (uint64_t)(tail[7]) << 56; /* fallthrough */
    case 7:  k1 ^= (uint64_t)(tail[6]) << 48; /* fallthrough */
    case 6:  k1 ^= (uint64_t)(tail[5]) << 40; /* fallthrough */
    case 5:  k1 ^= (uint64_t)(tail[4]) << 32; /* fallthrough */
    case 4:  k1 ^= (uint64_t)(tail[3]) << 24; /* fallthrough */
    case 3:  k1 ^= (uint64_t)(tail[2]) << 16; /* fallthrough */
    case 2:  k1 ^= (uint64_t)(tail[1]) << 8;  /* fallthrough */
    case 1:  k1 ^= (uint64_t)(tail[0]) << 0;
             k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
  }
  h1 ^= len; h2 ^= len;
  h1 += h2; h2 += h1;
  h1 = fmix64(h1); h2 = fmix64(h2);
  h1 += h2;
  return h1;
}

/* commonFunc.hpp:97-107 (bytes >= 127 index outside the reference's 127-entry table: treated as N) */
void makeUpperCaseAndValidDNA(char *seq, int len)
{
  for (int i = 0; i < len; i++) {
    if (seq[i] > 96 && seq[i] < 123) seq[i] -= 32;
    if (!(seq[i] == 'A' || seq[i] == 'C' || seq[i] == 'G' || seq[i] == 'T')) seq[i] = 'N';
  }
}
/* commonFunc.hpp:50-73 */
void reverseComplement(const char *src, char *dest, int length)
{
  for (int i = 0; i < length; i++) {
    char base = src[i];
    switch (base) {
      case 'A': base = 'T'; break;
      case 'C': base = 'G'; break;
      case 'G': base = 'C'; break;
      case 'T': base = 'A'; break;
      default: break;
    }
    dest[length - i - 1] = base;
  }
}

/* commonFunc.hpp:182-288 */
void sketchSequence(std::vector<orc_minmer> &minmerIndex, char *seq, int len, int kmerSize, int sketchSize, int seqCounter)
{
  makeUpperCaseAndValidDNA(seq, len);
  std::vector<char> seqRev((size_t)std::max(len, 1));
  reverseComplement(seq, seqRev.data(), len);
  std::unordered_map<hash_t, orc_minmer> sketched_vals;
  std::vector<hash_t> sketched_heap; /* max-heap */
  int ambig_kmer_count = 0;
  for (int i = kmerSize - 1; i >= 0; i--) {
    if (i < len && seq[i] == 'N') { ambig_kmer_count = i + 1; break; }
  }
  for (int i = 0; i < len - kmerSize + 1; i++) {
    if (seq[i + kmerSize - 1] == 'N') ambig_kmer_count = kmerSize;
    hash_t hashFwd = getHash(seq + i, kmerSize);
    hash_t hashBwd = getHash(seqRev.data() + len - i - kmerSize, kmerSize);
    if (hashBwd != hashFwd && ambig_kmer_count == 0) {
      hash_t currentKmer = std::min(hashFwd, hashBwd);
      int currentStrand = hashFwd < hashBwd ? FWD : REV;
      if ((int)sketched_heap.size() < sketchSize || currentKmer <= sketched_heap.front()) {
        if (sketched_heap.empty() || sketched_vals.find(currentKmer) == sketched_vals.end()) {
          if ((int)sketched_vals.size() < sketchSize || currentKmer < sketched_heap.front()) {
            orc_minmer mi; mi.hash = currentKmer; mi.wpos = i; mi.wpos_end = i; mi.seqId = seqCounter; mi.strand = (int16_t)currentStrand; mi._pad = 0;
            sketched_vals[currentKmer] = mi;
            sketched_heap.push_back(currentKmer);
            std::push_heap(sketched_heap.begin(), sketched_heap.end());
          }
          if ((int)sketched_vals.size() > sketchSize) {
            sketched_vals.erase(sketched_heap[0]);
            std::pop_heap(sketched_heap.begin(), sketched_heap.end());
            sketched_heap.pop_back();
          }
        } else {
          sketched_vals[currentKmer].wpos_end = i;
          sketched_vals[currentKmer].strand += currentStrand == FWD ? 1 : -1;
        }
      }
    }
    if (ambig_kmer_count > 0) ambig_kmer_count--;
  }
  minmerIndex.resize(sketched_heap.size());
  for (auto rev_it = minmerIndex.rbegin(); rev_it != minmerIndex.rend(); rev_it++) {
    *rev_it = sketched_vals[sketched_heap.front()];
    rev_it->strand = rev_it->strand > 0 ? FWD : (rev_it->strand == 0 ? AMBIG : REV);
    std::pop_heap(sketched_heap.begin(), sketched_heap.end());
    sketched_heap.pop_back();
  }
}

/* ---- map_stats.hpp ---- */
float j2md(float j, int k)
{ /* :45-55 */
  if (j == 0) return 1.0;
  if (j == 1) return 0.0;
  float mash_dist = 1 - std::pow(2 * j / (1 + j), 1.0 / k);
  return mash_dist;
}
float md2j(float d, int k)
{ /* :63-68 */
  float sim = 1 - d;
  float jaccard = std::pow(sim, k) / (2 - std::pow(sim, k));
  return jaccard;
}
float md_lower_bound(float d, int s, int k, float ci)
{ /* :81-113 */
  float q2 = (1.0 - ci) / 2;
  int x = std::max(int(ceil(s * md2j(d, k))), 1);
  while (x <= s) {
    double cdf_complement = gsl_cdf_binomial_Q(x - 1, md2j(d, k), s);
    if (cdf_complement < q2) { x--; break; }
    x++;
  }
  float jaccard = float(x) / s;
  return j2md(jaccard, k);
}
int estimateMinimumHits(int s, int k, float perc_identity)
{ /* :122-133 */
  float mash_dist = 1.0 - perc_identity;
  float jaccard = md2j(mash_dist, k);
  return ceil(1.0 * s * jaccard);
}
int estimateMinimumHitsRelaxed(int s, int k, float perc_identity, float ci)
{ /* :144-169 */
  int first = estimateMinimumHits(s, k, perc_identity);
  int relaxed = first;
  for (int i = first; i >= 0; i--) {
    float jaccard = 1.0 * i / s;
    float d = j2md(jaccard, k);
    float d_lower = md_lower_bound(d, s, k, ci);
    float id_upper = 1.0 - d_lower;
    if (id_upper >= perc_identity) relaxed = i; else break;
  }
  return relaxed;
}

struct Ctx {
  orc_params p;
  std::vector<orc_minmer> minmerIndex;                 /* winSketch.hpp:102 */
  std::vector<hash_t> keys;                            /* minmerPosLookupIndex keys, ascending */
  std::vector<uint64_t> offs;
  std::vector<orc_ipoint> pts;
  std::vector<uint8_t> isFreq;
  std::vector<int> contigLen;
  std::vector<std::string> contigName;
  std::vector<int> sketchCutoffs;                      /* computeMap.hpp:109 */
  std::vector<int> refIdGroup;
  std::map<int, int> minHitsMemo;
};

/* computeMap.hpp:178-258 (setProbs) */
void setProbs(Ctx &c)
{
  const orc_params &param = c.p;
  int ss = std::min<double>(param.sketchSize, SS_TABLE_MAX);
  c.sketchCutoffs.assign(ss + 1, 1);
  if (!param.stage1_topANI_filter) return;
  float deltaANI = param.ANIDiff;
  float min_p = 1 - param.ANIDiffConf;
  std::vector<std::vector<double>> sketchProbs(ss + 1, std::vector<double>(ss + 1.0));
  for (int ci = 0; ci <= ss; ci++)
    for (double y = 0; y <= ci; y++) sketchProbs[ci][y] = gsl_ran_hypergeometric_pdf(y, ss, ss - ci, ci);
  auto distDiff = [&](int cmax, int ci) {
    double prAboveCutoff = 0;
    for (double ymax = 0; ymax <= cmax; ymax++) {
      double pymax = sketchProbs[cmax][ymax];
      double yi_cutoff = deltaANI == 0 ? ymax : (std::floor(md2j(j2md(ymax / ss, param.kmerSize) + deltaANI, param.kmerSize) * ss));
      double pi_acc = (yi_cutoff - 1) >= 0 ? gsl_cdf_hypergeometric_P(yi_cutoff - 1, ss, ss - ci, ci) : 0;
      pi_acc = 1 - pi_acc;
      prAboveCutoff += pymax * pi_acc;
      if (prAboveCutoff > min_p) return true;
    }
    return prAboveCutoff > min_p;
  };
  std::vector<int> ss_range(ss + 1);
  std::iota(ss_range.begin(), ss_range.end(), 0);
  for (int cmax = 1; cmax <= ss; cmax++) {
    int ci = std::distance(ss_range.begin(), std::upper_bound(ss_range.begin(), ss_range.begin() + ss, false,
                                                              [&](bool, int ci2) { return distDiff(cmax, ci2); }));
    c.sketchCutoffs[cmax] = ci;
    if (c.sketchCutoffs[cmax] == 0) c.sketchCutoffs[cmax] = 1;
  }
}

bool isFreqSeed(const Ctx &c, hash_t h)
{ /* winSketch.hpp:506-509 */
  auto it = std::lower_bound(c.keys.begin(), c.keys.end(), h);
  return it != c.keys.end() && *it == h && c.isFreq[it - c.keys.begin()];
}

struct Query {
  std::vector<char> seq;
  int len, seqCounter, refGroup, nameId;
  std::vector<orc_minmer> minmerTableQuery;
  int sketchSize = 0;
  float kmerComplexity = 0;
};

inline bool ipLess(const orc_ipoint &a, const orc_ipoint &b)
{ /* base_types.hpp:75-78 */
  return std::tie(a.seqId, a.pos, a.side) < std::tie(b.seqId, b.pos, b.side);
}

/* computeMap.hpp:817-843 */
void getSeedHits(const Ctx &c, Query &Q)
{
  sketchSequence(Q.minmerTableQuery, Q.seq.data(), Q.len, c.p.kmerSize, c.p.sketchSize, Q.seqCounter);
  if (Q.minmerTableQuery.size() == 0) { Q.sketchSize = 0; return; }
  const double max_hash_01 = (long double)(Q.minmerTableQuery.back().hash) / std::numeric_limits<hash_t>::max();
  Q.kmerComplexity = (double(Q.minmerTableQuery.size()) / max_hash_01) / ((Q.len - c.p.kmerSize + 1) * 2);
  auto new_end = std::remove_if(Q.minmerTableQuery.begin(), Q.minmerTableQuery.end(), [&](orc_minmer &mi) { return isFreqSeed(c, mi.hash); });
  Q.minmerTableQuery.erase(new_end, Q.minmerTableQuery.end());
  Q.sketchSize = Q.minmerTableQuery.size();
}

/* computeMap.hpp:856-912 (k-way heap merge of the per-hash point lists) */
void getSeedIntervalPoints(const Ctx &c, Query &Q, std::vector<orc_ipoint> &intervalPoints)
{
  if (Q.minmerTableQuery.size() == 0) return;
  struct BoundPtr { const orc_ipoint *it, *end; };
  std::vector<BoundPtr> pq;
  auto heap_cmp = [](const BoundPtr &a, const BoundPtr &b) { return ipLess(*b.it, *a.it); };
  for (auto &mi : Q.minmerTableQuery) {
    auto kit = std::lower_bound(c.keys.begin(), c.keys.end(), mi.hash);
    if (kit != c.keys.end() && *kit == mi.hash) {
      size_t ki = kit - c.keys.begin();
      pq.push_back(BoundPtr{c.pts.data() + c.offs[ki], c.pts.data() + c.offs[ki + 1]});
    }
  }
  std::make_heap(pq.begin(), pq.end(), heap_cmp);
  while (!pq.empty()) {
    const orc_ipoint *ip = pq.front().it;
    /* :891-893; skip_self compares names: the caller passes the id of the reference name equal to the query's */
    if ((!c.p.skip_self || Q.nameId < 0 || c.contigName[ip->seqId] != c.contigName[Q.nameId]) &&
        (!c.p.skip_prefix || c.refIdGroup[ip->seqId] != Q.refGroup) && (!c.p.lower_triangular || Q.seqCounter > ip->seqId))
      intervalPoints.push_back(*ip);
    std::pop_heap(pq.begin(), pq.end(), heap_cmp);
    pq.back().it++;
    if (pq.back().it >= pq.back().end) pq.pop_back();
    else std::push_heap(pq.begin(), pq.end(), heap_cmp);
  }
}

/* computeMap.hpp:915-1116 (windowLen == 0 for every fragment the device path accepts; the general code is kept) */
void computeL1CandidateRegions(const Ctx &c, Query &Q, const orc_ipoint *ip_begin, const orc_ipoint *ip_end, int minimumHits,
                               std::vector<orc_l1> &l1Mappings)
{
  const orc_params &param = c.p;
  int overlapCount = 0, bestIntersectionSize = 0;
  std::vector<orc_l1> localOpts;
  int windowLen = std::max<int>(0, Q.len - param.segLength);
  const orc_ipoint *trailingIt = ip_begin, *leadingIt = ip_begin;
  int clusterLen = param.segLength;
  std::unordered_map<hash_t, int> hash_to_freq;
  if (param.stage1_topANI_filter) {
    while (leadingIt != ip_end) {
      while (trailingIt != ip_end && ((trailingIt->seqId == leadingIt->seqId && trailingIt->pos <= leadingIt->pos - windowLen) ||
                                      trailingIt->seqId < leadingIt->seqId)) {
        if (trailingIt->side == CLOSE) {
          if (windowLen != 0) hash_to_freq[trailingIt->hash]--;
          if (windowLen == 0 || hash_to_freq[trailingIt->hash] == 0) overlapCount--;
        }
        trailingIt++;
      }
      auto currentPos = leadingIt->pos;
      while (leadingIt != ip_end && leadingIt->pos == currentPos) {
        if (leadingIt->side == OPEN) {
          if (windowLen == 0 || hash_to_freq[leadingIt->hash] == 0) overlapCount++;
          if (windowLen != 0) hash_to_freq[leadingIt->hash]++;
        }
        leadingIt++;
      }
      bestIntersectionSize = std::max(bestIntersectionSize, overlapCount);
    }
    if (bestIntersectionSize < minimumHits) return;
    minimumHits = std::max(c.sketchCutoffs[int(std::min(bestIntersectionSize, Q.sketchSize) / std::max<double>(1, param.sketchSize / SS_TABLE_MAX))],
                           minimumHits);
  }
  hash_to_freq.clear();
  bestIntersectionSize = std::min(bestIntersectionSize, Q.sketchSize);
  bool in_candidate = false;
  orc_l1 l1_out = {};
  trailingIt = ip_begin; leadingIt = ip_begin;
  overlapCount = 0;
  int prevOverlap = 0;
  struct SeqCoord { int seqId, pos; };
  SeqCoord prevPos = {0, 0};
  SeqCoord currentPos{leadingIt->seqId, leadingIt->pos};
  while (leadingIt != ip_end) {
    prevOverlap = overlapCount;
    while (trailingIt != ip_end && ((trailingIt->seqId == leadingIt->seqId && trailingIt->pos <= leadingIt->pos - windowLen) ||
                                    trailingIt->seqId < leadingIt->seqId)) {
      if (trailingIt->side == CLOSE) {
        if (windowLen != 0) hash_to_freq[trailingIt->hash]--;
        if (windowLen == 0 || hash_to_freq[trailingIt->hash] == 0) overlapCount--;
      }
      trailingIt++;
    }
    if (leadingIt->pos != currentPos.pos) {
      prevPos = currentPos;
      currentPos = SeqCoord{leadingIt->seqId, leadingIt->pos};
    }
    while (leadingIt != ip_end && leadingIt->pos == currentPos.pos) {
      if (leadingIt->side == OPEN) {
        if (windowLen == 0 || hash_to_freq[leadingIt->hash] == 0) overlapCount++;
        if (windowLen != 0) hash_to_freq[leadingIt->hash]++;
      }
      leadingIt++;
    }
    if (prevOverlap >= minimumHits) {
      if (l1_out.seqId != prevPos.seqId && in_candidate) {
        localOpts.push_back(l1_out);
        l1_out = {};
        in_candidate = false;
      }
      if (!in_candidate) {
        l1_out.rangeStartPos = prevPos.pos - windowLen;
        l1_out.rangeEndPos = prevPos.pos - windowLen;
        l1_out.seqId = prevPos.seqId;
        l1_out.intersectionSize = prevOverlap;
        in_candidate = true;
      } else { /* stage2_full_scan is always true (parseCmdArgs.hpp:590) */
        l1_out.intersectionSize = std::max(l1_out.intersectionSize, prevOverlap);
        l1_out.rangeEndPos = prevPos.pos - windowLen;
      }
    } else {
      if (in_candidate) { localOpts.push_back(l1_out); l1_out = {}; }
      in_candidate = false;
    }
  }
  if (in_candidate) localOpts.push_back(l1_out);
  for (auto &lo : localOpts) {
    if (l1Mappings.empty() || lo.seqId != l1Mappings.back().seqId || lo.rangeStartPos > l1Mappings.back().rangeEndPos + clusterLen) {
      l1Mappings.push_back(lo);
    } else {
      l1Mappings.back().rangeEndPos = lo.rangeEndPos;
      l1Mappings.back().intersectionSize = std::max(lo.intersectionSize, l1Mappings.back().intersectionSize);
    }
  }
}

int minimumHitsFor(Ctx &c, int qs)
{
  auto it = c.minHitsMemo.find(qs);
  if (it != c.minHitsMemo.end()) return it->second;
  int v = estimateMinimumHitsRelaxed(qs, c.p.kmerSize, c.p.percentageIdentity, CONFIDENCE);
  c.minHitsMemo[qs] = v;
  return v;
}

/* computeMap.hpp:1129-1166 */
void doL1Mapping(Ctx &c, Query &Q, std::vector<orc_ipoint> &intervalPoints, std::vector<orc_l1> &l1Mappings, int *minimumHitsOut)
{
  getSeedHits(c, Q);
  if (minimumHitsOut) *minimumHitsOut = 0;
  if (Q.sketchSize == 0 || Q.kmerComplexity < c.p.kmerComplexityThreshold) return;
  getSeedIntervalPoints(c, Q, intervalPoints);
  int minimumHits = minimumHitsFor(c, Q.sketchSize);
  if (minimumHitsOut) *minimumHitsOut = minimumHits;
  const orc_ipoint *ip_begin = intervalPoints.data(), *ip_end = intervalPoints.data(), *end = intervalPoints.data() + intervalPoints.size();
  while (ip_end != end) {
    if (c.p.skip_prefix) {
      int currGroup = c.refIdGroup[ip_begin->seqId];
      ip_end = std::find_if_not(ip_begin, end, [&](const orc_ipoint &ip) { return currGroup == c.refIdGroup[ip.seqId]; });
    } else {
      ip_end = end;
    }
    computeL1CandidateRegions(c, Q, ip_begin, ip_end, minimumHits, l1Mappings);
    ip_begin = ip_end;
  }
}

/* slidingMap.hpp:27-212 */
struct SlideMapper {
  struct Val { hash_t hash_val; int16_t q_strand; int16_t strand_vote; unsigned int num_before_inc; bool active; };
  const Query &Q;
  std::vector<Val> v;
  size_t pivot, pivRank;
  int sharedSketchElements = 0, strand_votes = 0, intersectionSize = 0;
  explicit SlideMapper(const Query &Q_) : Q(Q_), v(Q_.sketchSize + 1)
  {
    int idx = 1;
    for (auto &mi : Q.minmerTableQuery) v[idx++] = Val{mi.hash, mi.strand, 0, 1, false};
    pivot = v.size() - 1;
    pivRank = v.size() - 1;
  }
  size_t locate(hash_t h) const
  {
    return std::lower_bound(v.begin() + 1, v.end(), h, [](const Val &a, hash_t b) { return a.hash_val < b; }) - v.begin();
  }
  void insert_minmer(const orc_minmer &mi)
  { /* :125-165 */
    size_t loc = locate(mi.hash);
    if (loc == v.size()) return;
    if (v[loc].hash_val == mi.hash) {
      v[loc].active = true;
      v[loc].strand_vote += (v[loc].q_strand * mi.strand);
      intersectionSize++;
      if (v[loc].hash_val <= v[pivot].hash_val) { sharedSketchElements++; strand_votes += v[loc].strand_vote; }
    } else {
      v[loc].num_before_inc++;
      if (v[loc].hash_val <= v[pivot].hash_val) pivRank++;
      if (pivRank > (size_t)Q.sketchSize) {
        sharedSketchElements -= v[pivot].active;
        strand_votes -= v[pivot].strand_vote;
        pivRank -= v[pivot].num_before_inc;
        pivot--;
      }
    }
  }
  void delete_minmer(const orc_minmer &mi)
  { /* :171-211 */
    size_t loc = locate(mi.hash);
    if (loc == v.size()) return;
    if (v[loc].hash_val == mi.hash) {
      if (v[loc].hash_val <= v[pivot].hash_val) { sharedSketchElements--; strand_votes -= v[loc].strand_vote; }
      v[loc].active = false;
      v[loc].strand_vote = 0;
      intersectionSize--;
    } else {
      v[loc].num_before_inc--;
      if (v[loc].hash_val <= v[pivot].hash_val) pivRank--;
      if (pivot + 1 != v.size() && pivRank + v[pivot + 1].num_before_inc <= (size_t)Q.sketchSize) {
        pivot++;
        sharedSketchElements += v[pivot].active;
        strand_votes += v[pivot].strand_vote;
        pivRank += v[pivot].num_before_inc;
      }
    }
  }
};

/* computeMap.hpp:1275-1451. std::next(windowIt) == end() is read out of bounds by the reference; it is
 * defined here as "another contig" (SURVEY A.6). */
void computeL2MappedRegions(const Ctx &c, Query &Q, const orc_l1 &cand, std::vector<orc_l2> &l2_vec_out)
{
  const auto &minmerIndex = c.minmerIndex;
  const orc_params &param = c.p;
  auto lessBySeqPos = [](const orc_minmer &a, const orc_minmer &b) { return std::tie(a.seqId, a.wpos) < std::tie(b.seqId, b.wpos); };
  orc_minmer first_minmer = {0, cand.rangeStartPos - param.segLength - 1, 0, cand.seqId, 0, 0};
  size_t windowIt = std::lower_bound(minmerIndex.begin(), minmerIndex.end(), first_minmer, lessBySeqPos) - minmerIndex.begin();
  const size_t END = minmerIndex.size();
  std::vector<orc_minmer> slidingWindow;
  auto heap_cmp = [](const orc_minmer &l, const orc_minmer &r) { return l.wpos_end > r.wpos_end; };
  int windowLen = std::max<int>(0, Q.len - param.segLength);
  std::unordered_map<hash_t, int> hash_to_freq;
  SlideMapper slideMap(Q);
  int bestSketchSize = 1;
  bool in_candidate = false;
  orc_l2 l2_out = {};
  auto nextWpos = [&](size_t it) { /* std::next(windowIt, next is on the same contig)->wpos */
    return (it + 1 < END && minmerIndex[it + 1].seqId == minmerIndex[it].seqId) ? minmerIndex[it + 1].wpos : minmerIndex[it].wpos;
  };
  while (windowIt != END && minmerIndex[windowIt].seqId == cand.seqId && minmerIndex[windowIt].wpos < cand.rangeStartPos) {
    if (minmerIndex[windowIt].wpos_end > cand.rangeStartPos) {
      if (windowLen > 0) hash_to_freq[minmerIndex[windowIt].hash]++;
      if (windowLen == 0 || hash_to_freq[minmerIndex[windowIt].hash] == 1) {
        slidingWindow.push_back(minmerIndex[windowIt]);
        std::push_heap(slidingWindow.begin(), slidingWindow.end(), heap_cmp);
        slideMap.insert_minmer(minmerIndex[windowIt]);
      }
    }
    windowIt++;
  }
  while (windowIt != END && minmerIndex[windowIt].seqId == cand.seqId && minmerIndex[windowIt].wpos <= cand.rangeEndPos + windowLen) {
    const orc_minmer &w = minmerIndex[windowIt];
    int prev_strand_votes = slideMap.strand_votes;
    while (!slidingWindow.empty() && slidingWindow.front().wpos_end <= w.wpos - windowLen) {
      if (windowLen > 0) hash_to_freq[slidingWindow.front().hash]--;
      if (windowLen == 0 || hash_to_freq[slidingWindow.front().hash] == 0) {
        slideMap.delete_minmer(slidingWindow.front());
        std::pop_heap(slidingWindow.begin(), slidingWindow.end(), heap_cmp);
        slidingWindow.pop_back();
      }
    }
    if (windowLen > 0) hash_to_freq[w.hash]++;
    if (windowLen == 0 || hash_to_freq[w.hash] == 1) {
      slideMap.insert_minmer(w);
      slidingWindow.push_back(w);
      std::push_heap(slidingWindow.begin(), slidingWindow.end(), heap_cmp);
    } else {
      windowIt++;
      continue;
    }
    if (slideMap.sharedSketchElements > bestSketchSize) {
      l2_vec_out.clear();
      in_candidate = true;
      bestSketchSize = slideMap.sharedSketchElements;
      l2_out.sharedSketchSize = slideMap.sharedSketchElements;
      l2_out.optimalStart = w.wpos;
      l2_out.optimalEnd = nextWpos(windowIt) - windowLen;
    } else if (slideMap.sharedSketchElements == bestSketchSize) {
      if (!in_candidate) {
        l2_out.sharedSketchSize = slideMap.sharedSketchElements;
        l2_out.optimalStart = w.wpos - windowLen;
      }
      in_candidate = true;
      l2_out.optimalEnd = nextWpos(windowIt) - windowLen;
    } else {
      if (in_candidate) {
        l2_out.optimalEnd = nextWpos(windowIt) - windowLen;
        l2_out.meanOptimalPos = (l2_out.optimalStart + l2_out.optimalEnd) / 2;
        l2_out.seqId = w.seqId;
        l2_out.strand = prev_strand_votes >= 0 ? FWD : REV;
        if (l2_vec_out.empty() || l2_vec_out.back().optimalEnd + param.segLength < l2_out.optimalStart) {
          l2_vec_out.push_back(l2_out);
        } else {
          l2_vec_out.back().optimalEnd = l2_out.optimalEnd;
          l2_vec_out.back().meanOptimalPos = (l2_vec_out.back().optimalStart + l2_vec_out.back().optimalEnd) / 2;
        }
        l2_out = orc_l2();
      }
      in_candidate = false;
    }
    windowIt++;
  }
  if (in_candidate) {
    l2_out.meanOptimalPos = (l2_out.optimalStart + l2_out.optimalEnd) / 2;
    l2_out.seqId = minmerIndex[windowIt - 1].seqId;
    l2_out.strand = slideMap.strand_votes >= 0 ? FWD : REV;
    if (l2_vec_out.empty() || l2_vec_out.back().optimalEnd + param.segLength < l2_out.optimalStart) {
      l2_vec_out.push_back(l2_out);
    } else {
      l2_vec_out.back().optimalEnd = l2_out.optimalEnd;
      l2_vec_out.back().meanOptimalPos = (l2_vec_out.back().optimalStart + l2_vec_out.back().optimalEnd) / 2;
    }
  }
}

/* computeMap.hpp:1181-1267 */
void doL2Mapping(Ctx &c, Query &Q, int fullLen, orc_l1 *l1_begin, orc_l1 *l1_end, std::vector<orc_mapping> &l2Mappings)
{
  const orc_params &param = c.p;
  auto cmp = [](const orc_l1 &a, const orc_l1 &b) { return a.intersectionSize < b.intersectionSize; };
  std::vector<orc_l2> l2_vec;
  double bestJaccardNumerator = 0;
  orc_l1 *loc_iterator = l1_begin;
  while (loc_iterator != l1_end) {
    orc_l1 &candidateLocus = *loc_iterator;
    if (param.stage1_topANI_filter) {
      double cutoff_ani = std::max(0.0, double((1 - j2md(bestJaccardNumerator / Q.sketchSize, param.kmerSize)) - param.ANIDiff));
      double cutoff_j = md2j(1 - cutoff_ani, param.kmerSize);
      if (double(candidateLocus.intersectionSize) / Q.sketchSize < cutoff_j) break;
    }
    l2_vec.clear();
    computeL2MappedRegions(c, Q, candidateLocus, l2_vec);
    for (auto &l2 : l2_vec) {
      float mash_dist = j2md(1.0 * l2.sharedSketchSize / Q.sketchSize, param.kmerSize);
      float nucIdentity = (1 - mash_dist);
      float nucIdentityUpperBound = 1 - md_lower_bound(mash_dist, Q.sketchSize, param.kmerSize, CONFIDENCE);
      if ((param.keep_low_pct_id && nucIdentityUpperBound >= param.percentageIdentity) || nucIdentity >= param.percentageIdentity) {
        bestJaccardNumerator = std::max<double>(bestJaccardNumerator, l2.sharedSketchSize);
        orc_mapping res;
        memset(&res, 0, sizeof(res));
        res.queryLen = Q.len;
        res.refStartPos = l2.meanOptimalPos;
        res.refEndPos = l2.meanOptimalPos + Q.len;
        res.queryStartPos = 0;
        res.queryEndPos = Q.len;
        res.refSeqId = l2.seqId;
        res.querySeqId = Q.seqCounter;
        res.nucIdentity = nucIdentity;
        res.nucIdentityUpperBound = nucIdentityUpperBound;
        res.sketchSize = Q.sketchSize;
        res.conservedSketches = l2.sharedSketchSize;
        res.blockLength = std::max(res.refEndPos - res.refStartPos, res.queryEndPos - res.queryStartPos);
        res.approxMatches = std::round(res.nucIdentity * res.blockLength / 100.0);
        res.strand = l2.strand;
        res.kmerComplexity = Q.kmerComplexity;
        res.selfMapFilter = ((param.skip_self || param.skip_prefix) && fullLen > c.contigLen[l2.seqId]);
        res.n_merged = 1; /* uninitialised in the reference (:1227): UB when read at :429-430; defined as 1 here */
        l2Mappings.push_back(res);
      }
    }
    if (param.stage1_topANI_filter) {
      std::pop_heap(l1_begin, l1_end, cmp);
      l1_end--;
    } else {
      loc_iterator++;
    }
  }
}

/* computeMap.hpp:755-815 */
void mapSingleQueryFrag(Ctx &c, Query &Q, int fullLen, std::vector<orc_ipoint> &intervalPoints, std::vector<orc_l1> &l1Mappings,
                        std::vector<orc_mapping> &l2Mappings)
{
  doL1Mapping(c, Q, intervalPoints, l1Mappings, nullptr);
  if (l1Mappings.size() == 0) return;
  auto cmp = [](const orc_l1 &a, const orc_l1 &b) { return a.intersectionSize < b.intersectionSize; };
  orc_l1 *l1_begin = l1Mappings.data(), *l1_end = l1Mappings.data(), *end = l1Mappings.data() + l1Mappings.size();
  while (l1_end != end) {
    if (c.p.skip_prefix) {
      int currGroup = c.refIdGroup[l1_begin->seqId];
      l1_end = std::find_if_not(l1_begin, end, [&](const orc_l1 &cand) { return currGroup == c.refIdGroup[cand.seqId]; });
    } else {
      l1_end = end;
    }
    if (c.p.stage1_topANI_filter) std::make_heap(l1_begin, l1_end, cmp);
    doL2Mapping(c, Q, fullLen, l1_begin, l1_end, l2Mappings);
    l1_begin = l1_end;
  }
  std::sort(l2Mappings.begin(), l2Mappings.end(),
            [](const orc_mapping &a, const orc_mapping &b) { return std::tie(a.refSeqId, a.refStartPos) < std::tie(b.refSeqId, b.refStartPos); });
}

/* ---- filter.hpp:102-160 (query axis) and :333-394 (reference axis) ---- */
struct QOrder {
  std::vector<orc_mapping> *vec;
  bool operator()(int x, int y) const
  {
    double xs = (*vec)[x].nucIdentity, ys = (*vec)[y].nucIdentity;
    return std::tie(xs, (*vec)[x].queryStartPos, (*vec)[x].refSeqId) > std::tie(ys, (*vec)[y].queryStartPos, (*vec)[y].refSeqId);
  }
};
void filterQuery(std::vector<orc_mapping> &m, int secondaryToKeep)
{
  if (m.size() <= 1) return;
  for (auto &e : m) e.discard = 1;
  QOrder ord{&m};
  std::set<int, QOrder> bst(ord);
  typedef std::tuple<int, int, int> Ev;
  std::vector<Ev> ev(2 * m.size()); /* 2n zero tuples first (filter.hpp:122), SURVEY A.9 */
  for (int i = 0; i < (int)m.size(); i++) { ev.emplace_back(m[i].queryStartPos, 1, i); ev.emplace_back(m[i].queryEndPos, 2, i); }
  std::sort(ev.begin(), ev.end());
  for (auto it = ev.begin(); it != ev.end();) {
    auto it2 = std::find_if(it, ev.end(), [&](const Ev &e) { return std::get<0>(e) != std::get<0>(*it); });
    std::for_each(it, it2, [&](const Ev &e) { if (std::get<1>(e) == 1) bst.insert(std::get<2>(e)); else bst.erase(std::get<2>(e)); });
    int kept = 0; /* markGood, filter.hpp:69-93 */
    for (auto s = bst.begin(); s != bst.end(); s++) {
      bool lower = (double)m[*bst.begin()].nucIdentity > (double)m[*s].nucIdentity;
      if ((lower || m[*s].discard == 0) && kept > secondaryToKeep) break;
      m[*s].discard = 0;
      ++kept;
    }
    it = it2;
  }
  m.erase(std::remove_if(m.begin(), m.end(), [](orc_mapping &e) { return e.discard == 1; }), m.end());
}
struct ROrder {
  std::vector<orc_mapping> *vec;
  bool operator()(int x, int y) const
  {
    double xs = (*vec)[x].nucIdentity, ys = (*vec)[y].nucIdentity;
    return std::tie(xs, (*vec)[x].refStartPos) > std::tie(ys, (*vec)[y].refStartPos);
  }
};
void filterRef(const Ctx &c, std::vector<orc_mapping> &m, int secondaryToKeep)
{
  if (m.size() <= 1) return;
  for (auto &e : m) e.discard = 1;
  ROrder ord{&m};
  std::set<int, ROrder> bst(ord);
  typedef std::tuple<int, int, int, int> Ev;
  std::vector<Ev> ev(2 * m.size());
  for (int i = 0; i < (int)m.size(); i++) {
    ev.emplace_back(m[i].refSeqId, m[i].refStartPos, 1, i);
    Ev end = std::make_tuple(m[i].refSeqId, m[i].refEndPos, 2, i);
    if (std::get<1>(end) == c.contigLen[std::get<0>(end)] - 1) { std::get<0>(end) += 1; std::get<1>(end) = 0; }
    else std::get<1>(end) += 1;
    ev.push_back(end);
  }
  std::sort(ev.begin(), ev.end());
  for (auto it = ev.begin(); it != ev.end();) {
    auto it2 = std::find_if(it, ev.end(), [&](const Ev &e) { return std::tie(std::get<0>(e), std::get<1>(e)) != std::tie(std::get<0>(*it), std::get<1>(*it)); });
    std::for_each(it, it2, [&](const Ev &e) { if (std::get<2>(e) == 1) bst.insert(std::get<3>(e)); else bst.erase(std::get<3>(e)); });
    int kept = 0; /* markGood, filter.hpp:289-304 */
    for (auto s = bst.begin(); s != bst.end(); s++) {
      bool lower = (double)m[*bst.begin()].nucIdentity > (double)m[*s].nucIdentity;
      if ((lower || m[*s].discard == 0) && ++kept > secondaryToKeep) break;
      m[*s].discard = 0;
    }
    it = it2;
  }
  m.erase(std::remove_if(m.begin(), m.end(), [](orc_mapping &e) { return e.discard == 1; }), m.end());
}

/* computeMap.hpp:504-561 */
void filterByGroup(const Ctx &c, std::vector<orc_mapping> &unf, std::vector<orc_mapping> &fil, int n_mappings, bool filter_ref)
{
  std::sort(unf.begin(), unf.end(), [](const orc_mapping &a, const orc_mapping &b) { return std::tie(a.refSeqId, a.refStartPos) < std::tie(b.refSeqId, b.refStartPos); });
  auto sb = unf.begin(), se = unf.begin();
  if (c.p.filterMode == 1 || c.p.filterMode == 2) {
    std::vector<orc_mapping> tmp;
    while (se != unf.end()) {
      if (c.p.skip_prefix) {
        int g = c.refIdGroup[sb->refSeqId];
        se = std::find_if_not(sb, unf.end(), [&](const orc_mapping &x) { return g == c.refIdGroup[x.refSeqId]; });
      } else se = unf.end();
      tmp.insert(tmp.end(), sb, se);
      std::sort(tmp.begin(), tmp.end(), [](const orc_mapping &a, const orc_mapping &b) { return std::tie(a.queryStartPos, a.refSeqId, a.refStartPos) < std::tie(b.queryStartPos, b.refSeqId, b.refStartPos); });
      if (filter_ref) filterRef(c, tmp, (uint16_t)n_mappings); else filterQuery(tmp, (uint16_t)n_mappings);
      fil.insert(fil.end(), tmp.begin(), tmp.end());
      tmp.clear();
      sb = se;
    }
  }
  std::sort(fil.begin(), fil.end(), [](const orc_mapping &a, const orc_mapping &b) { return std::tie(a.queryStartPos, a.refSeqId, a.refStartPos) < std::tie(b.queryStartPos, b.refSeqId, b.refStartPos); });
}

/* dset64.hpp:62-124 */
struct DSU {
  std::vector<uint64_t> parent, rank;
  explicit DSU(size_t n) : parent(n), rank(n, 0) { std::iota(parent.begin(), parent.end(), 0); }
  uint64_t find(uint64_t x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; }
  void unite(uint64_t a, uint64_t b)
  {
    a = find(a); b = find(b);
    if (a == b) return;
    uint64_t r1 = rank[a], r2 = rank[b];
    if (r1 > r2 || (r1 == r2 && a < b)) { std::swap(r1, r2); std::swap(a, b); }
    parent[a] = b;
    if (r1 == r2) rank[b] = r2 + 1;
  }
};

/* computeMap.hpp:1579-1704 */
void mergeMappingsInRange(std::vector<orc_mapping> &rm, int max_dist)
{
  if (rm.size() < 2) return;
  std::sort(rm.begin(), rm.end(), [](const orc_mapping &a, const orc_mapping &b) { return std::tie(a.refSeqId, a.refStartPos, a.queryStartPos) < std::tie(b.refSeqId, b.refStartPos, b.queryStartPos); });
  for (size_t i = 0; i < rm.size(); i++) { rm[i].splitMappingId = i; rm[i].discard = 0; }
  DSU ds(rm.size());
  for (auto it = rm.begin(); it != rm.end(); it++) {
    std::vector<std::pair<double, uint64_t>> distances;
    for (auto it2 = std::next(it); it2 != rm.end(); it2++) {
      if (it2->refSeqId != it->refSeqId || it2->refStartPos > it->refEndPos + max_dist) break;
      if (it2->strand == it->strand) {
        int ref_dist = it2->refStartPos - it->refEndPos;
        int query_dist = 0;
        double dist = std::numeric_limits<double>::max(), score = std::numeric_limits<double>::max();
        if (it->strand == FWD && it->queryStartPos <= it2->queryStartPos) {
          query_dist = it2->queryStartPos - it->queryEndPos;
          dist = std::sqrt(std::pow(query_dist, 2) + std::pow(ref_dist, 2));
          score = std::pow(query_dist - ref_dist, 2);
        } else if (it->strand != FWD && it->queryEndPos >= it2->queryEndPos) {
          query_dist = it->queryStartPos - it2->queryEndPos;
          dist = std::sqrt(std::pow(query_dist, 2) + std::pow(ref_dist, 2));
          score = std::pow(query_dist - ref_dist, 2);
        }
        if (dist < max_dist) distances.push_back(std::make_pair(dist + score, (uint64_t)it2->splitMappingId));
      }
    }
    if (distances.size()) {
      std::sort(distances.begin(), distances.end());
      ds.unite(it->splitMappingId, distances.front().second);
    }
  }
  for (auto &m : rm) m.splitMappingId = ds.find(m.splitMappingId);
  std::sort(rm.begin(), rm.end(), [](const orc_mapping &a, const orc_mapping &b) { return a.splitMappingId < b.splitMappingId; });
  for (auto it = rm.begin(); it != rm.end();) {
    auto it_end = std::find_if(it, rm.end(), [&](const orc_mapping &e) { return e.splitMappingId != it->splitMappingId; });
    std::for_each(it, it_end, [&](orc_mapping &e) {
      it->queryStartPos = std::min(it->queryStartPos, e.queryStartPos);
      it->refStartPos = std::min(it->refStartPos, e.refStartPos);
      it->queryEndPos = std::max(it->queryEndPos, e.queryEndPos);
      it->refEndPos = std::max(it->refEndPos, e.refEndPos);
      it->blockLength = std::max(it->refEndPos - it->refStartPos, it->queryEndPos - it->queryStartPos);
      it->approxMatches = std::round(it->nucIdentity * it->blockLength / 100.0);
    });
    it->n_merged = std::distance(it, it_end);
    it->nucIdentity = (std::accumulate(it, it_end, 0.0, [](double x, orc_mapping &e) { return x + e.nucIdentity; })) / it->n_merged;
    it->kmerComplexity = (std::accumulate(it, it_end, 0.0, [](double x, orc_mapping &e) { return x + e.kmerComplexity; })) / it->n_merged;
    std::for_each(std::next(it), it_end, [&](orc_mapping &e) { e.discard = 1; });
    it = it_end;
  }
  rm.erase(std::remove_if(rm.begin(), rm.end(), [](orc_mapping &e) { return e.discard == 1; }), rm.end());
}

Query makeQuery(const char *seq, int len, int seqCounter, int nameId, int refGroup)
{
  Query Q;
  Q.seq.assign(seq, seq + len);
  Q.len = len; Q.seqCounter = seqCounter; Q.nameId = nameId; Q.refGroup = refGroup;
  return Q;
}

/* computeMap.hpp:570-714 (mapModule) */
void mapModule(Ctx &c, const char *seq, int len, int seqCounter, int nameId, int refGroup, std::vector<orc_mapping> &out)
{
  const orc_params &param = c.p;
  std::vector<orc_mapping> unfiltered, l2Mappings;
  std::vector<orc_ipoint> intervalPoints;
  std::vector<orc_l1> l1Mappings;
  bool split_mapping = true;
  if (!param.split || len <= param.segLength) {
    Query Q = makeQuery(seq, len, seqCounter, nameId, refGroup);
    mapSingleQueryFrag(c, Q, len, intervalPoints, l1Mappings, l2Mappings);
    unfiltered.insert(unfiltered.end(), l2Mappings.begin(), l2Mappings.end());
    split_mapping = false;
  } else {
    int noOverlapFragmentCount = len / param.segLength;
    for (int i = 0; i < noOverlapFragmentCount; i++) {
      Query Q = makeQuery(seq + i * param.segLength, param.segLength, seqCounter, nameId, refGroup);
      intervalPoints.clear(); l1Mappings.clear(); l2Mappings.clear();
      mapSingleQueryFrag(c, Q, len, intervalPoints, l1Mappings, l2Mappings);
      for (auto &e : l2Mappings) { e.queryLen = len; e.queryStartPos = i * param.segLength; e.queryEndPos = i * param.segLength + Q.len; }
      unfiltered.insert(unfiltered.end(), l2Mappings.begin(), l2Mappings.end());
    }
    if (noOverlapFragmentCount >= 1 && len % param.segLength != 0) {
      Query Q = makeQuery(seq + len - param.segLength, param.segLength, seqCounter, nameId, refGroup);
      intervalPoints.clear(); l1Mappings.clear(); l2Mappings.clear();
      mapSingleQueryFrag(c, Q, len, intervalPoints, l1Mappings, l2Mappings);
      for (auto &e : l2Mappings) { e.queryLen = len; e.queryStartPos = len - param.segLength; e.queryEndPos = len; }
      unfiltered.insert(unfiltered.end(), l2Mappings.begin(), l2Mappings.end());
    }
  }
  int n_mappings = (len < param.segLength ? param.numMappingsForShortSequence : param.numMappingsForSegment) - 1;
  if (split_mapping && param.mergeMappings) {
    mergeMappingsInRange(unfiltered, param.chain_gap);
    int64_t min_count = std::floor(param.block_length / param.segLength);
    unfiltered.erase(std::remove_if(unfiltered.begin(), unfiltered.end(), [&](orc_mapping &e) { return e.queryLen > e.blockLength && e.n_merged < min_count; }),
                     unfiltered.end());
  }
  if (param.filterMode == 1 || param.filterMode == 2) {
    std::vector<orc_mapping> tmp;
    filterByGroup(c, unfiltered, tmp, n_mappings, false);
    unfiltered = std::move(tmp);
  }
  out.swap(unfiltered);
  if (param.filterLengthMismatches) { /* :441-454 */
    out.erase(std::remove_if(out.begin(), out.end(), [&](orc_mapping &e) {
                int64_t q_l = (int64_t)e.queryEndPos - (int64_t)e.queryStartPos;
                int64_t r_l = (int64_t)e.refEndPos + 1 - (int64_t)e.refStartPos;
                uint64_t delta = std::abs(r_l - q_l);
                float len_id_bound = (1.0 - (float)delta / (float)q_l);
                return len_id_bound < std::min(0.7, std::pow(param.percentageIdentity, 3));
              }), out.end());
  }
  for (auto &e : out) { /* :1713-1750 */
    int rlen = c.contigLen[e.refSeqId];
    if (e.refStartPos < 0) e.refStartPos = 0;
    if (e.refStartPos >= rlen) e.refStartPos = rlen - 1;
    if (e.refEndPos < e.refStartPos) e.refEndPos = e.refStartPos;
    if (e.refEndPos >= rlen) e.refEndPos = rlen - 1;
    if (e.queryStartPos < 0) e.queryStartPos = 0;
    if (e.queryStartPos >= len) e.queryStartPos = len;
    if (e.queryEndPos < e.queryStartPos) e.queryEndPos = e.queryStartPos;
    if (e.queryEndPos >= len) e.queryEndPos = len;
  }
}

} // namespace

/* ------------------------------------------------------------------------------------------------ */

ORC_API uint64_t orc_hash(const char *seq, int k) { return getHash(seq, k); }

ORC_API int orc_sketch_sequence(const char *seq, int len, int k, int s, int seqId, orc_minmer *out, int cap)
{
  std::vector<char> buf(seq, seq + len);
  std::vector<orc_minmer> v;
  sketchSequence(v, buf.data(), len, k, s, seqId);
  if ((int)v.size() > cap) return -(int)v.size();
  if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(orc_minmer));
  return (int)v.size();
}

ORC_API int orc_min_hits(int s, int k, float pi) { return estimateMinimumHitsRelaxed(s, k, pi, CONFIDENCE); }

ORC_API void *orc_create(const orc_params *p)
{
  Ctx *c = new Ctx();
  c->p = *p;
  setProbs(*c);
  return c;
}
ORC_API void orc_destroy(void *cv) { delete (Ctx *)cv; }

ORC_API int orc_cutoffs(void *cv, int *out, int cap)
{
  Ctx *c = (Ctx *)cv;
  for (int i = 0; i < (int)c->sketchCutoffs.size() && i < cap; i++) out[i] = c->sketchCutoffs[i];
  return (int)c->sketchCutoffs.size();
}

/* index content as the reference holds it after Sketch::Sketch (winSketch.hpp:122-138) */
ORC_API void orc_set_index(void *cv, const orc_minmer *mi, uint64_t n_mi, const uint64_t *keys, const uint64_t *offs, uint64_t n_keys,
                           const orc_ipoint *pts, const uint8_t *is_freq, const int32_t *contig_len, const char **contig_names,
                           const int32_t *contig_group, int n_contigs)
{
  Ctx *c = (Ctx *)cv;
  c->minmerIndex.assign(mi, mi + n_mi);
  c->keys.assign(keys, keys + n_keys);
  c->offs.assign(offs, offs + n_keys + 1);
  c->pts.assign(pts, pts + offs[n_keys]);
  c->isFreq.assign(is_freq, is_freq + n_keys);
  c->contigLen.assign(contig_len, contig_len + n_contigs);
  c->contigName.clear();
  for (int i = 0; i < n_contigs; i++) c->contigName.push_back(contig_names ? contig_names[i] : std::to_string(i));
  c->refIdGroup.assign(n_contigs, 0);
  if (contig_group) c->refIdGroup.assign(contig_group, contig_group + n_contigs);
}

/* stage dump of one fragment, same shape as refh_map_fragment in ref_harness.cpp */
ORC_API int orc_map_fragment(void *cv, const char *seq, int len, int fullLen, int seqCounter, int nameId, int refGroup,
                             orc_minmer *sketch, int *n_sketch, float *kmerComplexity, int *raw_count, uint64_t *raw_max_hash,
                             orc_ipoint *ip, int64_t ip_cap, int64_t *n_ip, int *minimumHits,
                             orc_l1 *l1, int l1_cap, int *n_l1, orc_l2 *l2, int *l2_cand, int l2_cap, int *n_l2,
                             orc_mapping *maps, int map_cap, int *n_maps)
{
  Ctx &c = *(Ctx *)cv;
  int rc = 0;
  {
    std::vector<char> buf(seq, seq + len);
    std::vector<orc_minmer> raw;
    sketchSequence(raw, buf.data(), len, c.p.kmerSize, c.p.sketchSize, seqCounter);
    *raw_count = (int)raw.size();
    *raw_max_hash = raw.empty() ? 0 : raw.back().hash;
  }
  {
    Query Q = makeQuery(seq, len, seqCounter, nameId, refGroup);
    std::vector<orc_ipoint> points;
    std::vector<orc_l1> cands;
    doL1Mapping(c, Q, points, cands, minimumHits);
    *n_sketch = (int)Q.minmerTableQuery.size();
    if (*n_sketch) memcpy(sketch, Q.minmerTableQuery.data(), (size_t)*n_sketch * sizeof(orc_minmer));
    *kmerComplexity = Q.kmerComplexity;
    *n_ip = (int64_t)points.size();
    if (!points.empty()) memcpy(ip, points.data(), std::min<size_t>(points.size(), (size_t)ip_cap) * sizeof(orc_ipoint));
    if ((int64_t)points.size() > ip_cap) rc |= 1;
    *n_l1 = (int)cands.size();
    for (int i = 0; i < (int)cands.size() && i < l1_cap; i++) l1[i] = cands[i];
    if ((int)cands.size() > l1_cap) rc |= 2;
    int nl2 = 0;
    for (int ci = 0; ci < (int)cands.size(); ci++) {
      std::vector<orc_l2> loci;
      computeL2MappedRegions(c, Q, cands[ci], loci);
      for (auto &l : loci) {
        if (nl2 < l2_cap) { l2[nl2] = l; l2_cand[nl2] = ci; } else rc |= 4;
        nl2++;
      }
    }
    *n_l2 = nl2;
  }
  {
    Query Q = makeQuery(seq, len, seqCounter, nameId, refGroup);
    std::vector<orc_ipoint> points;
    std::vector<orc_l1> cands;
    std::vector<orc_mapping> res;
    mapSingleQueryFrag(c, Q, fullLen, points, cands, res);
    *n_maps = (int)res.size();
    for (int i = 0; i < (int)res.size() && i < map_cap; i++) maps[i] = res[i];
    if ((int)res.size() > map_cap) rc |= 8;
  }
  return rc;
}

/* cpu_baseline / --impl reference driver: maps n_reads reads of read_len bases (back to back in `bases`) with
 * `threads` worker threads, one read per task like the reference's pool (ThreadPool.hpp:176-215; mapModule is the
 * task, computeMap.hpp:275,340). Returns the number of mappings; *mapped_reads = reads with >= 1 mapping. */
ORC_API int64_t orc_map_reads_mt(void *cv, const char *bases, int64_t n_reads, int read_len, int first_seq_counter, int threads,
                                 int64_t *mapped_reads, int32_t *rows_out, int64_t cap_rows)
{
  Ctx &c0 = *(Ctx *)cv;
  for (int s = 1; s <= c0.p.sketchSize; s++) minimumHitsFor(c0, s); /* fill the memo before the threads read it */
  std::atomic<int64_t> next{0}, total{0}, mapped{0};
  /* optional copy of the mappings for the full-scale parity diff of bench.py: kept per read (read order is the
   * reference's output order), flattened after the threads have finished */
  std::vector<std::vector<orc_mapping>> keep(rows_out ? (size_t)n_reads : 0);
  auto work = [&]() {
    std::vector<orc_mapping> res;
    while (true) {
      const int64_t i = next.fetch_add(1);
      if (i >= n_reads) break;
      res.clear();
      mapModule(c0, bases + i * (int64_t)read_len, read_len, first_seq_counter + (int)i, -1, -1, res);
      total += (int64_t)res.size();
      if (!res.empty()) mapped++;
      if (rows_out) keep[(size_t)i] = res;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < std::max(1, threads); t++) pool.emplace_back(work);
  for (auto &th : pool) th.join();
  if (mapped_reads) *mapped_reads = mapped.load();
  if (rows_out) { /* same row layout as the product's skch_bm_results */
    int64_t n = 0;
    for (auto &v : keep)
      for (auto &m : v) {
        if (n < cap_rows) {
          int32_t *r = rows_out + n * 10;
          r[0] = m.querySeqId; r[1] = m.queryStartPos; r[2] = m.queryEndPos; r[3] = m.refSeqId; r[4] = m.refStartPos;
          r[5] = m.refEndPos; r[6] = m.strand; r[7] = m.conservedSketches; r[8] = m.blockLength;
          r[9] = (int32_t)(m.nucIdentity * 1e6f);
        }
        n++;
      }
  }
  return total.load();
}

ORC_API int orc_map_read(void *cv, const char *seq, int len, int seqCounter, int nameId, int refGroup, orc_mapping *out, int cap)
{
  Ctx &c = *(Ctx *)cv;
  std::vector<orc_mapping> res;
  mapModule(c, seq, len, seqCounter, nameId, refGroup, res);
  for (int i = 0; i < (int)res.size() && i < cap; i++) out[i] = res[i];
  return (int)res.size();
}
