/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * Stage-dump harness around the UNMODIFIED reference (marbl/MashMap v3.1.3).
 * This file contains no reference code: it #includes the reference headers from
 * where they lie (/root/reference/src, passed with -I by oracle/Makefile) and calls
 * the reference's own functions, including private members of skch::Map (reached
 * with the `#define private public` trick placed AFTER the standard headers).
 * Output: oracle/_ref/libmm_ref.so (git-ignored, travels to the GPU box).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.
 */
#include <algorithm>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstring>
#include <deque>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <queue>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <pthread.h>
#include <zlib.h>

#define private public
#include "map/include/map_parameters.hpp"
#include "map/include/base_types.hpp"
#include "map/include/parseCmdArgs.hpp"
#include "map/include/winSketch.hpp"
#include "map/include/computeMap.hpp"
#include "common/argvparser.hpp"
#undef private

#include "mm_oracle_types.h"

namespace {

struct Handle {
  skch::Parameters params;
  std::vector<std::string> querySequences;
  std::string outFileName;
  std::unique_ptr<skch::Sketch> sketch;
  std::unique_ptr<skch::Map> map;
  std::unique_ptr<progress_meter::ProgressMeter> progress;
  // flattened lookup index (built on demand)
  std::vector<uint64_t> keys, offs;
  std::vector<orc_ipoint> pts;
  std::vector<int> is_freq;
};

static_assert(sizeof(skch::MinmerInfo) == sizeof(orc_minmer), "MinmerInfo layout");
static_assert(sizeof(skch::IntervalPoint) == sizeof(orc_ipoint), "IntervalPoint layout");

void flatten(const skch::MappingResult &m, orc_mapping &o) {
  o.queryLen = m.queryLen; o.refStartPos = m.refStartPos; o.refEndPos = m.refEndPos;
  o.queryStartPos = m.queryStartPos; o.queryEndPos = m.queryEndPos;
  o.refSeqId = m.refSeqId; o.querySeqId = m.querySeqId; o.blockLength = m.blockLength;
  o.nucIdentity = m.nucIdentity; o.nucIdentityUpperBound = m.nucIdentityUpperBound;
  o.sketchSize = m.sketchSize; o.conservedSketches = m.conservedSketches;
  o.strand = m.strand; o.approxMatches = m.approxMatches; o.n_merged = m.n_merged;
  o.splitMappingId = m.splitMappingId; o.discard = m.discard; o.selfMapFilter = m.selfMapFilter;
  o.kmerComplexity = (double)m.kmerComplexity;
}

void unflatten(const orc_mapping &o, skch::MappingResult &m) {
  m.queryLen = o.queryLen; m.refStartPos = o.refStartPos; m.refEndPos = o.refEndPos;
  m.queryStartPos = o.queryStartPos; m.queryEndPos = o.queryEndPos;
  m.refSeqId = o.refSeqId; m.querySeqId = o.querySeqId; m.blockLength = o.blockLength;
  m.nucIdentity = o.nucIdentity; m.nucIdentityUpperBound = o.nucIdentityUpperBound;
  m.sketchSize = o.sketchSize; m.conservedSketches = o.conservedSketches;
  m.strand = o.strand; m.approxMatches = o.approxMatches; m.n_merged = o.n_merged;
  m.splitMappingId = o.splitMappingId; m.discard = o.discard; m.selfMapFilter = o.selfMapFilter;
  m.kmerComplexity = o.kmerComplexity;
}

void copy_minmers(const std::vector<skch::MinmerInfo> &v, orc_minmer *out) {
  for (size_t i = 0; i < v.size(); i++) {
    out[i].hash = v[i].hash; out[i].wpos = v[i].wpos; out[i].wpos_end = v[i].wpos_end;
    out[i].seqId = v[i].seqId; out[i].strand = v[i].strand; out[i]._pad = 0;
  }
}

void copy_points(const std::vector<skch::IntervalPoint> &v, orc_ipoint *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    memset(&out[i], 0, sizeof(orc_ipoint));
    out[i].pos = v[i].pos; out[i].hash = v[i].hash; out[i].seqId = v[i].seqId; out[i].side = v[i].side;
  }
}

} // namespace

#pragma GCC visibility push(default)
extern "C" {

/* argv as for the mashmap CLI, without the program name. The query list is stashed so the
 * Map constructor (which maps everything, computeMap.hpp:138) has nothing to do. */
void *refh_open(int argc, const char **argv)
{
  std::vector<char *> av;
  static char prog[] = "mashmap";
  av.push_back(prog);
  for (int i = 0; i < argc; i++) av.push_back(const_cast<char *>(argv[i]));
  CommandLineProcessing::ArgvParser cmd;
  skch::initCmdParser(cmd);
  Handle *h = new Handle();
  skch::parseandSave((int)av.size(), av.data(), cmd, h->params);
  h->querySequences = h->params.querySequences;
  h->outFileName = h->params.outFileName;
  h->params.querySequences.clear();
  h->params.outFileName = "/dev/null";
  h->sketch.reset(new skch::Sketch(h->params));
  h->map.reset(new skch::Map(h->params, *h->sketch));
  h->progress.reset(new progress_meter::ProgressMeter(1, "[refh]"));
  return h;
}

/* parse only: Parameters from the command line, no Sketch / Map (the reference file is not read) */
void *refh_parse(int argc, const char **argv)
{
  std::vector<char *> av;
  static char prog[] = "mashmap";
  av.push_back(prog);
  for (int i = 0; i < argc; i++) av.push_back(const_cast<char *>(argv[i]));
  CommandLineProcessing::ArgvParser cmd;
  skch::initCmdParser(cmd);
  Handle *h = new Handle();
  skch::parseandSave((int)av.size(), av.data(), cmd, h->params);
  return h;
}

void refh_close(void *hv)
{
  Handle *h = (Handle *)hv;
  if (!h) return;
  if (h->progress) { h->progress->completed.store(h->progress->total); h->progress->logger.join(); }
  delete h;
}

void refh_params(void *hv, orc_params *o)
{
  const skch::Parameters &p = ((Handle *)hv)->params;
  memset(o, 0, sizeof(*o));
  o->kmerSize = p.kmerSize; o->segLength = p.segLength; o->sketchSize = p.sketchSize;
  o->alphabetSize = p.alphabetSize; o->percentageIdentity = p.percentageIdentity;
  o->filterMode = p.filterMode; o->numMappingsForSegment = p.numMappingsForSegment;
  o->numMappingsForShortSequence = p.numMappingsForShortSequence;
  o->block_length = p.block_length; o->chain_gap = p.chain_gap; o->split = p.split;
  o->mergeMappings = p.mergeMappings; o->stage1_topANI_filter = p.stage1_topANI_filter;
  o->ANIDiff = p.ANIDiff; o->ANIDiffConf = p.ANIDiffConf; o->stage2_full_scan = p.stage2_full_scan;
  o->keep_low_pct_id = p.keep_low_pct_id; o->kmer_pct_threshold = p.kmer_pct_threshold;
  o->kmerComplexityThreshold = p.kmerComplexityThreshold; o->skip_self = p.skip_self;
  o->skip_prefix = p.skip_prefix; o->prefix_delim = p.prefix_delim;
  o->lower_triangular = p.lower_triangular; o->filterLengthMismatches = p.filterLengthMismatches;
  o->legacy_output = p.legacy_output; o->report_ANI_percentage = p.report_ANI_percentage;
  o->sparsity_hash_threshold = p.sparsity_hash_threshold; o->referenceSize = p.referenceSize;
}

int refh_n_contigs(void *hv) { return (int)((Handle *)hv)->sketch->metadata.size(); }
const char *refh_contig_name(void *hv, int i) { return ((Handle *)hv)->sketch->metadata[i].name.c_str(); }
int refh_contig_len(void *hv, int i) { return ((Handle *)hv)->sketch->metadata[i].len; }

/* minmerIndex after dropFreqSeedSet (winSketch.hpp:497-504) */
int64_t refh_index_size(void *hv) { return (int64_t)((Handle *)hv)->sketch->minmerIndex.size(); }
const orc_minmer *refh_index_data(void *hv)
{
  return reinterpret_cast<const orc_minmer *>(((Handle *)hv)->sketch->minmerIndex.data());
}

/* minmerPosLookupIndex flattened with keys ascending: keys[n], offs[n+1], pts[offs[n]], is_freq[n] */
int64_t refh_lookup_build(void *hv)
{
  Handle *h = (Handle *)hv;
  if (!h->keys.empty()) return (int64_t)h->keys.size();
  const auto &m = h->sketch->minmerPosLookupIndex;
  h->keys.reserve(m.size());
  for (const auto &e : m) h->keys.push_back(e.first);
  std::sort(h->keys.begin(), h->keys.end());
  h->offs.assign(h->keys.size() + 1, 0);
  for (size_t i = 0; i < h->keys.size(); i++) h->offs[i + 1] = h->offs[i] + m.find(h->keys[i])->second.size();
  h->pts.resize(h->offs.back());
  h->is_freq.resize(h->keys.size());
  for (size_t i = 0; i < h->keys.size(); i++) {
    const auto &v = m.find(h->keys[i])->second;
    copy_points(v, h->pts.data() + h->offs[i], v.size());
    h->is_freq[i] = h->sketch->isFreqSeed(h->keys[i]) ? 1 : 0;
  }
  return (int64_t)h->keys.size();
}
const uint64_t *refh_lookup_keys(void *hv) { return ((Handle *)hv)->keys.data(); }
const uint64_t *refh_lookup_offs(void *hv) { return ((Handle *)hv)->offs.data(); }
const orc_ipoint *refh_lookup_pts(void *hv) { return ((Handle *)hv)->pts.data(); }
const int *refh_lookup_isfreq(void *hv) { return ((Handle *)hv)->is_freq.data(); }

int refh_freq_threshold(void *hv) { return ((Handle *)hv)->sketch->getFreqThreshold(); }
int refh_is_freq(void *hv, uint64_t hash) { return ((Handle *)hv)->sketch->isFreqSeed(hash) ? 1 : 0; }

int refh_cutoffs(void *hv, int *out, int cap)
{
  const auto &c = ((Handle *)hv)->map->sketchCutoffs;
  for (int i = 0; i < (int)c.size() && i < cap; i++) out[i] = c[i];
  return (int)c.size();
}

/* ---- context-free reference functions ---- */

uint64_t refh_hash(const char *seq, int k) { return skch::CommonFunc::getHash(seq, k); }

void refh_normalise(char *seq, int len) { skch::CommonFunc::makeUpperCaseAndValidDNA(seq, len); }

int refh_min_hits(int s, int k, float pi) {
  return skch::Stat::estimateMinimumHitsRelaxed(s, k, pi, skch::fixed::confidence_interval);
}
float refh_j2md(float j, int k) { return skch::Stat::j2md(j, k); }
float refh_md2j(float d, int k) { return skch::Stat::md2j(d, k); }
float refh_md_lower_bound(float d, int s, int k) {
  return skch::Stat::md_lower_bound(d, s, k, skch::fixed::confidence_interval);
}
int64_t refh_recommended_sketch_size(int k, float pi, int64_t segLength, uint64_t refSize) {
  return skch::Stat::recommendedSketchSize(skch::fixed::pval_cutoff, skch::fixed::confidence_interval,
                                           k, 4, pi, segLength, refSize);
}

int refh_sketch_sequence(const char *seq, int len, int k, int s, int seqId, orc_minmer *out, int cap)
{
  std::string buf(seq, len);
  std::vector<skch::MinmerInfo> v;
  skch::CommonFunc::sketchSequence(v, &buf[0], len, k, 4, s, seqId);
  if ((int)v.size() > cap) return -(int)v.size();
  copy_minmers(v, out);
  return (int)v.size();
}

int64_t refh_add_minmers(const char *seq, int64_t len, int k, int w, int s, int seqId, orc_minmer *out, int64_t cap)
{
  std::string buf(seq, len);
  std::vector<skch::MinmerInfo> v;
  skch::CommonFunc::addMinmers(v, &buf[0], (skch::offset_t)len, k, w, 4, s, seqId);
  if ((int64_t)v.size() > cap) return -(int64_t)v.size();
  copy_minmers(v, out);
  return (int64_t)v.size();
}

/* ---- per-fragment stage dump: mapSingleQueryFrag (computeMap.hpp:755-815) and its pieces ---- */
int refh_map_fragment(void *hv, const char *name, const char *seq, int len, int fullLen, int seqCounter,
                      orc_minmer *sketch, int *n_sketch, float *kmerComplexity,
                      orc_ipoint *ip, int64_t ip_cap, int64_t *n_ip,
                      int *minimumHits,
                      orc_l1 *l1, int l1_cap, int *n_l1,
                      orc_l2 *l2, int *l2_cand, int l2_cap, int *n_l2,
                      orc_mapping *maps, int map_cap, int *n_maps)
{
  Handle *h = (Handle *)hv;
  skch::Map &M = *h->map;
  typedef skch::QueryMetaData<skch::Sketch::MI_Type> Q_t;
  std::string buf(seq, len);
  int rc = 0;
  {
    Q_t Q;
    Q.seq = &buf[0]; Q.len = len; Q.fullLen = fullLen; Q.seqCounter = seqCounter; Q.seqName = name;
    Q.refGroup = M.getRefGroup(Q.seqName);
    Q.sketchSize = 0; Q.kmerComplexity = 0;
    std::vector<skch::IntervalPoint> points;
    std::vector<skch::Map::L1_candidateLocus_t> cands;
    M.doL1Mapping(Q, points, cands);
    *n_sketch = (int)Q.minmerTableQuery.size();
    copy_minmers(Q.minmerTableQuery, sketch);
    *kmerComplexity = Q.kmerComplexity;
    *n_ip = (int64_t)points.size();
    copy_points(points, ip, std::min<size_t>(points.size(), (size_t)ip_cap));
    if ((int64_t)points.size() > ip_cap) rc |= 1;
    *minimumHits = Q.sketchSize > 0
        ? skch::Stat::estimateMinimumHitsRelaxed(Q.sketchSize, h->params.kmerSize, h->params.percentageIdentity,
                                                 skch::fixed::confidence_interval)
        : 0;
    *n_l1 = (int)cands.size();
    for (int i = 0; i < (int)cands.size() && i < l1_cap; i++) {
      l1[i].seqId = cands[i].seqId; l1[i].rangeStartPos = cands[i].rangeStartPos;
      l1[i].rangeEndPos = cands[i].rangeEndPos; l1[i].intersectionSize = cands[i].intersectionSize;
    }
    if ((int)cands.size() > l1_cap) rc |= 2;
    int nl2 = 0;
    for (int c = 0; c < (int)cands.size(); c++) {
      std::vector<skch::Map::L2_mapLocus_t> loci;
      M.computeL2MappedRegions(Q, cands[c], loci);
      for (auto &l : loci) {
        if (nl2 < l2_cap) {
          l2[nl2].seqId = l.seqId; l2[nl2].meanOptimalPos = l.meanOptimalPos;
          l2[nl2].optimalStart = l.optimalStart; l2[nl2].optimalEnd = l.optimalEnd;
          l2[nl2].sharedSketchSize = l.sharedSketchSize; l2[nl2].strand = l.strand;
          l2_cand[nl2] = c;
        } else rc |= 4;
        nl2++;
      }
    }
    *n_l2 = nl2;
  }
  {
    std::string buf2(seq, len);
    Q_t Q;
    Q.seq = &buf2[0]; Q.len = len; Q.fullLen = fullLen; Q.seqCounter = seqCounter; Q.seqName = name;
    Q.refGroup = M.getRefGroup(Q.seqName);
    std::vector<skch::IntervalPoint> points;
    std::vector<skch::Map::L1_candidateLocus_t> cands;
    skch::MappingResultsVector_t res;
    M.mapSingleQueryFrag(Q, points, cands, res);
    *n_maps = (int)res.size();
    for (int i = 0; i < (int)res.size() && i < map_cap; i++) {
      res[i].n_merged = 0; res[i].splitMappingId = 0; res[i].discard = 0; /* uninitialised in the reference at this stage */
      flatten(res[i], maps[i]);
    }
    if ((int)res.size() > map_cap) rc |= 8;
  }
  return rc;
}

/* ---- whole read: mapModule (computeMap.hpp:570-714) ---- */
int refh_map_read(void *hv, const char *name, const char *seq, int len, int seqCounter, orc_mapping *out, int cap)
{
  Handle *h = (Handle *)hv;
  std::string s(seq, len);
  skch::InputSeqProgContainer *in = new skch::InputSeqProgContainer(s, name, seqCounter, *h->progress);
  skch::MapModuleOutput *o = h->map->mapModule(in);
  delete in;
  int n = (int)o->readMappings.size();
  for (int i = 0; i < n && i < cap; i++) flatten(o->readMappings[i], out[i]);
  delete o;
  return n;
}

/* ---- output: the reference's own reportReadMappings (computeMap.hpp:1758-1805) on caller-provided mappings, written to
 * `path` through a std::ofstream as the reference does (per-read mode: every line starts with queryName) ---- */
int refh_report_mappings(void *hv, const orc_mapping *in, int64_t n, const char *queryName, const char *path)
{
  Handle *h = (Handle *)hv;
  skch::MappingResultsVector_t v((size_t)n);
  for (int64_t i = 0; i < n; i++) unflatten(in[i], v[(size_t)i]);
  std::ofstream os(path);
  h->map->reportReadMappings(v, queryName, os);
  return os.good() ? 0 : 1;
}

/* ---- input: the reference's own reader (common/seqiter.hpp:20-111) over a file: records, bases and an FNV-1a digest of
 * every (name, sequence) pair in order, for the product's two readers to be compared with ---- */
int refh_read_file_digest(const char *path, uint64_t *n_records, uint64_t *n_bases, uint64_t *digest)
{
  uint64_t h = 1469598103934665603ULL, nr = 0, nb = 0;
  auto eat = [&h](const std::string &s) {
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
    h ^= 0xFF; h *= 1099511628211ULL;
  };
  seqiter::for_each_seq_in_file(path, {}, "", [&](const std::string &name, const std::string &seq) {
    eat(name); eat(seq); nr++; nb += seq.size();
  });
  *n_records = nr; *n_bases = nb; *digest = h;
  return 0;
}

/* ---- -f one-to-one, the run-wide step of mapQuery (computeMap.hpp:358-405) on caller-provided mappings ----
 * The step is not a function of its own in the reference (it sits at the end of mapQuery), so its statements are repeated
 * here around the reference's OWN filterByGroup (:504-561, with Filter::ref::filterMappings and its std::sort calls inside)
 * and the final std::sort with the reference's comparator, for sessions without -Y (one query group). Returns the number of
 * mappings kept; out needs room for n. */
int64_t refh_one_to_one(void *hv, const orc_mapping *in, int64_t n, orc_mapping *out)
{
  Handle *h = (Handle *)hv;
  skch::Map &M = *h->map;
  skch::MappingResultsVector_t allReadMappings((size_t)n);
  for (int64_t i = 0; i < n; i++) unflatten(in[i], allReadMappings[(size_t)i]);
  int n_mappings = M.param.numMappingsForSegment - 1;
  skch::MappingResultsVector_t tmpMappings, filteredMappings;
  tmpMappings.insert(tmpMappings.end(), std::make_move_iterator(allReadMappings.begin()), std::make_move_iterator(allReadMappings.end()));
  M.filterByGroup(tmpMappings, filteredMappings, n_mappings, true);
  allReadMappings = std::move(filteredMappings);
  std::sort(allReadMappings.begin(), allReadMappings.end(), [](const skch::MappingResult &a, const skch::MappingResult &b) {
    return std::tie(a.querySeqId, a.queryStartPos, a.refSeqId, a.refStartPos) < std::tie(b.querySeqId, b.queryStartPos, b.refSeqId, b.refStartPos);
  });
  for (size_t i = 0; i < allReadMappings.size(); i++) flatten(allReadMappings[i], out[i]);
  return (int64_t)allReadMappings.size();
}

/* cpu_baseline / --impl reference driver around the reference's own mapModule: n_reads reads of read_len bases (back to
 * back in `bases`), one read per task on `threads` threads like the reference's pool (computeMap.hpp:275,340).
 * Same contract and row layout as orc_map_reads_mt of the port. */
int64_t refh_map_reads_mt(void *hv, const char *bases, int64_t n_reads, int read_len, int first_seq_counter, int threads,
                          int64_t *mapped_reads, int32_t *rows_out, int64_t cap_rows)
{
  Handle *h = (Handle *)hv;
  std::atomic<int64_t> next{0}, total{0}, mapped{0};
  std::vector<std::vector<skch::MappingResult>> keep(rows_out ? (size_t)n_reads : 0);
  auto work = [&]() {
    while (true) {
      const int64_t i = next.fetch_add(1);
      if (i >= n_reads) break;
      std::string s(bases + i * (int64_t)read_len, (size_t)read_len);
      const std::string name = "q" + std::to_string(first_seq_counter + i);
      skch::InputSeqProgContainer *in = new skch::InputSeqProgContainer(s, name, first_seq_counter + (int)i, *h->progress);
      skch::MapModuleOutput *o = h->map->mapModule(in);
      delete in;
      total += (int64_t)o->readMappings.size();
      if (!o->readMappings.empty()) mapped++;
      if (rows_out) keep[(size_t)i] = o->readMappings;
      delete o;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < std::max(1, threads); t++) pool.emplace_back(work);
  for (auto &th : pool) th.join();
  if (mapped_reads) *mapped_reads = mapped.load();
  if (rows_out) {
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

} // extern "C"
#pragma GCC visibility pop
