/*
 * skch_tail.hpp -- the host tail of the mapping path: everything the reference does AFTER the L2 loci of a
 * fragment are known (reference src/map/include/computeMap.hpp): the identity / confidence-bound test and
 * hypergeometric early break of doL2Mapping (:1181-1267), query-coordinate rewriting and fragment order of
 * mapModule (:587-672), mergeMappingsInRange (:1579-1704), filterWeakMappings (:423-433), filterByGroup and the
 * plane-sweep filters (:504-561, filter.hpp), filterFalseHighIdentity (:441-454),
 * mappingBoundarySanityCheck (:1713-1750), sparsifyMappings (:482-493), reportReadMappings (:1758-1805).
 * It consumes the records the device returns through the C ABI (mm_segment_result / mm_l1_candidate /
 * mm_l2_locus), so it can be exercised on the CPU with records produced by any checker.
 *
 * One defined deviation from the reference: a fragment mapping starts with n_merged = 1. The reference never
 * initialises MappingResult::n_merged (:1227) and reads it in filterWeakMappings (:429-430) when a split read
 * has exactly one fragment mapping (mergeMappingsInRange returns early, :1584): undefined behaviour whose
 * outcome differs between builds of the same source ("reference UB").
 */
#ifndef SKCH_TAIL_HPP
#define SKCH_TAIL_HPP

#include <ostream>
#include <string>
#include <unordered_map>
#include <vector>

#include "skch_types.hpp"

namespace skch {

struct ReadRec {
  std::string name;
  offset_t len;
  seqno_t seqCounter;
  uint64_t first_seg;
  uint32_t n_seg;
  int refGroup;
};

/* Pure functions of two small integers, memoised per worker thread in flat tables (one per query sketch size seen):
 *   (sharedSketchSize, Q.sketchSize) -> (nucIdentity, nucIdentityUpperBound), doL2Mapping :1211-1215;
 *   (best Jaccard numerator so far, Q.sketchSize) -> the Jaccard cut-off of the stage-1 top-ANI filter, :1195-1200
 * (two pow() per candidate otherwise: most of the tail's time per read). */
struct IdentityCache {
  int k = 19;
  float ANIDiff = 0.0f;  // Parameters::ANIDiff, part of the cut-off
  struct Table {
    std::vector<std::pair<float, float>> identity;  // by shared count; first < 0 = not computed yet
    std::vector<double> cutoff;                     // by best numerator; < 0 = not computed yet
  };
  std::unordered_map<int, Table> tables;
  Table *last = nullptr;
  int last_qs = -1;
  Table &table(int qs);
  void use(int kmerSize, float aniDiff)  // (re)binds the cache to a parameter set; the tables survive while it stays the same
  {
    if (kmerSize != k || aniDiff != ANIDiff) { tables.clear(); last = nullptr; last_qs = -1; }
    k = kmerSize; ANIDiff = aniDiff;
  }
  std::pair<float, float> get(int shared, int qs);
  double cutoffJaccard(int best, int qs);
  // per-worker scratch of MapTail::mapRead (cleared, never shrunk: no allocation per read once warm)
  MappingResultsVector_t unfiltered, l2Mappings, filtered;
  std::vector<mm_l1_candidate> work;
};

class MapTail {
 public:
  MapTail(const Parameters &p, const std::vector<ContigInfo> &meta, const std::vector<int> &groups)
      : param(p), metadata(meta), refIdGroup(groups) {}

  // the device batch the tail works on (not owned)
  const mm_segment *segs = nullptr;
  const mm_segment_result *segRes = nullptr;
  const mm_l1_candidate *cands = nullptr;
  const mm_l2_locus *loci = nullptr;
  const std::vector<ContigInfo> *qmetadata = nullptr;  // query names for one-to-one output

  void fragmentMappings(const mm_segment &sg, const mm_segment_result &sr, const ReadRec &rd, IdentityCache &idc,
                        std::vector<mm_l1_candidate> &work, MappingResultsVector_t &l2Mappings) const;
  void mergeMappingsInRange(MappingResultsVector_t &readMappings, int max_dist) const;
  void filterByGroup(MappingResultsVector_t &unfiltered, MappingResultsVector_t &filtered, int n_mappings, bool filter_ref) const;
  void mapRead(const ReadRec &rd, IdentityCache &idc, MappingResultsVector_t &out) const;
  int getRefGroup(const std::string &seqName) const;  // computeMap.hpp:164-177
  /* -f one-to-one, the run-wide step (computeMap.hpp:358-405): reference-axis sweep over ALL mappings, final order, PAF text */
  void finalizeOneToOne(MappingResultsVector_t &allReadMappings, const std::vector<ContigInfo> &qmeta, std::string &paf) const;
  void formatMappings(const MappingResultsVector_t &readMappings, const std::string &queryName, std::ostream &os) const;
  void formatMappings(const MappingResultsVector_t &readMappings, const std::string &queryName, std::string &out) const;  // appends
  void formatMappings(const MappingResult *first, size_t n, const std::string &queryName, std::string &out) const;     // appends
  void formatMappingsStream(const MappingResultsVector_t &readMappings, const std::string &queryName, std::ostream &os) const;
  /* the real-number text of formatMappings against snprintf("%g") on n values of every kind; returns the differences */
  static int64_t realTextSelftest(int64_t n, uint64_t seed);
  /* the threaded exact sort against std::sort on n (key, index) pairs of a given pattern; returns the differences */
  static int64_t sortSelftest(int64_t n, uint64_t seed, int threads, int pattern, int64_t *heap_branches);

 private:
  const Parameters &param;
  const std::vector<ContigInfo> &metadata;
  const std::vector<int> &refIdGroup;
  mutable std::vector<std::string> textParts;  // finalizeOneToOne's per-thread text, kept between calls (one caller at a time)
};

}  // namespace skch
#endif
