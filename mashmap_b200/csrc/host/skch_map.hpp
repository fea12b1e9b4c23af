/*
 * skch_map.hpp -- skch::Map: maps the query sequences on the device-resident reference index.
 *
 * Same constructor contract as the reference class (reference src/map/include/computeMap.hpp:53-139):
 * constructing it runs the whole mapping and writes param.outFileName; the optional callback is invoked
 * once per reported mapping, in output order, from the constructing thread (:100, :1802-1803).
 *
 * What moved to the GPU: everything mapSingleQueryFrag does per fragment up to the L2 loci
 * (computeMap.hpp:755-815 -> mm_map_segments of the C ABI). What stays on the host, restated:
 *   the identity / confidence-bound test and HG early break of doL2Mapping       (:1181-1267)
 *   read segmentation and query-coordinate rewriting of mapModule                (:587-672)
 *   mergeMappingsInRange, filterWeakMappings, filterByGroup + plane sweeps,
 *   filterFalseHighIdentity, mappingBoundarySanityCheck, sparsifyMappings        (:423-561, :1579-1750)
 *   reportReadMappings (PAF text)                                                (:1758-1805)
 * Reads are batched (param.batch_bases query bases per device call); the per-read host tail of a batch
 * runs on param.threads worker threads; output order == input order, as in the reference.
 */
#ifndef SKCH_MAP_HPP
#define SKCH_MAP_HPP

#include <functional>
#include <string>
#include <vector>

#include "skch_index.hpp"
#include "skch_types.hpp"

struct mm_ctx;

namespace skch {

class Map {
 public:
  struct L1_candidateLocus_t {  // computeMap.hpp:58-68
    seqno_t seqId;
    offset_t rangeStartPos;
    offset_t rangeEndPos;
    int intersectionSize;
  };
  struct L2_mapLocus_t {  // computeMap.hpp:76-84
    seqno_t seqId;
    offset_t meanOptimalPos;
    offset_t optimalStart;
    offset_t optimalEnd;
    int sharedSketchSize;
    strand_t strand;
  };
  typedef std::function<void(const MappingResult &)> PostProcessResultsFn_t;  // computeMap.hpp:100

  Map(const Parameters &p, const Sketch &refsketch, PostProcessResultsFn_t f = nullptr);  // :123-139
  ~Map();

  static void insertL2ResultsToVec(MappingResultsVector_t &v, const MappingResult &r) { v.push_back(r); }  // :1813

  // timing of the last run (seconds), for the driver program
  double secondsDevice = 0, secondsHostTail = 0, secondsInput = 0;
  uint64_t totalQueryBases = 0;

 private:
  struct Impl;
  Impl *impl;
};

}  // namespace skch
#endif
