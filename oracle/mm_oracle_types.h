/* ORACLE / TEST INFRASTRUCTURE ONLY -- never included by the product.
 *
 * Plain-C record layouts shared by the two checkers in this directory:
 *   ref_harness.cpp  (the unmodified reference, compiled from /root/reference
 *                     into oracle/_ref/libmm_ref.so)
 *   mm_oracle.cpp    (the CPU restatement, oracle/libmm_oracle.so)
 * Layouts follow the reference's own structs (base_types.hpp:31-79,
 * computeMap.hpp:58-84) so the harness can memcpy them out.
 */
#ifndef MM_ORACLE_TYPES_H
#define MM_ORACLE_TYPES_H
#include <stdint.h>

typedef struct {            /* skch::MinmerInfo, base_types.hpp:31-63 (24 B) */
  uint64_t hash;
  int32_t wpos;
  int32_t wpos_end;
  int32_t seqId;
  int16_t strand;
  int16_t _pad;
} orc_minmer;

typedef struct {            /* skch::IntervalPoint, base_types.hpp:66-79 (24 B) */
  int32_t pos;
  int32_t _pad0;
  uint64_t hash;
  int32_t seqId;
  int8_t side;
  int8_t _pad1[3];
} orc_ipoint;

typedef struct {            /* Map::L1_candidateLocus_t, computeMap.hpp:58-68 */
  int32_t seqId;
  int32_t rangeStartPos;
  int32_t rangeEndPos;
  int32_t intersectionSize;
} orc_l1;

typedef struct {            /* Map::L2_mapLocus_t, computeMap.hpp:76-84 */
  int32_t seqId;
  int32_t meanOptimalPos;
  int32_t optimalStart;
  int32_t optimalEnd;
  int32_t sharedSketchSize;
  int32_t strand;
} orc_l2;

typedef struct {            /* skch::MappingResult, base_types.hpp:154-206, flattened */
  int32_t queryLen;
  int32_t refStartPos;
  int32_t refEndPos;
  int32_t queryStartPos;
  int32_t queryEndPos;
  int32_t refSeqId;
  int32_t querySeqId;
  int32_t blockLength;
  float nucIdentity;
  float nucIdentityUpperBound;
  int32_t sketchSize;
  int32_t conservedSketches;
  int32_t strand;
  int32_t approxMatches;
  int32_t n_merged;
  int32_t splitMappingId;
  int32_t discard;
  int32_t selfMapFilter;
  double kmerComplexity;    /* reference keeps a long double; every value stored is a float or a mean of floats */
} orc_mapping;

typedef struct {            /* the skch::Parameters fields the path reads (map_parameters.hpp:32-80) */
  int32_t kmerSize;
  int32_t segLength;
  int32_t sketchSize;
  int32_t alphabetSize;
  float percentageIdentity;
  int32_t filterMode;               /* 1 map, 2 one-to-one, 3 none */
  int32_t numMappingsForSegment;
  int32_t numMappingsForShortSequence;
  int32_t block_length;
  int32_t chain_gap;
  int32_t split;
  int32_t mergeMappings;
  int32_t stage1_topANI_filter;
  float ANIDiff;
  float ANIDiffConf;
  int32_t stage2_full_scan;
  int32_t keep_low_pct_id;
  float kmer_pct_threshold;
  float kmerComplexityThreshold;
  int32_t skip_self;
  int32_t skip_prefix;
  int32_t prefix_delim;
  int32_t lower_triangular;
  int32_t filterLengthMismatches;
  int32_t legacy_output;
  int32_t report_ANI_percentage;
  uint64_t sparsity_hash_threshold;
  uint64_t referenceSize;
} orc_params;

#endif
