/*
 * skch_stats.hpp -- Jaccard <-> Mash distance, confidence bounds, minimum-hit and sketch-size
 * estimates, and the hypergeometric L1 cut-off table (host side; the device only reads the tables).
 *
 * Mirrors skch::Stat (reference src/map/include/map_stats.hpp:45-258) and Map::setProbs
 * (computeMap.hpp:178-258) with the same float/double conversion points, because the reported
 * identity is a float computed from double pow() and must agree to the last bit.
 * The three GSL distribution functions the reference calls are replaced by own implementations
 * (mode-anchored term recurrences, no log-gamma): GSL is an un-vendored, un-pinned system library of
 * the reference, and only integer decisions depend on these values.
 */
#ifndef SKCH_STATS_HPP
#define SKCH_STATS_HPP

#include <cstdint>
#include <vector>

namespace skch {
namespace Stat {

/* P(X > k), X ~ Binomial(n, p)  (stands in for gsl_cdf_binomial_Q, map_stats.hpp:98,213) */
double binomial_Q(unsigned int k, double p, unsigned int n);
/* pmf over the whole support of the hypergeometric distribution "k successes in t draws from n1
 * successes and n2 failures" (gsl_ran_hypergeometric_pdf, computeMap.hpp:194): out[k], k in [0, t] */
void hypergeometric_pmf_row(unsigned int n1, unsigned int n2, unsigned int t, std::vector<double> &out);

float j2md(float j, int k);                                   // map_stats.hpp:45-55
float md2j(float d, int k);                                   // map_stats.hpp:63-68
float md_lower_bound(float d, int s, int k, float ci);        // map_stats.hpp:81-113
int estimateMinimumHits(int s, int k, float perc_identity);   // map_stats.hpp:122-133
int estimateMinimumHitsRelaxed(int s, int k, float perc_identity, float confidence_interval);  // :144-169
double estimate_pvalue(int s, int k, int alphabetSize, float identity, int64_t lengthQuery,
                       uint64_t lengthReference, float confidence_interval);  // :182-218
int64_t recommendedSketchSize(double pValue_cutoff, float confidence_interval, int k, int alphabetSize,
                              float identity, int64_t segmentLength, uint64_t lengthReference);  // :234-258

/* Map::setProbs (computeMap.hpp:178-258): cutoffs[cmax] = smallest L1 intersection still plausibly
 * within deltaANI of a best intersection cmax; size min(sketchSize, 1000) + 1 (computeMap.hpp:128). */
std::vector<int> sketchCutoffs(int sketchSize, int kmerSize, float ANIDiff, float ANIDiffConf,
                               bool stage1_topANI_filter);

}  // namespace Stat
}  // namespace skch

#endif
