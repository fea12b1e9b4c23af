#include "skch_stats.hpp"

#include <algorithm>
#include <cmath>

#include "skch_types.hpp"

namespace skch {
namespace Stat {

/* Terms of the pmf relative to the term at the mode (value 1), walked outwards with the exact ratio
 * pmf(i+1)/pmf(i) = (n-i)/(i+1) * p/(1-p); tail / total needs no normalising constant. */
double binomial_Q(unsigned int k, double p, unsigned int n)
{
  if (!(p >= 0.0 && p <= 1.0)) return std::nan("");
  if (k >= n) return 0.0;
  if (p == 0.0) return 0.0;
  if (p == 1.0) return 1.0;
  const double odds = p / (1.0 - p);
  unsigned int mode = (unsigned int)std::floor((double)(n + 1) * p);
  if (mode > n) mode = n;
  double total = 1.0, upper = (mode > k) ? 1.0 : 0.0; /* upper = sum over i > k */
  double t = 1.0;
  for (unsigned int i = mode; i < n; i++) { /* i -> i+1 */
    t *= odds * (double)(n - i) / (double)(i + 1);
    total += t;
    if (i + 1 > k) upper += t;
    /* stop when nothing that follows can change either sum: a tail that starts far above the mode (k >> mode) is tiny
     * against the total but is the whole of `upper`, so the walk goes on until it has been collected */
    if (t < total * 1e-19 && i + 1 > k && t < upper * 1e-19) break;
  }
  t = 1.0;
  for (unsigned int i = mode; i > 0; i--) { /* i -> i-1 */
    t *= (double)i / (odds * (double)(n - i + 1));
    total += t;
    if (i - 1 > k) upper += t;
    if (t < total * 1e-19) break;
  }
  return upper / total;
}

void hypergeometric_pmf_row(unsigned int n1, unsigned int n2, unsigned int t, std::vector<double> &out)
{
  out.assign((size_t)t + 1, 0.0);
  if (t > n1 + n2) return;
  const unsigned int lo = t > n2 ? t - n2 : 0;
  const unsigned int hi = t < n1 ? t : n1;
  if (lo > hi) return;
  /* ratio pmf(i+1)/pmf(i) = (n1-i)(t-i) / ((i+1)(n2-t+i+1)); anchor at the mode */
  unsigned int mode = (unsigned int)std::floor(((double)t + 1.0) * ((double)n1 + 1.0) / ((double)n1 + (double)n2 + 2.0));
  mode = std::min(std::max(mode, lo), hi);
  out[mode] = 1.0;
  double total = 1.0, v = 1.0;
  for (unsigned int i = mode; i < hi; i++) {
    v *= ((double)(n1 - i) * (double)(t - i)) / ((double)(i + 1) * ((double)n2 - (double)t + (double)i + 1.0));
    out[i + 1] = v;
    total += v;
  }
  v = 1.0;
  for (unsigned int i = mode; i > lo; i--) {
    v *= ((double)i * ((double)n2 - (double)t + (double)i)) / ((double)(n1 - i + 1) * (double)(t - i + 1));
    out[i - 1] = v;
    total += v;
  }
  for (unsigned int i = lo; i <= hi; i++) out[i] /= total;
}

float j2md(float j, int k)
{
  if (j == 0) return 1.0;
  if (j == 1) return 0.0;
  float mash_dist = 1 - std::pow(2 * j / (1 + j), 1.0 / k); /* float ratio, double pow, float result */
  return mash_dist;
}

float md2j(float d, int k)
{
  float sim = 1 - d;
  float jaccard = std::pow(sim, k) / (2 - std::pow(sim, k)); /* pow(float,int) is evaluated in double */
  return jaccard;
}

float md_lower_bound(float d, int s, int k, float ci)
{
  float q2 = (1.0 - ci) / 2;
  int x = std::max(int(ceil(s * md2j(d, k))), 1);
  while (x <= s) {
    double cdf_complement = binomial_Q(x - 1, md2j(d, k), s);
    if (cdf_complement < q2) {
      x--;
      break;
    }
    x++;
  }
  float jaccard = float(x) / s;
  float low_d = j2md(jaccard, k);
  return low_d;
}

int estimateMinimumHits(int s, int k, float perc_identity)
{
  float mash_dist = 1.0 - perc_identity;
  float jaccard = md2j(mash_dist, k);
  int minimumSharedMinimizers = ceil(1.0 * s * jaccard);
  return minimumSharedMinimizers;
}

int estimateMinimumHitsRelaxed(int s, int k, float perc_identity, float confidence_interval)
{
  const int first = estimateMinimumHits(s, k, perc_identity);
  int relaxed = first;
  for (int i = first; i >= 0; i--) {
    float jaccard = 1.0 * i / s;
    float d = j2md(jaccard, k);
    float d_lower = md_lower_bound(d, s, k, confidence_interval);
    float id_upper = 1.0 - d_lower;
    if (id_upper >= perc_identity) relaxed = i;
    else break;
  }
  return relaxed;
}

double estimate_pvalue(int s, int k, int alphabetSize, float identity, int64_t lengthQuery, uint64_t lengthReference,
                       float confidence_interval)
{
  double kmerSpace = pow(alphabetSize, k);
  double pX, pY;
  pX = pY = 1. / (1. + kmerSpace / lengthQuery);
  double r = pX * pY / (pX + pY - pX * pY);
  int x = estimateMinimumHitsRelaxed(s, k, identity, confidence_interval);
  double cdf_complement;
  if (x == 0) cdf_complement = 1.0;
  else cdf_complement = binomial_Q(x - 1, r, s);
  double pVal = lengthReference * cdf_complement;
  return pVal;
}

int64_t recommendedSketchSize(double pValue_cutoff, float confidence_interval, int k, int alphabetSize, float identity,
                              int64_t segmentLength, uint64_t lengthReference)
{
  int64_t lengthQuery = segmentLength - k;
  int optimalSketchSize;
  for (optimalSketchSize = 10; optimalSketchSize < lengthQuery; optimalSketchSize += 10) {
    double pVal = estimate_pvalue(optimalSketchSize, k, alphabetSize, identity, lengthQuery, lengthReference,
                                  confidence_interval);
    if (pVal <= pValue_cutoff) break;
  }
  return optimalSketchSize;
}

std::vector<int> sketchCutoffs(int sketchSize, int kmerSize, float ANIDiff, float ANIDiffConf, bool enabled)
{
  const int ss = std::min<double>(sketchSize, fixed::ss_table_max);
  std::vector<int> cutoffs((size_t)ss + 1, 1);
  if (!enabled) return cutoffs;
  const float deltaANI = ANIDiff;
  const float min_p = 1 - ANIDiffConf;

  /* pmf[ci][y] = P(y | n1 = ss, n2 = ss - ci, t = ci) and its running sum (the cdf the reference asks
   * GSL for, computeMap.hpp:213) */
  std::vector<std::vector<double>> pmf((size_t)ss + 1), cdf((size_t)ss + 1);
  for (int ci = 0; ci <= ss; ci++) {
    hypergeometric_pmf_row(ss, ss - ci, ci, pmf[ci]);
    cdf[ci].resize(pmf[ci].size());
    double acc = 0;
    for (size_t y = 0; y < pmf[ci].size(); y++) { acc += pmf[ci][y]; cdf[ci][y] = acc > 1.0 ? 1.0 : acc; }
  }
  auto distDiff = [&](int cmax, int ci) { /* computeMap.hpp:199-226 */
    double prAboveCutoff = 0;
    for (double ymax = 0; ymax <= cmax; ymax++) {
      double pymax = pmf[cmax][(size_t)ymax];
      double yi_cutoff = deltaANI == 0 ? ymax : (std::floor(md2j(j2md(ymax / ss, kmerSize) + deltaANI, kmerSize) * ss));
      double pi_acc = 0;
      if ((yi_cutoff - 1) >= 0) {
        const unsigned int kk = (unsigned int)(yi_cutoff - 1);
        pi_acc = kk >= (unsigned int)ci ? 1.0 : cdf[ci][kk]; /* k >= min(t, n1) -> 1 */
      }
      pi_acc = 1 - pi_acc;
      prAboveCutoff += pymax * pi_acc;
      if (prAboveCutoff > min_p) return true;
    }
    return prAboveCutoff > min_p;
  };
  for (int cmax = 1; cmax <= ss; cmax++) {
    /* std::upper_bound(range[0..ss), false, (val, ci) -> distDiff(cmax, ci)) with libstdc++'s probe order */
    int first = 0, len = ss;
    while (len > 0) {
      const int half = len >> 1;
      const int middle = first + half;
      if (distDiff(cmax, middle)) len = half;
      else { first = middle + 1; len = len - half - 1; }
    }
    cutoffs[cmax] = first;
    if (cutoffs[cmax] == 0) cutoffs[cmax] = 1;
  }
  return cutoffs;
}

}  // namespace Stat
}  // namespace skch
