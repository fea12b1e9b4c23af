/* Driver of the host library's self-tests for the sanitizer builds (scripts/sanitize_host.sh): the run-wide one-to-one step
 * (threaded exact sorts, parallel sweep, sliced text), the sort on every input pattern, the PAF number formatter. */
#include <cstdint>
#include <cstdio>
#include <initializer_list>
extern "C" {
int64_t skch_one_to_one_selftest(int64_t n, uint64_t seed, int threads, int n_contigs, int n_queries, int span_every, double *sec_fast,
                                 double *sec_plain);
int64_t skch_sort_selftest(int64_t n, uint64_t seed, int threads, int pattern, int64_t *heap_branches);
int skch_format_selftest(int64_t n, uint64_t seed);
}
int main()
{
  long bad = 0;
  for (int span : {0, 50}) {
    double a, b;
    bad += skch_one_to_one_selftest(80000, 3, 6, 64, 40000, span, &a, &b);
  }
  for (int pat = 0; pat < 7; pat++) {
    int64_t hb;
    bad += skch_sort_selftest(70000, 9, 5, pat, &hb);
  }
  bad += skch_format_selftest(20000, 1);
  printf("differences: %ld\n", bad);
  return bad != 0;
}
