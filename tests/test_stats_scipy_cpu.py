"""The statistics behind the integer tables of the path (minimum shared sketches, automatic sketch size, L1 cut-offs),
anchored on a THIRD implementation.

The reference calls three GNU Scientific Library functions (gsl_cdf_binomial_Q: map_stats.hpp:98,213;
gsl_ran_hypergeometric_pdf: computeMap.hpp:194; gsl_cdf_hypergeometric_P: computeMap.hpp:213). GSL is a system dependency
that is not in this image, so oracle/_ref is built against a stand-in (oracle/gsl_shim) and the product has its own
implementation (skch_stats.cpp): two codes written here. scipy.stats (Boost.Math / cephes underneath) is independent of
both. These tests restate the reference's decision procedures in Python with scipy supplying the distribution values --
the float conversions come from the product's j2md / md2j, which tests/test_host_cpu.py pins bit for bit to the reference
-- and check
  (1) the distribution values themselves (product, and through it the stand-in) against scipy;
  (2) that every integer decision (estimateMinimumHitsRelaxed, recommendedSketchSize, sketchCutoffs) comes out the same
      when scipy's values are used, and
  (3) the margin of every threshold comparison behind those decisions: how close `cdf < q2` / `pVal <= 1e-3` /
      `prAboveCutoff > min_p` come to flipping. A disagreement at a threshold would need an error of that size in GSL.
"""
import ctypes as C
import math

import numpy as np
import pytest

scipy_stats = pytest.importorskip("scipy.stats")

from mashmap_b200 import hostlib  # noqa: E402

CI = np.float32(0.95)  # fixed::confidence_interval (map_parameters.hpp:92)


def _lib():
    L = hostlib.lib()
    L.skch_binomial_Q.restype = C.c_double
    L.skch_binomial_Q.argtypes = [C.c_uint, C.c_double, C.c_uint]
    L.skch_j2md.restype = C.c_float
    L.skch_j2md.argtypes = [C.c_float, C.c_int]
    L.skch_md2j.restype = C.c_float
    L.skch_md2j.argtypes = [C.c_float, C.c_int]
    L.skch_min_hits.restype = C.c_int
    L.skch_min_hits.argtypes = [C.c_int, C.c_int, C.c_float]
    L.skch_recommended_sketch_size.restype = C.c_int64
    L.skch_recommended_sketch_size.argtypes = [C.c_int, C.c_float, C.c_int64, C.c_uint64]
    return L


def test_binomial_upper_tail_equals_scipy():
    """gsl_cdf_binomial_Q(k, p, n) = P(X > k): the product's value against scipy.stats.binom.sf on the (n, p) the path
    really asks for -- sketch sizes 10..1000, p from the Jaccard of a 0.80..0.99 identity cut-off and from the random-match
    probability of estimate_pvalue (1e-9..1e-3)"""
    L = _lib()
    worst = 0.0
    for n in (10, 20, 70, 130, 199, 220, 310, 500, 1000):
        ps = [float(L.skch_md2j(np.float32(1 - pi), 19)) for pi in (0.80, 0.85, 0.90, 0.95, 0.99)] + [1e-9, 3.6e-8, 1e-6, 1e-3, 0.5]
        for p in ps:
            ks = np.unique(np.clip(np.concatenate([np.arange(0, 12), (n * p + np.arange(-15, 16)).astype(int), [n - 2, n - 1]]), 0, n - 1))
            want = scipy_stats.binom.sf(ks, n, p)
            for k, w in zip(ks.tolist(), want.tolist()):
                got = L.skch_binomial_Q(int(k), p, n)
                err = abs(got - w) / max(w, 1e-300)
                if w > 1e-290:
                    worst = max(worst, err)
                    assert err < 2e-9, (n, p, k, got, w)
    print(f"binomial tail: worst relative difference from scipy {worst:.2e}")


def _md_lower_bound(L, d, s, k, margins):
    """map_stats.hpp:81-111 (GSL branch) with scipy's tail; `margins` collects |cdf - q2| / q2 of every comparison"""
    q2 = float(np.float32((1.0 - float(CI)) / 2))  # float q2 = (1.0 - ci) / 2
    j = np.float32(L.skch_md2j(np.float32(d), k))
    x = max(int(math.ceil(float(np.float32(s) * j))), 1)  # int * float is a float product
    while x <= s:
        cdf = float(scipy_stats.binom.sf(x - 1, s, float(j)))
        margins.append(abs(cdf - q2) / q2)
        if cdf < q2:
            x -= 1
            break
        x += 1
    return np.float32(L.skch_j2md(np.float32(np.float32(x) / np.float32(s)), k))


def _min_hits_relaxed(L, s, k, pi, margins):
    """estimateMinimumHits + estimateMinimumHitsRelaxed (map_stats.hpp:120-165)"""
    pi = np.float32(pi)
    jac = np.float32(L.skch_md2j(np.float32(1.0 - float(pi)), k))  # float mash_dist = 1.0 - perc_identity
    first = int(math.ceil(1.0 * s * float(jac)))
    best = first
    for i in range(first, -1, -1):
        d = np.float32(L.skch_j2md(np.float32(1.0 * i / s), k))
        d_lower = _md_lower_bound(L, d, s, k, margins)
        id_upper = np.float32(1.0 - float(d_lower))
        if id_upper >= pi:
            best = i
        else:
            break
    return best


@pytest.mark.parametrize("pi", [0.85, 0.90, 0.95])
def test_min_hits_with_scipy_tails_equal_the_product(pi):
    """the minimum-hits table (a7) recomputed with scipy's binomial tail: same integers, and no comparison of the search is
    closer than 1e-6 (relative) to its threshold"""
    L = _lib()
    margins = []
    for k in (19, 16):
        for s in list(range(1, 60)) + [70, 100, 130, 199, 220, 310, 400, 777, 1000]:
            assert _min_hits_relaxed(L, s, k, pi, margins) == L.skch_min_hits(s, k, pi), (s, k, pi)
    print(f"pi={pi}: {len(margins)} threshold comparisons, closest relative margin {min(margins):.3e}")
    assert min(margins) > 1e-6


def _recommended_sketch_size(L, k, pi, seg, ref_size, margins):
    """estimate_pvalue + recommendedSketchSize (map_stats.hpp:178-258)"""
    length_query = seg - k
    kmer_space = float(4 ** k)
    px = 1.0 / (1.0 + kmer_space / length_query)
    r = px * px / (px + px - px * px)
    s = 10
    while s < length_query:
        x = L.skch_min_hits(s, k, pi)  # pinned above and against the reference
        cdf = 1.0 if x == 0 else float(scipy_stats.binom.sf(x - 1, s, r))
        pval = ref_size * cdf
        margins.append(abs(pval - 1e-3) / 1e-3)
        if pval <= 1e-3:
            break
        s += 10
    return s


def test_recommended_sketch_size_with_scipy_tails_equals_the_product():
    """the automatic sketch size (220 / 20 / 70 of the BASELINE configurations, 310 for the wrapped 3 GB file size) recomputed
    with scipy's tail: same values; the p-value that stops the search is never within 1 % of the cut-off"""
    L = _lib()
    margins = []
    wrapped = (3_050_000_016 + 2**31) % 2**32 - 2**31  # the reference's int32 referenceSize for a 3.05 GB file, sign-extended
    for size in (1_200_000, 12_400_000, 100_000_000, 3_050_000_016, wrapped % 2**64):
        for pi, seg in ((0.85, 5000), (0.95, 5000), (0.90, 10000), (0.85, 1000)):
            got = L.skch_recommended_sketch_size(19, pi, seg, size)
            assert _recommended_sketch_size(L, 19, pi, seg, float(size), margins) == got, (size, pi, seg)
    print(f"{len(margins)} p-value comparisons, closest relative margin {min(margins):.3e}")
    assert min(margins) > 1e-2


def _sketch_cutoffs(L, ss, k, delta_ani, conf, margins):
    """setProbs (computeMap.hpp:180-257) with scipy's hypergeometric pmf / cdf: sketchCutoffs[cmax] = the ci that
    std::upper_bound's probe sequence ends on with the predicate distDiff(cmax, ci) = Pr(ANI_i >= ANI_max - deltaANI) > min_p"""
    min_p = float(np.float32(1) - np.float32(conf))  # float min_p = 1 - param.ANIDiffConf
    delta = np.float32(delta_ani)
    ks = np.arange(ss + 1)
    # k successes in t = ci draws from n1 = ss successes and n2 = ss - ci failures  ->  scipy's (M = n1 + n2, n = n1, N = t)
    pmf = [scipy_stats.hypergeom.pmf(ks[: ci + 1], 2 * ss - ci, ss, ci) for ci in range(ss + 1)]
    cdf = [np.minimum(1.0, scipy_stats.hypergeom.cdf(ks[: ci + 1], 2 * ss - ci, ss, ci)) for ci in range(ss + 1)]
    if delta == 0:
        cut = ks.astype(np.float64)
    else:
        cut = np.array([math.floor(float(np.float32(L.skch_md2j(np.float32(np.float32(L.skch_j2md(np.float32(y / ss), k)) + delta), k))
                                         * np.float32(ss))) for y in ks], dtype=np.float64)

    def dist_diff(cmax, ci):
        kk = cut[: cmax + 1] - 1
        acc = np.where(kk >= 0, np.where(kk >= ci, 1.0, cdf[ci][np.clip(kk, 0, ci).astype(int)]), 0.0)
        pr = float(np.sum(pmf[cmax][: cmax + 1] * (1.0 - acc)))  # the partial sums only grow: the early return changes nothing
        margins.append(abs(pr - min_p) / min_p)
        return pr > min_p

    out = [1]
    for cmax in range(1, ss + 1):
        first, length = 0, ss
        while length > 0:
            half = length >> 1
            middle = first + half
            if dist_diff(cmax, middle):
                length = half
            else:
                first, length = middle + 1, length - half - 1
        out.append(first if first else 1)
    return out


@pytest.mark.parametrize("ss,delta,conf", [(220, 0.0, 0.999), (70, 0.0, 0.999), (20, 0.0, 0.999), (199, 0.0, 0.999), (130, 0.02, 0.99)])
def test_sketch_cutoffs_with_scipy_values_equal_the_product(ss, delta, conf):
    """the L1 cut-off table (a8) recomputed with scipy's hypergeometric distribution: same integers; the closest any probed
    Pr(...) comes to min_p is reported and must exceed 1e-7 (relative)"""
    L = _lib()
    margins = []
    want = _sketch_cutoffs(L, ss, 19, delta, conf, margins)
    got = hostlib.sketch_cutoffs(ss, 19, delta, conf, True)
    assert list(map(int, got)) == want
    print(f"ss={ss}: {len(margins)} probes, closest relative margin {min(margins):.3e}")
    assert min(margins) > 1e-7


def test_hypergeometric_values_equal_scipy():
    """gsl_ran_hypergeometric_pdf(k, n1 = ss, n2 = ss - ci, t = ci): the product's pmf rows against scipy"""
    L = hostlib.lib()
    if not hasattr(L, "skch_hypergeometric_pmf_row"):
        pytest.skip("the C view does not export the pmf row")
    L.skch_hypergeometric_pmf_row.restype = C.c_int
    L.skch_hypergeometric_pmf_row.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_double), C.c_int]
    worst = 0.0
    for ss in (20, 70, 220, 1000):
        for ci in sorted({0, 1, 2, ss // 7, ss // 3, ss // 2, ss - 1, ss}):
            buf = (C.c_double * (ci + 1))()
            n = L.skch_hypergeometric_pmf_row(ss, ss - ci, ci, buf, ci + 1)
            assert n == ci + 1
            want = scipy_stats.hypergeom.pmf(np.arange(ci + 1), 2 * ss - ci, ss, ci)
            got = np.array(buf[:])
            big = want > 1e-280
            if big.any():
                err = np.abs(got[big] - want[big]) / want[big]
                worst = max(worst, float(err.max()))
                assert err.max() < 1e-8, (ss, ci, float(err.max()))
            assert abs(got.sum() - 1.0) < 1e-9
    print(f"hypergeometric pmf: worst relative difference from scipy {worst:.2e}")
