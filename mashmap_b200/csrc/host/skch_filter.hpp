/*
 * skch_filter.hpp -- plane-sweep "best mapping per position" filters over the query axis and the reference
 * axis. Restates skch::Filter (reference src/map/include/filter.hpp:29-396) including its quirks
 * (SURVEY A.9): the event vector starts with 2n zero tuples that erase id 0 from the (still empty) sweep
 * status before any BEGIN; the query-axis marker keeps all score ties; the reference-axis marker
 * pre-increments its counter.
 */
#ifndef SKCH_FILTER_HPP
#define SKCH_FILTER_HPP

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <thread>
#include <tuple>
#include <vector>

#include "skch_types.hpp"

namespace skch {
namespace Filter {

namespace query {

struct Order {  // filter.hpp:37-57: descending (score, queryStartPos, refSeqId)
  const MappingResultsVector_t *vec;
  bool operator()(int x, int y) const
  {
    const double xs = (*vec)[x].nucIdentity, ys = (*vec)[y].nucIdentity;
    return std::tie(xs, (*vec)[x].queryStartPos, (*vec)[x].refSeqId) > std::tie(ys, (*vec)[y].queryStartPos, (*vec)[y].refSeqId);
  }
};

inline void filterMappings(MappingResultsVector_t &readMappings, int secondaryToKeep)
{  // liFilterAlgorithm, filter.hpp:102-160
  if (readMappings.size() <= 1) return;
  for (auto &e : readMappings) e.discard = 1;
  Order ord{&readMappings};
  std::set<int, Order> status(ord);
  typedef std::tuple<offset_t, int, int> Event;
  std::vector<Event> events(2 * readMappings.size());
  for (int i = 0; i < (int)readMappings.size(); i++) {
    events.emplace_back(readMappings[i].queryStartPos, event::BEGIN, i);
    events.emplace_back(readMappings[i].queryEndPos, event::END, i);
  }
  std::sort(events.begin(), events.end());
  for (auto it = events.begin(); it != events.end();) {
    auto it2 = std::find_if(it, events.end(), [&](const Event &e) { return std::get<0>(e) != std::get<0>(*it); });
    std::for_each(it, it2, [&](const Event &e) {
      if (std::get<1>(e) == event::BEGIN) status.insert(std::get<2>(e));
      else status.erase(std::get<2>(e));
    });
    // markGood, filter.hpp:69-93
    if (!status.empty()) {
      const int beg = *status.begin();
      int kept = 0;
      for (auto s = status.begin(); s != status.end(); s++) {
        const bool lower = (double)readMappings[beg].nucIdentity > (double)readMappings[*s].nucIdentity;
        if ((lower || readMappings[*s].discard == 0) && kept > secondaryToKeep) break;
        readMappings[*s].discard = 0;
        ++kept;
      }
    }
    it = it2;
  }
  readMappings.erase(std::remove_if(readMappings.begin(), readMappings.end(), [](MappingResult &e) { return e.discard == 1; }),
                     readMappings.end());
}

}  // namespace query

namespace ref {

struct Order {  // filter.hpp:252-270: descending (score, refStartPos)
  const MappingResultsVector_t *vec;
  bool operator()(int x, int y) const
  {
    const double xs = (*vec)[x].nucIdentity, ys = (*vec)[y].nucIdentity;
    return std::tie(xs, (*vec)[x].refStartPos) > std::tie(ys, (*vec)[y].refStartPos);
  }
};

inline void filterMappings(MappingResultsVector_t &readMappings, const std::vector<ContigInfo> &metadata, int secondaryToKeep)
{  // filter.hpp:333-394
  if (readMappings.size() <= 1) return;
  for (auto &e : readMappings) e.discard = 1;
  Order ord{&readMappings};
  std::set<int, Order> status(ord);
  typedef std::tuple<seqno_t, offset_t, int, int> Event;
  std::vector<Event> events(2 * readMappings.size());
  for (int i = 0; i < (int)readMappings.size(); i++) {
    events.emplace_back(readMappings[i].refSeqId, readMappings[i].refStartPos, event::BEGIN, i);
    Event end = std::make_tuple(readMappings[i].refSeqId, readMappings[i].refEndPos, event::END, i);
    // refPosDoPlusOne, filter.hpp:311-324
    if (std::get<1>(end) == metadata[std::get<0>(end)].len - 1) {
      std::get<0>(end) += 1;
      std::get<1>(end) = 0;
    } else {
      std::get<1>(end) += 1;
    }
    events.push_back(end);
  }
  std::sort(events.begin(), events.end());
  for (auto it = events.begin(); it != events.end();) {
    auto it2 = std::find_if(it, events.end(), [&](const Event &e) {
      return std::tie(std::get<0>(e), std::get<1>(e)) != std::tie(std::get<0>(*it), std::get<1>(*it));
    });
    std::for_each(it, it2, [&](const Event &e) {
      if (std::get<2>(e) == event::BEGIN) status.insert(std::get<3>(e));
      else status.erase(std::get<3>(e));
    });
    // markGood, filter.hpp:289-304
    if (!status.empty()) {
      const int beg = *status.begin();
      int kept = 0;
      for (auto s = status.begin(); s != status.end(); s++) {
        const bool lower = (double)readMappings[beg].nucIdentity > (double)readMappings[*s].nucIdentity;
        if ((lower || readMappings[*s].discard == 0) && ++kept > secondaryToKeep) break;
        readMappings[*s].discard = 0;
      }
    }
    it = it2;
  }
  readMappings.erase(std::remove_if(readMappings.begin(), readMappings.end(), [](MappingResult &e) { return e.discard == 1; }),
                     readMappings.end());
}

/*
 * The same sweep on `threads` threads (with -f one-to-one the reference runs this filter once, on the main
 * thread, over ALL mappings of the run, computeMap.hpp:358-405 -- at 8 GPUs that is longer than the mapping itself).
 * Events are ordered by (refSeqId, position, ...), an END event carries the mapping's own refSeqId (or the next one, at
 * position 0, for a mapping that ends on the last base of its contig), and within one group of equal (refSeqId, position)
 * BEGINs (1) come before ENDs (2). So the sweep status holds mappings of one contig at a time and contigs are
 * independent sub-problems -- with one exception that is kept exact: while the BEGINs of group (c, 0) are inserted, the
 * mappings of contig c-1 that end on its last base are still in the status, and std::set refuses an element that is
 * equivalent (same identity, same refStartPos) to one it holds. That can only involve a mapping of contig c-1 that starts
 * at 0 and ends at len-1; such a contig is swept together with its successor.
 * Mapping ids (positions in readMappings) are kept, so ties between events break as in the serial sweep; the 2n zero
 * events of the reference only erase from an empty status and are not needed here.
 */
inline void filterMappingsParallel(MappingResultsVector_t &readMappings, const std::vector<ContigInfo> &metadata, int secondaryToKeep,
                                   int threads, bool erase_discarded = true)
{
  const size_t n = readMappings.size();
  if (n <= 1) return;
  if (threads <= 1 || n < 4096) {
    filterMappings(readMappings, metadata, secondaryToKeep);  // erases
    return;
  }
  const bool trace_ = getenv("MM_TRACE") != nullptr;
  auto tt_ = std::chrono::steady_clock::now();
  auto lap_ = [&](const char *w) { if (!trace_) return; auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[trace]     sweep %s: %.1f ms\n", w, std::chrono::duration<double, std::milli>(t - tt_).count()); tt_ = t; };
  const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, n / 8192));
  auto in_slices = [&](auto fn) {  // fn(t, lo, hi) over [0, n) on T threads
    std::vector<std::thread> pool;
    for (int t = 1; t < T; t++) pool.emplace_back([&, t] { fn(t, n * (size_t)t / (size_t)T, n * (size_t)(t + 1) / (size_t)T); });
    fn(0, 0, n / (size_t)T);
    for (auto &th : pool) th.join();
  };
  /* mappings by contig: a counting sort that keeps id order inside a contig (slice t's ids of contig c follow slice t-1's) */
  const size_t nc = metadata.size();
  std::vector<std::vector<size_t>> cnt((size_t)T, std::vector<size_t>(nc, 0));
  in_slices([&](int t, size_t lo, size_t hi) {
    std::vector<size_t> &c = cnt[(size_t)t];
    for (size_t i = lo; i < hi; i++) { readMappings[i].discard = 1; c[(size_t)readMappings[i].refSeqId]++; }
  });
  std::vector<size_t> start(nc + 2, 0);
  for (size_t c = 0; c < nc; c++) {
    size_t at = start[c];
    for (int t = 0; t < T; t++) { const size_t k = cnt[(size_t)t][c]; cnt[(size_t)t][c] = at; at += k; }  // count -> first slot
    start[c + 1] = at;
  }
  start[nc + 1] = start[nc];
  std::vector<int> ids(n);
  in_slices([&](int t, size_t lo, size_t hi) {
    std::vector<size_t> &at = cnt[(size_t)t];
    for (size_t i = lo; i < hi; i++) ids[at[(size_t)readMappings[i].refSeqId]++] = (int)i;
  });
  /* sweep units: runs of contigs; contig c joins c-1 when c-1 has a mapping covering [0, len-1] */
  std::vector<uint8_t> link(nc, 0);
  {
    std::atomic<size_t> next_c{1};
    auto find_links = [&]() {
      while (true) {
        const size_t c = next_c.fetch_add(1);
        if (c >= nc) break;
        if (start[c + 1] == start[c]) continue;
        for (size_t j = start[c - 1]; j < start[c]; j++) {
          const MappingResult &m = readMappings[ids[j]];
          if (m.refStartPos == 0 && m.refEndPos == metadata[c - 1].len - 1) { link[c] = 1; break; }
        }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; t++) pool.emplace_back(find_links);
    find_links();
    for (auto &th : pool) th.join();
  }
  std::vector<std::pair<size_t, size_t>> units;  // [first contig, last contig]
  for (size_t c = 0; c < nc; c++) {
    if (link[c] && !units.empty() && units.back().second == c - 1) units.back().second = c;
    else units.emplace_back(c, c);
  }
  lap_("prepare");
  std::atomic<size_t> next{0};
  auto work = [&]() {
    Order ord{&readMappings};
    typedef std::tuple<seqno_t, offset_t, int, int> Event;
    std::vector<Event> events;
    while (true) {
      const size_t u = next.fetch_add(1);
      if (u >= units.size()) break;
      const size_t j0 = start[units[u].first], j1 = start[units[u].second + 1];
      if (j0 == j1) continue;
      events.clear();
      for (size_t j = j0; j < j1; j++) {
        const int i = ids[j];
        const MappingResult &m = readMappings[i];
        events.emplace_back(m.refSeqId, m.refStartPos, event::BEGIN, i);
        Event end = std::make_tuple(m.refSeqId, m.refEndPos, event::END, i);
        if (std::get<1>(end) == metadata[std::get<0>(end)].len - 1) { std::get<0>(end) += 1; std::get<1>(end) = 0; }
        else std::get<1>(end) += 1;
        events.push_back(end);
      }
      std::sort(events.begin(), events.end());
      std::set<int, Order> status(ord);
      for (auto it = events.begin(); it != events.end();) {
        auto it2 = std::find_if(it, events.end(), [&](const Event &e) {
          return std::tie(std::get<0>(e), std::get<1>(e)) != std::tie(std::get<0>(*it), std::get<1>(*it));
        });
        std::for_each(it, it2, [&](const Event &e) {
          if (std::get<2>(e) == event::BEGIN) status.insert(std::get<3>(e));
          else status.erase(std::get<3>(e));
        });
        if (!status.empty()) {
          const int beg = *status.begin();
          int kept = 0;
          for (auto s = status.begin(); s != status.end(); s++) {
            const bool lower = (double)readMappings[beg].nucIdentity > (double)readMappings[*s].nucIdentity;
            if ((lower || readMappings[*s].discard == 0) && ++kept > secondaryToKeep) break;
            readMappings[*s].discard = 0;
          }
        }
        it = it2;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; t++) pool.emplace_back(work);
  work();
  for (auto &th : pool) th.join();
  lap_("units");
  if (erase_discarded) {  // a caller that sorts next can drop the discarded records while it gathers (sortLikeStd)
    readMappings.erase(std::remove_if(readMappings.begin(), readMappings.end(), [](MappingResult &e) { return e.discard == 1; }),
                       readMappings.end());
    lap_("erase");
  }
}

}  // namespace ref
}  // namespace Filter
}  // namespace skch
#endif
