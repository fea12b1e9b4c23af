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
                                   int threads)
{
  const size_t n = readMappings.size();
  if (n <= 1) return;
  if (threads <= 1 || n < 4096) { filterMappings(readMappings, metadata, secondaryToKeep); return; }
  const bool trace_ = getenv("MM_TRACE") != nullptr;
  auto tt_ = std::chrono::steady_clock::now();
  auto lap_ = [&](const char *w) { if (!trace_) return; auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[trace]     sweep %s: %.1f ms\n", w, std::chrono::duration<double, std::milli>(t - tt_).count()); tt_ = t; };
  for (auto &e : readMappings) e.discard = 1;
  /* mappings by contig (counting sort keeps id order inside a contig) */
  const size_t nc = metadata.size();
  std::vector<size_t> start(nc + 2, 0);
  for (auto &e : readMappings) start[(size_t)e.refSeqId + 1]++;
  for (size_t c = 0; c < nc; c++) start[c + 1] += start[c];
  std::vector<int> ids(n);
  {
    std::vector<size_t> at(start.begin(), start.end() - 1);
    for (size_t i = 0; i < n; i++) ids[at[(size_t)readMappings[i].refSeqId]++] = (int)i;
  }
  /* sweep units: runs of contigs; contig c joins c-1 when c-1 has a mapping covering [0, len-1] */
  std::vector<std::pair<size_t, size_t>> units;  // [first contig, last contig]
  for (size_t c = 0; c < nc; c++) {
    bool link = false;
    if (c > 0 && start[c + 1] > start[c])
      for (size_t j = start[c - 1]; j < start[c] && !link; j++) {
        const MappingResult &m = readMappings[ids[j]];
        link = m.refStartPos == 0 && m.refEndPos == metadata[c - 1].len - 1;
      }
    if (link && !units.empty() && units.back().second == c - 1) units.back().second = c;
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
  readMappings.erase(std::remove_if(readMappings.begin(), readMappings.end(), [](MappingResult &e) { return e.discard == 1; }),
                     readMappings.end());
  lap_("erase");
}

}  // namespace ref
}  // namespace Filter
}  // namespace skch
#endif
