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
#include <set>
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

}  // namespace ref
}  // namespace Filter
}  // namespace skch
#endif
