#include "skch_map.hpp"
#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <sstream>
#include <thread>
#include <tuple>

#include "../../../include/mashmap_b200_nccl.h"
#include "skch_seqio.hpp"
#include "skch_stats.hpp"

namespace skch {

namespace {

typedef std::chrono::steady_clock Clock;
double since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

[[noreturn]] void die(const std::string &msg)
{
  std::cerr << "[mashmap-b200] ERROR: " << msg << std::endl;
  exit(1);
}

std::string prefix(const std::string &s, const char c) { return s.substr(0, s.find_last_of(c)); }  // computeMap.hpp:1170-1173

}  // namespace

/* ------------------------------------------------------------------------------------------------------ */

void BatchMapper::setRefGroups()
{  // computeMap.hpp:144-161
  refIdGroup.assign(refSketch.metadata.size(), 0);
  if (!param.skip_prefix) return;
  int group = 0;
  size_t start_idx = 0, idx = 0;
  while (start_idx < refSketch.metadata.size()) {
    const auto currPrefix = prefix(refSketch.metadata[start_idx].name, param.prefix_delim);
    idx = start_idx;
    while (idx < refSketch.metadata.size() && currPrefix == prefix(refSketch.metadata[idx].name, param.prefix_delim))
      refIdGroup[idx++] = group;
    group++;
    start_idx = idx;
  }
}

int BatchMapper::getRefGroup(const std::string &seqName) const
{  // computeMap.hpp:164-177
  const auto queryPrefix = prefix(seqName, param.prefix_delim);
  for (size_t i = 0; i < refSketch.metadata.size(); i++)
    if (queryPrefix == prefix(refSketch.metadata[i].name, param.prefix_delim)) return refIdGroup[i];
  return -1;
}

/* Persistent worker threads for the per-read host tail: a part's tail is a few milliseconds of work, creating a hundred
 * threads for it costs as much again. One caller at a time. */
class BatchMapper::WorkerPool {
 public:
  explicit WorkerPool(int n)
  {
    for (int i = 0; i < n; i++) threads_.emplace_back([this, i]() { loop(i); });
  }
  ~WorkerPool()
  {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cvStart_.notify_all();
    for (auto &t : threads_) t.join();
  }
  int size() const { return (int)threads_.size(); }
  /* runs fn() on min(n, size()) workers and returns when all of them are back */
  void run(int n, const std::function<void()> &fn)
  {
    n = std::min(n, size());
    if (n <= 1) { fn(); return; }
    std::unique_lock<std::mutex> lk(mu_);
    fn_ = &fn; want_ = n; done_ = 0; gen_++;
    cvStart_.notify_all();
    cvDone_.wait(lk, [&] { return done_ == want_; });
    fn_ = nullptr;
  }

 private:
  void loop(int idx)
  {
    uint64_t seen = 0;
    while (true) {
      const std::function<void()> *fn = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cvStart_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        if (idx < want_) fn = fn_;
      }
      if (fn) {
        (*fn)();
        std::lock_guard<std::mutex> lk(mu_);
        if (++done_ == want_) cvDone_.notify_all();
      }
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cvStart_, cvDone_;
  const std::function<void()> *fn_ = nullptr;
  uint64_t gen_ = 0;
  int want_ = 0, done_ = 0;
  bool stop_ = false;
};

/* Upload chunks and the L2 phase exclude each other; a waiting L2 phase has priority over the next chunk. */
struct BatchMapper::Gate {
  std::mutex mu;
  std::condition_variable cv;
  int uploading = 0, l2_active = 0, l2_waiting = 0;
};

void BatchMapper::phaseHook(void *user, int phase, int begin)
{
  Gate &g = *static_cast<Gate *>(user);
  std::unique_lock<std::mutex> lk(g.mu);
  if (phase == MM_PHASE_UPLOAD_CHUNK) {
    if (begin) {
      g.cv.wait(lk, [&] { return g.l2_active == 0 && g.l2_waiting == 0; });
      g.uploading++;
    } else {
      g.uploading--;
      g.cv.notify_all();
    }
  } else if (phase == MM_PHASE_L2) {
    if (begin) {
      g.l2_waiting++;
      g.cv.wait(lk, [&] { return g.uploading == 0; });
      g.l2_waiting--;
      g.l2_active++;
    } else {
      g.l2_active--;
      g.cv.notify_all();
    }
  }
}

BatchMapper::BatchMapper(const Parameters &p, const Sketch &refsketch) : param(p), refSketch(refsketch)
{
  // --noSplit (param.split == false): a query no longer than a segment is one fragment with or without the option
  // (computeMap.hpp:587-607), so it is accepted; addRead stops the run at the first longer query.
  // Map::Map (computeMap.hpp:123-139): setProbs, setRefGroups; plus the per-sketch-size minimum-hit table
  sketchCutoffs = Stat::sketchCutoffs(param.sketchSize, param.kmerSize, param.ANIDiff, param.ANIDiffConf, param.stage1_topANI_filter);
  setRefGroups();
  minHits.assign((size_t)param.sketchSize + 1, 0);
  for (int s = 1; s <= param.sketchSize; s++)
    minHits[s] = Stat::estimateMinimumHitsRelaxed(s, param.kmerSize, param.percentageIdentity, fixed::confidence_interval);
  contigNameId.resize(refSketch.metadata.size());
  for (size_t i = 0; i < refSketch.metadata.size(); i++) {
    auto it = refNameId.find(refSketch.metadata[i].name);
    if (it == refNameId.end()) it = refNameId.emplace(refSketch.metadata[i].name, (int)refNameId.size()).first;
    contigNameId[i] = it->second;
  }
  mm_params mp{};
  mp.kmer_size = param.kmerSize; mp.seg_length = param.segLength; mp.sketch_size = param.sketchSize;
  mp.stage1_topani_filter = param.stage1_topANI_filter; mp.skip_self = param.skip_self;
  mp.skip_prefix = param.skip_prefix; mp.lower_triangular = param.lower_triangular;
  int rc = mm_ctx_create(param.device, &mp, &ctx);
  if (rc != MM_OK) die(std::string("mm_ctx_create: ") + mm_last_error(nullptr));
  std::vector<int32_t> clen(refSketch.metadata.size());
  for (size_t i = 0; i < clen.size(); i++) clen[i] = refSketch.metadata[i].len;
  if (refSketch.deviceBuildPending()) {
    // skch::Sketch's build / index / computeFreqHist / dropFreqSeedSet on the device (mm_index_build.cu); the log lines
    // are the reference's (winSketch.hpp:228, :403, :418-449)
    mm_index_stats st;
    auto t0 = Clock::now();
    rc = mm_index_build(ctx, refSketch.deviceText(), 0, refSketch.deviceTextOffsets().data(), (int32_t)clen.size(), contigNameId.data(),
                        refIdGroup.data(), param.kmer_pct_threshold, 0, &st);
    if (rc != MM_OK) die(std::string("mm_index_build: ") + mm_last_error(ctx) + " (--hostIndex builds the index on the host)");
    std::cerr << "[mashmap-b200::skch::Sketch::build] minmer windows picked from reference = " << st.n_minmers_before_filter << std::endl;
    std::cerr << "[mashmap-b200::skch::Sketch::index] unique minmers = " << st.n_keys << std::endl;
    if (st.n_keys) {
      std::cerr << "[mashmap-b200::skch::Sketch::computeFreqHist] Frequency histogram of minmer interval points = (" << st.hist_min_count << ", "
                << st.hist_min_keys << ") ... (" << st.hist_max_count << ", " << st.hist_max_keys << ")" << std::endl;
      if (st.freq_threshold != std::numeric_limits<int>::max())
        std::cerr << "[mashmap-b200::skch::Sketch::computeFreqHist] With threshold " << param.kmer_pct_threshold
                  << "%, ignore minmers occurring >= " << st.freq_threshold << " times during lookup." << std::endl;
      else
        std::cerr << "[mashmap-b200::skch::Sketch::computeFreqHist] With threshold " << param.kmer_pct_threshold
                  << "%, consider all minmers during lookup." << std::endl;
    } else {
      std::cerr << "[mashmap-b200::skch::Sketch::computeFreqHist] No minmers." << std::endl;
    }
    std::cerr << "[mashmap-b200::skch::Sketch] index built on the device in " << since(t0) << " s (window scan " << st.ms_scan * 1e-3
              << " s over " << st.n_chunks << " chunks, " << st.n_fixed_chunks << " re-scanned exactly; records " << st.ms_post * 1e-3
              << " s; lookup + frequency filter " << st.ms_lookup * 1e-3 << " s)" << std::endl;
    refSketch.deviceBuildDone(st.freq_threshold);
  } else {
    rc = mm_index_upload(ctx, refSketch.minmerIndex.data(), refSketch.minmerIndex.size(), refSketch.lookupKeys.data(),
                         refSketch.lookupOffsets.data(), refSketch.lookupKeys.size(), refSketch.lookupPoints.data(),
                         refSketch.lookupPoints.size(), refSketch.lookupKeyIsFreq.data(), clen.data(), contigNameId.data(),
                         refIdGroup.data(), (int32_t)clen.size());
    if (rc != MM_OK) die(std::string("mm_index_upload: ") + mm_last_error(ctx));
  }
  rc = mm_tables_upload(ctx, sketchCutoffs.data(), (int32_t)sketchCutoffs.size(), minHits.data(), (int32_t)minHits.size());
  if (rc != MM_OK) die(std::string("mm_tables_upload: ") + mm_last_error(ctx));
  tail_ = new MapTail(param, refSketch.metadata, refIdGroup);
  // one group per device; the first one owns the uploaded image, the others receive a copy over NVLink
  std::vector<int> devs = param.devices.empty() ? std::vector<int>{param.device} : param.devices;
  std::vector<mm_ctx *> others;
  for (size_t d = 0; d < devs.size(); d++) {
    DeviceGroup *g = new DeviceGroup();
    g->device = devs[d];
    if (d == 0) g->owner = ctx;
    else {
      rc = mm_ctx_create(devs[d], &mp, &g->owner);
      if (rc != MM_OK) die(std::string("mm_ctx_create (device ") + std::to_string(devs[d]) + "): " + mm_last_error(nullptr));
      others.push_back(g->owner);
    }
    groups.push_back(g);
  }
  if (!others.empty()) {
    auto t0 = Clock::now();
    rc = mm_index_replicate(ctx, others.data(), (int)others.size());
    if (rc != MM_OK) die(std::string("mm_index_replicate: ") + mm_comm_last_error(nullptr));
    std::cerr << "[mashmap-b200::skch::BatchMapper] index image replicated to " << others.size() << " more device(s) in " << since(t0)
              << " s (one grouped NCCL broadcast)" << std::endl;
  }
  const int G = (int)groups.size();
  for (DeviceGroup *g : groups) {
    // of a device's share of the host threads, three drive its pipeline (upload / kernels / fetch): with CPUs to spare they
    // spin on the device (lowest latency) and the rest run the per-read tail; with eight or fewer threads per device (one
    // process per GPU on a host with few CPUs) they sleep on blocking events instead and every thread runs the tail.
    // Measured per 400 k reads: 2 threads 258 (sleep) vs 507 ms (spin), 8 threads 92 vs 109 ms, 16 threads 110 vs 108 ms.
    const int share = std::max(1, param.threads / G);
    g->blockingWaits = getenv("MM_BLOCKING_WAIT") ? getenv("MM_BLOCKING_WAIT")[0] == '1' : share <= 8;
    g->tailThreads = g->blockingWaits ? share : std::max(1, share - 3);
    if (const char *e = getenv("MM_TAIL_THREADS")) g->tailThreads = std::max(1, atoi(e));  // experiment
    g->tailPool = new WorkerPool(g->tailThreads);
    // further contexts share the device's index image: one lane per pipeline stage in flight (upload / kernels / fetch + tail)
    g->lanes[0].ctx = g->owner;
    g->nLanes = 1;
    const int max_lanes = getenv("MM_LANES") ? std::max(1, std::min<int>(MAX_LANES, atoi(getenv("MM_LANES")))) : MAX_LANES;  // experiment
    for (int l = 1; l < max_lanes; l++) {
      mm_ctx *c2 = nullptr;
      if (mm_ctx_create(g->device, &mp, &c2) != MM_OK) break;
      if (mm_ctx_share_index(c2, g->owner) != MM_OK) { mm_ctx_destroy(c2); break; }
      g->lanes[l].ctx = c2;
      g->nLanes = l + 1;
    }
    for (int l = 0; l < g->nLanes; l++) mm_ctx_set_wait_mode(g->lanes[l].ctx, g->blockingWaits ? 1 : 0);
    if (g->nLanes > 1 && !getenv("MM_NO_GATE")) {
      g->gate = new Gate();
      for (int l = 0; l < g->nLanes; l++) mm_ctx_set_phase_hook(g->lanes[l].ctx, &BatchMapper::phaseHook, g->gate);
    }
  }
}

BatchMapper::~BatchMapper()
{
  for (DeviceGroup *g : groups) {
    delete g->tailPool;
    for (int l = 0; l < MAX_LANES; l++) { g->lanes[l].segRes.release(); g->lanes[l].cands.release(); g->lanes[l].loci.release(); }
    for (int l = g->nLanes - 1; l >= 1; l--) mm_ctx_destroy(g->lanes[l].ctx);
    if (g->owner && g->owner != ctx) mm_ctx_destroy(g->owner);
    delete g->gate;
    delete g;
  }
  delete tail_;
  if (ctx) mm_ctx_destroy(ctx);
}

char *BatchMapper::allocBases(uint64_t n_bases)
{
  char *p = nullptr;
  const uint64_t bytes = n_bases / 2 + 256;
  if (mm_host_alloc((void **)&p, bytes) != MM_OK)
    die("cannot allocate the pinned batch buffer (" + std::to_string(bytes >> 20) + " MiB)");
  return p;
}
void BatchMapper::freeBases(char *p) { if (p) mm_host_free(p); }

void BatchMapper::addRead(ReadBatch &b, const std::string &name, const char *seq, offset_t len, seqno_t seqCounter) const
{
  ReadRec rd;
  rd.name = name; rd.len = len; rd.seqCounter = seqCounter; rd.first_seg = b.segs.size();
  rd.refGroup = param.skip_prefix ? getRefGroup(name) : -1;
  int name_id = -1;
  if (param.skip_self) {
    auto it = refNameId.find(name);
    if (it != refNameId.end()) name_id = it->second;
  }
  if (seq) seqio::pack_bases(seq, (uint64_t)len, b.nibbles(b.used));
  auto push = [&](offset_t start, offset_t flen) {
    mm_segment s;
    s.offset = b.used + (uint64_t)start; s.length = flen; s.seq_counter = seqCounter; s.name_id = name_id; s.ref_group = rd.refGroup;
    b.segs.push_back(s);
  };
  if (!param.split && len > param.segLength)
    die("--noSplit: query '" + name + "' (" + std::to_string(len) + " bp) is longer than the segment length (" + std::to_string(param.segLength) +
        " bp): the B200 path maps unsplit queries up to the segment length only (fragments longer than a segment -- windowLen > 0, "
        "computeMap.hpp:933,1306 -- are not implemented on the device); raise -s or drop --noSplit");
  if (len <= param.segLength) push(0, len);  // computeMap.hpp:587-607 (with or without --noSplit)
  else {
    const int n = len / param.segLength;  // :610-641
    for (int i = 0; i < n; i++) push(i * param.segLength, param.segLength);
    if (len % param.segLength != 0) push(len - param.segLength, param.segLength);  // :644-671
  }
  rd.n_seg = (uint32_t)(b.segs.size() - rd.first_seg);
  b.used += ((uint64_t)len + ReadBatch::READ_ALIGN - 1) / ReadBatch::READ_ALIGN * ReadBatch::READ_ALIGN;
  b.reads.push_back(std::move(rd));
}

void BatchMapper::finalizeOneToOne(MappingResultsVector_t &allReadMappings, const std::vector<ContigInfo> &qmetadata, std::string &paf) const
{
  tail_->finalizeOneToOne(allReadMappings, qmetadata, paf);
}

/* The three stages of one part (reads [r0, r1) of the batch) on one lane (= one device context with its own stream
 * and buffers). mapBatch runs them as a pipeline: uploads on one thread, kernels on another, fetch + host tail on a
 * third, so that the PCIe copy of part i+1 and the host tail of part i-1 are hidden behind the kernels of part i. */
void BatchMapper::laneUpload(Lane &ln, const ReadBatch &b, size_t r0, size_t r1)
{
  auto t0 = Clock::now();
  ln.r0 = r0; ln.r1 = r1;
  ln.s0 = b.reads[r0].first_seg;
  const size_t s1 = b.reads[r1 - 1].first_seg + b.reads[r1 - 1].n_seg;
  const uint64_t b0 = b.segs[ln.s0].offset;  // a read's first fragment starts at the read's first base (a multiple of READ_ALIGN)
  const uint64_t b1 = r1 < b.reads.size() ? b.segs[b.reads[r1].first_seg].offset : b.used;
  const mm_segment *segp = b.segs.data() + ln.s0;
  if (b0 != 0) {  // fragment offsets are relative to the buffer handed to the device call
    ln.segs.assign(b.segs.begin() + ln.s0, b.segs.begin() + s1);
    for (auto &sg : ln.segs) sg.offset -= b0;
    segp = ln.segs.data();
  }
  ln.nseg = s1 - ln.s0;
  int rc = mm_batch_upload_packed(ln.ctx, b.nibbles(b0), b1 - b0, segp, ln.nseg);
  if (rc != MM_OK) die(std::string("mm_batch_upload_packed: ") + mm_last_error(ln.ctx));
  ln.msUpload = since(t0) * 1e3;
  ln.secDevice += since(t0);
}

void BatchMapper::laneCompute(Lane &ln)
{
  auto t0 = Clock::now();
  int rc = mm_map_resident(ln.ctx, &ln.nc, &ln.nl);
  if (rc != MM_OK) die(std::string("mm_map_resident: ") + mm_last_error(ln.ctx));
  ln.msCompute = since(t0) * 1e3;
  ln.secDevice += since(t0);
}

void BatchMapper::laneFinish(DeviceGroup &g, Lane &ln, const ReadBatch &b, std::vector<MappingResultsVector_t> &results,
                             std::vector<std::string> *text, const std::vector<ContigInfo> *qmetadata)
{
  const int tail_threads = g.tailThreads;
  auto t0 = Clock::now();
  const size_t r0 = ln.r0, r1 = ln.r1;
  if (ln.segRes.size() < ln.nseg) ln.segRes.reserve(ln.nseg + ln.nseg / 8 + 1024);
  if (ln.cands.size() < ln.nc) ln.cands.reserve(ln.nc + ln.nc / 8 + 1024);
  if (ln.loci.size() < ln.nl) ln.loci.reserve(ln.nl + ln.nl / 8 + 1024);
  int rc = mm_batch_fetch(ln.ctx, ln.segRes.data(), ln.cands.data(), ln.cands.size(), ln.loci.data(), ln.loci.size());
  if (rc != MM_OK) die(std::string("mm_batch_fetch: ") + mm_last_error(ln.ctx));
  mm_last_stage_ms(ln.ctx, ln.stageMs);
  const double msFetch = since(t0) * 1e3;
  ln.secDevice += since(t0);
  t0 = Clock::now();

  MapTail tail(param, refSketch.metadata, refIdGroup);
  tail.segs = b.segs.data();              // absolute fragment indices (only the lengths are read)
  tail.segRes = ln.segRes.data() - ln.s0; // so that indexing by the absolute fragment index works
  tail.cands = ln.cands.data();
  tail.loci = ln.loci.data();
  tail.qmetadata = qmetadata;
  const size_t nreads = r1 - r0;
  const int nthreads = std::max(1, std::min<int>(tail_threads, (int)((nreads + 255) / 256)));
  std::atomic<size_t> next{r0};
  auto worker = [&]() {
    /* one cache per pool thread, kept across parts and batches: filling its tables costs a binomial search per distinct
     * shared-sketch count (about a millisecond per worker), which every part used to pay again */
    static thread_local IdentityCache idc;
    idc.use(param.kmerSize, param.ANIDiff);
    uint64_t bytes = 0, mapped = 0, maps = 0;
    while (true) {
      const size_t lo = next.fetch_add(256);
      if (lo >= r1) break;
      const size_t hi = std::min(r1, lo + 256);
      for (size_t r = lo; r < hi; r++) {
        results[r].clear();
        if (text) (*text)[r].clear();
        tail.mapRead(b.reads[r], idc, results[r]);
        if (results[r].empty()) continue;
        mapped++;
        maps += results[r].size();
        if (text) {
          tail.formatMappings(results[r], b.reads[r].name, (*text)[r]);
          bytes += (*text)[r].size();
        }
      }
    }
    lastTextBytes += bytes; lastMappedReads += mapped; lastMappings += maps;
  };
  g.tailPool->run(nthreads, worker);
  ln.secTail += since(t0);
  static const bool trace = getenv("MM_TRACE") != nullptr;
  if (trace)
    fprintf(stderr, "[trace] lane %d reads %zu-%zu segs %zu: upload %.2f ms (h2d %.2f) compute %.2f ms (kernels %.2f [k1 %.2f k2 %.2f k3 %.2f: prep %.2f scan %.2f]) "
            "fetch %.2f ms (d2h %.2f) tail %.2f ms (%d threads)\n", (int)(&ln - g.lanes), r0, r1, ln.nseg, ln.msUpload, ln.stageMs[3],
            ln.msCompute, ln.stageMs[5], ln.stageMs[0], ln.stageMs[1], ln.stageMs[2], ln.stageMs[6], ln.stageMs[7], msFetch, ln.stageMs[4], since(t0) * 1e3, nthreads);
}

void BatchMapper::mapBatch(const ReadBatch &b, std::vector<MappingResultsVector_t> &results, std::vector<std::string> *text,
                           const std::vector<ContigInfo> *qmetadata)
{
  const size_t nreads = b.reads.size();
  // no up-front clearing: the tail workers reset each read's slot themselves (a million small frees on one thread
  // would cost tens of milliseconds per batch), capacity is reused from the previous batch
  results.resize(nreads);
  if (text) text->resize(nreads);
  lastTextBytes = 0; lastMappedReads = 0; lastMappings = 0;
  if (nreads == 0) return;
  double d0 = 0, t0 = 0;
  for (DeviceGroup *g : groups)
    for (auto &ln : g->lanes) { d0 += ln.secDevice; t0 += ln.secTail; }
  // parts of ~SUB bases (a read is never split across parts)
  // A large batch ends with smaller parts: the last fetch + host tail has nothing left to hide behind. It starts with a
  // half-size part: the kernels wait for the first upload only, and an upload now takes about half as long as the
  // kernels of the same part, so the second (full) part is on the device before the first one's kernels end. (When the
  // upload was barely faster than the kernels, a short first part only made the compute thread wait for the second upload.)
  uint64_t SUB = std::max<uint64_t>(param.sub_batch_bases, 1);
  if (groups.size() > 1)  // several devices: enough parts for every device's three lanes
    SUB = std::max<uint64_t>(std::min<uint64_t>(SUB, b.used / (3 * groups.size()) + 1), (uint64_t)param.segLength);
  std::vector<uint64_t> targets;
  if (b.used >= 4 * SUB) {
    uint64_t left = b.used;
    if (!getenv("MM_NO_RAMP")) { targets.push_back(SUB / 2); left -= SUB / 2; }
    while (left > SUB + SUB / 2 + SUB / 4) { targets.push_back(SUB); left -= SUB; }
    targets.push_back(std::max<uint64_t>(left * 4 / 7, 1));  // the rest in two parts, the last one the smaller
    targets.push_back(~0ULL);
  }
  std::vector<std::pair<size_t, size_t>> parts;
  {
    size_t r0 = 0;
    uint64_t acc = 0;
    for (size_t r = 0; r < nreads; r++) {
      acc += (uint64_t)b.reads[r].len;
      const uint64_t want = parts.size() < targets.size() ? targets[parts.size()] : SUB;
      if (acc >= want || r + 1 == nreads) { parts.emplace_back(r0, r + 1); r0 = r + 1; acc = 0; }
    }
  }
  const size_t G = groups.size();
  if (G == 1) {
    runGroup(*groups[0], b, parts, 0, 1, results, text, qmetadata);
  } else {  // parts dealt round robin to the devices, every device runs its own pipeline
    std::vector<std::thread> drivers;
    for (size_t g = 1; g < G; g++)
      drivers.emplace_back([&, g] { runGroup(*groups[g], b, parts, g, G, results, text, qmetadata); });
    runGroup(*groups[0], b, parts, 0, G, results, text, qmetadata);
    for (auto &t : drivers) t.join();
  }
  memcpy(lastStageMs, groups[0]->lanes[0].stageMs, sizeof(lastStageMs));
  for (DeviceGroup *g : groups)
    for (auto &ln : g->lanes) { secondsDevice += ln.secDevice; secondsHostTail += ln.secTail; }
  secondsDevice -= d0; secondsHostTail -= t0;
}

/* the parts first, first + step, ... of the batch through the three-stage pipeline of one device */
void BatchMapper::runGroup(DeviceGroup &g, const ReadBatch &b, const std::vector<std::pair<size_t, size_t>> &all_parts, size_t first,
                           size_t step, std::vector<MappingResultsVector_t> &results, std::vector<std::string> *text,
                           const std::vector<ContigInfo> *qmetadata)
{
  std::vector<std::pair<size_t, size_t>> parts;
  for (size_t i = first; i < all_parts.size(); i += step) parts.push_back(all_parts[i]);
  const size_t np = parts.size();
  if (np == 0) return;
  Lane *lanes = g.lanes;
  const int nLanes = g.nLanes;
  if (np < 2 || nLanes < 2) {
    for (auto &p : parts) {
      laneUpload(lanes[0], b, p.first, p.second);
      laneCompute(lanes[0]);
      laneFinish(g, lanes[0], b, results, text, qmetadata);
    }
  } else {
    // part i lives on lane i % nLanes; state: 0 waiting, 1 uploaded, 2 computed, 3 finished (its lane is free again)
    const size_t NL = (size_t)nLanes;
    std::vector<int> state(np, 0);
    std::mutex mu;
    std::condition_variable cv;
    auto wait_for = [&](size_t i, int st) {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return state[i] >= st; });
    };
    auto publish = [&](size_t i, int st) {
      { std::lock_guard<std::mutex> lk(mu); state[i] = st; }
      cv.notify_all();
    };
    std::thread uploader([&] {
      for (size_t i = 0; i < np; i++) {
        if (i >= NL) wait_for(i - NL, 3);
        laneUpload(lanes[i % NL], b, parts[i].first, parts[i].second);
        publish(i, 1);
      }
    });
    std::thread finisher([&] {
      for (size_t i = 0; i < np; i++) {
        wait_for(i, 2);
        laneFinish(g, lanes[i % NL], b, results, text, qmetadata);
        publish(i, 3);
      }
    });
    const auto tp0 = Clock::now();
    double waitUpload = 0, tFirst = 0, tLast = 0;
    for (size_t i = 0; i < np; i++) {  // kernels of successive parts run back to back from this thread
      const auto tw = Clock::now();
      wait_for(i, 1);
      waitUpload += since(tw);
      if (i == 0) tFirst = since(tp0);
      laneCompute(lanes[i % NL]);
      publish(i, 2);
    }
    tLast = since(tp0);
    uploader.join();
    finisher.join();
    static const bool trace = getenv("MM_TRACE") != nullptr;
    if (trace)
      fprintf(stderr, "[trace] device %d: %zu parts; first kernels start at %.1f ms, last kernels end at %.1f ms, pipeline drained at %.1f ms; "
              "compute thread waited %.1f ms for uploads\n", g.device, np, tFirst * 1e3, tLast * 1e3, since(tp0) * 1e3, waitUpload * 1e3);
  }
}

/* ------------------------------------------------------------------------------------------------------ */

struct Map::Impl {
  const Parameters &param;
  const Sketch &refSketch;
  PostProcessResultsFn_t processMappingResults;
  Map &self;
  BatchMapper bm;
  std::vector<ContigInfo> qmetadata;  // computeMap.hpp:105 (one-to-one only)
  ReadBatch batch;
  std::ofstream outstrm;
  MappingResultsVector_t allReadMappings;
  seqno_t totalReadsMapped = 0, totalReadsPicked = 0, seqCounter = 0;

  // the batch being filled by the reader (`batch`) and the one being mapped by the worker thread (`inflight`)
  ReadBatch inflight;
  std::thread worker;
  bool workerActive = false;
  std::vector<MappingResultsVector_t> results;
  std::vector<std::string> text;

  Impl(const Parameters &p, const Sketch &s, PostProcessResultsFn_t f, Map &m, Clock::time_point tCtor = Clock::now())
      : param(p), refSketch(s), processMappingResults(f), self(m), bm(p, s)
  {
    const double secIndex = since(tCtor);  // the default argument is evaluated before the members are constructed
    auto t1 = Clock::now();
    batch.capacity = param.batch_bases + (uint64_t)param.segLength + 64;
    batch.bases = bm.allocBases(batch.capacity);
    inflight.capacity = batch.capacity;
    inflight.bases = bm.allocBases(inflight.capacity);
    std::cerr << "[mashmap-b200::skch::Map] device contexts + index upload " << secIndex << " s, pinned batch buffers " << since(t1)
              << " s" << std::endl;
  }
  ~Impl()
  {
    waitWorker();
    bm.freeBases(batch.bases);
    bm.freeBases(inflight.bases);
  }

  void waitWorker()
  {
    if (workerActive) { worker.join(); workerActive = false; }
  }

  void mapAndWrite(ReadBatch &b)
  {
    const bool report_now = param.filterMode != filter::ONETOONE;
    bm.mapBatch(b, results, report_now ? &text : nullptr, &qmetadata);
    for (size_t r = 0; r < results.size(); r++) {  // mapModuleHandleOutput (computeMap.hpp:724-747), in input order
      if (!results[r].empty()) totalReadsMapped++;
      if (!report_now) allReadMappings.insert(allReadMappings.end(), results[r].begin(), results[r].end());
      else {
        outstrm << text[r];
        if (processMappingResults != nullptr)
          for (auto &e : results[r]) processMappingResults(e);
      }
    }
    b.clear();
  }

  /* hands the filled batch to the worker thread (mapping + output, in batch order) and goes on reading into the other
   * buffer; one-to-one mode and user callbacks keep the reference's "everything from the calling thread" behaviour */
  void flushBatch()
  {
    if (batch.reads.empty()) return;
    waitWorker();
    std::swap(batch, inflight);
    const bool async = param.filterMode != filter::ONETOONE && processMappingResults == nullptr && !getenv("MM_SERIAL_INPUT");
    if (async) {
      worker = std::thread([this]() { mapAndWrite(inflight); });
      workerActive = true;
    } else {
      mapAndWrite(inflight);
    }
  }

  void onSequence(const std::string &name, const std::string &seq)
  {  // the body of mapQuery's per-sequence callback (computeMap.hpp:317-349)
    const offset_t len = seq.length();
    if (param.filterMode == filter::ONETOONE) qmetadata.push_back(ContigInfo{name, len});
    if (len < param.kmerSize) {
      std::cerr << std::endl << "WARNING, skch::Map::mapQuery, read " << name << " of " << len << "bp "
                << " is not long enough for mapping at segment length " << param.segLength << std::endl;
    } else {
      totalReadsPicked++;
      // a read lives in one batch: flush first if it does not fit in what is left
      if (batch.used + (uint64_t)len > param.batch_bases && !batch.reads.empty()) flushBatch();
      if ((uint64_t)len + 64 > batch.capacity) {  // a single sequence larger than the batch buffer: grow it
        flushBatch();
        bm.freeBases(batch.bases);
        batch.capacity = (uint64_t)len + 64;
        batch.bases = bm.allocBases(batch.capacity);
      }
      bm.addRead(batch, name, seq.data(), len, seqCounter);
      self.totalQueryBases += (uint64_t)len;
    }
    seqCounter++;
  }

  /* onSequence for every record of a mapped FASTA file; the bases go from the file mapping straight into the pinned
   * batch buffer, copied by all host threads just before the batch is mapped */
  void ingestMapped(const seqio::FastaFile &ff)
  {
    struct CopyJob { size_t rec; uint64_t dst; };
    std::vector<CopyJob> jobs;
    const auto &recs = ff.records();
    auto runCopies = [&]() {
      if (jobs.empty()) return;
      const int T = std::max(1, std::min<int>(param.threads, (int)(jobs.size() / 16 + 1)));
      std::atomic<size_t> next{0};
      auto worker = [&]() {
        while (true) {
          const size_t b = next.fetch_add(64);
          if (b >= jobs.size()) break;
          const size_t e = std::min(jobs.size(), b + 64);
          for (size_t j = b; j < e; j++) ff.pack_bases(recs[jobs[j].rec], batch.nibbles(jobs[j].dst));
        }
      };
      if (T == 1) worker();
      else {
        std::vector<std::thread> pool;
        for (int t = 0; t < T; t++) pool.emplace_back(worker);
        for (auto &th : pool) th.join();
      }
      jobs.clear();
    };
    for (size_t i = 0; i < recs.size(); i++) {
      const seqio::FastaRecord &r = recs[i];
      if (r.seq_len > (uint64_t)std::numeric_limits<offset_t>::max()) {
        std::cerr << "[mashmap-b200] ERROR: sequence " << ff.name(r) << " is longer than 2^31 bases" << std::endl;
        exit(1);
      }
      const offset_t len = (offset_t)r.seq_len;
      const std::string name = ff.name(r);
      if (param.filterMode == filter::ONETOONE) qmetadata.push_back(ContigInfo{name, len});
      if (len < param.kmerSize) {
        std::cerr << std::endl << "WARNING, skch::Map::mapQuery, read " << name << " of " << len << "bp "
                  << " is not long enough for mapping at segment length " << param.segLength << std::endl;
      } else {
        totalReadsPicked++;
        if (batch.used + (uint64_t)len > param.batch_bases && !batch.reads.empty()) { runCopies(); flushBatch(); }
        if ((uint64_t)len + 64 > batch.capacity) {
          runCopies();
          flushBatch();
          bm.freeBases(batch.bases);
          batch.capacity = (uint64_t)len + 64;
          batch.bases = bm.allocBases(batch.capacity);
        }
        jobs.push_back(CopyJob{i, batch.used});
        bm.addRead(batch, name, nullptr, len, seqCounter);
        self.totalQueryBases += (uint64_t)len;
      }
      seqCounter++;
    }
    runCopies();
  }

  void mapQuery()
  {  // computeMap.hpp:263-415
    outstrm.open(param.outFileName);
    auto t0 = Clock::now();
    for (const auto &fileName : param.querySequences) {
      seqio::FastaFile ff;
      if (!getenv("MM_SERIAL_INPUT") && ff.open(fileName, param.threads)) {  // plain FASTA: bulk path
        ingestMapped(ff);
        continue;
      }
      bool ok = seqio::for_each_seq_in_file(fileName, {}, "", [&](const std::string &name, const std::string &seq) { onSequence(name, seq); });
      if (!ok) exit(1);
    }
    const double secRead = since(t0);
    flushBatch();
    waitWorker();
    std::cerr << "[mashmap-b200::skch::Map::mapQuery] input read and handed over in " << secRead << " s, last batch done at "
              << since(t0) << " s" << std::endl;
    self.secondsDevice = bm.secondsDevice;
    self.secondsHostTail = bm.secondsHostTail;
    self.secondsInput = since(t0) - self.secondsDevice - self.secondsHostTail;

    if (param.filterMode == filter::ONETOONE) {  // :358-405
      std::string paf;
      bm.finalizeOneToOne(allReadMappings, qmetadata, paf);
      outstrm << paf;
      if (processMappingResults != nullptr)
        for (auto &e : allReadMappings) processMappingResults(e);
    }
    outstrm.close();
    std::cerr << "[mashmap-b200::skch::Map::mapQuery] count of mapped reads = " << totalReadsMapped
              << ", reads qualified for mapping = " << totalReadsPicked << ", total input reads = " << seqCounter
              << ", total input bp = " << self.totalQueryBases << std::endl;
  }
};

Map::Map(const Parameters &p, const Sketch &refsketch, PostProcessResultsFn_t f) : impl(new Impl(p, refsketch, f, *this))
{
  impl->mapQuery();
}

Map::~Map() { delete impl; }

}  // namespace skch
