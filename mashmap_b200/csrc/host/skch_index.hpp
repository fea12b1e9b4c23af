/*
 * skch_index.hpp -- skch::Sketch: the reference minmer index (reference src/map/include/winSketch.hpp).
 *
 * Same constructor contract and public members as the reference class (winSketch.hpp:57-511): the
 * constructor builds and indexes (blocking); `metadata`, `minmerIndex`, the frequent-seed predicate and
 * threshold are public. The hash -> interval-point map (`minmerPosLookupIndex`, winSketch.hpp:100-101)
 * is kept flattened (keys ascending / offsets / points), which is the form the device consumes.
 *
 * Round-1 builder: minmer windows are computed on the host by a step-for-step restatement of
 * CommonFunc::addMinmers (commonFunc.hpp:301-570), one task per contig, because every record's wpos is an
 * L2 evaluation point and the reference's record boundaries (vote-sum zero crossings, chunking, the
 * unstable sort's tie order) are only reproducible by following the same steps with the same libstdc++
 * (SURVEY 7.2, A.6). A device builder is the first "next" row of SURVEY 8(f).
 */
#ifndef SKCH_INDEX_HPP
#define SKCH_INDEX_HPP

#include <limits>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "skch_types.hpp"

namespace skch {

template <class T>
struct default_init_allocator : std::allocator<T> {
  template <class U> struct rebind { typedef default_init_allocator<U> other; };
  default_init_allocator() = default;
  template <class U> default_init_allocator(const default_init_allocator<U> &) {}
  template <class U> void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new ((void *)p) U; }
  template <class U, class... A> void construct(U *p, A &&...a) { ::new ((void *)p) U(std::forward<A>(a)...); }
};
template <class T> using BigVec = std::vector<T, default_init_allocator<T>>;

namespace CommonFunc {
/* commonFunc.hpp:301-570 */
void addMinmers(std::vector<MinmerInfo> &minmerIndex, char *seq, offset_t len, int kmerSize, int windowSize,
                int alphabetSize, int sketchSize, seqno_t seqCounter, bool stable_ties = false);
/* the post-processing of addMinmers (commonFunc.hpp:522-568) over records in emission order */
void finishMinmers(std::vector<MinmerInfo> &out, int windowSize, bool stable_ties = false);
/* the chunked + stitched scan the GPU builder performs, on the host (tests): returns the number of re-scanned chunks */
int addMinmersChunked(std::vector<MinmerInfo> &out, char *seq, offset_t len, int kmerSize, int windowSize, int sketchSize,
                      seqno_t seqCounter, offset_t chunk, offset_t warm);
/* commonFunc.hpp:591-603 */
uint64_t getReferenceSize(const std::vector<std::string> &refSequences);
}  // namespace CommonFunc

class Sketch {
 public:
  typedef std::vector<MinmerInfo> MI_Type;

  explicit Sketch(const Parameters &p);  // winSketch.hpp:122-138: build + index + frequency filter
  ~Sketch();
  // same pipeline on sequences already in memory (seqs[i] has metadata[i].len bases); used by bench.py
  Sketch(const Parameters &p, const std::vector<ContigInfo> &contigs, const std::vector<const char *> &seqs);
  // index + frequency filter over an existing minmer list (winSketch.hpp:379-504 without build())
  Sketch(const Parameters &p, const std::vector<ContigInfo> &contigs, MI_Type &&minmers);

  std::vector<ContigInfo> metadata;          // winSketch.hpp:79
  std::vector<int> sequencesByFileInfo;      // winSketch.hpp:88
  MI_Type minmerIndex;                       // winSketch.hpp:102 (after dropFreqSeedSet)

  // minmerPosLookupIndex (winSketch.hpp:101), flattened: keys ascending; points of keys[i] are
  // lookupPoints[lookupOffsets[i] .. lookupOffsets[i+1]) in reference per-key order
  // (BigVec: resize() leaves trivially-constructible elements uninitialised -- these arrays are filled by all threads
  //  right away, and zero-filling 12 GB of points on one thread first cost seconds at 3 Gbp)
  BigVec<hash_t> lookupKeys;
  BigVec<uint64_t> lookupOffsets;
  BigVec<IntervalPoint> lookupPoints;
  std::vector<uint8_t> lookupKeyIsFreq;      // frequentSeeds membership per key (winSketch.hpp:488-495)

  /* The index is built ON THE DEVICE by default (mm_index_build, called by skch::BatchMapper which owns the device
   * context): the constructor then only reads the contigs; minmerIndex and the lookup arrays stay empty on the host.
   * --hostIndex, --saveIndex and --loadIndex keep everything on the host as before. */
  bool deviceBuildPending() const { return deviceText_ != nullptr; }
  const char *deviceText() const { return deviceText_; }
  const std::vector<uint64_t> &deviceTextOffsets() const { return deviceTextOffsets_; }
  void deviceBuildDone(int freq_threshold) const;  // releases the text, records the threshold

  int getFreqThreshold() const { return freqThreshold; }   // winSketch.hpp:483-486
  bool isFreqSeed(hash_t h) const;                         // winSketch.hpp:506-509
  bool isMinmerIndexEnd(MI_Type::const_iterator it) const { return it == minmerIndex.end(); }
  MI_Type::const_iterator getMinmerIndexEnd() const { return minmerIndex.end(); }

  // --saveIndex / --loadIndex (winSketch.hpp:270-374): TSV and PREFIX.index/.map binary formats
  void saveIndexTSV(const std::string &path) const;
  void saveIndexBinary(const std::string &prefix) const;
  void savePosListBinary(const std::string &prefix) const;

 private:
  const Parameters &param;
  mutable int freqThreshold = std::numeric_limits<int>::max();
  bool saving_ = false;
  mutable char *deviceText_ = nullptr;           // contigs back to back (text), until the device has built the index
  mutable std::vector<uint64_t> deviceTextOffsets_;

  void build();
  void buildFromMemory(const std::vector<const char *> &seqs);
  void finish();
  void index();
  void computeFreqHist();
  void dropFreqSeedSet();
  bool loadIndexTSV(const std::string &path);
  bool loadIndexBinary(const std::string &prefix);
};

}  // namespace skch
#endif
