/*
 * skch_seqio.hpp -- FASTA / FASTQ (optionally gzip) reader.
 * Same observable behaviour as seqiter::for_each_seq_in_file (reference src/common/seqiter.hpp:20-111):
 * the record name is the header up to the first space (:82), sequence lines are concatenated, records
 * not matching keep_prefix / keep_seq are delivered with an empty sequence. Reads through zlib's gzFile
 * with a large buffer instead of the reference's 303-byte gzstream buffer (gzstream.h:50).
 */
#ifndef SKCH_SEQIO_HPP
#define SKCH_SEQIO_HPP

#include <cstdint>
#include <functional>
#include <string>
#include <unordered_set>
#include <vector>

namespace skch {
namespace seqio {

typedef std::function<void(const std::string &name, const std::string &seq)> SeqCallback;

/* returns false (after printing to stderr) if the file cannot be read or has an unknown format */
bool for_each_seq_in_file(const std::string &filename, const std::unordered_set<std::string> &keep_seq,
                          const std::string &keep_prefix, const SeqCallback &func);

/*
 * Bulk view of a plain (uncompressed) FASTA file: the file is mapped and cut into records by all host threads, so that
 * a caller can place the bases where it wants them (the pinned batch buffer) without going through one std::string per
 * record on one thread -- at tens of Gbp/s of mapping, a serial parser is the bottleneck of the program.
 * Same record semantics as for_each_seq_in_file: a record starts at a line whose first byte is '>', its name is the
 * header up to the first space, its sequence is every following line up to the next record, concatenated (only the
 * '\n' bytes are dropped).
 */
struct FastaRecord {
  uint64_t name_off;  /* file offset of the first byte of the name */
  uint32_t name_len;
  uint64_t seq_off;   /* file offset of the first sequence line */
  uint64_t raw_len;   /* bytes from seq_off to the next record (or EOF), newlines included */
  uint64_t seq_len;   /* bases = raw_len minus the newlines */
};

class FastaFile {
 public:
  FastaFile() = default;
  ~FastaFile();
  FastaFile(const FastaFile &) = delete;
  /* false if the file is not a plain FASTA file that can be mapped (gzip, FASTQ, pipe ...): use for_each_seq_in_file */
  bool open(const std::string &filename, int threads);
  const std::vector<FastaRecord> &records() const { return recs_; }
  const char *data() const { return data_; }
  std::string name(const FastaRecord &r) const { return std::string(data_ + r.name_off, r.name_len); }
  /* copies the record's bases to dst (seq_len bytes) */
  void copy_bases(const FastaRecord &r, char *dst) const;
  /* the record's bases as nibbles (pack_bases below) to dst ((seq_len + 1) / 2 bytes), newlines dropped on the way */
  void pack_bases(const FastaRecord &r, uint8_t *dst) const;

 private:
  const char *data_ = nullptr;
  uint64_t size_ = 0;
  int fd_ = -1;
  std::vector<FastaRecord> recs_;
};

/*
 * The device's input format (include/mashmap_b200.h, mm_map_segments_packed): one nibble per base, base i of the
 * sequence in byte i / 2 (low nibble first); nibble = 2-bit code (A 0, C 1, T 2, G 3: bits 1-2 of the upper-cased
 * letter) | 8 for every byte that is not ACGT after upper-casing -- makeUpperCaseAndValidDNA (reference
 * commonFunc.hpp:75-107) folded into the encoding. The parser touches every base once anyway; writing 4 bits instead of
 * 8 halves what crosses PCIe afterwards. dst gets (n + 1) / 2 bytes; an odd n leaves an 'N' nibble in the last byte.
 * AVX2 where the CPU has it (32 bases per step), plain C otherwise.
 */
void pack_bases(const char *src, uint64_t n, uint8_t *dst);

}  // namespace seqio
}  // namespace skch
#endif
