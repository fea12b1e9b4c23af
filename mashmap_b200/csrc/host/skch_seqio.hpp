/*
 * skch_seqio.hpp -- FASTA / FASTQ (optionally gzip) reader.
 * Same observable behaviour as seqiter::for_each_seq_in_file (reference src/common/seqiter.hpp:20-111):
 * the record name is the header up to the first space (:82), sequence lines are concatenated, records
 * not matching keep_prefix / keep_seq are delivered with an empty sequence. Reads through zlib's gzFile
 * with a large buffer instead of the reference's 303-byte gzstream buffer (gzstream.h:50).
 */
#ifndef SKCH_SEQIO_HPP
#define SKCH_SEQIO_HPP

#include <functional>
#include <string>
#include <unordered_set>

namespace skch {
namespace seqio {

typedef std::function<void(const std::string &name, const std::string &seq)> SeqCallback;

/* returns false (after printing to stderr) if the file cannot be read or has an unknown format */
bool for_each_seq_in_file(const std::string &filename, const std::unordered_set<std::string> &keep_seq,
                          const std::string &keep_prefix, const SeqCallback &func);

}  // namespace seqio
}  // namespace skch
#endif
