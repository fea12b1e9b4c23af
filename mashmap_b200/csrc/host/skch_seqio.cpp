#include "skch_seqio.hpp"

#include <zlib.h>

#include <cstring>
#include <iostream>
#include <vector>

namespace skch {
namespace seqio {

namespace {

class LineReader {
 public:
  explicit LineReader(const std::string &path) : buf_(1 << 20)
  {
    f_ = gzopen(path.c_str(), "rb");
    if (f_) gzbuffer(f_, 1 << 20);
  }
  ~LineReader() { if (f_) gzclose(f_); }
  bool ok() const { return f_ != nullptr; }
  /* std::getline semantics: false only when nothing could be read */
  bool getline(std::string &line)
  {
    line.clear();
    bool got = false;
    while (true) {
      if (pos_ == len_) {
        if (eof_) return got;
        int n = gzread(f_, buf_.data(), (unsigned)buf_.size());
        if (n <= 0) { eof_ = true; return got; }
        len_ = (size_t)n; pos_ = 0;
      }
      const char *b = buf_.data() + pos_;
      const char *nl = (const char *)memchr(b, '\n', len_ - pos_);
      if (nl) {
        line.append(b, nl - b);
        pos_ += (size_t)(nl - b) + 1;
        return true;
      }
      line.append(b, len_ - pos_);
      pos_ = len_;
      got = true;
    }
  }
  bool good() const { return !(eof_ && pos_ == len_); }

 private:
  gzFile f_ = nullptr;
  std::vector<char> buf_;
  size_t pos_ = 0, len_ = 0;
  bool eof_ = false;
};

}  // namespace

bool for_each_seq_in_file(const std::string &filename, const std::unordered_set<std::string> &keep_seq,
                          const std::string &keep_prefix, const SeqCallback &func)
{
  LineReader in(filename);
  if (!in.ok()) {
    std::cerr << "[mashmap-b200] ERROR: cannot open " << filename << std::endl;
    return false;
  }
  std::string line;
  in.getline(line);
  const bool is_fasta = !line.empty() && line[0] == '>';
  const bool is_fastq = !line.empty() && line[0] == '@';
  if (!is_fasta && !is_fastq) {
    std::cerr << "[mashmap-b200] unknown file format given to the sequence reader: " << filename << std::endl;
    return false;
  }
  auto wanted = [&](const std::string &name) {
    return (keep_prefix.empty() || name.compare(0, keep_prefix.length(), keep_prefix) == 0) &&
           (keep_seq.empty() || keep_seq.find(name) != keep_seq.end());
  };
  std::string seq;
  if (is_fasta) {
    bool more = true;
    while (more) {
      const std::string name = line.substr(1, line.find(' ') - 1); /* seqiter.hpp:82 */
      const bool keep = wanted(name);
      seq.clear();
      more = false;
      while (in.getline(line)) {
        if (!line.empty() && line[0] == '>') { more = true; break; }
        if (keep) seq.append(line);
      }
      func(name, seq);
    }
  } else {
    bool more = true;
    while (more) {
      const std::string name = line.substr(1, line.find(' ') - 1);
      const bool keep = wanted(name);
      std::string s, tmp;
      in.getline(s);
      in.getline(tmp);
      in.getline(tmp);
      more = in.getline(line) && !line.empty();
      func(name, keep ? s : std::string());
    }
  }
  return true;
}

}  // namespace seqio
}  // namespace skch
