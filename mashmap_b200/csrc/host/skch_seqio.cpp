#include "skch_seqio.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <cstring>
#include <iostream>
#include <thread>
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

/* ---- mapped FASTA ---- */

FastaFile::~FastaFile()
{
  if (data_) munmap((void *)data_, size_);
  if (fd_ >= 0) close(fd_);
}

bool FastaFile::open(const std::string &filename, int threads)
{
  fd_ = ::open(filename.c_str(), O_RDONLY);
  if (fd_ < 0) return false;
  struct stat st;
  if (fstat(fd_, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 2) return false;
  size_ = (uint64_t)st.st_size;
  void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
  if (m == MAP_FAILED) return false;
  data_ = (const char *)m;
  if (data_[0] != '>') return false; /* gzip magic, FASTQ, anything else: the line reader handles those */
  madvise(m, size_, MADV_SEQUENTIAL);
  const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(1, threads), size_ / (1 << 20) + 1));
  /* 1. record starts: '>' at the beginning of a line, found independently in T byte ranges */
  std::vector<std::vector<uint64_t>> starts((size_t)T);
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < T; t++) {
      pool.emplace_back([&, t]() {
        const uint64_t lo = size_ * (uint64_t)t / (uint64_t)T, hi = size_ * (uint64_t)(t + 1) / (uint64_t)T;
        uint64_t p = lo;
        while (p < hi) {
          const char *q = (const char *)memchr(data_ + p, '>', hi - p);
          if (!q) break;
          p = (uint64_t)(q - data_);
          if (p == 0 || data_[p - 1] == '\n') starts[(size_t)t].push_back(p);
          p++;
        }
      });
    }
    for (auto &th : pool) th.join();
  }
  std::vector<uint64_t> all;
  for (auto &v : starts) all.insert(all.end(), v.begin(), v.end());
  /* 2. per record: header, sequence region, base count */
  recs_.resize(all.size());
  {
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < T; t++) {
      pool.emplace_back([&]() {
        while (true) {
          const size_t b = next.fetch_add(4096);
          if (b >= all.size()) break;
          const size_t e = std::min(all.size(), b + 4096);
          for (size_t i = b; i < e; i++) {
            const uint64_t p = all[i], end = i + 1 < all.size() ? all[i + 1] : size_;
            const char *eol = (const char *)memchr(data_ + p, '\n', end - p);
            const uint64_t hdr_end = eol ? (uint64_t)(eol - data_) : end;
            const char *sp = (const char *)memchr(data_ + p + 1, ' ', hdr_end - (p + 1));
            FastaRecord &r = recs_[i];
            r.name_off = p + 1;
            r.name_len = (uint32_t)((sp ? (uint64_t)(sp - data_) : hdr_end) - (p + 1));
            r.seq_off = eol ? hdr_end + 1 : end;
            r.raw_len = end - r.seq_off;
            uint64_t nl = 0;
            for (const char *c = data_ + r.seq_off, *ce = data_ + end; c < ce;) {
              const char *q = (const char *)memchr(c, '\n', (size_t)(ce - c));
              if (!q) break;
              nl++;
              c = q + 1;
            }
            r.seq_len = r.raw_len - nl;
          }
        }
      });
    }
    for (auto &th : pool) th.join();
  }
  return true;
}

void FastaFile::copy_bases(const FastaRecord &r, char *dst) const
{
  const char *c = data_ + r.seq_off, *ce = c + r.raw_len;
  while (c < ce) {
    const char *q = (const char *)memchr(c, '\n', (size_t)(ce - c));
    const size_t n = (size_t)((q ? q : ce) - c);
    memcpy(dst, c, n);
    dst += n;
    c += n + 1;
  }
}

namespace {

struct NibTable {
  uint8_t t[256];
  NibTable()
  {
    for (int i = 0; i < 256; i++) t[i] = 8;
    t[(int)'A'] = t[(int)'a'] = 0; t[(int)'C'] = t[(int)'c'] = 1; t[(int)'T'] = t[(int)'t'] = 2; t[(int)'G'] = t[(int)'g'] = 3;
  }
};
const NibTable NIB;

void pack_scalar(const uint8_t *src, uint64_t n, uint8_t *dst)
{
  uint64_t i = 0;
  for (; i + 1 < n; i += 2) dst[i >> 1] = (uint8_t)(NIB.t[src[i]] | (NIB.t[src[i + 1]] << 4));
  if (i < n) dst[i >> 1] = (uint8_t)(NIB.t[src[i]] | 0x80);
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void pack_avx2(const uint8_t *src, uint64_t n, uint8_t *dst)
{
  const __m256i up = _mm256_set1_epi8((char)0xDF), three = _mm256_set1_epi8(3), eight = _mm256_set1_epi8(8);
  const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T');
  const __m256i mul = _mm256_set1_epi16(0x1001); /* low byte * 1 + high byte * 16 */
  uint64_t i = 0;
  for (; i + 32 <= n; i += 32) {
    const __m256i v = _mm256_loadu_si256((const __m256i *)(src + i));
    const __m256i x = _mm256_and_si256(v, up);
    const __m256i code = _mm256_and_si256(_mm256_srli_epi16(x, 1), three);
    const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(x, cA), _mm256_cmpeq_epi8(x, cC)),
                                       _mm256_or_si256(_mm256_cmpeq_epi8(x, cG), _mm256_cmpeq_epi8(x, cT)));
    const __m256i nib = _mm256_blendv_epi8(eight, code, ok);
    const __m256i w = _mm256_maddubs_epi16(nib, mul);                 /* 16 x (n0 + 16 n1) */
    const __m256i b = _mm256_packus_epi16(w, w);                      /* per 128-bit lane: 8 bytes, twice */
    const __m256i q = _mm256_permute4x64_epi64(b, 0x08);              /* lanes' low halves next to each other */
    _mm_storeu_si128((__m128i *)(dst + (i >> 1)), _mm256_castsi256_si128(q));
  }
  pack_scalar(src + i, n - i, dst + (i >> 1));
}
bool have_avx2()
{
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}
#endif

}  // namespace

void pack_bases(const char *src, uint64_t n, uint8_t *dst)
{
#if defined(__x86_64__)
  if (have_avx2()) { pack_avx2((const uint8_t *)src, n, dst); return; }
#endif
  pack_scalar((const uint8_t *)src, n, dst);
}

void FastaFile::pack_bases(const FastaRecord &r, uint8_t *dst) const
{
  /* lines are gathered into an even-sized stretch of text first (a line may have an odd length; nibble pairs must not
   * straddle two pack calls), then packed: the stretch stays in the L1/L2 cache */
  char buf[8192 + 64];
  size_t fill = 0;
  const char *c = data_ + r.seq_off, *ce = c + r.raw_len;
  while (c < ce) {
    const char *q = (const char *)memchr(c, '\n', (size_t)(ce - c));
    size_t n = (size_t)((q ? q : ce) - c);
    const char *next = c + n + 1;
    while (n) {
      const size_t take = std::min(n, (size_t)8192 - fill);
      memcpy(buf + fill, c, take);
      fill += take; c += take; n -= take;
      if (fill == 8192) { seqio::pack_bases(buf, fill, dst); dst += fill / 2; fill = 0; }
    }
    c = next;
  }
  if (fill) seqio::pack_bases(buf, fill, dst);
}

}  // namespace seqio
}  // namespace skch
