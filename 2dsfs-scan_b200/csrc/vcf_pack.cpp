// vcf_pack.cpp -- K0: multithreaded host packer, gzip VCF + popmap -> 2-bit genotype matrix (B32 layout) for the GPU.
//
// Replaces make_data_dict_vcf (uricchio/2DSFS-scan scripts/src/twoDSFS_class.py:36-138, twin scripts/sims_scan.py:18-120) for
// the two scanned populations, with the reference's exact ingest rules:
//   * popmap: line.strip().split("\t"), first two columns (:60-63)
//   * header: samples found in the popmap, IN ORDER, form a positional population list (:81-85) that is zipped with the
//     sample columns (:118) -- a header sample missing from the popmap shifts every later label one column to the left
//   * FILTER must be PASS or "." (:101); REF and ALT single A/C/G/T after upper-casing (:105-109)
//   * annotation = second '|' field of INFO, else "No annotation" (:94-99)
//   * per sample: count the characters '0' and '1' at EVEN offsets of the GT sub-field (:122-130)
//   * duplicate CHROM-POS keys: the last record wins (dict assignment :134)
// Calls whose (ref, alt) contribution is not (2,0) / (1,1) / (0,2) / (0,0) are stored as missing plus a sparse fix-up.
// Build: g++ -O3 -std=c++17 -shared -fPIC -pthread vcf_pack.cpp -lz   (see build.py)
#include <sched.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <future>
#include <mutex>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Fix { int64_t snp; int32_t pop, dref, dalt; };

struct Rec {                 // one accepted VCF record
  std::string chrom, pos_str, ann;
  long long pos = 0;
  size_t woff = 0;                     // its RW words (pop1 then pop2) start here in the flat word pool
  std::vector<Fix> fix;                // snp filled in later
  bool ok = false;                     // passed the gates
  std::string err;
};

struct Packed {
  int64_t S = 0;
  int W1 = 0, W2 = 0, ns1 = 0, ns2 = 0;
  std::vector<uint32_t> G;             // B32 layout
  std::vector<int32_t> pos;
  std::vector<std::string> chroms;
  std::vector<int64_t> off;
  std::vector<Fix> fix;
  std::vector<int32_t> ann_code;
  std::vector<std::string> ann_vocab;
  int64_t last_key_row = -1;           // row (sorted order) of the last-inserted key (T2D_scan quirk)
  int64_t n_records = 0, n_skipped = 0;
  std::string names_blob, vocab_blob;
};

thread_local std::string g_err;

const char* WS = " \t\n\r\f\v";

std::string strip(const std::string& s) {
  size_t a = s.find_first_not_of(WS);
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(WS);
  return s.substr(a, b - a + 1);
}

void split_char(const char* p, size_t n, char sep, std::vector<std::pair<const char*, size_t>>& out) {
  out.clear();
  const char* start = p;
  const char* end = p + n;
  for (const char* q = p; q < end; ++q)
    if (*q == sep) { out.emplace_back(start, (size_t)(q - start)); start = q + 1; }
  out.emplace_back(start, (size_t)(end - start));
}

struct Plan {                 // what each sample column contributes
  std::vector<int> col_pop;   // per zipped column: 0 = pop1, 1 = pop2, -1 = other population
  std::vector<int> col_slot;  // sample slot inside its population block
  int ns1 = 0, ns2 = 0, W1 = 1, W2 = 1;
};

// One record.  `s` keeps its line terminator (the reference splits the raw line on tabs, so the last column carries it).
// Single pass: the nine fixed columns are located with memchr, the sample columns are scanned in place.
void parse_line(const char* s, size_t n, const Plan& plan, Rec& r, uint32_t* words) {
  const char* end = s + n;
  std::pair<const char*, size_t> cols[9];
  const char* q = s;
  int nc = 0;
  bool more = false;  // sample columns follow the FORMAT column
  while (nc < 9) {
    const char* t = (const char*)memchr(q, '\t', (size_t)(end - q));
    cols[nc++] = {q, (size_t)((t ? t : end) - q)};
    if (!t) { q = end; break; }
    q = t + 1;
    if (nc == 9) more = true;
  }
  if (nc < 9) { r.err = "VCF record with fewer than 9 columns"; return; }
  auto S = [&](int i) { return std::string(cols[i].first, cols[i].second); };
  static thread_local std::vector<std::pair<const char*, size_t>> sub;
  // annotation (:94-99)
  split_char(cols[7].first, cols[7].second, '|', sub);
  r.ann = sub.size() >= 2 ? std::string(sub[1].first, sub[1].second) : "No annotation";
  if (!(cols[6].second == 4 && memcmp(cols[6].first, "PASS", 4) == 0) && !(cols[6].second == 1 && cols[6].first[0] == '.')) return;
  auto base_ok = [](const std::pair<const char*, size_t>& c) {
    if (c.second != 1) return false;
    char u = (char)toupper((unsigned char)c.first[0]);
    return u == 'A' || u == 'C' || u == 'G' || u == 'T';
  };
  if (!base_ok(cols[3]) || !base_ok(cols[4])) return;
  // GT index (:115)
  split_char(cols[8].first, cols[8].second, ':', sub);
  int gti = -1;
  for (size_t i = 0; i < sub.size(); ++i)
    if (sub[i].second == 2 && sub[i].first[0] == 'G' && sub[i].first[1] == 'T') { gti = (int)i; break; }
  if (gti < 0) { r.err = "'GT' is not in list"; return; }  // the reference raises ValueError
  r.chrom = S(0);
  r.pos_str = S(1);
  char* endp = nullptr;
  r.pos = strtoll(r.pos_str.c_str(), &endp, 10);
  if (endp == r.pos_str.c_str() || *endp) { r.err = "invalid literal for int(): POS '" + r.pos_str + "'"; return; }
  const int RW = plan.W1 + plan.W2;
  for (int w = 0; w < RW; ++w) words[w] = 0u;
  // every slot starts as MISSING (a column absent from this record contributes (0, 0)); real calls overwrite it
  for (int pop = 0; pop < 2; ++pop) {
    const int ns = pop ? plan.ns2 : plan.ns1;
    uint32_t* w = words + (pop ? plan.W1 : 0);
    for (int g = 0; g * 32 < ns; ++g) w[2 * g + 1] = ns - g * 32 >= 32 ? 0xFFFFFFFFu : ((1u << (ns - g * 32)) - 1u);  // code 2 = (lo 0, hi 1)
  }
  // zip(poplist, cols[9:]): the shorter of the two ends the loop.  One scan over the sample columns, character by
  // character (a call is ~4 bytes: library calls per column would dominate).
  const size_t nplan = plan.col_pop.size();
  for (size_t j = 0; more && j < nplan; ++j) {
    const int pop = plan.col_pop[j];
    const char* g = q;  // start of this sample column
    if (pop >= 0) {
      // sample.split(':')[gti]: skip gti sub-fields
      for (int k = 0; k < gti; ++k) {
        while (g < end && *g != ':' && *g != '\t') ++g;
        if (g == end || *g == '\t') { r.err = "list index out of range (sample without GT sub-field)"; return; }
        ++g;
      }
      int ref = 0, alt = 0;
      for (const char* f = g; g < end && *g != ':' && *g != '\t'; ++g) {
        if (((g - f) & 1) == 0) {  // gt[::2]
          ref += *g == '0';
          alt += *g == '1';
        }
      }
      // (ref, alt) -> 2-bit code without data-dependent branches: (2,0) -> 0, (1,1) -> 1, (0,2) -> 3, anything else is
      // stored as missing (2) plus, when it carries alleles, a fix-up
      static const uint8_t CODE[10] = {2, 2, 3, 2, 1, 2, 0, 2, 2, 2};  // index ref * 3 + alt for ref, alt <= 2, else 9
      const uint32_t code = CODE[(ref > 2 || alt > 2) ? 9 : ref * 3 + alt];
      if (code == 2 && (ref | alt)) r.fix.push_back({0, pop, ref, alt});
      const int slot = plan.col_slot[j];
      // samples 32g..32g+31 of a population: word 2g = lo bits of their codes, word 2g+1 = hi bits
      uint32_t* w = words + (pop ? plan.W1 : 0) + 2 * (slot >> 5);
      const uint32_t bit = 1u << (slot & 31);
      w[0] = (w[0] & ~bit) | ((code & 1u) ? bit : 0u);
      w[1] = (w[1] & ~bit) | ((code & 2u) ? bit : 0u);
    }
    // to the end of the column
    if (g < end && *g != '\t') {
      const char* t = (const char*)memchr(g, '\t', (size_t)(end - g));
      g = t ? t : end;
    }
    if (g == end) break;
    q = g + 1;
  }
  r.ok = true;
}

// ---------------------------------------------------------------------------------------------- worker pool
// Persistent threads (spawning a thread per batch costs more than the parsing of small batches: every new thread starts
// on a cold malloc arena).  run(n, fn) executes fn(0..n-1) on the workers and returns when all are done.
class Pool {
 public:
  explicit Pool(int n) {
    for (int i = 0; i < n; ++i) th_.emplace_back([this, i]() { loop(i); });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; ++gen_; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return (int)th_.size(); }
  void run(int n, const std::function<void(int)>& fn) {
    n = std::min(n, size());
    if (n <= 0) return;
    std::unique_lock<std::mutex> l(m_);
    fn_ = &fn; active_ = n; pending_ = n; ++gen_;
    cv_.notify_all();
    done_.wait(l, [&]() { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop(int id) {
    unsigned long long seen = 0;
    while (true) {
      const std::function<void(int)>* fn = nullptr;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&]() { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        if (id >= active_) continue;
        fn = fn_;
      }
      (*fn)(id);
      {
        std::lock_guard<std::mutex> l(m_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int active_ = 0, pending_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
};

// ---------------------------------------------------------------------------------------------- input: gzip or BGZF
// Uncompressed text arrives in chunks that end on a line boundary.  A plain gzip stream is inflated by one thread (zlib's
// gzread); a BGZF file (bgzip: independent <= 64 KB gzip members whose "BC" extra field holds the member size) is
// inflated member-parallel.
struct Input {
  gzFile gz = nullptr;          // plain gzip / uncompressed
  FILE* fp = nullptr;           // BGZF
  int nthreads = 1;
  Pool* pool = nullptr;         // inflates BGZF members
  std::string carry;            // partial last line of the previous chunk
  bool eof = false;
  std::string err;

  static bool bgzf_header(const unsigned char* h, size_t n, size_t& bsize) {
    if (n < 18 || h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) return false;
    const size_t xlen = h[10] | (h[11] << 8);
    if (n < 12 + xlen) return false;
    for (size_t o = 12; o + 4 <= 12 + xlen;) {
      const size_t slen = h[o + 2] | (h[o + 3] << 8);
      if (h[o] == 'B' && h[o + 1] == 'C' && slen == 2 && o + 6 <= 12 + xlen) { bsize = (size_t)(h[o + 4] | (h[o + 5] << 8)) + 1; return true; }
      o += 4 + slen;
    }
    return false;
  }

  bool open(const char* path, int nt) {
    nthreads = std::max(1, nt);
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    unsigned char h[64];
    const size_t n = fread(h, 1, sizeof h, f);
    size_t bs = 0;
    if (bgzf_header(h, n, bs) && nthreads > 1) {
      fp = f;
      fseek(fp, 0, SEEK_SET);
      return true;
    }
    fclose(f);
    gz = gzopen(path, "rb");
    if (!gz) return false;
    gzbuffer(gz, 1 << 20);
    return true;
  }
  void close() {
    if (gz) gzclose(gz);
    if (fp) fclose(fp);
    gz = nullptr; fp = nullptr;
  }

  // raw bytes of the next chunk appended to `out`; false at end of input
  bool read_raw(std::string& out, size_t want) {
    if (gz) {
      const size_t o = out.size();
      out.resize(o + want);
      const int got = gzread(gz, &out[o], (unsigned)want);
      if (got < 0) { int e; err = gzerror(gz, &e); out.resize(o); return false; }
      out.resize(o + (size_t)got);
      return got > 0;
    }
    // BGZF: collect members up to ~want uncompressed bytes, inflate them in parallel
    struct Member { std::vector<unsigned char> comp; size_t isize = 0, off = 0; };
    std::vector<Member> ms;
    size_t total = 0;
    while (total < want) {
      unsigned char h[18];
      const size_t n = fread(h, 1, 18, fp);
      if (n == 0) break;
      size_t bs = 0;
      // the "BC" subfield is the first one in files written by bgzip; anything else is read through zlib's own reader
      if (n < 18 || !(h[0] == 31 && h[1] == 139 && h[2] == 8 && (h[3] & 4) && h[12] == 'B' && h[13] == 'C' && h[14] == 2 && h[15] == 0)) {
        err = "BGZF member header not understood";
        return false;
      }
      bs = (size_t)(h[16] | (h[17] << 8)) + 1;
      const size_t xlen = h[10] | (h[11] << 8);
      if (bs < 12 + xlen + 8) { err = "BGZF member too short"; return false; }
      Member m;
      m.comp.resize(bs);
      memcpy(m.comp.data(), h, 18);
      if (fread(m.comp.data() + 18, 1, bs - 18, fp) != bs - 18) { err = "truncated BGZF member"; return false; }
      const unsigned char* tail = m.comp.data() + bs - 4;
      m.isize = (size_t)tail[0] | ((size_t)tail[1] << 8) | ((size_t)tail[2] << 16) | ((size_t)tail[3] << 24);
      m.off = total;
      total += m.isize;
      ms.push_back(std::move(m));
    }
    if (ms.empty()) return false;
    const size_t o = out.size();
    out.resize(o + total);
    const int nt = (int)std::min<size_t>((size_t)(pool ? pool->size() : 1), ms.size());
    std::vector<char> okv((size_t)nt, 1);
    auto work = [&](int t) {
        okv[(size_t)t] = [&]() {
        for (size_t i = (size_t)t; i < ms.size(); i += (size_t)nt) {
          const Member& m = ms[i];
          if (m.isize == 0) continue;
          const size_t xlen = m.comp[10] | (m.comp[11] << 8);
          z_stream zs;
          memset(&zs, 0, sizeof zs);
          if (inflateInit2(&zs, -15) != Z_OK) return false;
          zs.next_in = const_cast<unsigned char*>(m.comp.data()) + 12 + xlen;
          zs.avail_in = (uInt)(m.comp.size() - 12 - xlen - 8);
          zs.next_out = (unsigned char*)&out[o + m.off];
          zs.avail_out = (uInt)m.isize;
          const int rc = inflate(&zs, Z_FINISH);
          inflateEnd(&zs);
          if (rc != Z_STREAM_END || zs.avail_out != 0) return false;
        }
        return true;
        }() ? 1 : 0;
    };
    if (pool) pool->run(nt, work); else work(0);
    for (char k : okv)
      if (!k) { err = "corrupt BGZF member"; return false; }
    return true;
  }

  // next chunk of whole lines (the final line of the input may lack its terminator); false when the input is exhausted
  bool next_chunk(std::string& chunk, size_t want = 2u << 20) {
    chunk.clear();
    if (eof) return false;
    chunk.swap(carry);
    while (true) {
      if (!read_raw(chunk, want)) {
        eof = true;
        return err.empty() && !chunk.empty();
      }
      const size_t nl = chunk.rfind('\n');
      if (nl != std::string::npos) {
        carry.assign(chunk, nl + 1, std::string::npos);
        chunk.resize(nl + 1);
        return true;
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------- counts mode
// make_data_dict_vcf itself (reference :36-138): every population of the popmap, per-record (ref, alt) counts, the key text
// "CHROM-POS" as written, REF/ALT, annotation.  The Python side (tdsfs_pack.vcf_counts) turns the arrays into the dict.
struct CRec {
  std::string key, ann, err;
  char ref = 0, alt = 0;
  int ncols = 0;          // zipped sample columns of this record: populations first seen at a later column get no entry
  bool ok = false;
};

struct CountsOut {
  int64_t n = 0;
  int npop = 0;
  std::string pop_blob, keys_blob, vocab_blob, refalt;
  std::vector<int64_t> key_off;
  std::vector<int32_t> first_col, ann_code, cnt, ncols;
};

void parse_line_counts(const char* s, size_t n, const std::vector<int>& col_pop, int npop, CRec& r, int32_t* cnt) {
  const char* end = s + n;
  std::pair<const char*, size_t> cols[9];
  const char* q = s;
  int nc = 0;
  bool more = false;
  while (nc < 9) {
    const char* t = (const char*)memchr(q, '\t', (size_t)(end - q));
    cols[nc++] = {q, (size_t)((t ? t : end) - q)};
    if (!t) { q = end; break; }
    q = t + 1;
    if (nc == 9) more = true;
  }
  if (nc < 8) { r.err = "list index out of range (record with fewer than 8 columns)"; return; }
  static thread_local std::vector<std::pair<const char*, size_t>> sub;
  split_char(cols[7].first, cols[7].second, '|', sub);
  r.ann = sub.size() >= 2 ? std::string(sub[1].first, sub[1].second) : "No annotation";
  if (!(cols[6].second == 4 && memcmp(cols[6].first, "PASS", 4) == 0) && !(cols[6].second == 1 && cols[6].first[0] == '.')) return;
  auto base = [](const std::pair<const char*, size_t>& c) -> char {
    if (c.second != 1) return 0;
    const char u = (char)toupper((unsigned char)c.first[0]);
    return (u == 'A' || u == 'C' || u == 'G' || u == 'T') ? u : 0;
  };
  r.ref = base(cols[3]);
  r.alt = base(cols[4]);
  if (!r.ref || !r.alt) return;
  if (nc < 9) { r.err = "list index out of range (record without a FORMAT column)"; return; }
  split_char(cols[8].first, cols[8].second, ':', sub);
  int gti = -1;
  for (size_t i = 0; i < sub.size(); ++i)
    if (sub[i].second == 2 && sub[i].first[0] == 'G' && sub[i].first[1] == 'T') { gti = (int)i; break; }
  if (gti < 0) { r.err = "'GT' is not in list"; return; }
  r.key.assign(cols[0].first, cols[0].second);
  r.key.push_back('-');
  r.key.append(cols[1].first, cols[1].second);
  for (int p = 0; p < 2 * npop; ++p) cnt[p] = 0;
  const size_t nplan = col_pop.size();
  for (size_t j = 0; more && j < nplan; ++j) {
    const char* g = q;
    for (int k = 0; k < gti; ++k) {
      while (g < end && *g != ':' && *g != '\t') ++g;
      if (g == end || *g == '\t') { r.err = "list index out of range (sample without GT sub-field)"; return; }
      ++g;
    }
    int ref = 0, alt = 0;
    for (const char* f = g; g < end && *g != ':' && *g != '\t'; ++g) {
      if (((g - f) & 1) == 0) {  // gt[::2]
        ref += *g == '0';
        alt += *g == '1';
      }
    }
    cnt[2 * col_pop[j]] += ref;
    cnt[2 * col_pop[j] + 1] += alt;
    r.ncols = (int)j + 1;
    if (g < end && *g != '\t') {
      const char* t = (const char*)memchr(g, '\t', (size_t)(end - g));
      g = t ? t : end;
    }
    if (g == end) break;
    q = g + 1;
  }
  r.ok = true;
}

}  // namespace

extern "C" {

const char* tdsfs_pack_last_error(void) { return g_err.c_str(); }

// Returns an opaque handle (NULL on error).  pop1 / pop2: population labels of the popmap.
void* tdsfs_pack_vcf(const char* vcf_path, const char* popmap_path, const char* pop1, const char* pop2, int nthreads) {
  try {
    // ---- popmap
    std::unordered_map<std::string, std::string> popmap;
    {
      FILE* f = fopen(popmap_path, "r");
      if (!f) { g_err = std::string("cannot open popmap ") + popmap_path; return nullptr; }
      char* line = nullptr; size_t cap = 0; ssize_t n;
      std::vector<std::pair<const char*, size_t>> cols;
      while ((n = getline(&line, &cap, f)) >= 0) {
        std::string s = strip(std::string(line, (size_t)n));
        split_char(s.data(), s.size(), '\t', cols);
        if (cols.size() >= 2) popmap[std::string(cols[0].first, cols[0].second)] = std::string(cols[1].first, cols[1].second);
      }
      free(line);
      fclose(f);
    }
    if (nthreads <= 0) {
      // the CPUs this process may run on (a container's share), not the machine's
      cpu_set_t set;
      CPU_ZERO(&set);
      nthreads = sched_getaffinity(0, sizeof set, &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
      nthreads = std::max(1, nthreads);
    }
    Pool parse_pool(nthreads), inflate_pool(std::max(1, nthreads / 2));
    Input in;
    in.pool = &inflate_pool;
    if (!in.open(vcf_path, nthreads)) { g_err = std::string("cannot open VCF ") + vcf_path; return nullptr; }

    Plan plan;
    std::vector<std::string> poplist;
    auto rebuild_plan = [&]() {
      plan = Plan();
      for (auto& p : poplist) {
        int pop = p == pop1 ? 0 : (p == pop2 ? 1 : -1);
        // if pop1 == pop2 (1D helpers) the label maps to population 0 only
        plan.col_pop.push_back(pop);
        plan.col_slot.push_back(pop == 0 ? plan.ns1++ : (pop == 1 ? plan.ns2++ : 0));
      }
      plan.W1 = 2 * std::max(1, (plan.ns1 + 31) / 32);  // a (lo plane, hi plane) word pair per 32 samples
      plan.W2 = 2 * std::max(1, (plan.ns2 + 31) / 32);
    };
    rebuild_plan();

    const bool dbg = getenv("TDSFS_PACK_DEBUG") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_wait = 0, t_parse = 0, t_merge = 0, t_start = now();
    std::vector<Rec> all;                      // accepted records in file order
    std::vector<uint32_t> words_all;           // their 2-bit rows, RW words each (Rec::woff)
    std::vector<uint32_t> cw;                  // rows of the batch being parsed
    std::vector<Rec> recs;
    std::vector<std::pair<const char*, size_t>> batch;   // record lines of the current chunk
    int64_t n_records = 0, n_skipped = 0;
    std::string first_err;
    bool have_records = false;
    auto flush = [&]() {
      if (batch.empty()) return;
      const size_t RW = (size_t)(plan.W1 + plan.W2);
      recs.assign(batch.size(), Rec());
      cw.resize(batch.size() * RW);
      const int nt = (int)std::min<size_t>((size_t)nthreads, (batch.size() + 63) / 64);
      const double tp0 = now();
      parse_pool.run(nt, [&](int t) {  // contiguous slices: a thread walks its own part of the chunk
        const size_t a = batch.size() * (size_t)t / (size_t)nt, b = batch.size() * (size_t)(t + 1) / (size_t)nt;
        for (size_t i = a; i < b; ++i) parse_line(batch[i].first, batch[i].second, plan, recs[i], cw.data() + i * RW);
      });
      const double tp1 = now();
      t_parse += tp1 - tp0;
      for (size_t i = 0; i < recs.size(); ++i) {
        Rec& r = recs[i];
        ++n_records;
        if (!r.err.empty()) { if (first_err.empty()) first_err = r.err; continue; }
        if (!r.ok) { ++n_skipped; continue; }
        r.woff = words_all.size();
        words_all.insert(words_all.end(), cw.begin() + (long)(i * RW), cw.begin() + (long)((i + 1) * RW));
        all.push_back(std::move(r));
      }
      t_merge += now() - tp1;
      batch.clear();
    };
    std::string chunk, next;
    bool have = in.next_chunk(chunk);
    bool bad_header = false;
    while (have && !bad_header) {
      // the next chunk is read and inflated while this one is parsed
      std::future<bool> more = std::async(std::launch::async, [&]() { return in.next_chunk(next); });
      const char* p = chunk.data();
      const char* end = p + chunk.size();
      while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl + 1 : end;        // the line keeps its terminator, like Python's file iteration
        const size_t n = (size_t)(le - p);
        if (n >= 2 && p[0] == '#' && p[1] == '#') { p = le; continue; }
        if (n >= 1 && p[0] == '#') {
          flush();
          if (have_records) { bad_header = true; break; }
          // header_cols = line.split(); samples found in the popmap extend the positional list (:78-85)
          const std::string line(p, n);
          size_t i = 0; int col = 0;
          while (i < line.size()) {
            size_t a = line.find_first_not_of(WS, i);
            if (a == std::string::npos) break;
            size_t b = line.find_first_of(WS, a);
            if (b == std::string::npos) b = line.size();
            if (col >= 9) {
              auto it = popmap.find(line.substr(a, b - a));
              if (it != popmap.end()) poplist.push_back(it->second);
            }
            ++col; i = b;
          }
          rebuild_plan();
          p = le;
          continue;
        }
        have_records = true;
        batch.emplace_back(p, n);
        p = le;
      }
      flush();
      const double tw0 = now();
      have = more.get();
      t_wait += now() - tw0;
      chunk.swap(next);
    }
    in.close();
    const double t_read_done = now();
    if (bad_header) { g_err = "second header line after records is not supported by the packer"; return nullptr; }
    if (!in.err.empty()) { g_err = "reading VCF: " + in.err; return nullptr; }
    if (!first_err.empty()) { g_err = first_err; return nullptr; }

    // ---- dict semantics: the last record of a key wins, the key keeps its first insertion slot
    std::unordered_map<std::string, size_t> key_slot;
    std::vector<size_t> keep;            // index into `all` per distinct key, in first-insertion order
    for (size_t i = 0; i < all.size(); ++i) {
      std::string key = all[i].chrom + "-" + all[i].pos_str;
      auto it = key_slot.find(key);
      if (it == key_slot.end()) { key_slot.emplace(key, keep.size()); keep.push_back(i); }
      else keep[it->second] = i;
    }
    // ---- sort by (chromosome string, position), stable on insertion order
    std::vector<size_t> order(keep.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
      const Rec& x = all[keep[a]];
      const Rec& y = all[keep[b]];
      if (x.chrom != y.chrom) return x.chrom < y.chrom;
      return x.pos < y.pos;
    });
    Packed* P = new Packed();
    P->S = (int64_t)order.size();
    P->W1 = plan.W1; P->W2 = plan.W2; P->ns1 = plan.ns1; P->ns2 = plan.ns2;
    P->n_records = n_records; P->n_skipped = n_skipped;
    const int RW = plan.W1 + plan.W2;
    const int64_t nblk = (P->S + 31) / 32;
    P->G.assign((size_t)(nblk * RW * 32), 0u);
    P->pos.resize((size_t)P->S);
    P->ann_code.resize((size_t)P->S);
    std::map<std::string, int> vocab;
    for (int64_t s = 0; s < P->S; ++s) {
      const Rec& r = all[keep[order[(size_t)s]]];
      if (r.pos < 0 || r.pos > 2147483646LL) { g_err = "position out of int32 range"; delete P; return nullptr; }
      if (P->chroms.empty() || P->chroms.back() != r.chrom) { P->chroms.push_back(r.chrom); P->off.push_back(s); }
      P->pos[(size_t)s] = (int32_t)r.pos;
      for (int w = 0; w < RW; ++w) P->G[(size_t)(((s >> 5) * RW + w) * 32 + (s & 31))] = words_all[r.woff + (size_t)w];
      for (Fix f : r.fix) { f.snp = s; P->fix.push_back(f); }
      auto it = vocab.find(r.ann);
      if (it == vocab.end()) { it = vocab.emplace(r.ann, (int)P->ann_vocab.size()).first; P->ann_vocab.push_back(r.ann); }
      P->ann_code[(size_t)s] = it->second;
      if (order[(size_t)s] == keep.size() - 1) P->last_key_row = s;
    }
    P->off.push_back(P->S);
    if (dbg)
      fprintf(stderr, "[vcf_pack] threads %d: read+parse %.3f s (parse %.3f, merge %.3f, waiting for input %.3f), sort+layout %.3f s\n",
              nthreads, t_read_done - t_start, t_parse, t_merge, t_wait, now() - t_read_done);
    if (P->chroms.empty()) P->off.assign(1, 0);
    for (auto& c : P->chroms) { P->names_blob += c; P->names_blob.push_back('\n'); }
    for (auto& c : P->ann_vocab) { P->vocab_blob += c; P->vocab_blob.push_back('\n'); }
    return P;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

void tdsfs_pack_free(void* h) { delete (Packed*)h; }

// dims[8] = S, W1, W2, ns1, ns2, C, n_fix, last_key_row ; dims2[4] = n_records, n_skipped, n_vocab, G words
void tdsfs_pack_dims(void* h, int64_t* dims, int64_t* dims2) {
  Packed* P = (Packed*)h;
  dims[0] = P->S; dims[1] = P->W1; dims[2] = P->W2; dims[3] = P->ns1; dims[4] = P->ns2; dims[5] = (int64_t)P->chroms.size();
  dims[6] = (int64_t)P->fix.size(); dims[7] = P->last_key_row;
  dims2[0] = P->n_records; dims2[1] = P->n_skipped; dims2[2] = (int64_t)P->ann_vocab.size(); dims2[3] = (int64_t)P->G.size();
}
const uint32_t* tdsfs_pack_genotypes(void* h) { return ((Packed*)h)->G.data(); }
const int32_t* tdsfs_pack_positions(void* h) { return ((Packed*)h)->pos.data(); }
const int64_t* tdsfs_pack_chrom_off(void* h) { return ((Packed*)h)->off.data(); }
const int32_t* tdsfs_pack_ann_codes(void* h) { return ((Packed*)h)->ann_code.data(); }
const void* tdsfs_pack_fixups(void* h) { return ((Packed*)h)->fix.data(); }  // {int64 snp; int32 pop, dref, dalt} = tdsfs_fixup_t
const char* tdsfs_pack_chrom_names(void* h) { return ((Packed*)h)->names_blob.c_str(); }   // '\n'-separated
const char* tdsfs_pack_ann_vocab(void* h) { return ((Packed*)h)->vocab_blob.c_str(); }


// ---- counts mode (make_data_dict_vcf): opaque handle, NULL on error
void* tdsfs_vcf_counts(const char* vcf_path, const char* popmap_path, int nthreads) {
  try {
    std::unordered_map<std::string, std::string> popmap;
    {
      FILE* f = fopen(popmap_path, "r");
      if (!f) { g_err = std::string("cannot open popmap ") + popmap_path; return nullptr; }
      char* line = nullptr; size_t cap = 0; ssize_t n;
      std::vector<std::pair<const char*, size_t>> cols;
      while ((n = getline(&line, &cap, f)) >= 0) {
        std::string s = strip(std::string(line, (size_t)n));
        split_char(s.data(), s.size(), '\t', cols);
        if (cols.size() >= 2) popmap[std::string(cols[0].first, cols[0].second)] = std::string(cols[1].first, cols[1].second);
      }
      free(line);
      fclose(f);
    }
    if (nthreads <= 0) {
      cpu_set_t set;
      CPU_ZERO(&set);
      nthreads = sched_getaffinity(0, sizeof set, &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
      nthreads = std::max(1, nthreads);
    }
    Pool parse_pool(nthreads), inflate_pool(std::max(1, nthreads / 2));
    Input in;
    in.pool = &inflate_pool;
    if (!in.open(vcf_path, nthreads)) { g_err = std::string("cannot open VCF ") + vcf_path; return nullptr; }

    std::vector<int> col_pop;                       // positional population list (:81-85), as indices into `pops`
    std::vector<std::string> pops;                  // distinct labels in order of first appearance
    std::vector<int32_t> first_col;
    std::unordered_map<std::string, int> pop_id;
    CountsOut* P = new CountsOut();
    std::unique_ptr<CountsOut> guard(P);
    std::map<std::string, int> vocab;
    std::vector<std::string> vocab_list;
    std::vector<std::pair<const char*, size_t>> batch;
    std::vector<CRec> recs;
    std::vector<int32_t> cw;
    // records parsed with fewer populations than the final count are padded at the end (a header after records adds labels)
    std::vector<int> rec_npop;
    std::string first_err;
    auto flush = [&]() {
      if (batch.empty()) return;
      const size_t np2 = 2 * pops.size();
      recs.assign(batch.size(), CRec());
      cw.assign(batch.size() * std::max<size_t>(np2, 1), 0);
      const int nt = (int)std::min<size_t>((size_t)nthreads, (batch.size() + 63) / 64);
      parse_pool.run(nt, [&](int t) {
        const size_t a = batch.size() * (size_t)t / (size_t)nt, b = batch.size() * (size_t)(t + 1) / (size_t)nt;
        for (size_t i = a; i < b; ++i) parse_line_counts(batch[i].first, batch[i].second, col_pop, (int)pops.size(), recs[i], cw.data() + i * np2);
      });
      for (size_t i = 0; i < recs.size(); ++i) {
        CRec& r = recs[i];
        if (!r.err.empty()) { if (first_err.empty()) first_err = r.err; break; }  // the reference raises at this record
        if (!r.ok) continue;
        P->key_off.push_back((int64_t)P->keys_blob.size());
        P->keys_blob += r.key;
        P->refalt.push_back(r.ref);
        P->refalt.push_back(r.alt);
        auto it = vocab.find(r.ann);
        if (it == vocab.end()) { it = vocab.emplace(r.ann, (int)vocab_list.size()).first; vocab_list.push_back(r.ann); }
        P->ann_code.push_back(it->second);
        P->ncols.push_back(r.ncols);
        P->cnt.insert(P->cnt.end(), cw.begin() + (long)(i * np2), cw.begin() + (long)((i + 1) * np2));
        rec_npop.push_back((int)pops.size());
        ++P->n;
      }
      batch.clear();
    };
    std::string chunk, next;
    bool have = in.next_chunk(chunk);
    while (have && first_err.empty()) {
      std::future<bool> more = std::async(std::launch::async, [&]() { return in.next_chunk(next); });
      const char* p = chunk.data();
      const char* end = p + chunk.size();
      while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl + 1 : end;
        const size_t n = (size_t)(le - p);
        if (n >= 2 && p[0] == '#' && p[1] == '#') { p = le; continue; }
        if (n >= 1 && p[0] == '#') {
          flush();
          const std::string line(p, n);
          size_t i = 0; int col = 0;
          while (i < line.size()) {
            size_t a = line.find_first_not_of(WS, i);
            if (a == std::string::npos) break;
            size_t b = line.find_first_of(WS, a);
            if (b == std::string::npos) b = line.size();
            if (col >= 9) {
              auto it = popmap.find(line.substr(a, b - a));
              if (it != popmap.end()) {
                auto pi = pop_id.find(it->second);
                if (pi == pop_id.end()) {
                  pi = pop_id.emplace(it->second, (int)pops.size()).first;
                  pops.push_back(it->second);
                  first_col.push_back((int32_t)col_pop.size());
                }
                col_pop.push_back(pi->second);
              }
            }
            ++col; i = b;
          }
          p = le;
          continue;
        }
        batch.emplace_back(p, n);
        p = le;
      }
      flush();
      have = more.get();
      chunk.swap(next);
    }
    in.close();
    if (!first_err.empty()) { g_err = first_err; return nullptr; }
    if (!in.err.empty()) { g_err = "reading VCF: " + in.err; return nullptr; }
    // uniform [n][npop][2] layout: records parsed before a later header line knew fewer populations
    P->npop = (int)pops.size();
    bool ragged = false;
    for (int v : rec_npop) ragged = ragged || v != P->npop;
    if (ragged) {
      std::vector<int32_t> full((size_t)P->n * 2 * (size_t)P->npop, 0);
      size_t o = 0;
      for (int64_t i = 0; i < P->n; ++i) {
        const size_t w = 2 * (size_t)rec_npop[(size_t)i];
        std::copy(P->cnt.begin() + (long)o, P->cnt.begin() + (long)(o + w), full.begin() + (long)((size_t)i * 2 * (size_t)P->npop));
        o += w;
      }
      P->cnt.swap(full);
    }
    P->key_off.push_back((int64_t)P->keys_blob.size());
    P->first_col = first_col;
    for (auto& c : pops) { P->pop_blob += c; P->pop_blob.push_back('\n'); }
    for (auto& c : vocab_list) { P->vocab_blob += c; P->vocab_blob.push_back('\n'); }
    return guard.release();
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void tdsfs_vcf_counts_free(void* h) { delete (CountsOut*)h; }
// dims[4] = n records, n populations, bytes of the key blob, n vocabulary entries
void tdsfs_vcf_counts_dims(void* h, int64_t* dims) {
  CountsOut* P = (CountsOut*)h;
  dims[0] = P->n; dims[1] = P->npop; dims[2] = (int64_t)P->keys_blob.size(); dims[3] = (int64_t)P->first_col.size();
}
const char* tdsfs_vcf_counts_keys(void* h) { return ((CountsOut*)h)->keys_blob.data(); }
const int64_t* tdsfs_vcf_counts_key_off(void* h) { return ((CountsOut*)h)->key_off.data(); }
const char* tdsfs_vcf_counts_refalt(void* h) { return ((CountsOut*)h)->refalt.data(); }
const int32_t* tdsfs_vcf_counts_ann_codes(void* h) { return ((CountsOut*)h)->ann_code.data(); }
const int32_t* tdsfs_vcf_counts_cnt(void* h) { return ((CountsOut*)h)->cnt.data(); }
const int32_t* tdsfs_vcf_counts_ncols(void* h) { return ((CountsOut*)h)->ncols.data(); }
const int32_t* tdsfs_vcf_counts_first_col(void* h) { return ((CountsOut*)h)->first_col.data(); }
const char* tdsfs_vcf_counts_pops(void* h) { return ((CountsOut*)h)->pop_blob.c_str(); }
const char* tdsfs_vcf_counts_vocab(void* h) { return ((CountsOut*)h)->vocab_blob.c_str(); }

}  // extern "C"
