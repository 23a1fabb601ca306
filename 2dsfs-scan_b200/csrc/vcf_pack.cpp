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
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <future>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Fix { int64_t snp; int32_t pop, dref, dalt; };

struct Rec {                 // one accepted VCF record
  std::string chrom, pos_str, ann;
  long long pos = 0;
  std::vector<uint32_t> words;         // RW words, pop1 then pop2
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

void parse_line(const std::string& line, const Plan& plan, Rec& r) {
  std::vector<std::pair<const char*, size_t>> cols, sub;
  split_char(line.data(), line.size(), '\t', cols);
  if (cols.size() < 9) { r.err = "VCF record with fewer than 9 columns"; return; }
  auto S = [&](int i) { return std::string(cols[i].first, cols[i].second); };
  // annotation (:94-99)
  split_char(cols[7].first, cols[7].second, '|', sub);
  r.ann = sub.size() >= 2 ? std::string(sub[1].first, sub[1].second) : "No annotation";
  std::string filt = S(6);
  if (filt != "PASS" && filt != ".") return;
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
  r.words.assign(RW, 0u);
  // every slot starts as MISSING (a column absent from this record contributes (0, 0)); real calls overwrite it
  for (int pop = 0; pop < 2; ++pop) {
    const int ns = pop ? plan.ns2 : plan.ns1;
    uint32_t* w = r.words.data() + (pop ? plan.W1 : 0);
    for (int s = 0; s < ns; ++s) w[2 * (s >> 5) + 1] |= 1u << (s & 31);  // code 2 = (lo 0, hi 1)
  }
  const size_t ncol = std::min(plan.col_pop.size(), cols.size() - 9);
  for (size_t j = 0; j < ncol; ++j) {
    const int pop = plan.col_pop[j];
    if (pop < 0) continue;
    split_char(cols[9 + j].first, cols[9 + j].second, ':', sub);
    if ((int)sub.size() <= gti) { r.err = "list index out of range (sample without GT sub-field)"; return; }
    const char* gt = sub[gti].first;
    const size_t gn = sub[gti].second;
    int ref = 0, alt = 0;
    for (size_t q = 0; q < gn; q += 2) {  // gt[::2]
      if (gt[q] == '0') ++ref; else if (gt[q] == '1') ++alt;
    }
    uint32_t code;
    if (ref == 2 && alt == 0) code = 0; else if (ref == 1 && alt == 1) code = 1; else if (ref == 0 && alt == 2) code = 3;
    else {
      code = 2;
      if (ref || alt) r.fix.push_back({0, pop, ref, alt});
    }
    const int slot = plan.col_slot[j];
    // samples 32g..32g+31 of a population: word 2g = lo bits of their codes, word 2g+1 = hi bits
    uint32_t* w = r.words.data() + (pop ? plan.W1 : 0) + 2 * (slot >> 5);
    const uint32_t bit = 1u << (slot & 31);
    w[0] = (w[0] & ~bit) | ((code & 1u) ? bit : 0u);
    w[1] = (w[1] & ~bit) | ((code & 2u) ? bit : 0u);
  }
  r.ok = true;
}

bool read_line(gzFile f, std::string& out) {
  out.clear();
  char buf[1 << 16];
  while (gzgets(f, buf, sizeof buf)) {
    out.append(buf);
    if (!out.empty() && out.back() == '\n') return true;
  }
  return !out.empty();
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
    gzFile gz = gzopen(vcf_path, "rb");
    if (!gz) { g_err = std::string("cannot open VCF ") + vcf_path; return nullptr; }
    gzbuffer(gz, 1 << 20);
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());

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

    std::vector<Rec> all;                      // accepted records in file order
    std::vector<std::string> batch;
    const size_t BATCH = 2048;
    std::string line;
    int64_t n_records = 0, n_skipped = 0;
    std::string first_err;
    bool have_records = false;
    auto flush = [&]() {
      if (batch.empty()) return;
      std::vector<Rec> recs(batch.size());
      const int nt = (int)std::min<size_t>(nthreads, batch.size());
      std::vector<std::future<void>> fut;
      for (int t = 0; t < nt; ++t)
        fut.push_back(std::async(std::launch::async, [&, t]() {
          for (size_t i = t; i < batch.size(); i += nt) parse_line(batch[i], plan, recs[i]);
        }));
      for (auto& f : fut) f.get();
      for (auto& r : recs) {
        ++n_records;
        if (!r.err.empty()) { if (first_err.empty()) first_err = r.err; continue; }
        if (!r.ok) { ++n_skipped; continue; }
        all.push_back(std::move(r));
      }
      batch.clear();
    };
    while (read_line(gz, line)) {
      if (line.size() >= 2 && line[0] == '#' && line[1] == '#') continue;
      if (!line.empty() && line[0] == '#') {
        flush();
        if (have_records) { g_err = "second header line after records is not supported by the packer"; gzclose(gz); return nullptr; }
        // header_cols = line.split(); samples found in the popmap extend the positional list (:78-85)
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
        continue;
      }
      have_records = true;
      batch.push_back(line);
      if (batch.size() >= BATCH) flush();
    }
    flush();
    gzclose(gz);
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
      for (int w = 0; w < RW; ++w) P->G[(size_t)(((s >> 5) * RW + w) * 32 + (s & 31))] = r.words[(size_t)w];
      for (Fix f : r.fix) { f.snp = s; P->fix.push_back(f); }
      auto it = vocab.find(r.ann);
      if (it == vocab.end()) { it = vocab.emplace(r.ann, (int)P->ann_vocab.size()).first; P->ann_vocab.push_back(r.ann); }
      P->ann_code[(size_t)s] = it->second;
      if (order[(size_t)s] == keep.size() - 1) P->last_key_row = s;
    }
    P->off.push_back(P->S);
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

}  // extern "C"
