// Readers for the reference's index artefacts: what `PanGenie -f <prefix>` loads before the hot
// path starts (src/commands.cpp:760-790 the UniqueKmersMap archive, :858-874 / src/kmerparser.cpp:16-28 the per-chromosome
// k-mer tables).  Host code only; the result is the flat pg_panel form the rest of the library consumes.
//
// `<prefix>_UniqueKmersMap.cereal` is a cereal BinaryOutputArchive of `UniqueKmersMap` (src/commands.hpp:11-28): raw
// little-endian; size_t -> u64; string / vector / map = u64 length + elements; polymorphic shared_ptr = u32 polymorphic
// id (MSB set on first use, then a u64-length-prefixed class name) + u32 pointer id (MSB set = the object follows).
// Fields of the two concrete classes: src/biallelicuniquekmers.hpp:102-114, src/multiallelicuniquekmers.hpp:101-113,
// src/kmerpath.hpp:26-33, src/kmerpath16.hpp:26-33.
// `<prefix>_<chrom>_kmers.tsv.gz`: 5 tab-separated columns, comma lists of k-mers, "nan" if empty (header line starts
// with '#', src/stepwiseuniquekmercomputer.cpp:105).
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

struct ChromData {
  std::string name;
  uint32_t n_paths = 0;
  std::vector<uint64_t> positions;
  std::vector<uint16_t> path_to_allele, coverage, kmer_counts, allele_ids, allele_koff;
  std::vector<uint32_t> kmer_off, allele_off, allele_kmask, flank_off;
  std::vector<uint8_t> allele_undef;
  std::vector<uint64_t> kmer_codes, flank_codes;
  bool has_kmers = false;
};

struct Cursor {
  const uint8_t* p;
  size_t n, o = 0;
  bool ok = true;
  template <class T>
  T take() {
    T v{};
    if (o + sizeof(T) > n) {
      ok = false;
      return v;
    }
    memcpy(&v, p + o, sizeof(T));
    o += sizeof(T);
    return v;
  }
  std::string str() {
    const uint64_t len = take<uint64_t>();
    if (!ok || o + len > n) {
      ok = false;
      return std::string();
    }
    std::string s(reinterpret_cast<const char*>(p + o), (size_t)len);
    o += len;
    return s;
  }
};

bool read_all(const std::string& path, std::vector<uint8_t>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(sz > 0 ? (size_t)sz : 0);
  const size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
  fclose(f);
  return got == out.size();
}

struct AlleleRec {
  uint16_t id, koff;
  uint32_t mask;
  uint8_t undef;
};

}  // namespace

struct pg_index {
  uint32_t kmer_size = 0;
  bool add_reference = false;
  std::string segments_path;
  std::vector<ChromData> chroms;  // std::map order of the archive = the order `-f` processes them (src/commands.cpp:781-782)
};

namespace {

// 2-bit code of an ASCII k-mer, first base most significant; false if a character is not a base
bool encode_kmer(const char* s, size_t len, uint64_t& out) {
  uint64_t v = 0;
  for (size_t i = 0; i < len; ++i) {
    uint64_t c;
    switch (s[i]) {
      case 'A': case 'a': c = 0; break;
      case 'C': case 'c': c = 1; break;
      case 'G': case 'g': c = 2; break;
      case 'T': case 't': c = 3; break;
      default: return false;
    }
    v = (v << 2) | c;
  }
  out = v;
  return true;
}

int parse_archive(const std::string& path, pg_index* ix) {
  std::vector<uint8_t> buf;
  if (!read_all(path, buf)) return pg::fail(PG_ERR_IO, "File " + path + " cannot be opened.");
  Cursor r{buf.data(), buf.size()};
  ix->kmer_size = (uint32_t)r.take<uint64_t>();
  std::map<uint32_t, std::string> class_names;
  const uint64_t n_chrom = r.take<uint64_t>();
  for (uint64_t c = 0; c < n_chrom && r.ok; ++c) {
    ChromData cd;
    cd.name = r.str();
    const uint64_t V = r.take<uint64_t>();
    if (!r.ok || V > 0xfffffff0ull) break;
    cd.kmer_off.push_back(0);
    cd.allele_off.push_back(0);
    for (uint64_t v = 0; v < V && r.ok; ++v) {
      const uint32_t pid = r.take<uint32_t>();
      if (pid & 0x80000000u) class_names[pid & 0x7fffffffu] = r.str();
      const auto it = class_names.find(pid & 0x7fffffffu);
      if (!r.ok || it == class_names.end()) return pg::fail(PG_ERR_FORMAT, path + ": unknown polymorphic type id");
      const bool bi = it->second == "BiallelicUniqueKmers";
      if (!bi && it->second != "MultiallelicUniqueKmers") return pg::fail(PG_ERR_FORMAT, path + ": unexpected class " + it->second);
      const uint32_t ptr = r.take<uint32_t>();
      if (!(ptr & 0x80000000u)) return pg::fail(PG_ERR_FORMAT, path + ": shared UniqueKmers objects are not supported");
      cd.positions.push_back(r.take<uint64_t>());
      cd.coverage.push_back((uint16_t)r.take<float>());
      const uint64_t current_index = r.take<uint64_t>();
      const uint64_t n_counts = r.take<uint64_t>();
      if (!r.ok || n_counts > (1u << 20) || current_index != n_counts) return pg::fail(PG_ERR_FORMAT, path + ": corrupt k-mer count list");
      for (uint64_t i = 0; i < n_counts; ++i) cd.kmer_counts.push_back(r.take<uint16_t>());
      const uint64_t n_alleles = r.take<uint64_t>();
      if (!r.ok || n_alleles > 65536) return pg::fail(PG_ERR_FORMAT, path + ": corrupt allele map");
      std::vector<AlleleRec> al;
      for (uint64_t a = 0; a < n_alleles; ++a) {
        AlleleRec x;
        if (bi) {
          x.id = r.take<uint8_t>();
          x.koff = r.take<uint16_t>();
          x.mask = r.take<uint16_t>();
        } else {
          x.id = r.take<uint16_t>();
          x.koff = r.take<uint16_t>();
          x.mask = r.take<uint32_t>();
        }
        x.undef = r.take<uint8_t>();
        al.push_back(x);
      }
      std::sort(al.begin(), al.end(), [](const AlleleRec& a, const AlleleRec& b) { return a.id < b.id; });
      for (const AlleleRec& x : al) {
        cd.allele_ids.push_back(x.id);
        cd.allele_koff.push_back(x.koff);
        cd.allele_kmask.push_back(x.mask);
        cd.allele_undef.push_back(x.undef);
      }
      const uint64_t n_paths = r.take<uint64_t>();
      if (!r.ok || n_paths > 65535) return pg::fail(PG_ERR_FORMAT, path + ": corrupt path list");
      if (v == 0) cd.n_paths = (uint32_t)n_paths;
      else if (n_paths != cd.n_paths) return pg::fail(PG_ERR_FORMAT, path + ": variants of one chromosome are covered by different numbers of paths");
      for (uint64_t q = 0; q < n_paths; ++q) cd.path_to_allele.push_back(bi ? (uint16_t)r.take<uint8_t>() : r.take<uint16_t>());
      cd.kmer_off.push_back((uint32_t)cd.kmer_counts.size());
      cd.allele_off.push_back((uint32_t)cd.allele_ids.size());
    }
    ix->chroms.push_back(std::move(cd));
  }
  for (int m = 0; m < 2 && r.ok; ++m) {  // runtimes, sampling_runtimes
    const uint64_t n = r.take<uint64_t>();
    for (uint64_t i = 0; i < n && r.ok; ++i) {
      r.str();
      r.take<double>();
    }
  }
  ix->add_reference = r.take<uint8_t>() != 0;
  if (!r.ok) return pg::fail(PG_ERR_FORMAT, path + ": truncated archive");
  if (r.o != r.n) return pg::fail(PG_ERR_FORMAT, path + ": trailing bytes in archive");
  return PG_OK;
}

// splits a comma list ("nan" = empty) of k-mers and appends their codes
int append_kmers(const char* s, size_t len, std::vector<uint64_t>& out, uint32_t k, const std::string& path) {
  if (len == 3 && memcmp(s, "nan", 3) == 0) return PG_OK;
  size_t i = 0;
  while (i <= len) {
    size_t j = i;
    while (j < len && s[j] != ',') ++j;
    uint64_t code;
    if (j - i != k || !encode_kmer(s + i, j - i, code)) return pg::fail(PG_ERR_FORMAT, path + ": malformed k-mer in table");
    out.push_back(code);
    i = j + 1;
    if (j == len) break;
  }
  return PG_OK;
}

int parse_kmer_table(const std::string& path, uint32_t k, ChromData& cd) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) return pg::fail(PG_ERR_IO, "File " + path + " cannot be opened.");
  struct Closer {
    gzFile f;
    ~Closer() { gzclose(f); }
  } closer{f};
  std::string line;
  std::vector<char> chunk(1 << 16);
  size_t v = 0;
  cd.flank_off.assign(1, 0);
  cd.kmer_codes.clear();
  cd.flank_codes.clear();
  bool eof = false;
  while (!eof) {
    line.clear();
    while (true) {  // one logical line, however long
      if (!gzgets(f, chunk.data(), (int)chunk.size())) {
        eof = true;
        break;
      }
      line += chunk.data();
      if (!line.empty() && line.back() == '\n') break;
    }
    while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
    if (line.empty()) continue;
    if (line[0] == '#') continue;  // header
    size_t tab[4], nt = 0;
    for (size_t i = 0; i < line.size() && nt < 4; ++i)
      if (line[i] == '\t') tab[nt++] = i;
    if (nt != 4) return pg::fail(PG_ERR_FORMAT, path + ": expected 5 tab-separated columns");
    if (v >= cd.positions.size()) return pg::fail(PG_ERR_FORMAT, path + ": more variants than in the UniqueKmersMap");
    const uint64_t start = strtoull(line.c_str() + tab[0] + 1, nullptr, 10);
    if (start != cd.positions[v]) return pg::fail(PG_ERR_FORMAT, path + ": variant order differs from the UniqueKmersMap");
    const size_t before = cd.kmer_codes.size();
    PG_TRY(append_kmers(line.c_str() + tab[2] + 1, tab[3] - tab[2] - 1, cd.kmer_codes, k, path));
    if (cd.kmer_codes.size() - before != cd.kmer_off[v + 1] - cd.kmer_off[v])
      return pg::fail(PG_ERR_FORMAT, path + ": number of unique k-mers differs from the UniqueKmersMap");
    PG_TRY(append_kmers(line.c_str() + tab[3] + 1, line.size() - tab[3] - 1, cd.flank_codes, k, path));
    cd.flank_off.push_back((uint32_t)cd.flank_codes.size());
    ++v;
  }
  if (v != cd.positions.size()) return pg::fail(PG_ERR_FORMAT, path + ": fewer variants than in the UniqueKmersMap");
  cd.has_kmers = true;
  return PG_OK;
}

}  // namespace

extern "C" pg_index* pg_index_open_archive(const char* archive_path) {
  pg::clear_error();
  if (!archive_path) {
    pg::fail(PG_ERR_ARG, "null path");
    return nullptr;
  }
  pg_index* ix = new pg_index();
  if (parse_archive(archive_path, ix) != PG_OK) {
    delete ix;
    return nullptr;
  }
  return ix;
}

extern "C" pg_index* pg_index_open(const char* prefix, int with_kmers) {
  pg::clear_error();
  if (!prefix) {
    pg::fail(PG_ERR_ARG, "null prefix");
    return nullptr;
  }
  const std::string pre(prefix);
  pg_index* ix = pg_index_open_archive((pre + "_UniqueKmersMap.cereal").c_str());
  if (!ix) return nullptr;
  ix->segments_path = pre + "_path_segments.fasta";
  if (with_kmers) {
    if (ix->kmer_size < 1 || ix->kmer_size > 32) {
      pg::fail(PG_ERR_ARG, "k must be in [1,32]");
      delete ix;
      return nullptr;
    }
    for (ChromData& cd : ix->chroms) {
      if (parse_kmer_table(pre + "_" + cd.name + "_kmers.tsv.gz", ix->kmer_size, cd) != PG_OK) {
        delete ix;
        return nullptr;
      }
    }
  }
  return ix;
}

extern "C" void pg_index_close(pg_index* ix) { delete ix; }
extern "C" uint32_t pg_index_kmer_size(const pg_index* ix) { return ix ? ix->kmer_size : 0; }
extern "C" uint32_t pg_index_n_chromosomes(const pg_index* ix) { return ix ? (uint32_t)ix->chroms.size() : 0; }
extern "C" int pg_index_add_reference(const pg_index* ix) { return ix && ix->add_reference ? 1 : 0; }
extern "C" const char* pg_index_segments_path(const pg_index* ix) { return ix ? ix->segments_path.c_str() : ""; }
extern "C" const char* pg_index_chromosome_name(const pg_index* ix, uint32_t i) {
  return ix && i < ix->chroms.size() ? ix->chroms[i].name.c_str() : "";
}

extern "C" int pg_index_panel(pg_index* ix, uint32_t i, pg_panel* out) {
  pg::clear_error();
  if (!ix || !out || i >= ix->chroms.size()) return pg::fail(PG_ERR_ARG, "invalid argument");
  ChromData& cd = ix->chroms[i];
  memset(out, 0, sizeof(*out));
  out->n_variants = (uint32_t)cd.positions.size();
  out->n_paths = cd.n_paths;
  out->positions = cd.positions.data();
  out->path_to_allele = cd.path_to_allele.data();
  out->coverage = cd.coverage.data();
  out->kmer_offsets = cd.kmer_off.data();
  out->kmer_counts = cd.kmer_counts.data();
  out->allele_offsets = cd.allele_off.data();
  out->allele_ids = cd.allele_ids.data();
  out->allele_undefined = cd.allele_undef.data();
  out->allele_kmer_offset = cd.allele_koff.data();
  out->allele_kmer_mask = cd.allele_kmask.data();
  if (cd.has_kmers) {
    out->kmer_codes = cd.kmer_codes.data();
    out->flank_offsets = cd.flank_off.data();
    out->flank_codes = cd.flank_codes.data();
  }
  return PG_OK;
}
