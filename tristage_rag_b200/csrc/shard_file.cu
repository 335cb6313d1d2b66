// Shard files: persistence of one corpus shard (ts_index) or one token shard
// (ts_tokstore), the on-disk side of the row-sharded layout.
//
// Replaces faiss.write_index / faiss.read_index
// (/root/reference/src/stage1_retriever.py:436,463); the token-shard file has
// no reference equivalent (the reference re-encodes every candidate at query
// time, src/stage2_rescorer.py:255-259).  The pickle side-car with
// documents / doc_metadata / bm25_index stays in Python
// (tristage_rag_b200/stage1_retriever.py::save_index).
//
// File layout (little endian, every section starts on a 4096-byte boundary so
// the payload can be mmap'ed and handed to the DMA engine page by page):
//
//   [0, 4096)      ShardHeader, zero padded
//   table section  index:    inv_norm fp32[n]          (TS_METRIC_COSINE only)
//                  tokstore: doc_off int64[n] (first padded row of each doc,
//                            relative to the payload), then doc_len int32[n]
//   payload        index:    rows[n][ld]   storage dtype (pad columns zero)
//                  tokstore: tok[nrows][dim] storage dtype, docs padded to 8 rows
//
// Both sections carry an XXH64 digest in the header.  A file is written to
// "<path>.tmp" with a zeroed header, the real header goes in last and the
// file is renamed into place, so a torn write never parses.
//
// Loading maps the file, and moves the payload through two pinned staging
// buffers so the page-cache read of chunk i+1 overlaps the H2D copy of chunk
// i.  Any [lo, lo+n) row (doc) range of a file can be appended to a handle:
// that is what lets a corpus saved by G ranks be loaded by G' ranks
// (tristage_rag_b200/dist.py::plan_reshard).
//
// ts_file_probe / ts_file_verify / ts_file_write_*_host are host-only (no CUDA
// call) so tools and the CPU test-suite can inspect and produce shard files.
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "ts_handles.h"

using namespace ts;

namespace {

constexpr uint64_t kAlign = 4096;
#ifdef TS_HOSTSIM
constexpr size_t kStageBytes = 48 * 1024 + 32;   // tests/hostsim: small + odd so the ping-pong loops run many times
#else
constexpr size_t kStageBytes = 32ull << 20;   // per pinned staging buffer
#endif
const char kMagic[8] = {'T', 'S', 'S', 'H', 'A', 'R', 'D', '2'};

struct ShardHeader {
  char magic[8];
  uint32_t version, kind;
  int32_t dim, ld, dtype, metric;
  int64_t n, nrows, ntokens, id_base;
  uint64_t table_offset, table_bytes, payload_offset, payload_bytes;
  uint64_t table_hash, payload_hash;
};
static_assert(sizeof(ShardHeader) <= kAlign, "header must fit its page");

uint64_t align_up(uint64_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

// ------------------------------------------------------------- XXH64 -------
// Streaming XXH64 (Y. Collet's public algorithm), restated; any split of the
// input into update() calls gives the same digest.
struct XXH64 {
  static constexpr uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
                            P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;
  uint64_t v[4];
  uint64_t total = 0;
  unsigned char buf[32];
  int nbuf = 0;
  uint64_t seed;
  explicit XXH64(uint64_t s = 0) : seed(s) { v[0] = s + P1 + P2; v[1] = s + P2; v[2] = s; v[3] = s - P1; }
  static uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
  static uint64_t rd64(const unsigned char* p) { uint64_t x; memcpy(&x, p, 8); return x; }
  static uint32_t rd32(const unsigned char* p) { uint32_t x; memcpy(&x, p, 4); return x; }
  static uint64_t round(uint64_t acc, uint64_t in) { return rotl(acc + in * P2, 31) * P1; }
  static uint64_t merge(uint64_t h, uint64_t val) { return (h ^ round(0, val)) * P1 + P4; }
  void stripe(const unsigned char* p) {
    v[0] = round(v[0], rd64(p)); v[1] = round(v[1], rd64(p + 8));
    v[2] = round(v[2], rd64(p + 16)); v[3] = round(v[3], rd64(p + 24));
  }
  void update(const void* data, size_t n) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    total += n;
    if (nbuf) {
      const size_t take = (size_t)(32 - nbuf) < n ? (size_t)(32 - nbuf) : n;
      memcpy(buf + nbuf, p, take);
      nbuf += (int)take; p += take; n -= take;
      if (nbuf < 32) return;
      stripe(buf);
      nbuf = 0;
    }
    for (; n >= 32; p += 32, n -= 32) stripe(p);
    if (n) { memcpy(buf, p, n); nbuf = (int)n; }
  }
  uint64_t digest() const {
    uint64_t h;
    if (total >= 32) {
      h = rotl(v[0], 1) + rotl(v[1], 7) + rotl(v[2], 12) + rotl(v[3], 18);
      h = merge(h, v[0]); h = merge(h, v[1]); h = merge(h, v[2]); h = merge(h, v[3]);
    } else {
      h = seed + P5;
    }
    h += total;
    const unsigned char* p = buf;
    int n = nbuf;
    for (; n >= 8; p += 8, n -= 8) h = rotl(h ^ round(0, rd64(p)), 27) * P1 + P4;
    if (n >= 4) { h = rotl(h ^ ((uint64_t)rd32(p) * P1), 23) * P2 + P3; p += 4; n -= 4; }
    for (; n > 0; ++p, --n) h = rotl(h ^ ((uint64_t)*p * P5), 11) * P1;
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
  }
};

// ------------------------------------------------------- mapped reader -----
struct MappedFile {
  int fd = -1;
  const unsigned char* base = nullptr;
  size_t size = 0;
  ShardHeader hd{};
  ~MappedFile() { close_(); }
  void close_() {
    if (base) munmap(const_cast<unsigned char*>(base), size);
    if (fd >= 0) ::close(fd);
    base = nullptr; fd = -1;
  }
  // open + map + validate the header and section bounds
  int open_(const char* path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) { set_error("cannot open %s: %s", path, strerror(errno)); return TS_ERR_IO; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (uint64_t)sb.st_size < kAlign) { set_error("%s is not a tristage shard file (too short)", path); return TS_ERR_IO; }
    size = (size_t)sb.st_size;
    void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { set_error("mmap(%s) failed: %s", path, strerror(errno)); return TS_ERR_IO; }
    base = static_cast<const unsigned char*>(m);
    madvise(m, size, MADV_SEQUENTIAL);
    memcpy(&hd, base, sizeof(hd));
    if (memcmp(hd.magic, kMagic, 8) != 0 || hd.version != 2) { set_error("%s is not a tristage shard file (bad magic/version)", path); return TS_ERR_IO; }
    const bool kind_ok = hd.kind == TS_FILE_INDEX || hd.kind == TS_FILE_TOKSTORE;
    const bool dt_ok = hd.dtype == TS_F32 || hd.dtype == TS_BF16 || hd.dtype == TS_F16;
    if (!kind_ok || !dt_ok || hd.dim <= 0 || hd.ld < hd.dim || hd.n < 0 || hd.nrows < 0 || hd.ntokens < 0) {
      set_error("%s: corrupt shard header", path); return TS_ERR_IO;
    }
    const uint64_t esz = (uint64_t)dtype_size(hd.dtype);
    const uint64_t want_payload = (uint64_t)hd.nrows * (uint64_t)hd.ld * esz;
    uint64_t want_table = 0;
    if (hd.kind == TS_FILE_INDEX) {
      if (hd.nrows != hd.n || hd.ld != row_pitch(hd.dim, hd.dtype)) { set_error("%s: corrupt shard header (rows/pitch)", path); return TS_ERR_IO; }
      want_table = hd.metric == TS_METRIC_COSINE ? (uint64_t)hd.n * 4 : 0;
    } else {
      if (hd.ld != hd.dim) { set_error("%s: corrupt shard header (pitch)", path); return TS_ERR_IO; }
      want_table = (uint64_t)hd.n * 12;
    }
    if (hd.payload_bytes != want_payload || hd.table_bytes != want_table ||
        hd.table_offset % kAlign || hd.payload_offset % kAlign ||
        hd.table_offset + hd.table_bytes > size || hd.payload_offset + hd.payload_bytes > size ||
        hd.table_offset < kAlign || hd.payload_offset < kAlign) {
      set_error("%s: section table does not match the file (truncated?)", path); return TS_ERR_IO;
    }
    return TS_OK;
  }
  const unsigned char* table() const { return base + hd.table_offset; }
  const unsigned char* payload() const { return base + hd.payload_offset; }
};

// ------------------------------------------------------- atomic writer -----
struct ShardWriter {
  FILE* f = nullptr;
  std::string path, tmp;
  uint64_t pos = 0;
  XXH64 hash;
  int begin(const char* p) {
    path = p; tmp = path + ".tmp";
    f = fopen(tmp.c_str(), "wb");
    if (!f) { set_error("cannot open %s for writing: %s", tmp.c_str(), strerror(errno)); return TS_ERR_IO; }
    setvbuf(f, nullptr, _IONBF, 0);
    std::vector<char> zero(kAlign, 0);
    if (fwrite(zero.data(), 1, kAlign, f) != kAlign) return fail();
    pos = kAlign;
    return TS_OK;
  }
  int fail() { set_error("write to %s failed: %s", tmp.c_str(), strerror(errno)); abort_(); return TS_ERR_IO; }
  void abort_() { if (f) { fclose(f); f = nullptr; remove(tmp.c_str()); } }
  void start_section() { hash = XXH64(); }
  int write(const void* p, size_t n) {
    if (n == 0) return TS_OK;
    if (fwrite(p, 1, n, f) != n) return fail();
    hash.update(p, n);
    pos += n;
    return TS_OK;
  }
  int pad() {
    const uint64_t to = align_up(pos);
    if (to > pos) {
      std::vector<char> zero((size_t)(to - pos), 0);
      if (fwrite(zero.data(), 1, zero.size(), f) != zero.size()) return fail();
      pos = to;
    }
    return TS_OK;
  }
  int finish(const ShardHeader& hd) {
    if (fseek(f, 0, SEEK_SET) != 0 || fwrite(&hd, sizeof(hd), 1, f) != 1) return fail();
    if (fflush(f) != 0) return fail();
    // the data must be durable BEFORE the name appears: fsync the file, rename, fsync the directory --
    // otherwise a crash can leave a renamed file whose pages never reached the disk
    if (fsync(fileno(f)) != 0) return fail();
    if (fclose(f) != 0) { f = nullptr; remove(tmp.c_str()); set_error("closing %s failed", tmp.c_str()); return TS_ERR_IO; }
    f = nullptr;
    if (rename(tmp.c_str(), path.c_str()) != 0) { set_error("rename %s -> %s failed: %s", tmp.c_str(), path.c_str(), strerror(errno)); remove(tmp.c_str()); return TS_ERR_IO; }
    const size_t slash = path.find_last_of('/');
    const std::string dir = slash == std::string::npos ? std::string(".") : (slash == 0 ? std::string("/") : path.substr(0, slash));
    const int dfd = open(dir.c_str(), O_RDONLY | O_DIRECTORY);
    if (dfd >= 0) { (void)fsync(dfd); close(dfd); }   // best effort: some filesystems refuse fsync on a directory
    return TS_OK;
  }
  ~ShardWriter() { abort_(); }
};

ShardHeader make_header(int kind, int dim, int ld, int dtype, int metric, int64_t n, int64_t nrows, int64_t ntokens,
                        int64_t id_base) {
  ShardHeader hd;
  memset(&hd, 0, sizeof(hd));
  memcpy(hd.magic, kMagic, 8);
  hd.version = 2; hd.kind = (uint32_t)kind; hd.dim = dim; hd.ld = ld; hd.dtype = dtype; hd.metric = metric;
  hd.n = n; hd.nrows = nrows; hd.ntokens = ntokens; hd.id_base = id_base;
  return hd;
}

// --------------------------------------------- pinned double buffering -----
struct Staging {
  void* buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2];
  bool pinned = false, made = false;
  bool pending[2] = {false, false};
  int init() {
    for (int i = 0; i < 2; ++i) {
      if (cudaHostAlloc(&buf[i], kStageBytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        for (int j = 0; j < i; ++j) cudaFreeHost(buf[j]);
        buf[0] = buf[1] = nullptr;
        break;
      }
    }
    pinned = buf[0] != nullptr;
    if (!pinned) {   // pageable fallback for the STAGING memory only (still the same CUDA copies)
      for (int i = 0; i < 2; ++i) { buf[i] = malloc(kStageBytes); if (!buf[i]) { set_error("staging malloc failed"); return TS_ERR_NOMEM; } }
    }
    for (int i = 0; i < 2; ++i) TS_CUDA_OK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    made = true;
    return TS_OK;
  }
  int wait(int i) {
    if (pending[i]) { TS_CUDA_OK(cudaEventSynchronize(ev[i])); pending[i] = false; }
    return TS_OK;
  }
  ~Staging() {
    if (made) for (int i = 0; i < 2; ++i) cudaEventDestroy(ev[i]);
    for (int i = 0; i < 2; ++i) if (buf[i]) { if (pinned) cudaFreeHost(buf[i]); else free(buf[i]); }
  }
};

// host bytes -> device, chunked through the staging pair (overlaps the page-cache read with the DMA)
int copy_in(Staging& sg, void* dst_dev, const unsigned char* src, size_t bytes, cudaStream_t st, XXH64* hash) {
  int which = 0;
  for (size_t off = 0; off < bytes; off += kStageBytes, which ^= 1) {
    const size_t m = bytes - off < kStageBytes ? bytes - off : kStageBytes;
    int rc = sg.wait(which);
    if (rc) return rc;
    memcpy(sg.buf[which], src + off, m);
    if (hash) hash->update(sg.buf[which], m);
    TS_CUDA_OK(cudaMemcpyAsync((char*)dst_dev + off, sg.buf[which], m, cudaMemcpyHostToDevice, st));
    TS_CUDA_OK(cudaEventRecord(sg.ev[which], st));
    sg.pending[which] = true;
  }
  int rc = sg.wait(0);
  if (rc) return rc;
  return sg.wait(1);
}

// device bytes -> file section (D2H of chunk i+1 overlaps hashing + write of chunk i)
int copy_out(Staging& sg, ShardWriter& w, const void* src_dev, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return TS_OK;
  auto issue = [&](size_t off, int which) -> int {
    const size_t m = bytes - off < kStageBytes ? bytes - off : kStageBytes;
    TS_CUDA_OK(cudaMemcpyAsync(sg.buf[which], (const char*)src_dev + off, m, cudaMemcpyDeviceToHost, st));
    TS_CUDA_OK(cudaEventRecord(sg.ev[which], st));
    sg.pending[which] = true;
    return TS_OK;
  };
  int rc = issue(0, 0);
  if (rc) return rc;
  int which = 0;
  for (size_t off = 0; off < bytes; off += kStageBytes, which ^= 1) {
    const size_t m = bytes - off < kStageBytes ? bytes - off : kStageBytes;
    if (off + kStageBytes < bytes && (rc = issue(off + kStageBytes, which ^ 1))) return rc;
    if ((rc = sg.wait(which))) return rc;
    if ((rc = w.write(sg.buf[which], m))) return rc;
  }
  return TS_OK;
}

int verify_section(const char* path, const char* what, const unsigned char* p, uint64_t bytes, uint64_t want) {
  XXH64 h;
  for (uint64_t off = 0; off < bytes; off += (64ull << 20)) h.update(p + off, (size_t)(bytes - off < (64ull << 20) ? bytes - off : (64ull << 20)));
  if (h.digest() != want) { set_error("%s: %s checksum mismatch (file is corrupt)", path, what); return TS_ERR_IO; }
  return TS_OK;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------ host only ----
int ts_file_probe(const char* path, ts_file_info* out) {
  if (!path || !out) { set_error("ts_file_probe: invalid argument"); return TS_ERR_INVALID; }
  MappedFile mf;
  int rc = mf.open_(path);
  if (rc) return rc;
  const ShardHeader& h = mf.hd;
  memset(out, 0, sizeof(*out));
  out->kind = (int32_t)h.kind; out->version = (int32_t)h.version; out->dim = h.dim; out->ld = h.ld; out->dtype = h.dtype;
  out->metric = h.metric; out->n = h.n; out->nrows = h.nrows; out->ntokens = h.ntokens; out->id_base = h.id_base;
  out->table_offset = h.table_offset; out->table_bytes = h.table_bytes;
  out->payload_offset = h.payload_offset; out->payload_bytes = h.payload_bytes;
  out->table_hash = h.table_hash; out->payload_hash = h.payload_hash;
  return TS_OK;
}

int ts_file_verify(const char* path) {
  if (!path) { set_error("ts_file_verify: invalid argument"); return TS_ERR_INVALID; }
  MappedFile mf;
  int rc = mf.open_(path);
  if (rc) return rc;
  if ((rc = verify_section(path, "table", mf.table(), mf.hd.table_bytes, mf.hd.table_hash))) return rc;
  if ((rc = verify_section(path, "payload", mf.payload(), mf.hd.payload_bytes, mf.hd.payload_hash))) return rc;
  if (mf.hd.kind == TS_FILE_TOKSTORE) {
    // the doc table must describe the payload: ascending 8-row aligned offsets, lengths in range
    const int64_t* off = reinterpret_cast<const int64_t*>(mf.table());
    const int32_t* len = reinterpret_cast<const int32_t*>(mf.table() + (size_t)mf.hd.n * 8);
    int64_t expect = 0, tokens = 0;
    for (int64_t i = 0; i < mf.hd.n; ++i) {
      if (off[i] != expect || len[i] < 1 || len[i] > TS_S2_MAX_LD) { set_error("%s: doc table entry %lld is inconsistent", path, (long long)i); return TS_ERR_IO; }
      expect += (len[i] + 7) & ~7;
      tokens += len[i];
    }
    if (expect != mf.hd.nrows || tokens != mf.hd.ntokens) { set_error("%s: doc table does not add up to the payload", path); return TS_ERR_IO; }
  }
  return TS_OK;
}

int ts_file_write_index_host(const char* path, int dim, int storage_dtype, int metric, int64_t n, int64_t id_base,
                             const void* rows_storage, const float* inv_norm) {
  if (!path || dim <= 0 || n < 0 || (n > 0 && !rows_storage) ||
      (storage_dtype != TS_F32 && storage_dtype != TS_BF16 && storage_dtype != TS_F16) ||
      (metric != TS_METRIC_IP && metric != TS_METRIC_COSINE) || (metric == TS_METRIC_COSINE && n > 0 && !inv_norm)) {
    set_error("ts_file_write_index_host: invalid argument");
    return TS_ERR_INVALID;
  }
  const int ld = row_pitch(dim, storage_dtype);
  ShardHeader hd = make_header(TS_FILE_INDEX, dim, ld, storage_dtype, metric, n, n, 0, id_base);
  ShardWriter w;
  int rc = w.begin(path);
  if (rc) return rc;
  hd.table_offset = w.pos;
  w.start_section();
  if (metric == TS_METRIC_COSINE && (rc = w.write(inv_norm, (size_t)n * 4))) return rc;
  hd.table_bytes = w.pos - hd.table_offset; hd.table_hash = w.hash.digest();
  if ((rc = w.pad())) return rc;
  hd.payload_offset = w.pos;
  w.start_section();
  const size_t row_b = (size_t)ld * dtype_size(storage_dtype);
  const int64_t chunk = (int64_t)((64ull << 20) / row_b) + 1;
  for (int64_t s = 0; s < n; s += chunk) {
    const int64_t m = n - s < chunk ? n - s : chunk;
    if ((rc = w.write((const char*)rows_storage + (size_t)s * row_b, (size_t)m * row_b))) return rc;
  }
  hd.payload_bytes = w.pos - hd.payload_offset; hd.payload_hash = w.hash.digest();
  return w.finish(hd);
}

int ts_file_write_tokstore_host(const char* path, int dim, int storage_dtype, int64_t n_docs, int64_t id_base,
                                const int32_t* lens, const void* tok_storage) {
  if (!path || dim <= 0 || n_docs < 0 || (n_docs > 0 && (!lens || !tok_storage)) ||
      (storage_dtype != TS_F32 && storage_dtype != TS_BF16 && storage_dtype != TS_F16)) {
    set_error("ts_file_write_tokstore_host: invalid argument");
    return TS_ERR_INVALID;
  }
  std::vector<int64_t> off((size_t)n_docs);
  int64_t rows = 0, tokens = 0;
  for (int64_t i = 0; i < n_docs; ++i) {
    if (lens[i] < 1 || lens[i] > TS_S2_MAX_LD) { set_error("ts_file_write_tokstore_host: doc %lld has %d tokens (allowed 1..%d)", (long long)i, lens[i], TS_S2_MAX_LD); return TS_ERR_INVALID; }
    off[(size_t)i] = rows;
    rows += (lens[i] + 7) & ~7;
    tokens += lens[i];
  }
  ShardHeader hd = make_header(TS_FILE_TOKSTORE, dim, dim, storage_dtype, TS_METRIC_IP, n_docs, rows, tokens, id_base);
  ShardWriter w;
  int rc = w.begin(path);
  if (rc) return rc;
  hd.table_offset = w.pos;
  w.start_section();
  if ((rc = w.write(off.data(), (size_t)n_docs * 8))) return rc;
  if ((rc = w.write(lens, (size_t)n_docs * 4))) return rc;
  hd.table_bytes = w.pos - hd.table_offset; hd.table_hash = w.hash.digest();
  if ((rc = w.pad())) return rc;
  hd.payload_offset = w.pos;
  w.start_section();
  const size_t row_b = (size_t)dim * dtype_size(storage_dtype);
  std::vector<char> zero(7 * row_b, 0);
  const char* src = static_cast<const char*>(tok_storage);
  for (int64_t i = 0; i < n_docs; ++i) {
    const size_t L = (size_t)lens[i], padr = (size_t)(((lens[i] + 7) & ~7) - lens[i]);
    if ((rc = w.write(src, L * row_b))) return rc;
    if (padr && (rc = w.write(zero.data(), padr * row_b))) return rc;
    src += L * row_b;
  }
  hd.payload_bytes = w.pos - hd.payload_offset; hd.payload_hash = w.hash.digest();
  return w.finish(hd);
}

// --------------------------------------------------------------- index -----
int ts_index_save(const ts_index* h, const char* path) {
  if (!h || !path) { set_error("ts_index_save: invalid argument"); return TS_ERR_INVALID; }
  TS_CUDA_OK(cudaSetDevice(h->device));
  TS_CUDA_OK(cudaDeviceSynchronize());   // rows may still be in flight on the caller's stream
  Staging sg;
  int rc = sg.init();
  if (rc) return rc;
  ShardHeader hd = make_header(TS_FILE_INDEX, h->dim, h->ld, h->dtype, h->metric, h->n, h->n, 0, h->id_base);
  ShardWriter w;
  if ((rc = w.begin(path))) return rc;
  hd.table_offset = w.pos;
  w.start_section();
  if (h->metric == TS_METRIC_COSINE && (rc = copy_out(sg, w, h->inv_norm, (size_t)h->n * 4, 0))) return rc;
  hd.table_bytes = w.pos - hd.table_offset; hd.table_hash = w.hash.digest();
  if ((rc = w.pad())) return rc;
  hd.payload_offset = w.pos;
  w.start_section();
  if ((rc = copy_out(sg, w, h->rows, (size_t)h->n * h->ld * dtype_size(h->dtype), 0))) return rc;
  hd.payload_bytes = w.pos - hd.payload_offset; hd.payload_hash = w.hash.digest();
  return w.finish(hd);
}

int ts_index_append_file(ts_index* h, const char* path, int64_t row_lo, int64_t n_rows, void* stream) {
  if (!h || !path || row_lo < 0 || n_rows < 0) { set_error("ts_index_append_file: invalid argument"); return TS_ERR_INVALID; }
  MappedFile mf;
  int rc = mf.open_(path);
  if (rc) return rc;
  const ShardHeader& hd = mf.hd;
  if (hd.kind != TS_FILE_INDEX) { set_error("%s is not an index shard file", path); return TS_ERR_IO; }
  if (hd.dim != h->dim || hd.ld != h->ld || hd.dtype != h->dtype || hd.metric != h->metric) {
    set_error("%s: dim/dtype/metric (%d/%d/%d) do not match the index (%d/%d/%d)", path, hd.dim, hd.dtype, hd.metric, h->dim, h->dtype, h->metric);
    return TS_ERR_INVALID;
  }
  if (row_lo + n_rows > hd.n) { set_error("%s: rows [%lld, %lld) outside the file's %lld rows", path, (long long)row_lo, (long long)(row_lo + n_rows), (long long)hd.n); return TS_ERR_INVALID; }
  if (n_rows == 0) return TS_OK;
  if (h->n + n_rows > 0xFFFFFFF0ll) { set_error("ts_index_append_file: shard limited to 2^32 rows"); return TS_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->device));
  if ((rc = index_reserve(h, h->n + n_rows, st))) return rc;
  Staging sg;
  if ((rc = sg.init())) return rc;
  const size_t row_b = (size_t)h->ld * dtype_size(h->dtype);
  const bool whole = (row_lo == 0 && n_rows == hd.n);
  // a partial range (re-sharded load) cannot be checked while it streams: check the whole file first, so a
  // corrupt or torn shard never loads silently on the path this format exists for
  if (!whole) {
    if ((rc = verify_section(path, "table", mf.table(), hd.table_bytes, hd.table_hash))) return rc;
    if ((rc = verify_section(path, "payload", mf.payload(), hd.payload_bytes, hd.payload_hash))) return rc;
  }
  XXH64 hash;
  if ((rc = copy_in(sg, (char*)h->rows + (size_t)h->n * row_b, mf.payload() + (size_t)row_lo * row_b, (size_t)n_rows * row_b, st,
                    whole ? &hash : nullptr))) return rc;
  if (whole && hash.digest() != hd.payload_hash) { set_error("%s: payload checksum mismatch (file is corrupt)", path); return TS_ERR_IO; }
  if (h->metric == TS_METRIC_COSINE) {
    XXH64 th;
    if ((rc = copy_in(sg, h->inv_norm + h->n, mf.table() + (size_t)row_lo * 4, (size_t)n_rows * 4, st, whole ? &th : nullptr))) return rc;
    if (whole && th.digest() != hd.table_hash) { set_error("%s: table checksum mismatch (file is corrupt)", path); return TS_ERR_IO; }
  }
  h->n += n_rows;
  return TS_OK;
}

int ts_index_load(ts_index** out, int device, const char* path) {
  if (!out || !path) { set_error("ts_index_load: invalid argument"); return TS_ERR_INVALID; }
  ts_file_info fi;
  int rc = ts_file_probe(path, &fi);
  if (rc) return rc;
  if (fi.kind != TS_FILE_INDEX) { set_error("%s is not an index shard file", path); return TS_ERR_IO; }
  ts_index* h = nullptr;
  if ((rc = ts_index_create(&h, device, fi.dim, fi.dtype, fi.metric, fi.n))) return rc;
  if ((rc = ts_index_append_file(h, path, 0, fi.n, nullptr))) { ts_index_destroy(h); return rc; }
  h->id_base = fi.id_base;
  *out = h;
  return TS_OK;
}

// ------------------------------------------------------------ tokstore -----
int ts_tokstore_save(const ts_tokstore* h, const char* path) {
  if (!h || !path) { set_error("ts_tokstore_save: invalid argument"); return TS_ERR_INVALID; }
  TS_CUDA_OK(cudaSetDevice(h->device));
  TS_CUDA_OK(cudaDeviceSynchronize());
  Staging sg;
  int rc = sg.init();
  if (rc) return rc;
  ShardHeader hd = make_header(TS_FILE_TOKSTORE, h->dim, h->dim, h->dtype, TS_METRIC_IP, h->ndocs, h->nrows, h->ntokens, h->id_base);
  ShardWriter w;
  if ((rc = w.begin(path))) return rc;
  hd.table_offset = w.pos;
  w.start_section();
  if ((rc = copy_out(sg, w, h->doc_off, (size_t)h->ndocs * 8, 0))) return rc;
  if ((rc = copy_out(sg, w, h->doc_len, (size_t)h->ndocs * 4, 0))) return rc;
  hd.table_bytes = w.pos - hd.table_offset; hd.table_hash = w.hash.digest();
  if ((rc = w.pad())) return rc;
  hd.payload_offset = w.pos;
  w.start_section();
  // shard files hold ONE image whatever the HBM layout is: row-major rows, zero pad rows.  A tile-layout
  // shard is converted in place for the duration of the copy (two extra HBM passes; saving is rare and disk-bound)
  const bool tile = (h->layout == kTokTile);
  if (tile && (rc = launch_tok_relayout(h->tok, h->dtype, h->doc_off, h->doc_len, 0, h->ndocs, h->dim, kTokRowMajor, 0))) return rc;
  if (tile) TS_CUDA_OK(cudaDeviceSynchronize());
  rc = copy_out(sg, w, h->tok, (size_t)h->nrows * h->dim * dtype_size(h->dtype), 0);
  if (tile) {
    const int rc2 = launch_tok_relayout(h->tok, h->dtype, h->doc_off, h->doc_len, 0, h->ndocs, h->dim, kTokTile, 0);
    if (rc2 == TS_OK) { TS_CUDA_OK(cudaDeviceSynchronize()); }
    if (!rc) rc = rc2;
  }
  if (rc) return rc;
  hd.payload_bytes = w.pos - hd.payload_offset; hd.payload_hash = w.hash.digest();
  return w.finish(hd);
}

int ts_tokstore_append_file(ts_tokstore* h, const char* path, int64_t doc_lo, int64_t n_docs, void* stream) {
  if (!h || !path || doc_lo < 0 || n_docs < 0) { set_error("ts_tokstore_append_file: invalid argument"); return TS_ERR_INVALID; }
  MappedFile mf;
  int rc = mf.open_(path);
  if (rc) return rc;
  const ShardHeader& hd = mf.hd;
  if (hd.kind != TS_FILE_TOKSTORE) { set_error("%s is not a token-store shard file", path); return TS_ERR_IO; }
  if (hd.dim != h->dim || hd.dtype != h->dtype) { set_error("%s: dim/dtype (%d/%d) do not match the store (%d/%d)", path, hd.dim, hd.dtype, h->dim, h->dtype); return TS_ERR_INVALID; }
  if (doc_lo + n_docs > hd.n) { set_error("%s: docs [%lld, %lld) outside the file's %lld docs", path, (long long)doc_lo, (long long)(doc_lo + n_docs), (long long)hd.n); return TS_ERR_INVALID; }
  if (n_docs == 0) return TS_OK;
  if (!(doc_lo == 0 && n_docs == hd.n)) {    // partial range: check the whole file before trusting any of it
    if ((rc = verify_section(path, "table", mf.table(), hd.table_bytes, hd.table_hash))) return rc;
    if ((rc = verify_section(path, "payload", mf.payload(), hd.payload_bytes, hd.payload_hash))) return rc;
  }
  const int64_t* off = reinterpret_cast<const int64_t*>(mf.table());
  const int32_t* len = reinterpret_cast<const int32_t*>(mf.table() + (size_t)hd.n * 8);
  // the doc table is trusted for addressing: validate the slice before using it
  std::vector<int64_t> dst_off((size_t)n_docs);
  const int64_t row_lo = off[doc_lo];
  int64_t expect = row_lo, tokens = 0;
  if (row_lo < 0 || (row_lo & 7)) { set_error("%s: corrupt doc table", path); return TS_ERR_IO; }
  for (int64_t i = 0; i < n_docs; ++i) {
    const int64_t o = off[doc_lo + i];
    const int32_t L = len[doc_lo + i];
    if (o != expect || L < 1 || L > TS_S2_MAX_LD) { set_error("%s: corrupt doc table at doc %lld", path, (long long)(doc_lo + i)); return TS_ERR_IO; }
    dst_off[(size_t)i] = h->nrows + (o - row_lo);
    expect += (L + 7) & ~7;
    tokens += L;
  }
  const int64_t n_rows = expect - row_lo;
  if (expect > hd.nrows) { set_error("%s: doc table points past the payload", path); return TS_ERR_IO; }
  if (h->nrows + n_rows > 0x7FFFFF00ll) { set_error("ts_tokstore_append_file: shard limited to 2^31 token rows"); return TS_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->device));
  if ((rc = tok_reserve(h, h->ndocs + n_docs, h->nrows + n_rows, st))) return rc;
  Staging sg;
  if ((rc = sg.init())) return rc;
  const bool whole = (doc_lo == 0 && n_docs == hd.n);
  if (whole && (rc = verify_section(path, "table", mf.table(), hd.table_bytes, hd.table_hash))) return rc;
  TS_CUDA_OK(cudaMemcpyAsync(h->doc_off + h->ndocs, dst_off.data(), (size_t)n_docs * 8, cudaMemcpyHostToDevice, st));
  if ((rc = copy_in(sg, h->doc_len + h->ndocs, reinterpret_cast<const unsigned char*>(len + doc_lo), (size_t)n_docs * 4, st, nullptr))) return rc;
  const size_t row_b = (size_t)h->dim * dtype_size(h->dtype);
  XXH64 hash;
  if ((rc = copy_in(sg, (char*)h->tok + (size_t)h->nrows * row_b, mf.payload() + (size_t)row_lo * row_b, (size_t)n_rows * row_b, st,
                    whole ? &hash : nullptr))) return rc;
  TS_CUDA_OK(cudaStreamSynchronize(st));   // dst_off (host vector) was read asynchronously
  if (whole && hash.digest() != hd.payload_hash) { set_error("%s: payload checksum mismatch (file is corrupt)", path); return TS_ERR_IO; }
  if (h->layout == kTokTile) {             // the file image is row-major: bring the appended docs into the shard's layout
    if ((rc = launch_tok_relayout(h->tok, h->dtype, h->doc_off, h->doc_len, h->ndocs, n_docs, h->dim, kTokTile, st))) return rc;
    TS_CUDA_OK(cudaStreamSynchronize(st));
  }
  h->ndocs += n_docs; h->nrows += n_rows; h->ntokens += tokens;
  return TS_OK;
}

int ts_tokstore_load(ts_tokstore** out, int device, const char* path) {
  if (!out || !path) { set_error("ts_tokstore_load: invalid argument"); return TS_ERR_INVALID; }
  ts_file_info fi;
  int rc = ts_file_probe(path, &fi);
  if (rc) return rc;
  if (fi.kind != TS_FILE_TOKSTORE) { set_error("%s is not a token-store shard file", path); return TS_ERR_IO; }
  ts_tokstore* h = nullptr;
  if ((rc = ts_tokstore_create(&h, device, fi.dim, fi.dtype, fi.n, fi.ntokens))) return rc;
  if ((rc = ts_tokstore_append_file(h, path, 0, fi.n, nullptr))) { ts_tokstore_destroy(h); return rc; }
  h->id_base = fi.id_base;
  *out = h;
  return TS_OK;
}

}  // extern "C"
