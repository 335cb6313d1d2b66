// TEST INFRASTRUCTURE ONLY (see tests/hostsim/cuda_runtime.h): drives the host side of the C ABI
// under AddressSanitizer/UBSan -- growth by doubling, staged host adds, shard-file save / load /
// range append with the double-buffered copy loops, error paths, and that every handle frees
// everything it allocated.  Exit code 0 = all checks passed.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/tristage.h"
#include "cuda_runtime.h"

#define CHECK(c)                                                                    \
  do {                                                                              \
    if (!(c)) { fprintf(stderr, "CHECK failed %s:%d: %s  [%s]\n", __FILE__, __LINE__, #c, ts_last_error()); exit(1); } \
  } while (0)

static std::vector<char> slurp(const std::string& p) {
  FILE* f = fopen(p.c_str(), "rb");
  CHECK(f);
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> b((size_t)n);
  CHECK(fread(b.data(), 1, (size_t)n, f) == (size_t)n);
  fclose(f);
  return b;
}

static float frand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return ((*s >> 8) & 0xffff) / 32768.0f - 1.0f; }

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : "/tmp";
  unsigned seed = 7;
  // ---------------------------------------------------------------- index ----
  {
    const int dim = 70;   // ld = 72: pad columns must be zero
    const int64_t n1 = 3000, n2 = 5000;
    std::vector<float> x((size_t)(n1 + n2) * dim);
    for (auto& v : x) v = frand(&seed);
    ts_index* h = nullptr;
    CHECK(ts_index_create(&h, 0, dim, TS_BF16, TS_METRIC_IP, 0) == TS_OK);
    CHECK(ts_index_add(h, x.data(), n1, TS_F32, 0, 1, nullptr) == TS_OK);       // 1024 -> 4096 rows
    CHECK(ts_index_add(h, x.data() + n1 * dim, n2, TS_F32, 0, 1, nullptr) == TS_OK);  // -> 8192 rows, copies the old ones
    CHECK(ts_index_ntotal(h) == n1 + n2 && ts_index_dim(h) == dim && ts_index_dtype(h) == TS_BF16);
    CHECK(ts_index_add(h, x.data(), 5, TS_F16, 0, 0, nullptr) == TS_ERR_INVALID);   // src dtype must be f32 or storage
    std::vector<float> back((size_t)(n1 + n2) * dim);
    CHECK(ts_index_get_rows(h, 0, n1 + n2, back.data()) == TS_OK);
    for (int64_t r = 0; r < n1 + n2; r += 97) {   // rows are unit length after x/(|x|+1e-8) and bf16 rounding
      double ss = 0;
      for (int c = 0; c < dim; ++c) ss += (double)back[(size_t)r * dim + c] * back[(size_t)r * dim + c];
      CHECK(ss > 0.98 && ss < 1.02);
    }
    CHECK(ts_index_get_rows(h, n1 + n2 - 1, 2, back.data()) == TS_ERR_INVALID);
    CHECK(ts_index_set_id_base(h, 123456789012ll) == TS_OK);
    const std::string p = dir + "/hs_index.tsshard";
    CHECK(ts_index_save(h, p.c_str()) == TS_OK);
    CHECK(ts_file_verify(p.c_str()) == TS_OK);
    ts_file_info fi;
    CHECK(ts_file_probe(p.c_str(), &fi) == TS_OK);
    CHECK(fi.kind == TS_FILE_INDEX && fi.n == n1 + n2 && fi.ld == 72 && fi.dim == dim && fi.id_base == 123456789012ll);
    CHECK(fi.payload_bytes == (uint64_t)(n1 + n2) * 72 * 2);
    ts_index* h2 = nullptr;
    CHECK(ts_index_load(&h2, 0, p.c_str()) == TS_OK);
    CHECK(ts_index_ntotal(h2) == n1 + n2);
    std::vector<float> back2((size_t)(n1 + n2) * dim);
    CHECK(ts_index_get_rows(h2, 0, n1 + n2, back2.data()) == TS_OK);
    CHECK(memcmp(back.data(), back2.data(), back.size() * 4) == 0);
    // loaded + re-saved file is byte-identical (pad columns included)
    const std::string p2 = dir + "/hs_index2.tsshard";
    CHECK(ts_index_save(h2, p2.c_str()) == TS_OK);
    CHECK(slurp(p) == slurp(p2));
    // range appends, in pieces, rebuild the same rows
    ts_index* h3 = nullptr;
    CHECK(ts_index_create(&h3, 0, dim, TS_BF16, TS_METRIC_IP, 10) == TS_OK);   // tiny reservation: must grow
    CHECK(ts_index_append_file(h3, p.c_str(), 0, 1, nullptr) == TS_OK);
    CHECK(ts_index_append_file(h3, p.c_str(), 1, 4999, nullptr) == TS_OK);
    CHECK(ts_index_append_file(h3, p.c_str(), 5000, 3000, nullptr) == TS_OK);
    CHECK(ts_index_append_file(h3, p.c_str(), 8000, 0, nullptr) == TS_OK);
    CHECK(ts_index_append_file(h3, p.c_str(), 7999, 2, nullptr) == TS_ERR_INVALID);
    CHECK(ts_index_ntotal(h3) == 8000);
    CHECK(ts_index_set_id_base(h3, 123456789012ll) == TS_OK);
    const std::string p3 = dir + "/hs_index3.tsshard";
    CHECK(ts_index_save(h3, p3.c_str()) == TS_OK);
    CHECK(slurp(p) == slurp(p3));
    // mismatching handle / corrupt file / allocation failure
    ts_index* hb = nullptr;
    CHECK(ts_index_create(&hb, 0, dim, TS_F16, TS_METRIC_IP, 0) == TS_OK);
    CHECK(ts_index_append_file(hb, p.c_str(), 0, 10, nullptr) == TS_ERR_INVALID);
    CHECK(ts_index_ntotal(hb) == 0);
    std::vector<char> raw = slurp(p);
    raw[fi.payload_offset + 31337] ^= 1;
    const std::string pbad = dir + "/hs_bad.tsshard";
    { FILE* f = fopen(pbad.c_str(), "wb"); CHECK(f); fwrite(raw.data(), 1, raw.size(), f); fclose(f); }
    ts_index* hbad = nullptr;
    CHECK(ts_index_load(&hbad, 0, pbad.c_str()) == TS_ERR_IO && hbad == nullptr);
    CHECK(ts_file_verify(pbad.c_str()) == TS_ERR_IO);
    hostsim_fail_malloc_over = 1 << 20;
    CHECK(ts_index_load(&hbad, 0, p.c_str()) == TS_ERR_NOMEM && hbad == nullptr);
    const long long before = ts_index_ntotal(h3);
    CHECK(ts_index_add(h3, x.data(), n1, TS_F32, 0, 1, nullptr) == TS_ERR_NOMEM);   // growth refused: state unchanged
    CHECK(ts_index_ntotal(h3) == before);
    hostsim_fail_malloc_over = 0;
    CHECK(ts_index_reset(h3) == TS_OK && ts_index_ntotal(h3) == 0);
    ts_index_destroy(h); ts_index_destroy(h2); ts_index_destroy(h3); ts_index_destroy(hb);
  }
  // ------------------------------------------------------- cosine + fp16 -----
  {
    const int dim = 16; const int64_t n = 777;
    std::vector<float> x((size_t)n * dim);
    for (auto& v : x) v = 3.f * frand(&seed);
    ts_index* h = nullptr;
    CHECK(ts_index_create(&h, 0, dim, TS_F16, TS_METRIC_COSINE, 0) == TS_OK);
    CHECK(ts_index_add(h, x.data(), n, TS_F32, 0, 0, nullptr) == TS_OK);
    const std::string p = dir + "/hs_cos.tsshard";
    CHECK(ts_index_save(h, p.c_str()) == TS_OK && ts_file_verify(p.c_str()) == TS_OK);
    ts_file_info fi;
    CHECK(ts_file_probe(p.c_str(), &fi) == TS_OK && fi.table_bytes == (uint64_t)n * 4 && fi.metric == TS_METRIC_COSINE);
    ts_index* h2 = nullptr;
    CHECK(ts_index_create(&h2, 0, dim, TS_F16, TS_METRIC_COSINE, 0) == TS_OK);
    CHECK(ts_index_append_file(h2, p.c_str(), 0, 400, nullptr) == TS_OK);
    CHECK(ts_index_append_file(h2, p.c_str(), 400, 377, nullptr) == TS_OK);
    const std::string p2 = dir + "/hs_cos2.tsshard";
    CHECK(ts_index_save(h2, p2.c_str()) == TS_OK);
    CHECK(slurp(p) == slurp(p2));
    ts_index_destroy(h); ts_index_destroy(h2);
  }
  // ------------------------------------------------------------- tokstore ----
  {
    const int dim = 24; const int n_docs = 2500;
    std::vector<int32_t> lens(n_docs);
    int64_t total = 0;
    for (auto& L : lens) { seed = seed * 1664525u + 1013904223u; L = 1 + (int)((seed >> 10) % 256); total += L; }
    std::vector<float> tok((size_t)total * dim);
    for (auto& v : tok) v = frand(&seed);
    ts_tokstore* s = nullptr;
    CHECK(ts_tokstore_create(&s, 0, dim, TS_BF16, 0, 0) == TS_OK);
    // three adds: the doc tables and the token rows both have to grow
    int64_t t0 = 0; for (int i = 0; i < 700; ++i) t0 += lens[i];
    int64_t t1 = t0; for (int i = 700; i < 1900; ++i) t1 += lens[i];
    CHECK(ts_tokstore_add(s, tok.data(), TS_F32, 0, lens.data(), 700, 1, nullptr) == TS_OK);
    CHECK(ts_tokstore_add(s, tok.data() + t0 * dim, TS_F32, 0, lens.data() + 700, 1200, 1, nullptr) == TS_OK);
    CHECK(ts_tokstore_add(s, tok.data() + t1 * dim, TS_F32, 0, lens.data() + 1900, 600, 1, nullptr) == TS_OK);
    CHECK(ts_tokstore_ndocs(s) == n_docs && ts_tokstore_ntokens(s) == total && ts_tokstore_dim(s) == dim);
    int32_t bad_len[1] = {257};
    CHECK(ts_tokstore_add(s, tok.data(), TS_F32, 0, bad_len, 1, 1, nullptr) == TS_ERR_INVALID);
    CHECK(ts_tokstore_set_id_base(s, 40) == TS_OK);
    const std::string p = dir + "/hs_tok.tsshard";
    CHECK(ts_tokstore_save(s, p.c_str()) == TS_OK && ts_file_verify(p.c_str()) == TS_OK);
    ts_file_info fi;
    CHECK(ts_file_probe(p.c_str(), &fi) == TS_OK && fi.kind == TS_FILE_TOKSTORE && fi.n == n_docs && fi.ntokens == total && fi.id_base == 40);
    ts_tokstore* s2 = nullptr;
    CHECK(ts_tokstore_load(&s2, 0, p.c_str()) == TS_OK);
    CHECK(ts_tokstore_ndocs(s2) == n_docs && ts_tokstore_ntokens(s2) == total && ts_tokstore_dtype(s2) == TS_BF16);
    const std::string p2 = dir + "/hs_tok2.tsshard";
    CHECK(ts_tokstore_save(s2, p2.c_str()) == TS_OK);
    CHECK(slurp(p) == slurp(p2));
    // doc ranges appended piecewise rebuild the same store (offsets are rebased per piece)
    ts_tokstore* s3 = nullptr;
    CHECK(ts_tokstore_create(&s3, 0, dim, TS_BF16, 0, 0) == TS_OK);
    CHECK(ts_tokstore_append_file(s3, p.c_str(), 0, 1000, nullptr) == TS_OK);
    CHECK(ts_tokstore_append_file(s3, p.c_str(), 1000, 1, nullptr) == TS_OK);
    CHECK(ts_tokstore_append_file(s3, p.c_str(), 1001, 1499, nullptr) == TS_OK);
    CHECK(ts_tokstore_append_file(s3, p.c_str(), 2499, 2, nullptr) == TS_ERR_INVALID);
    CHECK(ts_tokstore_ndocs(s3) == n_docs && ts_tokstore_ntokens(s3) == total);
    CHECK(ts_tokstore_set_id_base(s3, 40) == TS_OK);
    const std::string p3 = dir + "/hs_tok3.tsshard";
    CHECK(ts_tokstore_save(s3, p3.c_str()) == TS_OK);
    CHECK(slurp(p) == slurp(p3));
    // the host writer produces the same file from the same (already normalised, bf16) rows
    std::vector<char> raw = slurp(p);
    std::vector<uint16_t> unp((size_t)total * dim);
    {
      const int64_t* off = (const int64_t*)(raw.data() + fi.table_offset);
      const uint16_t* body = (const uint16_t*)(raw.data() + fi.payload_offset);
      size_t at = 0;
      for (int d = 0; d < n_docs; ++d) { memcpy(&unp[at], body + (size_t)off[d] * dim, (size_t)lens[d] * dim * 2); at += (size_t)lens[d] * dim; }
    }
    const std::string p4 = dir + "/hs_tok4.tsshard";
    CHECK(ts_file_write_tokstore_host(p4.c_str(), dim, TS_BF16, n_docs, 40, lens.data(), unp.data()) == TS_OK);
    CHECK(slurp(p) == slurp(p4));
    ts_tokstore* sb = nullptr;
    CHECK(ts_tokstore_create(&sb, 0, dim + 8, TS_BF16, 0, 0) == TS_OK);
    CHECK(ts_tokstore_append_file(sb, p.c_str(), 0, 10, nullptr) == TS_ERR_INVALID);
    ts_index* wrong = nullptr;
    CHECK(ts_index_load(&wrong, 0, p.c_str()) == TS_ERR_IO);     // a token shard is not an index shard
    CHECK(ts_tokstore_reset(s3) == TS_OK && ts_tokstore_ndocs(s3) == 0);
    ts_tokstore_destroy(s); ts_tokstore_destroy(s2); ts_tokstore_destroy(s3); ts_tokstore_destroy(sb);
  }
  CHECK(hostsim_live_allocs == 0);
  CHECK(hostsim_live_pinned == 0);
  printf("hostsim ok\n");
  return 0;
}
