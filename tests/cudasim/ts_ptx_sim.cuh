// TEST INFRASTRUCTURE ONLY (tests/cudasim) -- functional stand-ins for csrc/ts_ptx.cuh so the
// tensor-path kernels (s1_umma_kernel, maxsim_umma_kernel) can be EXECUTED on the CPU emulator:
//   * mbarrier: arrival count + transaction bytes + phase bit, try_wait by phase parity; a waiter
//     that finds its phase incomplete yields to the other simulated threads;
//   * TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B): the box is copied at issue time, out-of-range
//     elements read as zero, every 16-byte chunk lands at the XOR-swizzled shared-memory address,
//     the full box size is credited to the mbarrier;
//   * tcgen05.mma kind::f16 (M128, N from the instruction descriptor, K16): operands are fetched
//     through the SAME swizzle from the K-major SWIZZLE_128B shared-memory descriptors (start
//     address, SBO) the kernels build, products accumulate in fp32 into a 128-lane x 512-column
//     TMEM array; tcgen05.commit arrives at once; tcgen05.ld 32x32b.x32 reads the caller's lane;
//   * bar.sync id, n: counting barrier among the arriving threads.
// Everything "asynchronous" completes at issue time: one legal schedule of the real machine.
// That checks data movement, descriptor arithmetic, phase bookkeeping, masks and the fused
// selection -- not latency, not the memory model.  Operation order inside a dot product differs
// from the tensor core's, so scores agree with hardware only to rounding (the parity rule has a
// tolerance); variants compared with each other on the emulator are bit-comparable.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <functional>

#include "cudasim.h"

struct alignas(64) CUtensorMap {
  const void* base;
  int64_t rows;
  int dim, ld, box_cols, box_rows, dtype;
  char pad_[128 - 8 - 8 - 5 * 4];
};

namespace cudasim {
// per-CTA emulation state (reset at the start of every simulated block)
unsigned char* smem_base();
uint32_t* tmem();                        // [128][512]
void mbar_init(uint32_t addr, uint32_t count);
void mbar_arrive(uint32_t addr, uint32_t expect_tx_bytes);
void mbar_complete_tx(uint32_t addr, uint32_t bytes);
void mbar_expect_tx(uint32_t addr, uint32_t bytes);   // more pending bytes, no arrival
bool mbar_phase_done(uint32_t addr, uint32_t parity);
void named_barrier(int id, int nthreads);
// thread-block clusters (launch_cluster)
int cluster_rank();
void cluster_barrier();
unsigned char* smem_base_of(int cta);
uint32_t* tmem_of(int cta);
void mbar_arrive_at(int cta, uint32_t addr, uint32_t expect_tx_bytes);
void mbar_complete_tx_at(int cta, uint32_t addr, uint32_t bytes);
void yield_spin();                       // give the other simulated threads a turn (no progress made)
// asynchronous engines: executed at once, or -- under CUDASIM_ASYNC -- some scheduler rounds later (TMA operations
// independently, tensor-core operations in issue order), so a kernel that touches data before its barrier fails
void defer_tma(std::function<void()> fn);
void defer_mma(std::function<void()> fn);
}  // namespace cudasim

namespace ts {
namespace ptx {

inline uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(static_cast<const unsigned char*>(p) - cudasim::smem_base());
}
inline uint32_t sw128(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }   // 16-byte chunk index ^= row within the 8-row atom

// ------------------------------------------------------------- mbarrier ----
inline void mbar_init(uint64_t* bar, uint32_t count) { cudasim::mbar_init(smem_u32(bar), count); }
inline void fence_mbar_init() {}
inline void fence_proxy_async_smem() {}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { cudasim::mbar_arrive(smem_u32(bar), bytes); }
inline void mbar_arrive(uint64_t* bar) { cudasim::mbar_arrive(smem_u32(bar), 0); }
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  if (cudasim::mbar_phase_done(smem_u32(bar), parity)) return true;
  cudasim::yield_spin();
  return cudasim::mbar_phase_done(smem_u32(bar), parity);
}
#ifndef TS_WAIT_TIMEOUT_CYCLES
#define TS_WAIT_TIMEOUT_CYCLES (8000000000ll)
#endif
inline void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  (void)tag;
  while (!mbar_try_wait(bar, parity)) {}   // a wait that can never complete is reported by the scheduler (deadlock)
}
inline void named_bar_sync(int id, int nthreads) { cudasim::named_barrier(id, nthreads); }

// ------------------------------------------------------------------ TMA ----
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;
inline void prefetch_tmap(const CUtensorMap*) {}

inline int sim_make_tmap_2d(CUtensorMap* out, const void* base, int dtype, int64_t rows, int dim, int ld, int box_cols, int box_rows) {
  memset(out, 0, sizeof(*out));
  out->base = base; out->rows = rows; out->dim = dim; out->ld = ld; out->box_cols = box_cols; out->box_rows = box_rows; out->dtype = dtype;
  const int esz = (dtype == 0 /* TS_F32, mapped as TFLOAT32 */) ? 4 : 2;
  if (box_cols * esz != 128 || box_rows < 1 || box_rows > 256 || (reinterpret_cast<uintptr_t>(base) & 15) || (ld * esz) % 16) {
    fprintf(stderr, "[cudasim] tensor map violates the SWIZZLE_128B / alignment rules\n");
    abort();
  }
  return 0;
}

// fp32 -> tf32 as a TFLOAT32 tensor map delivers it: round to nearest (even) on the 13 dropped mantissa bits
inline uint32_t sim_round_tf32(uint32_t u) {
  if ((u & 0x7F800000u) == 0x7F800000u) return u;      // inf / nan
  return (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
}

inline void tma_load_2d(void* smem_dst, const CUtensorMap* mp, uint64_t* bar, int c0, int c1, uint64_t) {
  const uint32_t dst = smem_u32(smem_dst), bar_addr = smem_u32(bar);
  if (dst & 1023u) { fprintf(stderr, "[cudasim] TMA destination %u is not 1024-byte aligned (swizzle atom)\n", dst); abort(); }
  const CUtensorMap map = *mp;
  cudasim::defer_tma([=]() {
  const CUtensorMap* m = &map;
  unsigned char* sm = cudasim::smem_base();
  if (m->dtype == 0) {                                  // fp32 rows through a TFLOAT32 map: 32 elements per 128-byte row
    const uint32_t* g32 = static_cast<const uint32_t*>(m->base);
    for (int r = 0; r < m->box_rows; ++r) {
      const int64_t row = (int64_t)c1 + r;
      for (int c = 0; c < m->box_cols; ++c) {
        const int col = c0 + c;
        uint32_t v = 0;
        if (row >= 0 && row < m->rows && col >= 0 && col < m->dim) v = sim_round_tf32(g32[row * m->ld + col]);
        memcpy(sm + sw128(dst + (uint32_t)r * 128u + (uint32_t)c * 4u), &v, 4);
      }
    }
    cudasim::mbar_complete_tx(bar_addr, (uint32_t)m->box_rows * 128u);
    return;
  }
  const uint16_t* g = static_cast<const uint16_t*>(m->base);
  for (int r = 0; r < m->box_rows; ++r) {
    const int64_t row = (int64_t)c1 + r;
    for (int c = 0; c < m->box_cols; ++c) {
      const int col = c0 + c;
      uint16_t v = 0;                                   // out-of-bounds elements are filled with zeros
      if (row >= 0 && row < m->rows && col >= 0 && col < m->dim) v = g[row * m->ld + col];
      memcpy(sm + sw128(dst + (uint32_t)r * 128u + (uint32_t)c * 2u), &v, 2);
    }
  }
  cudasim::mbar_complete_tx(bar_addr, (uint32_t)m->box_rows * 128u);   // the whole box counts, also past the end
  });
}

// cp.async.bulk global -> shared: contiguous bytes, the size is credited to the mbarrier on completion
inline void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t) {
  const uint32_t dst = smem_u32(smem_dst), bar_addr = smem_u32(bar);
  if ((dst & 15u) || (reinterpret_cast<uintptr_t>(gsrc) & 15u) || (bytes & 15u) || bytes == 0) {
    fprintf(stderr, "[cudasim] cp.async.bulk needs 16-byte aligned addresses and size (dst %u bytes %u)\n", dst, bytes); abort();
  }
  cudasim::defer_tma([=]() {
    memcpy(cudasim::smem_base() + dst, gsrc, bytes);
    cudasim::mbar_complete_tx(bar_addr, bytes);
  });
}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { cudasim::mbar_expect_tx(smem_u32(bar), bytes); }
inline uint32_t warp_or(uint32_t v) {
  for (int o = 16; o >= 1; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// -------------------------------------------------------------- tcgen05 ----
inline void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  if (ncols != 512) { fprintf(stderr, "[cudasim] tmem_alloc(%u): the model has 512 columns\n", ncols); abort(); }
  *smem_dst = 0;
}
inline void tmem_relinquish() {}
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tc_fence_before() {}
inline void tc_fence_after() {}

inline float sim_elem(uint32_t addr, bool bf16) {
  uint16_t h;
  memcpy(&h, cudasim::smem_base() + addr, 2);
  if (bf16) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
  _Float16 x; memcpy(&x, &h, 2); return (float)x;
}

// address of element (row m, K index k of this instruction) of a K-major operand: SWIZZLE_128B (layout type 2:
// 128-byte rows, 8-row atoms SBO apart, XOR swizzle) or no swizzle (type 0: core matrices of 8 rows x 16 bytes,
// LBO apart along K, SBO apart along the rows) -- the fields csrc/ts_ptx.cuh documents
inline uint32_t sim_operand_addr(uint64_t desc, int m, int k, int esz) {
  const uint32_t base = (uint32_t)(desc & 0x3FFF) << 4;
  const uint32_t lbo = (uint32_t)((desc >> 16) & 0x3FFF) << 4, sbo = (uint32_t)((desc >> 32) & 0x3FFF) << 4;
  if ((desc >> 61) == 2) return sw128(base + (uint32_t)(m >> 3) * sbo + (uint32_t)(m & 7) * 128u + (uint32_t)k * (uint32_t)esz);
  const int per = 16 / esz;                                  // elements per 16-byte core-matrix row
  return base + (uint32_t)(m >> 3) * sbo + (uint32_t)(m & 7) * 16u + (uint32_t)(k / per) * lbo + (uint32_t)(k % per) * (uint32_t)esz;
}

// descriptor fields as documented in csrc/ts_ptx.cuh
inline void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const int N = (int)((idesc >> 17) & 0x3F) << 3, M = (int)((idesc >> 24) & 0x1F) << 4;
  const bool bf16 = ((idesc >> 7) & 7) == 1;
  if (M != 128 || N < 8 || N > 256 || (N & 15) || ((idesc >> 4) & 3) != 1 || ((idesc >> 10) & 7) != ((idesc >> 7) & 7)) {
    fprintf(stderr, "[cudasim] unsupported instruction descriptor %08x (M %d N %d)\n", idesc, M, N); abort();
  }
  for (uint64_t d : {adesc, bdesc})
    if (((d >> 61) != 2 && (d >> 61) != 0) || ((d >> 46) & 3) != 1) { fprintf(stderr, "[cudasim] shared-memory descriptor is neither K-major SWIZZLE_128B nor un-swizzled\n"); abort(); }
  const uint32_t col0 = tmem_d & 0xFFFFu, lane0 = tmem_d >> 16;
  if (lane0 != 0 || col0 + (uint32_t)N > 512) { fprintf(stderr, "[cudasim] accumulator outside TMEM\n"); abort(); }
  cudasim::defer_mma([=]() {
  uint32_t* T = cudasim::tmem();
  float a[128][16];
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 16; ++k) a[m][k] = sim_elem(sim_operand_addr(adesc, m, k, 2), bf16);
  for (int n = 0; n < N; ++n) {
    float b[16];
    for (int k = 0; k < 16; ++k) b[k] = sim_elem(sim_operand_addr(bdesc, n, k, 2), bf16);
    for (int m = 0; m < 128; ++m) {
      float acc = 0.f;
      for (int k = 0; k < 16; ++k) acc += a[m][k] * b[k];
      uint32_t* cell = T + (size_t)m * 512 + col0 + (uint32_t)n;
      float prev;
      memcpy(&prev, cell, 4);
      const float out = accumulate ? prev + acc : acc;
      memcpy(cell, &out, 4);
    }
  }
  });
}
// kind::tf32: 8 fp32 elements (32 bytes) of K per instruction; the tensor core ignores the 13 low mantissa bits
inline void umma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const int N = (int)((idesc >> 17) & 0x3F) << 3, M = (int)((idesc >> 24) & 0x1F) << 4;
  if (M != 128 || N < 8 || N > 256 || (N & 15) || ((idesc >> 4) & 3) != 1 || ((idesc >> 7) & 7) != 2 || ((idesc >> 10) & 7) != 2) {
    fprintf(stderr, "[cudasim] unsupported tf32 instruction descriptor %08x (M %d N %d)\n", idesc, M, N); abort();
  }
  for (uint64_t d : {adesc, bdesc})
    if ((d >> 61) != 2 || ((d >> 46) & 3) != 1) { fprintf(stderr, "[cudasim] shared-memory descriptor is not K-major SWIZZLE_128B\n"); abort(); }
  const uint32_t a0 = (uint32_t)(adesc & 0x3FFF) << 4, b0 = (uint32_t)(bdesc & 0x3FFF) << 4;
  const uint32_t a_sbo = (uint32_t)((adesc >> 32) & 0x3FFF) << 4, b_sbo = (uint32_t)((bdesc >> 32) & 0x3FFF) << 4;
  const uint32_t col0 = tmem_d & 0xFFFFu, lane0 = tmem_d >> 16;
  if (lane0 != 0 || col0 + (uint32_t)N > 512) { fprintf(stderr, "[cudasim] accumulator outside TMEM\n"); abort(); }
  cudasim::defer_mma([=]() {
  uint32_t* T = cudasim::tmem();
  auto elem = [](uint32_t addr) { uint32_t u; memcpy(&u, cudasim::smem_base() + addr, 4); u &= 0xFFFFE000u; float f; memcpy(&f, &u, 4); return f; };
  float a[128][8];
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 8; ++k) a[m][k] = elem(sw128(a0 + (uint32_t)(m >> 3) * a_sbo + (uint32_t)(m & 7) * 128u + (uint32_t)k * 4u));
  for (int n = 0; n < N; ++n) {
    float b[8];
    for (int k = 0; k < 8; ++k) b[k] = elem(sw128(b0 + (uint32_t)(n >> 3) * b_sbo + (uint32_t)(n & 7) * 128u + (uint32_t)k * 4u));
    for (int m = 0; m < 128; ++m) {
      float acc = 0.f;
      for (int k = 0; k < 8; ++k) acc += a[m][k] * b[k];
      uint32_t* cell = T + (size_t)m * 512 + col0 + (uint32_t)n;
      float prev;
      memcpy(&prev, cell, 4);
      const float out = accumulate ? prev + acc : acc;
      memcpy(cell, &out, 4);
    }
  }
  });
}
inline void umma_commit(uint64_t* bar) {
  const uint32_t addr = smem_u32(bar);
  cudasim::defer_mma([=]() { cudasim::mbar_arrive(addr, 0); });   // after every MMA issued before it
}

inline void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31u), col = taddr & 0xFFFFu;
  if ((taddr >> 16) != 32u * ((threadIdx.x >> 5) & 3u)) { fprintf(stderr, "[cudasim] warp %u may not read TMEM lanes from %u\n", threadIdx.x >> 5, taddr >> 16); abort(); }
  if (lane >= 128 || col + 32 > 512) { fprintf(stderr, "[cudasim] tcgen05.ld outside TMEM (lane %u col %u)\n", lane, col); abort(); }
  memcpy(r, cudasim::tmem() + (size_t)lane * 512 + col, 32 * 4);
}
inline void tmem_ld_wait() {}

// ------------------------------------------------ CTA pair (cta_group::2) ----
// Model (assumptions spelled out; they follow the CUTLASS 2-SM GEMM data flow): every CTA of the
// pair holds ITS 128 rows of A and ITS N/2 rows of B at the descriptor offsets in its own shared
// memory; CTA r's TMEM lane m receives A_r[m] . B[n] for all N columns, where columns [0, N/2)
// come from rank 0's B rows and [N/2, N) from rank 1's.  tcgen05.commit multicasts its arrive to
// the barrier at the same offset in every CTA of the mask; a 2-SM TMA load credits the leader's barrier.
inline uint32_t cluster_ctarank() { return (uint32_t)cudasim::cluster_rank(); }
inline void cluster_sync() { cudasim::cluster_barrier(); }
inline void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) { cudasim::mbar_arrive_at((int)cta, smem_u32(bar), 0); }
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

inline void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* mp, uint64_t* bar, int c0, int c1, uint64_t) {
  const uint32_t dst = smem_u32(smem_dst), bar_addr = smem_u32(bar);
  if (dst & 1023u) { fprintf(stderr, "[cudasim] TMA destination %u is not 1024-byte aligned\n", dst); abort(); }
  const CUtensorMap map = *mp;
  cudasim::defer_tma([=]() {
  const CUtensorMap* m = &map;
  unsigned char* sm = cudasim::smem_base();               // data lands in the ISSUER's shared memory
  const uint16_t* g = static_cast<const uint16_t*>(m->base);
  for (int r = 0; r < m->box_rows; ++r) {
    const int64_t row = (int64_t)c1 + r;
    for (int c = 0; c < m->box_cols; ++c) {
      const int col = c0 + c;
      uint16_t v = 0;
      if (row >= 0 && row < m->rows && col >= 0 && col < m->dim) v = g[row * m->ld + col];
      memcpy(sm + sw128(dst + (uint32_t)r * 128u + (uint32_t)c * 2u), &v, 2);
    }
  }
  cudasim::mbar_complete_tx_at(0, bar_addr, (uint32_t)m->box_rows * 128u);   // ... the bytes count on the LEADER's barrier
  });
}
inline void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) { tmem_alloc(smem_dst, ncols); }
inline void tmem_relinquish_2sm() {}
inline void tmem_dealloc_2sm(uint32_t, uint32_t) {}

inline float sim_elem_of(int cta, uint32_t addr, bool bf16) {
  uint16_t h;
  memcpy(&h, cudasim::smem_base_of(cta) + addr, 2);
  if (bf16) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
  _Float16 x; memcpy(&x, &h, 2); return (float)x;
}
inline void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (cudasim::cluster_rank() != 0) { fprintf(stderr, "[cudasim] cta_group::2 MMA issued by the non-leader CTA\n"); abort(); }
  const int N = (int)((idesc >> 17) & 0x3F) << 3, M = (int)((idesc >> 24) & 0x1F) << 4;
  const bool bf16 = ((idesc >> 7) & 7) == 1;
  if (M != 256 || N < 32 || N > 256 || (N & 31)) { fprintf(stderr, "[cudasim] unsupported cta_group::2 descriptor (M %d N %d)\n", M, N); abort(); }
  const uint32_t a0 = (uint32_t)(adesc & 0x3FFF) << 4, b0 = (uint32_t)(bdesc & 0x3FFF) << 4;
  const uint32_t a_sbo = (uint32_t)((adesc >> 32) & 0x3FFF) << 4, b_sbo = (uint32_t)((bdesc >> 32) & 0x3FFF) << 4;
  const uint32_t col0 = tmem_d & 0xFFFFu;
  cudasim::defer_mma([=]() {
  for (int cta = 0; cta < 2; ++cta) {
    uint32_t* T = cudasim::tmem_of(cta);
    float a[128][16];
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < 16; ++k) a[m][k] = sim_elem_of(cta, sw128(a0 + (uint32_t)(m >> 3) * a_sbo + (uint32_t)(m & 7) * 128u + (uint32_t)k * 2u), bf16);
    for (int n = 0; n < N; ++n) {
      const int src = n / (N / 2), nn = n % (N / 2);       // which CTA's B rows feed this column
      float b[16];
      for (int k = 0; k < 16; ++k) b[k] = sim_elem_of(src, sw128(b0 + (uint32_t)(nn >> 3) * b_sbo + (uint32_t)(nn & 7) * 128u + (uint32_t)k * 2u), bf16);
      for (int m = 0; m < 128; ++m) {
        float acc = 0.f;
        for (int k = 0; k < 16; ++k) acc += a[m][k] * b[k];
        uint32_t* cell = T + (size_t)m * 512 + col0 + (uint32_t)n;
        float prev;
        memcpy(&prev, cell, 4);
        const float out = accumulate ? prev + acc : acc;
        memcpy(cell, &out, 4);
      }
    }
  }
  });
}
inline void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  const uint32_t addr = smem_u32(bar);
  cudasim::defer_mma([=]() {
    for (int cta = 0; cta < 2; ++cta)
      if (cta_mask & (1u << cta)) cudasim::mbar_arrive_at(cta, addr, 0);
  });
}

inline uint64_t make_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
inline uint64_t make_desc_kmajor_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
constexpr uint64_t kDescKStep = 2;
constexpr uint32_t make_idesc_f16(int M, int N, bool bf16) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

}  // namespace ptx
}  // namespace ts
