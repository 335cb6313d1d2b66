// TEST INFRASTRUCTURE ONLY -- scheduler of the SIMT emulator (see cudasim.h).
#include "cudasim.h"
#include <unistd.h>

#include <algorithm>
#include <unordered_map>
#include <utility>
#include <vector>

// Context switch.  x86-64: a 12-instruction stack switch (callee-saved registers only);
// elsewhere: ucontext.  A launch of a few thousand simulated threads makes tens of millions of
// switches, and swapcontext() pays a signal-mask system call on each.
#if defined(__x86_64__)
extern "C" void cudasim_swap(void** save_sp, void* new_sp);
asm(R"(
.text
.globl cudasim_swap
.type cudasim_swap,@function
cudasim_swap:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size cudasim_swap,.-cudasim_swap
)");
struct Context { void* sp = nullptr; };
static inline void ctx_switch(Context* from, Context* to) { cudasim_swap(&from->sp, to->sp); }
static inline void ctx_make(Context* c, char* stack, size_t bytes, void (*entry)()) {
  uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + bytes) & ~uintptr_t(15);
  void** sp = reinterpret_cast<void**>(top);
  *--sp = nullptr;                               // return address slot of `entry` (it never returns)
  *--sp = reinterpret_cast<void*>(entry);        // popped by `ret` in cudasim_swap
  for (int i = 0; i < 6; ++i) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
  c->sp = sp;
}
#else
#include <ucontext.h>
struct Context { ucontext_t uc; };
static inline void ctx_switch(Context* from, Context* to) { swapcontext(&from->uc, &to->uc); }
static inline void ctx_make(Context* c, char* stack, size_t bytes, void (*entry)()) {
  getcontext(&c->uc);
  c->uc.uc_stack.ss_sp = stack;
  c->uc.uc_stack.ss_size = bytes;
  c->uc.uc_link = nullptr;
  makecontext(&c->uc, entry, 0);
}
#endif

namespace cudasim {

ThreadCtx* g_cur = nullptr;
unsigned long long g_launches = 0, g_blocks_run = 0, g_switches = 0;

namespace {

constexpr size_t kStackBytes = 96 * 1024;
enum State { kRunnable, kWaitBarrier, kWaitCollective, kDone };

struct Fiber {
  Context ctx;
  ThreadCtx tc;
  State state = kRunnable;
  unsigned wait_gen = 0;
  // pending warp collective
  Collective kind = kSyncWarp;
  unsigned mask = 0;
  uint64_t val = 0, result = 0;
  int aux = 0;
  bool result_ready = false;
  char* stack = nullptr;
};

struct Block {
  std::vector<Fiber> fibers;
  std::vector<unsigned> warp_arrived;   // lanes waiting in a collective, per warp
  Context sched;
  int n = 0, live = 0, bar_arrived = 0, cur = 0;
  unsigned bar_gen = 0;
  unsigned long long progress = 0;
  unsigned char* smem = nullptr;         // 1024-byte aligned (swizzle atoms), kSmemMax bytes
  const std::function<void()>* body = nullptr;
  // tensor-path emulation (ts_ptx_sim.cuh)
  struct MBar { uint32_t init = 0, pending = 0, phase = 0; long long tx = 0; };
  struct NamedBar { int arrived = 0; unsigned gen = 0; };
  std::unordered_map<uint32_t, MBar> mbars;
  NamedBar named[16];
  std::vector<uint32_t> tmem;
  // asynchronous engines (CUDASIM_ASYNC=<seed>): operations issued by TMA / tensor-core wrappers are queued and
  // completed some scheduler rounds later -- TMA operations independently of each other, tensor-core operations
  // (MMAs and the commits that follow them) strictly in issue order
  struct Deferred { unsigned long long due; std::function<void()> fn; };
  std::vector<Deferred> tma_q;
  std::vector<Deferred> mma_q;
  size_t mma_head = 0;
  // thread-block cluster (launch_cluster): rank inside the cluster and the cluster's blocks
  int cluster_rank = 0, cluster_size = 1;
  Block** cluster = nullptr;
  int cl_arrived = 0;            // cluster barrier state lives in rank 0's block
  unsigned cl_gen = 0;
};
constexpr size_t kSmemMax = 232448;

// CUDASIM_ASYNC=<seed != 0>: adversarial timing -- deferred completion of asynchronous operations and a shuffled
// thread order every scheduler round.  Off (0 / unset): everything completes at issue time, round-robin order.
unsigned long long g_async_seed = 0, g_rng = 0, g_round = 0;
inline unsigned long long rnd() {
  g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17;
  return g_rng;
}

std::vector<Block*> g_blocks;     // blocks alive at the same time: 1, or the whole grid of a cooperative launch
Block* g_blk = nullptr;           // block of the running simulated thread
bool g_in_launch = false;
std::vector<char*> g_stacks;

[[noreturn]] void die(const char* what) {
  Block& B = *g_blk;
  fprintf(stderr, "[cudasim] %s (block %u,%u: %d threads, %d live, %d at __syncthreads)\n", what,
          B.fibers.empty() ? 0 : B.fibers[0].tc.blockIdx_.x, B.fibers.empty() ? 0 : B.fibers[0].tc.blockIdx_.y, B.n, B.live,
          B.bar_arrived);
  for (int i = 0; i < B.n && i < 1024; ++i) {
    const Fiber& f = B.fibers[i];
    if (f.state == kWaitCollective)
      fprintf(stderr, "  thread %d waits in a warp collective kind %d mask %08x\n", i, (int)f.kind, f.mask);
  }
  abort();
}

void release_barrier_if_complete() {
  Block& B = *g_blk;
  if (B.bar_arrived > 0 && B.bar_arrived == B.live) { B.bar_arrived = 0; ++B.bar_gen; }
}

void yield() {
  Block& B = *g_blk;
  ++g_switches;
  ctx_switch(&B.fibers[B.cur].ctx, &B.sched);
}

void fiber_main() {
  Block& B = *g_blk;
  (*B.body)();
  Fiber& f = B.fibers[B.cur];
  f.state = kDone;
  --B.live;
  ++B.progress;
  release_barrier_if_complete();   // exited threads no longer count towards __syncthreads
  ctx_switch(&f.ctx, &B.sched);
  abort();
}

unsigned existing_lanes(int warp) {
  const int first = warp * 32, n = g_blk->n - first;
  return n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
}

void complete_collective(int warp, unsigned mask) {
  Block& B = *g_blk;
  Fiber* lanes = &B.fibers[(size_t)warp * 32];
  const int lead = __builtin_ctz(mask);
  unsigned ballot = 0;
  for (int l = 0; l < 32; ++l)
    if (mask & (1u << l)) {
      if (lanes[l].mask != lanes[lead].mask || lanes[l].kind != lanes[lead].kind) die("lanes of one warp met in different collectives");
      if (lanes[l].kind == kBallot && lanes[l].val) ballot |= 1u << l;
    }
  for (int l = 0; l < 32; ++l) {
    if (!(mask & (1u << l))) continue;
    Fiber& f = lanes[l];
    int src = l;
    switch (f.kind) {
      case kShflIdx: src = f.aux; break;
      case kShflXor: src = l ^ f.aux; break;
      case kShflUp: src = l - f.aux; break;
      case kShflDown: src = l + f.aux; break;
      default: break;
    }
    if (src < 0 || src > 31 || !(mask & (1u << src))) src = l;   // out of range / inactive source: own value
    f.result = f.kind == kBallot ? ballot : (f.kind == kSyncWarp ? 0 : lanes[src].val);
  }
  for (int l = 0; l < 32; ++l)
    if (mask & (1u << l)) { lanes[l].result_ready = true; lanes[l].state = kRunnable; }
  B.warp_arrived[warp] &= ~mask;
}

}  // namespace

void* dyn_smem() { return g_blk->smem; }
unsigned char* smem_base() { return g_blk->smem; }
uint32_t* tmem() { return g_blk->tmem.data(); }

void yield_spin() { yield(); }

// asynchronous engines: run now (default) or queue for a later scheduler round (CUDASIM_ASYNC)
void defer_tma(std::function<void()> fn) {
  if (!g_async_seed) { fn(); return; }
  g_blk->tma_q.push_back({g_round + 1 + rnd() % 4, std::move(fn)});
}
void defer_mma(std::function<void()> fn) {
  if (!g_async_seed) { fn(); return; }
  g_blk->mma_q.push_back({g_round + rnd() % 3, std::move(fn)});
}

int cluster_rank() { return g_blk->cluster_rank; }
static Block& block_of(int cta) {
  Block& B = *g_blk;
  if (cta < 0 || cta >= B.cluster_size) die("cluster rank out of range");
  return B.cluster_size == 1 ? B : *B.cluster[cta];
}
unsigned char* smem_base_of(int cta) { return block_of(cta).smem; }
uint32_t* tmem_of(int cta) { return block_of(cta).tmem.data(); }
void cluster_barrier() {
  Block& B = *g_blk;
  Block& R = block_of(0);
  const unsigned gen = R.cl_gen;
  const int want = B.cluster_size * B.n;
  ++B.progress;
  if (++R.cl_arrived == want) { R.cl_arrived = 0; ++R.cl_gen; return; }
  while (R.cl_gen == gen) yield();
}

static Block::MBar& mbar_at(uint32_t addr, int cta = -1) {
  Block& T = cta < 0 ? *g_blk : block_of(cta);
  auto it = T.mbars.find(addr);
  if (it == T.mbars.end()) { fprintf(stderr, "[cudasim] mbarrier at shared offset %u used before mbarrier.init\n", addr); abort(); }
  return it->second;
}
static void mbar_check_complete(Block::MBar& b) {
  if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.init; }
}
void mbar_init(uint32_t addr, uint32_t count) {
  Block::MBar b;
  b.init = b.pending = count;
  g_blk->mbars[addr] = b;
  ++g_blk->progress;
}
void mbar_arrive_at(int cta, uint32_t addr, uint32_t expect_tx_bytes) {
  Block::MBar& b = mbar_at(addr, cta);
  if (b.pending == 0) die("mbarrier received more arrivals than its count");
  b.tx += expect_tx_bytes;
  --b.pending;
  mbar_check_complete(b);
  ++g_blk->progress;
  ++block_of(cta).progress;
}
void mbar_complete_tx_at(int cta, uint32_t addr, uint32_t bytes) {
  Block::MBar& b = mbar_at(addr, cta);
  b.tx -= bytes;                 // may run ahead of the expect_tx of the same phase (partner CTA's loads)
  mbar_check_complete(b);
  ++g_blk->progress;
  ++block_of(cta).progress;
}
void mbar_arrive(uint32_t addr, uint32_t expect_tx_bytes) {
  Block::MBar& b = mbar_at(addr);
  if (b.pending == 0) die("mbarrier received more arrivals than its count");
  b.tx += expect_tx_bytes;
  --b.pending;
  mbar_check_complete(b);
  ++g_blk->progress;
}
void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
  Block::MBar& b = mbar_at(addr);
  b.tx += bytes;
  mbar_check_complete(b);
  ++g_blk->progress;
}
void mbar_complete_tx(uint32_t addr, uint32_t bytes) {
  Block::MBar& b = mbar_at(addr);
  b.tx -= bytes;
  if (b.tx < 0) die("mbarrier transaction count went negative (more bytes delivered than expected)");
  mbar_check_complete(b);
  ++g_blk->progress;
}
bool mbar_phase_done(uint32_t addr, uint32_t parity) { return mbar_at(addr).phase != (parity & 1u); }

void named_barrier(int id, int nthreads) {
  Block& B = *g_blk;
  if (id < 0 || id >= 16) die("named barrier id out of range");
  Block::NamedBar& nb = B.named[id];
  const unsigned gen = nb.gen;
  ++B.progress;
  if (++nb.arrived == nthreads) { nb.arrived = 0; ++nb.gen; return; }
  if (nb.arrived > nthreads) die("named barrier over-subscribed");
  while (nb.gen == gen) yield();
}

void sync_threads() {
  Block& B = *g_blk;
  Fiber& f = B.fibers[B.cur];
  const unsigned gen = B.bar_gen;
  ++B.bar_arrived;
  ++B.progress;
  release_barrier_if_complete();
  if (B.bar_gen != gen) return;
  f.state = kWaitBarrier;
  f.wait_gen = gen;
  while (B.bar_gen == gen) yield();
  f.state = kRunnable;
}

uint64_t warp_collective(Collective kind, unsigned mask, uint64_t value, int aux) {
  Block& B = *g_blk;
  Fiber& f = B.fibers[B.cur];
  const int warp = B.cur >> 5, lane = B.cur & 31;
  mask &= existing_lanes(warp);       // a full mask in a partial last warp names only the lanes that exist
  if (!(mask & (1u << lane))) die("a lane executed a *_sync primitive whose mask does not name it");
  f.kind = kind; f.mask = mask; f.val = value; f.aux = aux; f.result_ready = false;
  B.warp_arrived[warp] |= 1u << lane;
  ++B.progress;
  if ((B.warp_arrived[warp] & mask) == mask) complete_collective(warp, mask);
  if (!f.result_ready) {
    f.state = kWaitCollective;
    while (!f.result_ready) yield();
  }
  f.state = kRunnable;
  return f.result;
}

// run the deferred operations of one block that are due (all of them when `drain`)
static bool run_deferred(Block& B, bool drain) {
  bool did = false;
  Block* saved = g_blk;
  g_blk = &B;
  for (size_t i = 0; i < B.tma_q.size();) {
    if (drain || B.tma_q[i].due <= g_round) {
      auto fn = std::move(B.tma_q[i].fn);
      B.tma_q.erase(B.tma_q.begin() + (long)i);
      fn();
      did = true;
    } else {
      ++i;
    }
  }
  while (B.mma_head < B.mma_q.size() && (drain || B.mma_q[B.mma_head].due <= g_round)) {
    auto fn = std::move(B.mma_q[B.mma_head].fn);
    ++B.mma_head;
    fn();
    did = true;
  }
  if (B.mma_head == B.mma_q.size()) { B.mma_q.clear(); B.mma_head = 0; }
  if (did) ++B.progress;
  g_blk = saved;
  return did;
}

static void setup_block(Block& B, int n, dim3 grid, dim3 block, unsigned bx, unsigned by, unsigned bz, size_t dyn_smem_bytes,
                        const std::function<void()>* body, size_t stack0) {
  B.tma_q.clear(); B.mma_q.clear(); B.mma_head = 0;
  B.n = B.live = n;
  B.bar_arrived = 0; B.bar_gen = 0; B.progress = 0;
  B.body = body;
  B.fibers.assign((size_t)n, Fiber());
  B.warp_arrived.assign((size_t)(n + 31) / 32, 0u);
  if (!B.smem && posix_memalign(reinterpret_cast<void**>(&B.smem), 1024, kSmemMax) != 0) { fprintf(stderr, "[cudasim] smem allocation failed\n"); abort(); }
  (void)dyn_smem_bytes;
  memset(B.smem, 0xCD, kSmemMax);               // never zeroed on a GPU either
  B.mbars.clear();
  for (auto& nb : B.named) nb = Block::NamedBar();
  B.tmem.assign(128 * 512, 0x7fc00000u);        // NaN until an MMA overwrites it
  B.cluster_rank = 0; B.cluster_size = 1; B.cluster = nullptr; B.cl_arrived = 0; B.cl_gen = 0;
  while (g_stacks.size() < stack0 + (size_t)n) g_stacks.push_back(static_cast<char*>(malloc(kStackBytes)));
  for (int t = 0; t < n; ++t) {
    Fiber& f = B.fibers[t];
    f.tc.threadIdx_ = {(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
    f.tc.blockIdx_ = {bx, by, bz};
    f.tc.blockDim_ = block;
    f.tc.gridDim_ = grid;
    f.stack = g_stacks[stack0 + (size_t)t];
    ctx_make(&f.ctx, f.stack, kStackBytes, fiber_main);
  }
}

// run the given blocks until every simulated thread has exited (round-robin over all of them; with
// CUDASIM_ASYNC the visiting order is reshuffled every round and some runnable threads sit a round out)
static void run_blocks(Block** blocks, int nb) {
  unsigned long long last_progress = ~0ull;
  int idle_rounds = 0, external_idle = 0;
  bool skipped_any = false;      // a runnable thread sat the previous round out: that round proves nothing
  std::vector<std::pair<int, int>> order;
  for (;;) {
    ++g_round;
    unsigned long long progress = 0;
    int live = 0;
    size_t pending = 0;
    for (int i = 0; i < nb; ++i) {
      run_deferred(*blocks[i], false);
      progress += blocks[i]->progress; live += blocks[i]->live;
      pending += blocks[i]->tma_q.size() + (blocks[i]->mma_q.size() - blocks[i]->mma_head);
    }
    if (live == 0) { for (int i = 0; i < nb; ++i) run_deferred(*blocks[i], true); break; }
    idle_rounds = (progress == last_progress && pending == 0 && !skipped_any) ? idle_rounds + 1 : 0;
    skipped_any = false;
    g_blk = blocks[0];
    if (idle_rounds > 3) {
      // CUDASIM_EXTERNAL_WAITS=1: the kernel may be spinning on memory another PROCESS writes (the multi-rank tests of
      // the peer-memory exchange map the receive buffers into several emulator processes): wait instead of giving up,
      // but not for ever
      static const bool external = getenv("CUDASIM_EXTERNAL_WAITS") && getenv("CUDASIM_EXTERNAL_WAITS")[0] == '1';
      if (!external || ++external_idle > 400000) die(external ? "timeout: a peer process never wrote what a simulated thread waits for"
                                                              : "deadlock: no simulated thread can make progress");
      usleep(50);
      idle_rounds = 0;
    } else if (progress != last_progress) {
      external_idle = 0;
    }
    last_progress = progress;
    order.clear();
    for (int i = 0; i < nb; ++i)
      for (int t = 0; t < blocks[i]->n; ++t) order.emplace_back(i, t);
    if (g_async_seed)
      for (size_t i = order.size(); i > 1; --i) std::swap(order[i - 1], order[rnd() % i]);
    for (auto [i, t] : order) {
      Block& B = *blocks[i];
      Fiber& f = B.fibers[t];
      if (f.state == kDone) continue;
      if (f.state == kWaitBarrier && B.bar_gen == f.wait_gen) continue;
      if (f.state == kWaitCollective && !f.result_ready) continue;
      if (g_async_seed && (rnd() & 7) == 0) { skipped_any = true; continue; }   // a slow warp: skips this round
      g_blk = &B;
      B.cur = t;
      g_cur = &f.tc;
      ++g_switches;
      ctx_switch(&B.sched, &f.ctx);
    }
  }
}

static void read_async_env() {
  const char* e = getenv("CUDASIM_ASYNC");
  const unsigned long long seed = e ? strtoull(e, nullptr, 10) : 0;
  if (seed != g_async_seed) { g_async_seed = seed; g_rng = seed * 0x9E3779B97F4A7C15ull + 1; }
}

static void launch_impl(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body, bool cooperative) {
  read_async_env();
  if (g_in_launch) { fprintf(stderr, "[cudasim] nested launch\n"); abort(); }
  const int n = (int)(block.x * block.y * block.z);
  if (n <= 0 || n > 1024) { fprintf(stderr, "[cudasim] bad block size %d\n", n); abort(); }
  if (dyn_smem_bytes > kSmemMax) { fprintf(stderr, "[cudasim] dynamic shared memory request exceeds 227 KB\n"); abort(); }
  const size_t n_blocks = (size_t)grid.x * grid.y * grid.z;
  const size_t alive = cooperative ? n_blocks : 1;
  if (cooperative && n_blocks * (size_t)n > 16384) { fprintf(stderr, "[cudasim] cooperative grid too large to emulate (%zu threads): lower HOSTSIM_SM_COUNT\n", n_blocks * n); abort(); }
  while (g_blocks.size() < alive) g_blocks.push_back(new Block());
  g_in_launch = true;
  ++g_launches;
  size_t bi = 0;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx, ++bi) {
        ++g_blocks_run;
        if (cooperative) {
          setup_block(*g_blocks[bi], n, grid, block, bx, by, bz, dyn_smem_bytes, &body, bi * (size_t)n);
        } else {
          setup_block(*g_blocks[0], n, grid, block, bx, by, bz, dyn_smem_bytes, &body, 0);
          Block* one = g_blocks[0];
          run_blocks(&one, 1);
        }
      }
  if (cooperative) run_blocks(g_blocks.data(), (int)n_blocks);
  g_in_launch = false;
  g_blk = nullptr;
  g_cur = nullptr;
}

void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body) {
  launch_impl(grid, block, dyn_smem_bytes, body, false);
}
// clusters of `csize` consecutive blocks (1-D grid) run together: distributed shared memory, remote
// mbarrier arrives and the cluster barrier work inside a cluster; clusters run one after another
void launch_cluster(dim3 grid, dim3 block, size_t dyn_smem_bytes, int csize, const std::function<void()>& body) {
  read_async_env();
  if (g_in_launch) { fprintf(stderr, "[cudasim] nested launch\n"); abort(); }
  const int n = (int)(block.x * block.y * block.z);
  if (grid.y != 1 || grid.z != 1 || csize < 1 || csize > 8 || grid.x % (unsigned)csize) {
    fprintf(stderr, "[cudasim] launch_cluster: grid %u not a multiple of cluster size %d\n", grid.x, csize); abort();
  }
  while ((int)g_blocks.size() < csize) g_blocks.push_back(new Block());
  g_in_launch = true;
  ++g_launches;
  for (unsigned b0 = 0; b0 < grid.x; b0 += (unsigned)csize) {
    for (int r = 0; r < csize; ++r) {
      ++g_blocks_run;
      setup_block(*g_blocks[r], n, grid, block, b0 + (unsigned)r, 0, 0, dyn_smem_bytes, &body, (size_t)r * (size_t)n);
      g_blocks[r]->cluster_rank = r; g_blocks[r]->cluster_size = csize; g_blocks[r]->cluster = g_blocks.data();
    }
    run_blocks(g_blocks.data(), csize);
  }
  g_in_launch = false;
  g_blk = nullptr;
  g_cur = nullptr;
}
// every CTA of the grid is alive at once (cudaLaunchCooperativeKernel): a grid-wide barrier can complete
void launch_cooperative(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body) {
  launch_impl(grid, block, dyn_smem_bytes, body, true);
}

}  // namespace cudasim

extern "C" {
unsigned long long cudasim_launches(void) { return cudasim::g_launches; }
unsigned long long cudasim_blocks(void) { return cudasim::g_blocks_run; }
unsigned long long cudasim_switches(void) { return cudasim::g_switches; }
}
