// smm_api.cu -- C ABI of libsmm_b200.so (include/smm_b200.h): handle management, the iteration
// loop of run! (AlgoAbstract.jl:27-76) as a stream of kernel launches, NCCL plumbing, host copies.
#include <cuda_runtime.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "smm_device.cuh"

namespace smm {
// launchers defined in smm_kernels.cu
size_t pairs_smem_bytes(int N, int n_s);
size_t exch_smem_bytes(int N);
cudaError_t configure_kernels(int N, int n_s);
int eval_max_blocks_per_sm();
int persistent_max_blocks_per_sm(int N, int D, int M, int cta_seg);
int persistent_max_cta_seg();
cudaError_t configure_persistent(int N, int D, int M, int cta_seg);
cudaError_t launch_persistent(const DevProblem &pb, const DevState &st, int iter0, int n_iters, int sched_iter0,
                              int n_s, int part_len, int max_seg, int cta_seg, int grid, bool flow,
                              unsigned long long done_base, cudaStream_t s);
void launch_eval(const DevProblem &pb, const DevState &st, int iter, int n_split, int part_len, cudaStream_t s);
void launch_pairs(const DevProblem &pb, const DevState &st, int iter0, int n_iters, int n_s, cudaStream_t s);
void launch_exchange(const DevProblem &pb, const DevState &st, int iter, int sched_idx, int n_s, cudaStream_t s);
void launch_objective(const DevProblem &pb, const double *params, int B, int noseed, uint32_t rep0, int n_split,
                      int part_len, double *partials, unsigned *arrive, double *value, double *moments, int *status,
                      cudaStream_t s);
size_t panel_smem_bytes(int K, int P, int variant);
cudaError_t configure_panel(int K, int P, int variant);
int panel_max_blocks_per_sm(int K, int P, int variant);
void launch_propose(const DevProblem &pb, const DevState &st, int iter, int zero_len, cudaStream_t s);
void launch_panel_chains(const DevProblem &pb, const DevState &st, int iter, int grid, int variant, cudaStream_t s);
void launch_panel_batch(const DevProblem &pb, const DevState &st, const double *params, int B, int noseed, uint32_t uid0,
                        uint32_t rep0, unsigned long long *acc, unsigned *done, unsigned *unit_ctr, double *value,
                        double *moments, int *status, int grid, int variant, cudaStream_t s);
void launch_debug_normals(uint64_t seed, uint32_t k, uint32_t c2, uint32_t c3, int n_pairs, int zig, double *out,
                          cudaStream_t s);
void launch_rng_throughput(long long n_per_thread, int blocks, double *out, cudaStream_t s);
cudaError_t launch_barrier_bench(const DevProblem &pb, const DevState &st, int variant, int n, int grid, cudaStream_t s);
void launch_sim_throughput(const DevProblem &pb, int n_per_thread, int blocks, int threads, int dyn, double *out,
                           cudaStream_t s);
namespace ll {  // smm_kernels.cu compiled with -DSMM_LL_TU: the persistent kernel with the flag-in-data hand-over (exchange_mode 3)
int persistent_max_blocks_per_sm(int N, int D, int M, int cta_seg);
cudaError_t configure_persistent(int N, int D, int M, int cta_seg);
cudaError_t launch_persistent(const DevProblem &pb, const DevState &st, int iter0, int n_iters, int sched_iter0,
                              int n_s, int part_len, int max_seg, int cta_seg, int grid, bool flow,
                              unsigned long long done_base, cudaStream_t s);
}  // namespace ll
// smm_stats.cu
int stats_smem_cap();
cudaError_t launch_accepted_stats(const DevState &st, int L, int P, int it_lo, int it_hi, const double *probs, int n_probs,
                                  long long *count, double *mean, double *quant, double *scratch, int cap2, cudaStream_t s);
cudaError_t launch_chain_summary(const DevState &st, int L, int N, int iter, long long *n_exchanged, int *most_with,
                                 double *best_val, cudaStream_t s);
}  // namespace smm

using namespace smm;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail(SMM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));             \
  } while (0)

#define NCCL_TRY(expr)                                                                          \
  do {                                                                                          \
    ncclResult_t r__ = (expr);                                                                  \
    if (r__ != ncclSuccess)                                                                     \
      return fail(SMM_E_NCCL, std::string(#expr) + ": " + ncclGetErrorString(r__));             \
  } while (0)

// Device buffers come from the device's default stream-ordered memory pool (cudaMallocAsync on the legacy stream)
// whose release threshold is raised once per device, so that the memory of a destroyed handle stays mapped and
// the next smm_bgp_create / smm_bgp_eval_batch gets it back in microseconds: estimations are run repeatedly
// (cudaMalloc + cudaFree of the ~45 buffers of a handle cost tens to hundreds of milliseconds).  Buffers that are
// exported over CUDA IPC (fused multi-GPU mode) must be plain cudaMalloc allocations: alloc(count, true).
void pool_setup_once(int device) {
  static bool done[64] = {false};
  if (device < 0 || device >= 64 || done[device]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
  done[device] = true;
}

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  bool plain = false;
  bool owned = true;  // false: carved out of the handle's slab (or the exchange arena) by SlabPlan
  cudaError_t alloc(size_t count, bool ipc = false) {
    n = count;
    plain = ipc;
    owned = true;
    if (ipc) return cudaMalloc((void **)&p, sizeof(T) * (count ? count : 1));
    return cudaMallocAsync((void **)&p, sizeof(T) * (count ? count : 1), (cudaStream_t)0);
  }
  void free() {
    if (p && owned) {
      if (plain)
        cudaFree(p);
      else
        cudaFreeAsync(p, (cudaStream_t)0);
    }
    p = nullptr;
  }
};

// ---- one allocation, one upload, one initialisation kernel per handle ---------------------------------------------
// smm_bgp_create used to make ~45 pool allocations, 8 blocking cudaMemcpy and ~30 fill launches (0.4 ms, a third of a
// short run).  Now every buffer of a handle is a 256-byte aligned piece of ONE stream-ordered allocation (space 0) --
// the buffers the peers write into live in the process-wide exchange arena instead (space 1, CUDA IPC) -- the
// problem definition travels in ONE cudaMemcpyAsync, and ONE kernel writes every initial value.
constexpr int kMaxInitSegs = 56;
struct InitSeg {
  void *p;
  unsigned long long n16;  // 16-byte words to fill
  unsigned long long pat;  // 8-byte pattern (narrower element values are replicated)
};
struct InitTable {
  int n;
  InitSeg seg[kMaxInitSegs];
};

__global__ void __launch_bounds__(256) init_kernel(InitTable t) {
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  for (int s = 0; s < t.n; ++s) {
    ulonglong2 *p = (ulonglong2 *)t.seg[s].p;
    const ulonglong2 v = make_ulonglong2(t.seg[s].pat, t.seg[s].pat);
    for (size_t i = gtid; i < t.seg[s].n16; i += gsz) p[i] = v;
  }
}

template <typename T>
unsigned long long fill_pattern(T v) {
  unsigned long long pat = 0;
  unsigned char *b = (unsigned char *)&pat;
  static_assert(sizeof(T) <= 8 && 8 % sizeof(T) == 0, "element size must divide 8");
  for (size_t i = 0; i < 8; i += sizeof(T)) memcpy(b + i, &v, sizeof(T));
  return pat;
}

struct SlabPlan {
  struct Item {
    void **pp;
    size_t bytes, off;
    int space;
    bool fill;
    unsigned long long pat;
    const void *src;  // host data to upload (space 0 only), or null
  };
  std::vector<Item> items;
  size_t total[2] = {0, 0};
  static size_t pad(size_t b) { return (b + 255) / 256 * 256; }
  template <typename T>
  void add(DevBuf<T> &b, size_t n, int space, bool fill, T v, const T *src = nullptr) {
    b.n = n;
    b.owned = false;
    b.plain = false;
    const size_t bytes = pad(sizeof(T) * (n ? n : 1));
    items.push_back({(void **)&b.p, bytes, total[space], space, fill && n > 0, fill_pattern(v), (const void *)src});
    if (src) items.back().bytes = sizeof(T) * n;  // uploaded pieces: exact length (the padding is not copied)
    total[space] += bytes;
  }
  template <typename T>
  void filled(DevBuf<T> &b, size_t n, T v, int space = 0) { add(b, n, space, true, v); }
  template <typename T>
  void upload(DevBuf<T> &b, const T *src, size_t n) { add(b, n, 0, false, T(), src); }
  void assign(void *base0, void *base1) {
    for (Item &it : items) *it.pp = (char *)(it.space ? base1 : base0) + it.off;
  }
};

// ---- process-wide caches: streams/events per device, communicator + exchange arena per (device, world, rank) -------
// A C2 ensemble is at most ~1000 iterations (60 ms), so estimations create handles over and over: stream / event
// creation, ncclCommInitRank (~1 s at 8 ranks) and the CUDA-IPC mapping of the peers' gather buffers (~0.3 s) must
// not be paid per handle.  They are made once per process and handed from handle to handle; smm_shutdown() (or
// process exit) ends them.
struct StreamSet {
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int *h_err = nullptr;  // page-locked word the sticky device error flag is copied into
};

constexpr size_t kArenaHead = 256;  // bytes at the head of an arena that no handle re-initialises: the barrier slots
struct Arena {  // cudaMalloc'ed memory of this rank that every peer has mapped (and vice versa)
  void *base = nullptr;
  size_t bytes = 0;
  void *peer[kMaxWorld] = {nullptr};  // [rank] = base
  unsigned long long epoch = 0;       // barriers done on this arena (the same number on every rank)
};

// Cross-rank barrier on the device, over the mapped arenas: thread r stores this barrier's epoch into slot [rank] of
// rank r's arena head and waits until rank r's epoch has arrived in slot [r] of our own.  Stream ordered: everything
// enqueued before it on this rank (the handle's initialisation) has completed before any peer passes the barrier.
struct PeerBarrierArgs {
  unsigned long long *slots[kMaxWorld];  // every rank's arena head
  int world, rank;
  unsigned long long epoch;
  int *err;
};
__global__ void peer_barrier_kernel(PeerBarrierArgs a) {
  const int r = threadIdx.x;
  if (r >= a.world) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.slots[r] + a.rank), "l"(a.epoch) : "memory");
  const unsigned long long *mine = a.slots[a.rank] + r;
  unsigned long long t0 = 0, v = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (unsigned spins = 0;; ++spins) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
    if (v >= a.epoch) break;
    if ((spins & 1023u) == 1023u) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) {  // 20 s: a rank that never created its handle becomes an error, not a hang
        atomicOr(a.err, kErrTimeout);
        break;
      }
    }
  }
}

struct RankCtx {
  int device = 0, world = 1, rank = 0;
  ncclComm_t comm = nullptr;
  Arena arena;
  bool arena_busy = false;  // one live handle at a time uses the cached arena; a second one maps its own
};

std::mutex g_cache_mu;
std::vector<StreamSet> g_stream_cache[64];
std::vector<RankCtx *> g_rank_ctx;

cudaError_t stream_set_acquire(int device, StreamSet &out) {
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (device >= 0 && device < 64 && !g_stream_cache[device].empty()) {
      out = g_stream_cache[device].back();
      g_stream_cache[device].pop_back();
      return cudaSuccess;
    }
  }
  out = StreamSet();
  cudaError_t e = cudaStreamCreateWithFlags(&out.stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&out.copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&out.ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&out.ev1);
  if (e == cudaSuccess) e = cudaHostAlloc((void **)&out.h_err, sizeof(int), cudaHostAllocDefault);
  return e;
}

void stream_set_destroy(StreamSet &ss) {
  if (ss.ev0) cudaEventDestroy(ss.ev0);
  if (ss.ev1) cudaEventDestroy(ss.ev1);
  if (ss.copy_stream) cudaStreamDestroy(ss.copy_stream);
  if (ss.stream) cudaStreamDestroy(ss.stream);
  if (ss.h_err) cudaFreeHost(ss.h_err);
  ss = StreamSet();
}

void stream_set_release(int device, StreamSet &ss) {  // the streams are idle (the caller synchronised them)
  if (!ss.stream) return;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  if (device >= 0 && device < 64 && g_stream_cache[device].size() < 4) {
    g_stream_cache[device].push_back(ss);
    ss = StreamSet();
  } else {
    stream_set_destroy(ss);
  }
}

void arena_destroy(Arena &a, int rank) {
  for (int r = 0; r < kMaxWorld; ++r) {
    if (a.peer[r] && r != rank) cudaIpcCloseMemHandle(a.peer[r]);
    a.peer[r] = nullptr;
  }
  if (a.base) cudaFree(a.base);
  a.base = nullptr;
  a.bytes = 0;
}

// collective over the communicator: allocate `bytes` on every rank and map everybody's allocation everywhere
int arena_create(Arena &a, ncclComm_t comm, int world, int rank, size_t bytes, cudaStream_t s) {
  CUDA_TRY(cudaMalloc(&a.base, bytes));
  CUDA_TRY(cudaMemset(a.base, 0, kArenaHead));  // barrier slots: zero before any peer can reach them
  CUDA_TRY(cudaDeviceSynchronize());
  a.bytes = bytes;
  a.epoch = 0;
  a.peer[rank] = a.base;
  cudaIpcMemHandle_t mine;
  CUDA_TRY(cudaIpcGetMemHandle(&mine, a.base));
  char *d_all = nullptr;  // the handles travel over the communicator
  CUDA_TRY(cudaMalloc((void **)&d_all, sizeof(mine) * (size_t)world));
  cudaError_t ce = cudaMemcpyAsync(d_all + sizeof(mine) * (size_t)rank, &mine, sizeof mine, cudaMemcpyHostToDevice, s);
  ncclResult_t nr = ncclSuccess;
  if (ce == cudaSuccess)
    nr = ncclAllGather(d_all + sizeof(mine) * (size_t)rank, d_all, sizeof(mine), ncclChar, comm, s);
  std::vector<cudaIpcMemHandle_t> all(world);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  if (ce == cudaSuccess) ce = cudaMemcpy(all.data(), d_all, sizeof(mine) * (size_t)world, cudaMemcpyDeviceToHost);
  cudaFree(d_all);
  NCCL_TRY(nr);
  CUDA_TRY(ce);
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    CUDA_TRY(cudaIpcOpenMemHandle(&a.peer[r], all[r], cudaIpcMemLazyEnablePeerAccess));
  }
  return 0;
}

// the communicator of (device, world, rank): made from `id` on first use, reused afterwards (the id of later handles is
// not looked at -- every rank of a job makes the same sequence of handles, so the cache hits on all ranks or on none)
int rank_ctx_get(int device, int world, int rank, const uint8_t *id_bytes, RankCtx **out) {
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (RankCtx *c : g_rank_ctx)
      if (c->device == device && c->world == world && c->rank == rank) {
        *out = c;
        return 0;
      }
  }
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof id);
  ncclComm_t comm = nullptr;
  NCCL_TRY(ncclCommInitRank(&comm, world, id, rank));
  RankCtx *c = new RankCtx();
  c->device = device;
  c->world = world;
  c->rank = rank;
  c->comm = comm;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  g_rank_ctx.push_back(c);
  *out = c;
  return 0;
}

}  // namespace

struct smm_bgp {
  int device = 0;
  int P = 0, M = 0, N = 0, L = 0, R = 0, max_iter = 0, world = 1, rank = 0;
  int n_s = 0;       // pairs per iteration
  int n_split = 1, part_len = 0;
  double eval_param_limit = 0.0;  // |param| bound for which the fixed-point accumulators are sized
  bool panel = false;     // SMM_OBJ_PANEL: propose kernel + panel simulation kernel instead of bgp_eval_kernel
  int panel_grid = 0;     // CTAs of the panel simulation kernel (one resident wave)
  int panel_variant = 2;  // K = 8 instantiation: 2..5 thread per individual (2 = default), 6..8 eight lanes per individual
                          // (smm_panel.cuh; SMM_PANEL_VARIANT picks another one for experiments)
  std::vector<double> h_lb, h_ub;
  uint64_t seed_algo = 0, seed_sim = 0;
  int mode = 0;           // 0 = multi-launch (+ NCCL), 1 = persistent kernel (+ fused peer-store all-gather),
                          // 2 = persistent kernel without grid barriers (one completion counter per rank),
                          // 3 = the same with flag-in-data words for what the next iteration's critical path needs
  int grid = 0, max_seg = 1, cta_seg = 1;  // persistent kernel: CTAs, partial slots per chain, chains per CTA share
  int iter = 0;      // iterations completed (algo.i)
  unsigned long long done_base = 0;  // exchange_mode 2: value of the completion counter once everything enqueued has run
  int sched_iter0 = -1, sched_n = 0;
  StreamSet ss;      // from the per-device cache
  cudaStream_t stream = nullptr, copy_stream = nullptr;  // = ss.stream, ss.copy_stream
  std::vector<cudaEvent_t> win_ev;  // window boundaries of smm_bgp_run
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  RankCtx *ctx = nullptr;    // world > 1: cached communicator (+ exchange arena)
  ncclComm_t comm = nullptr;  // = ctx->comm
  Arena own_arena;            // only when the cached arena was busy
  bool uses_ctx_arena = false;
  void *slab = nullptr;       // every private device buffer of the handle
  DevProblem pb{};
  DevState st{};
  smm_counters ctr{};
  bool profiling = false;
  std::vector<cudaEvent_t> prof_ev;          // pairs of events, reused
  std::vector<int> prof_kind;                // kind of each recorded pair in the current step
  double prof_ms[4] = {0, 0, 0, 0};
  int64_t prof_iters = 0;                    // iterations covered by the kind-0 launches
  int64_t prof_n[4] = {0, 0, 0, 0};
  // device memory (pieces of `slab` / of the exchange arena)
  DevBuf<double> lb, ub, init, data, w, acc_tuner, min_improve;
  DevBuf<double> sigma, accept_rate, la_cur, la_pub, la_all, val_all, pp;
  DevBuf<GridBarrier> bar;
  DevBuf<unsigned long long> sync_seq, flags, ll;
  DevBuf<int> n_noex, n_acc;
  DevBuf<double> t_value, t_prob, t_curr, t_best, t_params, t_mom;
  DevBuf<uint8_t> t_acc;
  DevBuf<int> t_status, t_exch, t_bestid;
  DevBuf<double> partials;
  DevBuf<unsigned> arrive, unit_ctr, applied;
  DevBuf<int> sched_ij, sched_off, sched_nlev, err;
  DevBuf<unsigned long long> counters, phase_ts;

  void release() {
    // the slab goes back to the pool in stream order: drain this handle's own streams first
    if (stream) cudaStreamSynchronize(stream);
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    if (uses_ctx_arena && ctx) {
      std::lock_guard<std::mutex> lk(g_cache_mu);
      ctx->arena_busy = false;
    }
    uses_ctx_arena = false;
    arena_destroy(own_arena, rank);
    comm = nullptr;  // owned by the cache
    if (slab) {
      if (stream)
        cudaFreeAsync(slab, stream);
      else
        cudaFree(slab);
    }
    slab = nullptr;
    for (cudaEvent_t e : prof_ev) cudaEventDestroy(e);
    prof_ev.clear();
    for (cudaEvent_t e : win_ev) cudaEventDestroy(e);
    win_ev.clear();
    stream_set_release(device, ss);
    stream = copy_stream = nullptr;
    ev0 = ev1 = nullptr;
  }
};

namespace {

int check_config(const smm_bgp_config *cfg) {
  if (!cfg) return fail(SMM_E_ARG, "null config");
  if (cfg->abi_version != SMM_ABI_VERSION) return fail(SMM_E_ARG, "abi_version mismatch");
  if (cfg->n_params < 1 || cfg->n_params > SMM_MAX_PARAMS) return fail(SMM_E_ARG, "n_params out of range");
  if (cfg->n_moments < 1 || cfg->n_moments > SMM_MAX_MOMENTS) return fail(SMM_E_ARG, "n_moments out of range");
  if (cfg->n_chains < 1 || cfg->max_iter < 1) return fail(SMM_E_ARG, "n_chains / max_iter must be positive");
  if (cfg->max_iter > (int)SMM_ITER_MASK) return fail(SMM_E_ARG, "max_iter exceeds the 28-bit iteration index");
  if (!cfg->lb || !cfg->ub || !cfg->init || !cfg->data_mom || !cfg->data_w || !cfg->sigma0 || !cfg->acc_tuner ||
      !cfg->min_improve)
    return fail(SMM_E_ARG, "null array in config");
  if (cfg->batch_size < 1 || cfg->n_params % cfg->batch_size != 0)
    return fail(SMM_E_UNSUPPORTED_SHAPE,
                "batch_size must divide n_params (upstream's batches are ill-formed otherwise, AlgoBGP.jl:95-103)");
  const int P = cfg->n_params, M = cfg->n_moments;
  switch (cfg->objective_id) {
    case SMM_OBJ_NORM:
    case SMM_OBJ_NORM_SLOW:
      if (P != M)
        return fail(SMM_E_UNSUPPORTED_SHAPE, "objfunc_norm needs n_params == n_moments (ObjExamples.jl:77-78)");
      break;
    case SMM_OBJ_NORM_MV:
      if (M != 2 * P) return fail(SMM_E_UNSUPPORTED_SHAPE, "norm_mv needs n_moments == 2*n_params");
      break;
    case SMM_OBJ_PANEL: {
      const int K = cfg->panel_K;
      if (K < 1 || K > kPanelMaxK || P != 2 * K + 4 || M != 4 * K + 8 || cfg->panel_T < 7 || cfg->panel_N < 1 ||
          cfg->panel_N > (1 << 22))
        return fail(SMM_E_UNSUPPORTED_SHAPE, "panel needs 1 <= K <= 16, P == 2K+4, M == 4K+8, T >= 7, 1 <= N_ind <= 2^22");
      // theta = (rho, beta[K], phi[K], sigma_alpha, sigma_eps, mu0): the simulator is stationary only for |rho|, |phi| < 1
      for (int k = 0; k <= 2 * K; ++k) {
        if (k >= 1 && k <= K) continue;
        if (!(std::fabs(cfg->lb[k]) < 1.0 && std::fabs(cfg->ub[k]) < 1.0))
          return fail(SMM_E_UNSUPPORTED_SHAPE, "panel needs |rho| < 1 and |phi_k| < 1 over the whole sampling box");
      }
      if (cfg->exchange_mode != 0)
        return fail(SMM_E_UNSUPPORTED_SHAPE, "the panel objective runs with exchange_mode 0 (its own simulation kernel)");
      break;
    }
    case SMM_OBJ_FAILS:
      break;
    default:
      return fail(SMM_E_ARG, "unknown objective_id");
  }
  if (cfg->n_sim < 2) return fail(SMM_E_ARG, "n_sim must be >= 2");
  if (cfg->sigma_update_steps < 1) return fail(SMM_E_ARG, "sigma_update_steps must be >= 1");
  if (cfg->smpl_iters < 1) return fail(SMM_E_ARG, "smpl_iters must be >= 1");
  for (int k = 0; k < P; ++k)
    if (!(cfg->ub[k] > cfg->lb[k])) return fail(SMM_E_ARG, "need ub > lb (mprob.jl:82)");
  if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size)
    return fail(SMM_E_ARG, "bad rank / world_size");
  if (cfg->n_chains % cfg->world_size != 0)
    return fail(SMM_E_UNSUPPORTED_SHAPE, "n_chains must be a multiple of world_size");
  if (cfg->n_chains > 8192) return fail(SMM_E_UNSUPPORTED_SHAPE, "n_chains > 8192 not supported");
  return 0;
}

int choose_split(int L, int n_sm, int blocks_per_sm, int n_blocks_philox, int requested) {
  if (requested > 0) return requested > kMaxSplit ? kMaxSplit : requested;
  // one wave: L * n_split CTAs <= resident capacity, as even as possible over the SMs
  const int cap = n_sm * (blocks_per_sm > 0 ? blocks_per_sm : 1);
  int best = 1;
  double best_score = -1.0;
  for (int s = 1; s <= kMaxSplit; ++s) {
    if ((long long)L * s > cap && s > 1) break;
    if (n_blocks_philox / s < 8) break;
    const double per_sm = (double)L * s / n_sm;
    const double balance = per_sm / std::ceil(per_sm);  // tail efficiency of the busiest SM
    const double score = balance + 1e-3 * s;            // prefer finer splits on ties
    if (score > best_score) {
      best_score = score;
      best = s;
    }
  }
  return best;
}

int device_error_to_rc(int flags) {
  if (flags & kErrNegative)
    return fail(SMM_E_NEGATIVE_OBJECTIVE,
                "AlgoBGP assumes that your objective function returns a non-negative number (AlgoBGP.jl:341)");
  if (flags & kErrExhausted)
    return fail(SMM_E_SAMPLER_EXHAUSTED, "no draw in support after smpl_iters trials (AlgoBGP.jl:409)");
  if (flags & kErrTimeout)
    return fail(SMM_E_CUDA, "persistent kernel: barrier / peer-flag wait timed out (a rank did not reach the exchange)");
  return 0;
}

}  // namespace

extern "C" {

int smm_abi_version(void) { return SMM_ABI_VERSION; }

const char *smm_last_error(void) { return g_err.c_str(); }

int smm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int smm_nccl_unique_id(uint8_t out[SMM_NCCL_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) <= SMM_NCCL_ID_BYTES, "ncclUniqueId does not fit");
  if (!out) return fail(SMM_E_ARG, "null output");
  ncclUniqueId id;
  NCCL_TRY(ncclGetUniqueId(&id));
  memset(out, 0, SMM_NCCL_ID_BYTES);
  memcpy(out, &id, sizeof id);
  return 0;
}

void smm_bgp_destroy(smm_bgp *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  h->release();
  delete h;
}

int smm_bgp_create(const smm_bgp_config *cfg, smm_bgp **out) {
  if (!out) return fail(SMM_E_ARG, "null output handle");
  const bool timing = getenv("SMM_TIMING") != nullptr;
  const auto tc0 = std::chrono::steady_clock::now();
  auto stamp = [&](const char *what) {
    if (timing)
      fprintf(stderr, "[smm_bgp_create] %-28s %8.3f ms\n", what,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count());
  };
  *out = nullptr;
  if (int rc = check_config(cfg)) return rc;
  if (smm_device_count() <= cfg->device || cfg->device < 0)
    return fail(SMM_E_CUDA, "no such CUDA device (this library has no CPU fallback)");
  CUDA_TRY(cudaSetDevice(cfg->device));
  pool_setup_once(cfg->device);
  smm_bgp *h = new smm_bgp();
  struct Guard {
    smm_bgp *h;
    bool ok = false;
    ~Guard() {
      if (!ok) {
        h->release();
        delete h;
      }
    }
  } guard{h};
  h->device = cfg->device;
  h->P = cfg->n_params;
  h->M = cfg->n_moments;
  h->N = cfg->n_chains;
  h->world = cfg->world_size;
  h->rank = cfg->rank;
  h->L = h->N / h->world;
  h->R = rec_len(h->P, h->M);
  h->max_iter = cfg->max_iter;
  h->n_s = h->N < 3 ? h->N - 1 : h->N;
  const int P = h->P, M = h->M, N = h->N, L = h->L, R = h->R, I = h->max_iter;

  CUDA_TRY(stream_set_acquire(cfg->device, h->ss));
  h->stream = h->ss.stream;
  h->copy_stream = h->ss.copy_stream;
  h->ev0 = h->ss.ev0;
  h->ev1 = h->ss.ev1;
  *h->ss.h_err = 0;
  stamp("stream + events");

  if (cfg->exchange_mode < 0 || cfg->exchange_mode > 3) return fail(SMM_E_ARG, "exchange_mode must be 0, 1, 2 or 3");
  h->mode = cfg->exchange_mode;
  if (h->world > kMaxWorld) return fail(SMM_E_ARG, "world_size > 8");
  h->seed_algo = cfg->seed_algo;
  h->seed_sim = cfg->seed_sim;

  // ---- kernel configuration (host-side arithmetic and function attributes; no device work) ----
  struct {
    int multiProcessorCount = 0;
  } prop;  // (cudaGetDeviceProperties costs milliseconds; one attribute is all that is needed)
  CUDA_TRY(cudaDeviceGetAttribute(&prop.multiProcessorCount, cudaDevAttrMultiProcessorCount, cfg->device));
  h->part_len = 2 * P;
  const int n_blocks_philox = zig_blocks(cfg->n_sim);  // Philox blocks per simulated row (three draws each)
  h->n_split = choose_split(L, prop.multiProcessorCount, eval_max_blocks_per_sm(), n_blocks_philox, cfg->n_split);
  if (cfg->objective_id == SMM_OBJ_FAILS) h->n_split = 1;
  h->panel = cfg->objective_id == SMM_OBJ_PANEL;
  h->h_lb.assign(cfg->lb, cfg->lb + P);
  h->h_ub.assign(cfg->ub, cfg->ub + P);
  if (h->panel) {
    h->n_split = 1;
    h->part_len = 2 * panel_na(cfg->panel_K);  // (hi, lo) fixed-point words per raw sum
    if (const char *v = getenv("SMM_PANEL_VARIANT")) h->panel_variant = (atoi(v) >= 2 && atoi(v) <= 8) ? atoi(v) : 2;
    CUDA_TRY(configure_panel(cfg->panel_K, P, h->panel_variant));
    const int per_sm = panel_max_blocks_per_sm(cfg->panel_K, P, h->panel_variant);
    if (per_sm < 1) return fail(SMM_E_CUDA, "panel simulation kernel does not fit on an SM");
    h->panel_grid = prop.multiProcessorCount * per_sm;
    if (cfg->n_split > 0 && cfg->n_split < h->panel_grid) h->panel_grid = cfg->n_split;  // n_split caps the CTA count
  }
  if (N > 1) CUDA_TRY(configure_kernels(N, h->n_s));
  h->max_seg = h->n_split;
  if (h->mode >= 1) {
    int coop = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg->device));
    if (!coop) return fail(SMM_E_CUDA, "device does not support cooperative launches (exchange_mode 1)");
    if (P > 32)
      return fail(SMM_E_UNSUPPORTED_SHAPE, "exchange_mode 1/2 (persistent kernel) needs n_params <= 32; use exchange_mode 0");
    h->grid = prop.multiProcessorCount;  // one 768-thread CTA per SM
    if (cfg->n_split > 0 && cfg->n_split < h->grid) h->grid = cfg->n_split;  // n_split caps the CTA count in this mode
    const long long Tj = (long long)L * n_blocks_philox;
    const long long gw = Tj < h->grid ? Tj : h->grid;
    h->cta_seg = (int)((Tj / gw + 1 + n_blocks_philox - 1) / n_blocks_philox + 1);  // chains one CTA's share can touch
    if (h->cta_seg > persistent_max_cta_seg())
      return fail(SMM_E_UNSUPPORTED_SHAPE, "exchange_mode 1: too many chains per SM for the persistent kernel; use exchange_mode 0");
    CUDA_TRY(h->mode == 3 ? ll::configure_persistent(N, P, M, h->cta_seg) : configure_persistent(N, P, M, h->cta_seg));
    if ((h->mode == 3 ? ll::persistent_max_blocks_per_sm(N, P, M, h->cta_seg) : persistent_max_blocks_per_sm(N, P, M, h->cta_seg)) < 1)
      return fail(SMM_E_UNSUPPORTED_SHAPE,
                  "persistent kernel does not fit on an SM (its shared memory grows with the chain count); use exchange_mode 0");
    const int per_chain = (int)((gw + L - 1) / L + 1);
    if (per_chain > h->max_seg) h->max_seg = per_chain;
  }
  stamp("kernel configuration");

  // ---- memory plan: every buffer of the handle is a piece of one allocation ----
  const double nan = std::numeric_limits<double>::quiet_NaN(), inf = std::numeric_limits<double>::infinity();
  const bool fused_peers = h->world > 1 && h->mode >= 1;  // the peers store into la_all / val_all / flags
  const int xs = fused_peers ? 1 : 0;                     // ... which then live in the exchange arena
  SlabPlan plan;
  if (fused_peers) plan.total[1] = kArenaHead;  // the arena's head holds the barrier slots
  plan.upload(h->lb, cfg->lb, P);
  plan.upload(h->ub, cfg->ub, P);
  plan.upload(h->init, cfg->init, P);
  plan.upload(h->data, cfg->data_mom, M);
  plan.upload(h->w, cfg->data_w, M);
  plan.upload(h->acc_tuner, cfg->acc_tuner, N);
  plan.upload(h->min_improve, cfg->min_improve, N);
  std::vector<double> sigma_local(L);  // local chain c is global chain c * world + rank (round robin, smm_device.cuh)
  for (int c = 0; c < L; ++c) sigma_local[c] = cfg->sigma0[(size_t)c * h->world + h->rank];
  plan.upload(h->sigma, sigma_local.data(), L);
  const size_t upload_end = plan.total[0];
  plan.filled(h->accept_rate, (size_t)L, 0.0);
  plan.filled(h->n_noex, (size_t)L, 0);
  plan.filled(h->n_acc, (size_t)L, 0);
  plan.filled(h->la_cur, (size_t)L * R, nan);
  plan.filled(h->la_pub, (size_t)L * R, nan);
  const bool have_la_all = h->world > 1 || h->mode >= 2;
  if (have_la_all) plan.filled(h->la_all, (size_t)(h->mode ? 2 : 1) * N * R, nan, xs);
  // [2][N] values by iteration parity, followed by [N] 64-bit words whose first is the rank's completion counter
  // (exchange_mode 2), zero = nothing done
  plan.add(h->val_all, (size_t)3 * N, xs, false, 0.0);
  plan.filled(h->flags, (size_t)kMaxWorld, 0ull, xs);
  if (h->mode == 3) plan.filled(h->ll, (size_t)2 * N * ll_words(P), 0ull, xs);  // tag 0 = nothing arrived
  plan.filled(h->applied, (size_t)L, 0u);
  plan.filled(h->pp, (size_t)L * P, nan);
  plan.filled(h->bar, 1, GridBarrier{0u, 0u});
  plan.filled(h->sync_seq, 1, 0ull);
  // trace: unrun slots look like a fresh BGPChain (AlgoBGP.jl:81-89); Eval slots are `undef` -> NaN
  const size_t IL = (size_t)I * L;
  plan.filled(h->t_value, IL, nan);
  plan.filled(h->t_prob, IL, nan);
  plan.filled(h->t_curr, IL, inf);
  plan.filled(h->t_best, IL, inf);
  plan.filled(h->t_params, IL * P, nan);
  plan.filled(h->t_mom, IL * M, nan);
  plan.filled(h->t_acc, IL, (uint8_t)0);
  plan.filled(h->t_status, IL, 0);
  plan.filled(h->t_exch, IL, 0);
  plan.filled(h->t_bestid, IL, -1);
  plan.filled(h->unit_ctr, 1, 0u);
  plan.filled(h->partials, (size_t)L * h->max_seg * h->part_len, 0.0);
  plan.filled(h->arrive, (size_t)L, 0u);
  if (h->N > 1) {
    plan.filled(h->sched_ij, (size_t)kPairChunk * h->n_s * 2, 0);
    plan.filled(h->sched_off, (size_t)kPairChunk * (h->n_s + 1), 0);
    plan.filled(h->sched_nlev, (size_t)kPairChunk, 0);
  }
  plan.filled(h->err, 1, 0);
  plan.filled(h->counters, 4, 0ull);
  const bool want_phase_ts = getenv("SMM_PHASE_TS") != nullptr;
  if (want_phase_ts) {
    size_t slots = (size_t)L * h->n_split;
    if ((size_t)h->grid * 6 + L > slots) slots = (size_t)h->grid * 6 + L;  // + two publish rows per CTA
    plan.filled(h->phase_ts, slots * 4, 0ull);
  }

  // ---- communicator and exchange arena (world > 1): cached per process ----
  void *arena_base = nullptr;
  Arena *arena = nullptr;
  if (h->world > 1) {
    if (int rc = rank_ctx_get(cfg->device, h->world, h->rank, cfg->nccl_id, &h->ctx)) return rc;
    h->comm = h->ctx->comm;
    if (fused_peers) {
      const size_t need = plan.total[1];
      bool busy;
      {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        busy = h->ctx->arena_busy;
        if (!busy) h->ctx->arena_busy = true;
      }
      if (!busy) {
        h->uses_ctx_arena = true;
        arena = &h->ctx->arena;
        if (arena->bytes < need) {  // (re)made collectively: every rank sees the same sizes in the same order
          arena_destroy(*arena, h->rank);
          size_t min_cap = (size_t)4 << 20;  // holds 8192 chains of the MvNormal shapes; SMM_ARENA_MIN_BYTES: tests
          if (const char *v = getenv("SMM_ARENA_MIN_BYTES")) min_cap = (size_t)atoll(v) > 4096 ? (size_t)atoll(v) : 4096;
          const size_t cap = need * 2 > min_cap ? need * 2 : min_cap;
          if (int rc = arena_create(*arena, h->comm, h->world, h->rank, cap, h->stream)) return rc;
        }
      } else {  // a second live fused handle in this process: its own mapping (slow path, seconds at 8 ranks)
        arena = &h->own_arena;
        if (int rc = arena_create(*arena, h->comm, h->world, h->rank, need, h->stream)) return rc;
      }
      arena_base = arena->base;
    }
  }
  stamp("communicator + arena");

  CUDA_TRY(cudaMallocAsync(&h->slab, plan.total[0], h->stream));
  plan.assign(h->slab, arena_base);
  {
    // one upload: the problem definition sits at the head of the slab (pageable source: the call returns once the
    // bytes are staged, so the vector may die right after it)
    std::vector<char> stage(upload_end, 0);
    for (const SlabPlan::Item &it : plan.items)
      if (it.src) memcpy(stage.data() + it.off, it.src, it.bytes);
    CUDA_TRY(cudaMemcpyAsync(h->slab, stage.data(), upload_end, cudaMemcpyHostToDevice, h->stream));
  }
  {
    InitTable tab;
    tab.n = 0;
    auto seg = [&](void *p, size_t bytes, unsigned long long pat) {
      tab.seg[tab.n++] = InitSeg{p, (unsigned long long)((bytes + 15) / 16), pat};
    };
    for (const SlabPlan::Item &it : plan.items)
      if (it.fill) seg(*it.pp, it.bytes, it.pat);
    seg(h->val_all.p, sizeof(double) * 2 * (size_t)N, fill_pattern(nan));
    seg(h->val_all.p + 2 * (size_t)N, sizeof(double) * (size_t)N, 0ull);
    if (tab.n > kMaxInitSegs) return fail(SMM_E_STATE, "init table overflow");
    size_t words = 0;
    for (int i = 0; i < tab.n; ++i) words += tab.seg[i].n16;
    int blocks = (int)((words + 1023) / 1024);
    blocks = blocks < 1 ? 1 : (blocks > 8 * prop.multiProcessorCount ? 8 * prop.multiProcessorCount : blocks);
    init_kernel<<<blocks, 256, 0, h->stream>>>(tab);
    CUDA_TRY(cudaGetLastError());
    h->ctr.kernel_launches++;
  }
  stamp("slab + upload + init kernel");

  DevProblem &pb = h->pb;
  pb.P = P; pb.M = M; pb.S = cfg->n_sim; pb.obj = cfg->objective_id; pb.noseed = cfg->noseed;
  pb.N = N; pb.L = L; pb.max_iter = I; pb.world = h->world; pb.rank = h->rank;
  {
    uint32_t k0 = (uint32_t)cfg->seed_sim, k1 = (uint32_t)(cfg->seed_sim >> 32);
    for (int r = 0; r < 10; ++r) {
      pb.rk_sim0[r] = k0;
      pb.rk_sim1[r] = k1;
      k0 += SMM_PHILOX_W0;
      k1 += SMM_PHILOX_W1;
    }
  }
  pb.sigma_update_steps = cfg->sigma_update_steps; pb.smpl_iters = cfg->smpl_iters; pb.batch_size = cfg->batch_size;
  pb.panel_T = cfg->panel_T; pb.panel_N = cfg->panel_N; pb.panel_K = cfg->panel_K;
  pb.sigma_adjust_by = cfg->sigma_adjust_by; pb.slow_seconds = cfg->slow_seconds;
  pb.seed_sim = cfg->seed_sim; pb.seed_algo = cfg->seed_algo;
  {
    // fixed-point grids of the order-invariant accumulators (smm_kernels.cu): |x| <= xmax = box + 14 sigma-units
    // (a ziggurat draw is at most R + 52 ln2 / R < 13.3 in magnitude),
    // totals S*xmax and S*xmax^2 must stay below 2^62, single terms -- the sums of a block's three values -- below
    // 2^(51-F) (4x headroom for eval_batch)
    double box = 0.0;
    for (int k = 0; k < P; ++k) {
      box = std::fmax(box, std::fabs(cfg->lb[k]));
      box = std::fmax(box, std::fabs(cfg->ub[k]));
    }
    const double xmax = 4.0 * (box + 14.0);
    auto pick = [&](double term_max) {
      int f_total = 62 - (int)std::ceil(std::log2((double)cfg->n_sim * term_max));
      int f_term = 51 - (int)std::ceil(std::log2(SMM_ZIG_PER_BLOCK * term_max));
      int f = f_total < f_term ? f_total : f_term;
      return f > 60 ? 60 : f;
    };
    const int f_sum = pick(xmax), f_sq = pick(xmax * xmax);
    pb.magic_sum = std::ldexp(1.5, 52 - f_sum);
    pb.magic_sq = std::ldexp(1.5, 52 - f_sq);
    pb.scale_sum = std::ldexp(1.0, -f_sum);
    pb.scale_sq = std::ldexp(1.0, -f_sq);
    h->eval_param_limit = xmax - 14.0;
  }
  if (h->panel) {
    // Two-word fixed point of the per-individual sums (smm_panel.cuh).  Hard bounds from the sampling box with
    // |z| <= sqrt(2 * 52 ln 2) < 8.5 (the Box-Muller radius of a 52-bit uniform): |x_k| <= zmax / (1 - |phi_k|),
    // |y| <= (|mu0| + (sigma_alpha + sigma_eps) zmax + sum_k |beta_k| |x_k|) / (1 - |rho|); a per-individual sum has
    // at most T + 2 such products (HG_l has 2T + 1 terms of |y|).  N_ind of them must stay below 2^62 at grid 2^-Fhi.
    const int K = cfg->panel_K;
    auto amax = [&](int k) { return std::fmax(std::fabs(cfg->lb[k]), std::fabs(cfg->ub[k])); };
    const double zmax = 8.5;
    double bx = 0.0, bxb = 0.0;
    for (int k = 0; k < K; ++k) {
      const double b = zmax / (1.0 - amax(1 + K + k));
      bx = std::fmax(bx, b);
      bxb += amax(1 + k) * b;
    }
    const double by = (amax(3 + 2 * K) + (amax(1 + 2 * K) + amax(2 + 2 * K)) * zmax + bxb) / (1.0 - amax(0));
    const double term = std::fmax(std::fmax(by * by, bx * bx), std::fmax(by * bx, std::fmax(std::fmax(by, bx), 1.0)));
    const double bound = (2.0 * cfg->panel_T + 2.0) * term;
    const int fhi = 61 - (int)std::ceil(std::log2((double)cfg->panel_N)) - (int)std::ceil(std::log2(bound));
    if (fhi < -8) return fail(SMM_E_UNSUPPORTED_SHAPE, "panel: the sampling box allows sums too large for the exact accumulators");
    const int fh = fhi > 40 ? 40 : fhi;
    pb.pan_hi_scale = std::ldexp(1.0, fh);
    pb.pan_hi_inv = std::ldexp(1.0, -fh);
    pb.pan_lo_scale = std::ldexp(1.0, fh + 40);
    pb.pan_lo_inv = std::ldexp(1.0, -(fh + 40));
  }
  pb.lb = h->lb.p; pb.ub = h->ub.p; pb.init = h->init.p; pb.data = h->data.p; pb.w = h->w.p;
  pb.acc_tuner = h->acc_tuner.p; pb.min_improve = h->min_improve.p;

  DevState &st = h->st;
  st.sigma = h->sigma.p; st.accept_rate = h->accept_rate.p; st.n_noex = h->n_noex.p; st.n_acc = h->n_acc.p;
  st.la_cur = h->la_cur.p; st.la_pub = h->la_pub.p;
  st.la_all = have_la_all ? h->la_all.p : h->la_pub.p;
  st.applied = h->applied.p;
  st.t_value = h->t_value.p; st.t_prob = h->t_prob.p; st.t_curr = h->t_curr.p; st.t_best = h->t_best.p;
  st.t_params = h->t_params.p; st.t_mom = h->t_mom.p; st.t_acc = h->t_acc.p; st.t_status = h->t_status.p;
  st.t_exch = h->t_exch.p; st.t_bestid = h->t_bestid.p;
  st.partials = h->partials.p; st.arrive = h->arrive.p; st.unit_ctr = h->unit_ctr.p;
  st.sched_ij = h->sched_ij.p; st.sched_off = h->sched_off.p; st.sched_nlev = h->sched_nlev.p;
  st.err = h->err.p; st.counters = h->counters.p;
  st.val_all = h->val_all.p; st.pp = h->pp.p; st.bar = h->bar.p; st.sync_seq = h->sync_seq.p; st.flags = h->flags.p;
  for (int r = 0; r < kMaxWorld; ++r) {
    st.peer_la_all[r] = nullptr;
    st.peer_val_all[r] = nullptr;
    st.peer_flags[r] = nullptr;
  }
  st.ll = h->mode == 3 ? h->ll.p : nullptr;
  for (int r = 0; r < kMaxWorld; ++r) st.peer_ll[r] = nullptr;
  if (h->mode >= 2) {  // the barrier-free kernel always uses the gather layout; with one rank the "peer" is this GPU
    st.peer_la_all[0] = h->la_all.p;
    st.peer_val_all[0] = h->val_all.p;
    st.peer_ll[0] = st.ll;
  }
  if (fused_peers) {
    // the same plan on every rank: a peer's buffers sit at the same offsets of its arena
    const size_t o_la = (char *)h->la_all.p - (char *)arena_base, o_val = (char *)h->val_all.p - (char *)arena_base,
                 o_fl = (char *)h->flags.p - (char *)arena_base;
    for (int r = 0; r < h->world; ++r) {
      st.peer_la_all[r] = (double *)((char *)arena->peer[r] + o_la);
      st.peer_val_all[r] = (double *)((char *)arena->peer[r] + o_val);
      st.peer_flags[r] = (unsigned long long *)((char *)arena->peer[r] + o_fl);
      if (st.ll) st.peer_ll[r] = (unsigned long long *)((char *)arena->peer[r] + ((char *)h->ll.p - (char *)arena_base));
    }
  }
  st.phase_ts = want_phase_ts ? h->phase_ts.p : nullptr;

  if (h->N > 1 && I >= 2) {
    // Pairs[2 .. ] and their level schedules are data independent: the first window is computed right here, behind the
    // initialisation, instead of in front of the first iteration
    const int w = I - 1 < kPairChunk ? I - 1 : kPairChunk;
    launch_pairs(h->pb, h->st, 2, w, h->n_s, h->stream);
    CUDA_TRY(cudaGetLastError());
    h->ctr.kernel_launches++;
    h->sched_iter0 = 2;
    h->sched_n = w;
  }
  if (fused_peers) {
    // Cross-rank barrier on the device: a peer's first kernel must not store into this rank's arena before the
    // initialisation above has run.  (The other direction needs nothing: a handle's last kernel ends only after every
    // rank's last record has arrived, so nobody still writes into an arena whose handle was destroyed.)  Stream
    // ordered, no host synchronisation: the first iteration simply starts behind it.
    PeerBarrierArgs ba{};
    for (int r = 0; r < h->world; ++r) ba.slots[r] = (unsigned long long *)arena->peer[r];
    ba.world = h->world;
    ba.rank = h->rank;
    ba.epoch = ++arena->epoch;
    ba.err = h->err.p;
    peer_barrier_kernel<<<1, 32, 0, h->stream>>>(ba);
    CUDA_TRY(cudaGetLastError());
    h->ctr.kernel_launches++;
  }
  stamp("pairs + barrier");
  guard.ok = true;
  *out = h;
  return 0;
}

/* ends the process-wide caches (communicators, exchange arenas, streams); every handle must have been destroyed */
void smm_shutdown(void) {
  std::vector<RankCtx *> ctxs;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    ctxs.swap(g_rank_ctx);
  }
  for (RankCtx *c : ctxs) {
    cudaSetDevice(c->device);
    arena_destroy(c->arena, c->rank);
    if (c->comm) ncclCommDestroy(c->comm);
    delete c;
  }
  for (int d = 0; d < 64; ++d) {
    std::vector<StreamSet> v;
    {
      std::lock_guard<std::mutex> lk(g_cache_mu);
      v.swap(g_stream_cache[d]);
    }
    if (v.empty()) continue;
    cudaSetDevice(d);
    for (StreamSet &ss : v) stream_set_destroy(ss);
  }
}

int smm_bgp_iteration(const smm_bgp *h) { return h ? h->iter : -1; }
int smm_bgp_local_chains(const smm_bgp *h) { return h ? h->L : -1; }
void *smm_bgp_stream(smm_bgp *h) { return h ? (void *)h->stream : nullptr; }

}  // extern "C"

namespace {

cudaError_t prof_begin(smm_bgp *h, int kind) {
  if (!h->profiling) return cudaSuccess;
  const size_t i = h->prof_kind.size();
  while (h->prof_ev.size() < 2 * (i + 1)) {
    cudaEvent_t e;
    cudaError_t rc = cudaEventCreate(&e);
    if (rc != cudaSuccess) return rc;
    h->prof_ev.push_back(e);
  }
  h->prof_kind.push_back(kind);
  return cudaEventRecord(h->prof_ev[2 * i], h->stream);
}
cudaError_t prof_end(smm_bgp *h) {
  if (!h->profiling) return cudaSuccess;
  return cudaEventRecord(h->prof_ev[2 * (h->prof_kind.size() - 1) + 1], h->stream);
}

// enqueue iterations i+1 .. i+n_iters on the handle's stream (no synchronisation)
int enqueue_iterations(smm_bgp *h, int n_iters) {
  cudaStream_t s = h->stream;
  const bool exchange = h->N > 1;
  if (h->mode >= 1) {
    int left = n_iters;
    while (left > 0) {
      const int it0 = h->iter + 1;
      int n = left < kPairChunk ? left : kPairChunk;
      if (exchange) {
        // the launch needs the schedules of exchanges max(it0, 2) .. it0 + n - 1: keep a window of kPairChunk
        // iterations precomputed (one pairs launch per window, also across one-iteration step() calls)
        const int first_ex = it0 < 2 ? 2 : it0;
        if (it0 + n - 1 >= 2 &&
            (h->sched_iter0 < 0 || first_ex < h->sched_iter0 || first_ex >= h->sched_iter0 + h->sched_n)) {
          int w = h->max_iter - first_ex + 1;
          if (w > kPairChunk) w = kPairChunk;
          CUDA_TRY(prof_begin(h, 2));
          launch_pairs(h->pb, h->st, first_ex, w, h->n_s, s);
          CUDA_TRY(prof_end(h));
          h->ctr.kernel_launches++;
          h->sched_iter0 = first_ex;
          h->sched_n = w;
        }
        if (h->sched_iter0 >= 0 && it0 + n > h->sched_iter0 + h->sched_n) n = h->sched_iter0 + h->sched_n - it0;
      }
      CUDA_TRY(prof_begin(h, 0));
      CUDA_TRY((h->mode == 3 ? ll::launch_persistent : launch_persistent)(
          h->pb, h->st, it0, n, h->sched_iter0 < 0 ? 2 : h->sched_iter0, h->n_s, h->part_len, h->max_seg, h->cta_seg,
          h->grid, h->mode >= 2, h->done_base, s));
      if (h->mode >= 2) h->done_base += (unsigned long long)h->N * (unsigned)n;  // every chain of every rank, n times
      CUDA_TRY(prof_end(h));
      h->prof_iters += h->profiling ? n : 0;
      h->ctr.kernel_launches++;
      h->iter += n;
      left -= n;
    }
  }
  for (int k = 0; h->mode == 0 && k < n_iters; ++k) {
    const int it = h->iter + 1;
    if (exchange && it >= 2 && (h->sched_iter0 < 0 || it >= h->sched_iter0 + h->sched_n)) {
      // precompute Pairs[it .. it+chunk) and their level schedules
      int n = h->max_iter - it + 1;
      if (n > kPairChunk) n = kPairChunk;
      CUDA_TRY(prof_begin(h, 2));
      launch_pairs(h->pb, h->st, it, n, h->n_s, s);
      CUDA_TRY(prof_end(h));
      h->sched_iter0 = it;
      h->sched_n = n;
      h->ctr.kernel_launches++;
    }
    CUDA_TRY(prof_begin(h, 0));
    if (h->panel) {
      launch_propose(h->pb, h->st, it, h->part_len, s);
      launch_panel_chains(h->pb, h->st, it, h->panel_grid, h->panel_variant, s);
      h->ctr.kernel_launches++;
    } else {
      launch_eval(h->pb, h->st, it, h->n_split, h->part_len, s);
    }
    CUDA_TRY(prof_end(h));
    h->prof_iters += h->profiling ? 1 : 0;
    h->ctr.kernel_launches++;
    if (exchange && it >= 2) {  // AlgoBGP.jl:637
      if (h->world > 1) {
        CUDA_TRY(prof_begin(h, 3));
        NCCL_TRY(ncclAllGather(h->st.la_pub, h->st.la_all, (size_t)h->L * h->R, ncclDouble, h->comm, s));
        CUDA_TRY(prof_end(h));
        h->ctr.collectives++;
      }
      CUDA_TRY(prof_begin(h, 1));
      launch_exchange(h->pb, h->st, it, it - h->sched_iter0, h->n_s, s);
      CUDA_TRY(prof_end(h));
      h->ctr.kernel_launches++;
    }
    h->iter = it;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// after the stream drained: per-kernel times (profiling), counters, the sticky device error flag
int finish_step(smm_bgp *h) {
  for (size_t i = 0; i < h->prof_kind.size(); ++i) {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
    h->prof_ms[h->prof_kind[i]] += ms;
    h->prof_n[h->prof_kind[i]] += 1;
  }
  h->ctr.iterations = h->iter;
  h->ctr.evaluations = (int64_t)h->iter * h->L;
  return device_error_to_rc(*h->ss.h_err);  // copied behind the last kernel by the caller (page-locked word)
}

// D2H of trace rows [iter_lo, iter_hi] into `out`, whose row 0 is iteration `out_iter0`, on stream s
int enqueue_trace_copy(smm_bgp *h, int iter_lo, int iter_hi, int out_iter0, const smm_trace_view *out, cudaStream_t s) {
  const size_t L = h->L, off = (size_t)(iter_lo - 1) * L, n = (size_t)(iter_hi - iter_lo + 1) * L;
  const size_t ooff = (size_t)(iter_lo - out_iter0) * L;
#define COPY(dst, src, T, mult)                                                                              \
  if (out->dst)                                                                                              \
    CUDA_TRY(cudaMemcpyAsync(out->dst + ooff * (mult), h->st.src + off * (mult), sizeof(T) * n * (mult),     \
                             cudaMemcpyDeviceToHost, s))
  COPY(value, t_value, double, 1);
  COPY(prob, t_prob, double, 1);
  COPY(curr_val, t_curr, double, 1);
  COPY(best_val, t_best, double, 1);
  COPY(params, t_params, double, (size_t)h->P);
  COPY(sim_moments, t_mom, double, (size_t)h->M);
  COPY(accepted, t_acc, uint8_t, 1);
  COPY(status, t_status, int32_t, 1);
  COPY(exchanged, t_exch, int32_t, 1);
  COPY(best_id, t_bestid, int32_t, 1);
#undef COPY
  return 0;
}

}  // namespace

extern "C" {

int smm_bgp_step(smm_bgp *h, int32_t n_iters, float *elapsed_ms) {
  if (!h) return fail(SMM_E_ARG, "null handle");
  if (n_iters < 0 || h->iter + n_iters > h->max_iter)
    return fail(SMM_E_ARG, "step would exceed max_iter (use restart!/extend to grow the chains)");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  CUDA_TRY(cudaEventRecord(h->ev0, s));
  h->prof_kind.clear();
  if (int rc = enqueue_iterations(h, n_iters)) return rc;
  CUDA_TRY(cudaEventRecord(h->ev1, s));
  CUDA_TRY(cudaMemcpyAsync(h->ss.h_err, h->st.err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
  return finish_step(h);
}

// run!(algo) with the trace streamed to the host: iterations are enqueued in windows; the rows of a finished
// window travel to `host_out` on the copy stream while the next window computes.
int smm_bgp_run(smm_bgp *h, int32_t n_iters, int32_t window, const smm_trace_view *host_out, float *elapsed_ms) {
  if (!h) return fail(SMM_E_ARG, "null handle");
  if (n_iters < 0 || h->iter + n_iters > h->max_iter)
    return fail(SMM_E_ARG, "run would exceed max_iter (use restart!/extend to grow the chains)");
  CUDA_TRY(cudaSetDevice(h->device));
  if (window <= 0) window = kPairChunk;
  cudaStream_t s = h->stream;
  CUDA_TRY(cudaEventRecord(h->ev0, s));
  h->prof_kind.clear();
  const int first = h->iter + 1;
  size_t wi = 0;
  for (int left = n_iters; left > 0;) {
    const int it0 = h->iter + 1;
    int n = left < window ? left : window;
    // keep windows aligned with the precomputed pair-schedule windows, so a launch never has to be cut in two
    if (h->mode >= 1 && h->N > 1 && h->sched_iter0 >= 0 && it0 >= h->sched_iter0 &&
        it0 < h->sched_iter0 + h->sched_n && it0 + n > h->sched_iter0 + h->sched_n)
      n = h->sched_iter0 + h->sched_n - it0;
    if (int rc = enqueue_iterations(h, n)) return rc;
    left -= n;
    if (host_out) {
      while (h->win_ev.size() <= wi) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->win_ev.push_back(e);
      }
      CUDA_TRY(cudaEventRecord(h->win_ev[wi], s));
      CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->win_ev[wi], 0));
      if (int rc = enqueue_trace_copy(h, it0, it0 + n - 1, first, host_out, h->copy_stream)) return rc;
      ++wi;
    }
  }
  CUDA_TRY(cudaEventRecord(h->ev1, s));
  CUDA_TRY(cudaMemcpyAsync(h->ss.h_err, h->st.err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (host_out) CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
  return finish_step(h);
}

int smm_host_alloc(int64_t nbytes, void **out) {
  if (!out || nbytes < 0) return fail(SMM_E_ARG, "bad argument");
  *out = nullptr;
  CUDA_TRY(cudaHostAlloc(out, (size_t)(nbytes ? nbytes : 1), cudaHostAllocDefault));
  return 0;
}

void smm_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int smm_bgp_read_trace(smm_bgp *h, int32_t iter_lo, int32_t iter_hi, const smm_trace_view *out) {
  if (!h || !out) return fail(SMM_E_ARG, "null argument");
  if (iter_lo < 1 || iter_hi < iter_lo || iter_hi > h->max_iter) return fail(SMM_E_ARG, "bad iteration range");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  if (int rc = enqueue_trace_copy(h, iter_lo, iter_hi, iter_lo, out, s)) return rc;
  CUDA_TRY(cudaStreamSynchronize(s));
  return 0;
}

int smm_bgp_read_chain_state(smm_bgp *h, double *sigma, double *accept_rate) {
  if (!h) return fail(SMM_E_ARG, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (sigma) CUDA_TRY(cudaMemcpy(sigma, h->st.sigma, sizeof(double) * h->L, cudaMemcpyDeviceToHost));
  if (accept_rate)
    CUDA_TRY(cudaMemcpy(accept_rate, h->st.accept_rate, sizeof(double) * h->L, cudaMemcpyDeviceToHost));
  return 0;
}

int smm_bgp_get_counters(smm_bgp *h, smm_counters *out) {
  if (!h || !out) return fail(SMM_E_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  unsigned long long c[4];
  CUDA_TRY(cudaMemcpy(c, h->st.counters, sizeof c, cudaMemcpyDeviceToHost));
  h->ctr.accepted = (int64_t)c[0];
  h->ctr.swaps = (int64_t)c[1];
  h->ctr.proposal_attempts = (int64_t)c[2];
  *out = h->ctr;
  return 0;
}

int smm_bgp_eval_batch(smm_bgp *h, const double *params, int32_t B, int32_t noseed, uint32_t rep0, double *value,
                       double *moments, int32_t *status) {
  if (!h || !params) return fail(SMM_E_ARG, "null argument");
  if (B < 1) return fail(SMM_E_ARG, "B must be positive");
  if (h->panel) {
    for (int64_t i = 0; i < (int64_t)B * h->P; ++i)
      if (!(params[i] >= h->h_lb[i % h->P] && params[i] <= h->h_ub[i % h->P]))
        return fail(SMM_E_ARG, "panel: parameter outside the sampling box (the exact accumulators are sized from the box)");
  } else {
    for (int64_t i = 0; i < (int64_t)B * h->P; ++i)
      if (!(std::fabs(params[i]) <= h->eval_param_limit))
        return fail(SMM_E_ARG, "parameter far outside the sampling box (> 4x): accumulators are sized from the box");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  const int P = h->P, M = h->M;
  if (h->panel) {
    DevBuf<double> d_params, d_value, d_mom;
    DevBuf<unsigned long long> d_acc;
    DevBuf<int> d_status;
    DevBuf<unsigned> d_done;  // [B] + the queue head
    struct FreeP {
      DevBuf<double> &a, &b, &c;
      DevBuf<unsigned long long> &d;
      DevBuf<int> &e;
      DevBuf<unsigned> &f;
      cudaStream_t st;
      ~FreeP() {
        cudaStreamSynchronize(st);
        a.free(); b.free(); c.free(); d.free(); e.free(); f.free();
      }
    } frp{d_params, d_value, d_mom, d_acc, d_status, d_done, h->stream};
    CUDA_TRY(d_params.alloc((size_t)B * P));
    CUDA_TRY(d_value.alloc(B));
    CUDA_TRY(d_mom.alloc((size_t)B * M));
    CUDA_TRY(d_acc.alloc((size_t)B * h->part_len));
    CUDA_TRY(d_status.alloc(B));
    CUDA_TRY(d_done.alloc((size_t)B + 1));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0));  // pool allocations are ordered on the legacy stream
    cudaStream_t s = h->stream;
    CUDA_TRY(cudaMemcpyAsync(d_params.p, params, sizeof(double) * B * P, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(d_acc.p, 0, sizeof(unsigned long long) * (size_t)B * h->part_len, s));
    CUDA_TRY(cudaMemsetAsync(d_done.p, 0, sizeof(unsigned) * ((size_t)B + 1), s));
    launch_panel_batch(h->pb, h->st, d_params.p, B, noseed, 0u, rep0, d_acc.p, d_done.p, d_done.p + B, d_value.p, d_mom.p,
                       d_status.p, h->panel_grid, h->panel_variant, s);
    h->ctr.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    if (value) CUDA_TRY(cudaMemcpyAsync(value, d_value.p, sizeof(double) * B, cudaMemcpyDeviceToHost, s));
    if (moments) CUDA_TRY(cudaMemcpyAsync(moments, d_mom.p, sizeof(double) * B * M, cudaMemcpyDeviceToHost, s));
    if (status) CUDA_TRY(cudaMemcpyAsync(status, d_status.p, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
  }
  int n_split = h->n_split;
  DevBuf<double> d_params, d_value, d_mom, d_part;
  DevBuf<int> d_status;
  DevBuf<unsigned> d_arrive;
  struct Free {
    DevBuf<double> &a, &b, &c, &d;
    DevBuf<int> &e;
    DevBuf<unsigned> &f;
    cudaStream_t st;
    ~Free() {
      cudaStreamSynchronize(st);
      a.free(); b.free(); c.free(); d.free(); e.free(); f.free();
    }
  } fr{d_params, d_value, d_mom, d_part, d_status, d_arrive, h->stream};
  CUDA_TRY(d_params.alloc((size_t)B * P));
  CUDA_TRY(d_value.alloc(B));
  CUDA_TRY(d_mom.alloc((size_t)B * M));
  CUDA_TRY(d_part.alloc((size_t)B * n_split * h->part_len));
  CUDA_TRY(d_status.alloc(B));
  CUDA_TRY(d_arrive.alloc(B));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0));  // pool allocations are ordered on the legacy stream
  cudaStream_t s = h->stream;
  CUDA_TRY(cudaMemcpyAsync(d_params.p, params, sizeof(double) * B * P, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemsetAsync(d_arrive.p, 0, sizeof(unsigned) * B, s));
  // grid.y is limited to 65535
  for (int b0 = 0; b0 < B; b0 += 32768) {
    const int nb = (B - b0) < 32768 ? (B - b0) : 32768;
    launch_objective(h->pb, d_params.p + (size_t)b0 * P, nb, noseed, rep0 + (uint32_t)b0, n_split, h->part_len,
                     d_part.p + (size_t)b0 * n_split * h->part_len, d_arrive.p + b0, d_value.p + b0,
                     d_mom.p + (size_t)b0 * M, d_status.p + b0, s);
    h->ctr.kernel_launches++;
  }
  CUDA_TRY(cudaGetLastError());
  if (value) CUDA_TRY(cudaMemcpyAsync(value, d_value.p, sizeof(double) * B, cudaMemcpyDeviceToHost, s));
  if (moments) CUDA_TRY(cudaMemcpyAsync(moments, d_mom.p, sizeof(double) * B * M, cudaMemcpyDeviceToHost, s));
  if (status) CUDA_TRY(cudaMemcpyAsync(status, d_status.p, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return 0;
}

// ---- accepted-only statistics and summary(c), reduced on the device (AlgoBGP.jl:174-206) -----------------------------
int smm_bgp_accepted_stats(smm_bgp *h, int32_t iter_lo, int32_t iter_hi, const double *probs, int32_t n_probs,
                           int64_t *count, double *mean, double *quantiles) {
  if (!h) return fail(SMM_E_ARG, "null handle");
  if (iter_lo < 1 || iter_hi < iter_lo || iter_hi > h->iter) return fail(SMM_E_ARG, "bad iteration range (1 <= lo <= hi <= iterations run)");
  if (n_probs < 0 || (n_probs > 0 && (!probs || !quantiles))) return fail(SMM_E_ARG, "bad quantile arguments");
  for (int q = 0; q < n_probs; ++q)
    if (!(probs[q] >= 0.0 && probs[q] <= 1.0)) return fail(SMM_E_ARG, "quantile probabilities must lie in [0, 1]");
  CUDA_TRY(cudaSetDevice(h->device));
  const int L = h->L, P = h->P, n = iter_hi - iter_lo + 1;
  int cap2 = 1;
  while (cap2 < n) cap2 <<= 1;
  const bool need_scratch = cap2 > stats_smem_cap();  // more accepted values than the shared-memory sort holds
  const size_t nq = (size_t)L * P * (n_probs > 0 ? n_probs : 1);
  cudaStream_t s = h->stream;
  char *d = nullptr;
  const size_t o_cnt = 0, o_mean = o_cnt + 8 * (size_t)L, o_q = o_mean + 8 * (size_t)L * P, o_p = o_q + 8 * nq,
               o_scr = o_p + 8 * (size_t)(n_probs > 0 ? n_probs : 1),
               total = o_scr + (need_scratch ? 8 * (size_t)L * P * cap2 : 8);
  CUDA_TRY(cudaMallocAsync((void **)&d, total, s));
  cudaError_t e = cudaSuccess;
  if (n_probs > 0) e = cudaMemcpyAsync(d + o_p, probs, 8 * (size_t)n_probs, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess)
    e = launch_accepted_stats(h->st, L, P, iter_lo, iter_hi, (const double *)(d + o_p), n_probs, (long long *)(d + o_cnt),
                              (double *)(d + o_mean), (double *)(d + o_q), (double *)(d + o_scr), cap2, s);
  h->ctr.kernel_launches++;
  if (e == cudaSuccess && count) e = cudaMemcpyAsync(count, d + o_cnt, 8 * (size_t)L, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && mean) e = cudaMemcpyAsync(mean, d + o_mean, 8 * (size_t)L * P, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && n_probs > 0) e = cudaMemcpyAsync(quantiles, d + o_q, 8 * nq, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFreeAsync(d, s);
  CUDA_TRY(e);
  return 0;
}

int smm_bgp_chain_summary(smm_bgp *h, int64_t *n_exchanged, int32_t *exchanged_most_with, double *best_val) {
  if (!h) return fail(SMM_E_ARG, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  const int L = h->L;
  cudaStream_t s = h->stream;
  char *d = nullptr;
  CUDA_TRY(cudaMallocAsync((void **)&d, 24 * (size_t)L, s));
  cudaError_t e = launch_chain_summary(h->st, L, h->N, h->iter, (long long *)d, (int *)(d + 16 * (size_t)L),
                                       (double *)(d + 8 * (size_t)L), s);
  h->ctr.kernel_launches++;
  if (e == cudaSuccess && n_exchanged) e = cudaMemcpyAsync(n_exchanged, d, 8 * (size_t)L, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && best_val) e = cudaMemcpyAsync(best_val, d + 8 * (size_t)L, 8 * (size_t)L, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && exchanged_most_with)
    e = cudaMemcpyAsync(exchanged_most_with, d + 16 * (size_t)L, 4 * (size_t)L, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFreeAsync(d, s);
  CUDA_TRY(e);
  return 0;
}

int smm_stream_acc_uniforms(uint64_t seed_algo, uint32_t chain, int32_t iter_lo, int32_t iter_hi, double *out) {
  if (!out || iter_lo < 1 || iter_hi < iter_lo) return fail(SMM_E_ARG, "bad argument");
  for (int32_t it = iter_lo; it <= iter_hi; ++it) out[it - iter_lo] = smm_acc_uniform(seed_algo, chain, (uint32_t)it);
  return 0;
}

int smm_bgp_set_profiling(smm_bgp *h, int32_t enabled) {
  if (!h) return fail(SMM_E_ARG, "null handle");
  h->profiling = enabled != 0;
  h->prof_iters = 0;
  for (int k = 0; k < 4; ++k) {
    h->prof_ms[k] = 0.0;
    h->prof_n[k] = 0;
  }
  return 0;
}

int smm_bgp_kernel_times(smm_bgp *h, double ms_sum[4], int64_t launches[4]) {
  if (!h || !ms_sum || !launches) return fail(SMM_E_ARG, "null argument");
  for (int k = 0; k < 4; ++k) {
    ms_sum[k] = h->prof_ms[k];
    launches[k] = h->prof_n[k];
  }
  return (int)h->prof_iters;  /* iterations covered by the kind-0 launches (>= 0) */
}

// ---- checkpoint ----------------------------------------------------------------------------------
// layout: header {magic, P, M, L, R, iter, N, world, rank, seed_algo, seed_sim} (11 x int64) | sigma[L] accept_rate[L] | n_noex[L] n_acc[L] (int32)
//         | la_cur[L][R] | trace rows 1..iter of every column
namespace {
struct StateHeader {
  int64_t magic, P, M, L, R, iter;
  int64_t N, world, rank;        // whose chains these are: a blob of another rank / ensemble / seed is rejected
  uint64_t seed_algo, seed_sim;  // the streams the trace was drawn from
};
const int64_t kMagic = 0x534d4d4232303032ll;  // "SMMB2002"
}  // namespace

int64_t smm_bgp_state_bytes(const smm_bgp *h) {
  if (!h) return -1;
  const int64_t L = h->L, R = h->R, it = h->iter, P = h->P, M = h->M;
  return (int64_t)sizeof(StateHeader) + 8 * 2 * L + 4 * 2 * L + 8 * L * R + it * L * (8 * (4 + P + M) + 1 + 4 * 3);
}

int smm_bgp_export_state(smm_bgp *h, void *buf, int64_t nbytes) {
  if (!h || !buf) return fail(SMM_E_ARG, "null argument");
  if (nbytes < smm_bgp_state_bytes(h)) return fail(SMM_E_ARG, "buffer too small");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  char *p = (char *)buf;
  StateHeader hd{kMagic, h->P, h->M, h->L, h->R, h->iter, h->N, h->world, h->rank, h->seed_algo, h->seed_sim};
  memcpy(p, &hd, sizeof hd);
  p += sizeof hd;
  const size_t L = h->L, n = (size_t)h->iter * L;
#define OUT(src, T, count)                                                               \
  do {                                                                                   \
    CUDA_TRY(cudaMemcpy(p, src, sizeof(T) * (count), cudaMemcpyDeviceToHost));           \
    p += sizeof(T) * (count);                                                            \
  } while (0)
  OUT(h->st.sigma, double, L);
  OUT(h->st.accept_rate, double, L);
  OUT(h->st.n_noex, int, L);
  OUT(h->st.n_acc, int, L);
  OUT(h->st.la_cur, double, L * h->R);
  OUT(h->st.t_value, double, n);
  OUT(h->st.t_prob, double, n);
  OUT(h->st.t_curr, double, n);
  OUT(h->st.t_best, double, n);
  OUT(h->st.t_params, double, n * h->P);
  OUT(h->st.t_mom, double, n * h->M);
  OUT(h->st.t_acc, uint8_t, n);
  OUT(h->st.t_status, int, n);
  OUT(h->st.t_exch, int, n);
  OUT(h->st.t_bestid, int, n);
#undef OUT
  return 0;
}

int smm_bgp_import_state(smm_bgp *h, const void *buf, int64_t nbytes) {
  if (!h || !buf) return fail(SMM_E_ARG, "null argument");
  if (nbytes < (int64_t)sizeof(StateHeader)) return fail(SMM_E_ARG, "buffer too small");
  const char *p = (const char *)buf;
  StateHeader hd;
  memcpy(&hd, p, sizeof hd);
  p += sizeof hd;
  if (hd.magic != kMagic || hd.P != h->P || hd.M != h->M || hd.L != h->L || hd.R != h->R)
    return fail(SMM_E_STATE, "checkpoint does not match this handle's shape");
  if (hd.N != h->N || hd.world != h->world || hd.rank != h->rank)
    return fail(SMM_E_STATE, "checkpoint belongs to another rank or ensemble (n_chains / world_size / rank differ)");
  if (hd.seed_algo != h->seed_algo || hd.seed_sim != h->seed_sim)
    return fail(SMM_E_STATE, "checkpoint was drawn from other streams (seed_algo / seed_sim differ)");
  if (hd.iter < 0 || hd.iter > h->max_iter) return fail(SMM_E_STATE, "checkpoint has more iterations than max_iter");
  const size_t L = h->L, n = (size_t)hd.iter * L;
  const int64_t need = (int64_t)sizeof(StateHeader) + 8 * 2 * (int64_t)L + 4 * 2 * (int64_t)L + 8 * (int64_t)L * h->R +
                       (int64_t)n * (8 * (4 + h->P + h->M) + 1 + 4 * 3);
  if (nbytes < need) return fail(SMM_E_ARG, "truncated checkpoint");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
#define IN(dst, T, count)                                                                \
  do {                                                                                   \
    CUDA_TRY(cudaMemcpyAsync(dst, p, sizeof(T) * (count), cudaMemcpyHostToDevice, h->stream)); \
    p += sizeof(T) * (count);                                                            \
  } while (0)
  IN(h->st.sigma, double, L);
  IN(h->st.accept_rate, double, L);
  IN(h->st.n_noex, int, L);
  IN(h->st.n_acc, int, L);
  IN(h->st.la_cur, double, L * h->R);
  IN(h->st.t_value, double, n);
  IN(h->st.t_prob, double, n);
  IN(h->st.t_curr, double, n);
  IN(h->st.t_best, double, n);
  IN(h->st.t_params, double, n * h->P);
  IN(h->st.t_mom, double, n * h->M);
  IN(h->st.t_acc, uint8_t, n);
  IN(h->st.t_status, int, n);
  IN(h->st.t_exch, int, n);
  IN(h->st.t_bestid, int, n);
#undef IN
  // completion counter / applied marks of exchange_mode 2 refer to the run that is being replaced.  (With several ranks
  // every rank imports between the same two steps, and a step ends only when all ranks have finished it.)
  CUDA_TRY(cudaMemsetAsync(h->val_all.p + 2 * (size_t)h->N, 0, sizeof(double) * h->N, h->stream));
  h->done_base = 0;
  CUDA_TRY(cudaMemsetAsync(h->applied.p, 0, sizeof(unsigned) * L, h->stream));
  // exchange_mode 3: stale flag-in-data words of the replaced run could carry the tags of the iterations to come
  if (h->ll.p) CUDA_TRY(cudaMemsetAsync(h->ll.p, 0, sizeof(unsigned long long) * 2 * (size_t)h->N * ll_words(h->P), h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->iter = (int)hd.iter;
  h->sched_iter0 = -1;
  h->sched_n = 0;
  return 0;
}

// ---- diagnostics ---------------------------------------------------------------------------------
static int debug_normals_impl(int32_t device, uint64_t seed, uint32_t k, uint32_t c2, uint32_t c3, int32_t n_pairs,
                              int zig, double *out) {
  if (!out || n_pairs < 1) return fail(SMM_E_ARG, "bad argument");
  const size_t per = zig ? SMM_ZIG_PER_BLOCK : 2;  // normals per Philox block
  if (smm_device_count() <= device) return fail(SMM_E_CUDA, "no such CUDA device");
  CUDA_TRY(cudaSetDevice(device));
  DevBuf<double> d;
  CUDA_TRY(d.alloc(per * n_pairs, true));
  launch_debug_normals(seed, k, c2, c3, n_pairs, zig, d.p, 0);
  cudaError_t e = cudaMemcpy(out, d.p, sizeof(double) * per * n_pairs, cudaMemcpyDeviceToHost);
  d.free();
  CUDA_TRY(e);
  return 0;
}
int smm_debug_normals(int32_t device, uint64_t seed, uint32_t k, uint32_t c2, uint32_t c3, int32_t n_pairs,
                      double *out) {
  return debug_normals_impl(device, seed, k, c2, c3, n_pairs, 0, out);
}
int smm_debug_zig_normals(int32_t device, uint64_t seed, uint32_t k, uint32_t c2, uint32_t c3, int32_t n_pairs,
                          double *out) {
  return debug_normals_impl(device, seed, k, c2, c3, n_pairs, 1, out);
}

int smm_debug_pairs(smm_bgp *h, int32_t iter, int32_t *ij, int32_t *level_offsets, int32_t *n_levels) {
  if (!h || !ij || !level_offsets || !n_levels) return fail(SMM_E_ARG, "null argument");
  if (h->N < 2) return fail(SMM_E_ARG, "no exchange with a single chain");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  launch_pairs(h->pb, h->st, iter, 1, h->n_s, h->stream);
  h->sched_iter0 = -1;  // the cached schedule was overwritten
  h->sched_n = 0;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemcpy(ij, h->st.sched_ij, sizeof(int) * 2 * h->n_s, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(level_offsets, h->st.sched_off, sizeof(int) * (h->n_s + 1), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(n_levels, h->st.sched_nlev, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

int smm_debug_phase_ts(smm_bgp *h, uint64_t *out, int64_t n) {
  if (!h || !out) return fail(SMM_E_ARG, "null argument");
  if (!h->st.phase_ts) return fail(SMM_E_STATE, "set SMM_PHASE_TS=1 before smm_bgp_create");
  // persistent modes: 4 rows per CTA, then one row per local chain {publish start, tag time, finishing CTA, iteration},
  // then 2 rows per CTA {before the system fence, after it, after the completion adds} (world > 1, exchange_mode 2)
  const int64_t blocks = h->mode >= 1 ? (int64_t)h->grid * 6 + h->L : (int64_t)h->L * h->n_split;
  const int64_t have = blocks * 4;
  CUDA_TRY(cudaMemcpy(out, h->st.phase_ts, sizeof(uint64_t) * (n < have ? n : have), cudaMemcpyDeviceToHost));
  return (int)blocks;
}

int smm_debug_barrier_bench(smm_bgp *h, int32_t variant, int32_t n, float *elapsed_ms) {
  if (!h || n < 1) return fail(SMM_E_ARG, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, h->device));
  cudaStream_t s = h->stream;
  CUDA_TRY(launch_barrier_bench(h->pb, h->st, variant, 10, prop.multiProcessorCount, s));
  CUDA_TRY(cudaEventRecord(h->ev0, s));
  CUDA_TRY(launch_barrier_bench(h->pb, h->st, variant, n, prop.multiProcessorCount, s));
  CUDA_TRY(cudaEventRecord(h->ev1, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
  return 0;
}

int smm_debug_sim_throughput(smm_bgp *h, int32_t n_pairs_per_thread, int32_t blocks, int32_t threads, int32_t dynamic,
                             float *elapsed_ms) {
  if (!h || threads < 32 || threads > 1024 || threads % 32 || blocks < 1) return fail(SMM_E_ARG, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  DevBuf<double> d;
  CUDA_TRY(d.alloc((size_t)blocks * threads * 2, true));
  cudaStream_t s = h->stream;
  launch_sim_throughput(h->pb, n_pairs_per_thread, blocks, threads, dynamic, d.p, s);  // warm-up
  CUDA_TRY(cudaEventRecord(h->ev0, s));
  launch_sim_throughput(h->pb, n_pairs_per_thread, blocks, threads, dynamic, d.p, s);
  CUDA_TRY(cudaEventRecord(h->ev1, s));
  cudaError_t e = cudaStreamSynchronize(s);
  d.free();
  CUDA_TRY(e);
  CUDA_TRY(cudaGetLastError());
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
  return 0;
}

int smm_debug_rng_throughput(int32_t device, int64_t n_pairs_per_thread, int32_t blocks, int32_t threads,
                             float *elapsed_ms, double *checksum) {
  (void)threads;
  if (smm_device_count() <= device) return fail(SMM_E_CUDA, "no such CUDA device");
  CUDA_TRY(cudaSetDevice(device));
  DevBuf<double> d;
  const size_t n = (size_t)blocks * kEvalThreads * 2;
  CUDA_TRY(d.alloc(n, true));
  cudaEvent_t a, b;
  CUDA_TRY(cudaEventCreate(&a));
  CUDA_TRY(cudaEventCreate(&b));
  launch_rng_throughput(n_pairs_per_thread, blocks, d.p, 0);  // warm-up
  CUDA_TRY(cudaEventRecord(a, 0));
  launch_rng_throughput(n_pairs_per_thread, blocks, d.p, 0);
  CUDA_TRY(cudaEventRecord(b, 0));
  CUDA_TRY(cudaEventSynchronize(b));
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, a, b));
  std::vector<double> hbuf(n);
  cudaError_t e = cudaMemcpy(hbuf.data(), d.p, sizeof(double) * n, cudaMemcpyDeviceToHost);
  d.free();
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  CUDA_TRY(e);
  if (checksum) {
    double cs = 0.0;
    for (double v : hbuf) cs += v;
    *checksum = cs;
  }
  return 0;
}

}  // extern "C"
