// pbf_slab.cu — host side of the x-slab decomposition: slab
// planning, the two transports (NCCL between processes, peer copies inside one process), the
// per-substep driver that interleaves the kernels of kernels/slab.cu with the solver passes, and
// the C ABI around them.  The reference has no multi-GPU path; parity is pinned by requiring the
// slab result to be bit-identical to the single-GPU result (tests/test_gpu_slabs.py).
#include <nccl.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "pbf_context.h"

using namespace pbf;

// ================================================================== slab planning (host only)
namespace {

// static_cast<int>(std::floor(x * inv)) with inv = 1.0f / h — reference core.cpp:28-34; the
// same expression the device uses (cell_coord in pbf_device.cuh).
inline int host_cell(float x, float inv_h) {
  const float f = std::floor(x * inv_h);
  if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT_MIN;
  return (int)f;
}

// ---- work-weighted cuts -------------------------------------------------------------------------
// A slab does not only pay for the particles it owns: the first ghost layer on each interior side
// gets neighbour lists and a lambda per iteration (about half the cost of an owned particle:
// neighbours + lambda are ~50 % of a substep), the second layer is only copied.
// Modelled cost of slab r owning layers [a, b):
//     sum hist[a..b)  +  w * (g(a-1) + g(b))  +  w/4 * (g(a-2) + g(b+1))
// with g(l) = hist[l] on a side that faces another slab and 0 at the two ends of the scene.  For
// 0 <= w <= 1 the cost grows when the interval grows on either side.  Whether every slab can stay
// under a bound T is decided by a sweep over the reachable cut positions (the >= 2 layers rule
// rules out the plain greedy fill), and a binary search over T gives the cuts that minimise the
// most expensive slab.  w = 0 balances owned particles only.  Costs are integers in units of
// 1/256 particle.
constexpr unsigned long long kCostUnit = 256;

struct CutPlanner {
  const std::vector<size_t>& hist;
  const int nranks;
  const long long layers;
  unsigned long long w1, w2;               // ghost weights in 1/256
  std::vector<unsigned long long> prefix;  // prefix[l] = particles in layers [0, l)

  CutPlanner(const std::vector<size_t>& h, int r, float ghost_weight) : hist(h), nranks(r), layers((long long)h.size()) {
    const float w = std::min(1.0f, std::max(0.0f, ghost_weight));
    w1 = (unsigned long long)std::lround((double)w * (double)kCostUnit);
    w2 = w1 / 4;
    prefix.assign(hist.size() + 1, 0);
    for (size_t l = 0; l < hist.size(); ++l) prefix[l + 1] = prefix[l] + hist[l];
  }
  unsigned long long at(long long l) const { return (l >= 0 && l < layers) ? (unsigned long long)hist[(size_t)l] : 0ull; }
  // cost of slab r owning layers [a, b)
  unsigned long long cost(int r, long long a, long long b) const {
    unsigned long long c = (prefix[(size_t)b] - prefix[(size_t)a]) * kCostUnit;
    if (r > 0) c += w1 * at(a - 1) + w2 * at(a - 2);
    if (r < nranks - 1) c += w1 * at(b) + w2 * at(b + 1);
    return c;
  }
  // reach[r][b] != 0: slabs 0..r can own the layers [0, b), each at least two layers wide and at
  // most T expensive, with room for two layers per remaining slab (r = 0 .. nranks-2).  Returns
  // whether the last slab can then take the rest.  For a fixed end the cheapest start of slab r is
  // the LARGEST reachable one (cost shrinks with the interval), so one running maximum per slab
  // decides every end: O(nranks * layers).
  bool sweep(unsigned long long T, std::vector<std::vector<char>>& reach) const {
    reach.assign((size_t)nranks - 1, std::vector<char>((size_t)layers + 1, 0));
    for (int r = 0; r + 1 < nranks; ++r) {
      const long long first = 2LL * (r + 1), last = layers - 2LL * (nranks - 1 - r);
      long long start = (r == 0) ? 0 : -1;
      for (long long b = first; b <= last; ++b) {
        if (r > 0 && reach[(size_t)r - 1][(size_t)(b - 2)]) start = b - 2;
        reach[(size_t)r][(size_t)b] = (start >= 0 && cost(r, start, b) <= T) ? 1 : 0;
      }
    }
    for (long long a = layers - 2; a >= 2LL * (nranks - 1); --a)
      if (reach[(size_t)nranks - 2][(size_t)a]) return cost(nranks - 1, a, layers) <= T;
    return false;
  }
};

// Cuts on x-cell boundaries that minimise the modelled cost of the most expensive slab; every slab
// at least two cells wide so that both ghost layers of a slab come from its direct neighbour.
// hist[l] = particles in the x-layer lo + l.  Among the plans that reach the optimum, each cut is
// the one closest to its equal-share quantile, so that the cheaper slabs are balanced as well and
// a re-plan on a slightly different histogram moves few layers.
int plan_cuts_hist(const std::vector<size_t>& hist, int lo, int nranks, float ghost_weight, std::vector<int>& cuts,
                   std::string& err) {
  cuts.assign((size_t)nranks + 1, 0);
  cuts[0] = INT_MIN;
  cuts[nranks] = INT_MAX;
  if (nranks == 1) return PBF_OK;
  const long long layers = (long long)hist.size();
  unsigned long long n = 0;
  for (size_t c : hist) n += c;
  if (n == 0 || layers < 2LL * nranks) {
    char buf[160];
    std::snprintf(buf, sizeof(buf), "slab plan: %lld x-layers of cells can not be split into %d slabs of >= 2 layers",
                  layers, nranks);
    err = buf;
    return PBF_E_INVALID;
  }
  const CutPlanner pl(hist, nranks, ghost_weight);
  std::vector<std::vector<char>> reach;
  // owned and ghost layers of a slab are disjoint and every weight is <= 1: no slab costs more than n
  unsigned long long lo_t = 0, hi_t = n * kCostUnit;
  while (lo_t < hi_t) {
    const unsigned long long mid = lo_t + (hi_t - lo_t) / 2;
    if (pl.sweep(mid, reach)) hi_t = mid; else lo_t = mid + 1;
  }
  const unsigned long long T = hi_t;
  if (!pl.sweep(T, reach)) {  // unreachable: T = n admits every plan of >= 2 layers per slab
    err = "slab plan: internal error (no feasible plan)";
    return PBF_E_INVALID;
  }
  // Walk back from the end: the start of slab r is any reachable cut that keeps the slab under T;
  // take the one closest to the equal-share quantile n * r / nranks.
  long long b = layers;
  for (int r = nranks - 1; r >= 1; --r) {
    const unsigned long long target = n * (unsigned)r;
    long long pick = -1;
    unsigned long long pick_dist = 0;
    for (long long a = b - 2; a >= 2LL * r; --a) {
      if (pl.cost(r, a, b) > T) break;  // cost only grows as the start moves left
      if (!reach[(size_t)r - 1][(size_t)a]) continue;
      const unsigned long long have = pl.prefix[(size_t)a] * (unsigned)nranks;
      const unsigned long long dist = have > target ? have - target : target - have;
      if (pick < 0 || dist < pick_dist) {
        pick = a;
        pick_dist = dist;
      }
    }
    if (pick < 0) {
      err = "slab plan: internal error (back-tracking lost the plan)";
      return PBF_E_INVALID;
    }
    cuts[(size_t)r] = (int)(lo + pick);
    b = pick;
  }
  return PBF_OK;
}

// Default ghost weight of every planner call of this process (environment PBF_SLAB_GHOST_WEIGHT,
// 0 = balance owned particles only).
float default_ghost_weight() {  // read on every call: an A/B run changes it between two uploads of one process
  const char* e = std::getenv("PBF_SLAB_GHOST_WEIGHT");
  if (e && *e) {
    char* end = nullptr;
    const float v = std::strtof(e, &end);
    if (end != e && v >= 0.0f && v <= 1.0f) return v;
  }
  return 0.5f;
}

int plan_cuts(size_t n, const float* px, float h, int nranks, std::vector<int>& cuts, std::string& err) {
  if (nranks == 1) return plan_cuts_hist({}, 0, 1, 0.0f, cuts, err);
  if (n == 0) {
    err = "slab plan: no particles";
    return PBF_E_INVALID;
  }
  const float inv_h = 1.0f / h;
  int lo = INT_MAX, hi = INT_MIN;
  for (size_t i = 0; i < n; ++i) {
    const int c = host_cell(px[i], inv_h);
    lo = std::min(lo, c);
    hi = std::max(hi, c);
  }
  const long long layers = (long long)hi - lo + 1;
  if (lo == INT_MIN || layers > (1LL << 24)) {
    err = "slab plan: non-finite or absurdly spread x coordinates";
    return PBF_E_INVALID;
  }
  std::vector<size_t> hist((size_t)layers, 0);
  for (size_t i = 0; i < n; ++i) hist[(size_t)(host_cell(px[i], inv_h) - lo)]++;
  return plan_cuts_hist(hist, lo, nranks, default_ghost_weight(), cuts, err);
}

}  // namespace

// ================================================================== transports
namespace {

#define PBF_NCCL(ctx, expr)                                                                   \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess)                                                                    \
      return fail(ctx, PBF_E_COMM, std::string(#expr) + ": " + ncclGetErrorString(_r));       \
  } while (0)

// One process per GPU: nearest-neighbour ncclSend/ncclRecv, grouped per exchange.
struct NcclTransport : Transport {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int repeat = 1;
  NcclTransport() {
    if (const char* e = std::getenv("PBF_SLAB_EXCHANGE_REPEAT")) repeat = std::max(1, std::atoi(e));
  }
  ~NcclTransport() override {
    if (comm) ncclCommDestroy(comm);
  }
  int exchange(pbf_ctx* ctx, float4* const send[2], float4* const recv[2], size_t bytes) override {
    if (nranks == 1) return PBF_OK;
    for (int rep = 0; rep < repeat; ++rep) {  // repeat > 1: measurement aid (PBF_SLAB_EXCHANGE_REPEAT)
      PBF_NCCL(ctx, ncclGroupStart());
      if (rank > 0) {
        PBF_NCCL(ctx, ncclSend(send[0], bytes, ncclChar, rank - 1, comm, ctx->stream));
        PBF_NCCL(ctx, ncclRecv(recv[0], bytes, ncclChar, rank - 1, comm, ctx->stream));
      }
      if (rank + 1 < nranks) {
        PBF_NCCL(ctx, ncclSend(send[1], bytes, ncclChar, rank + 1, comm, ctx->stream));
        PBF_NCCL(ctx, ncclRecv(recv[1], bytes, ncclChar, rank + 1, comm, ctx->stream));
      }
      PBF_NCCL(ctx, ncclGroupEnd());
    }
    return PBF_OK;
  }
  int reduce_status_device(pbf_ctx* ctx, unsigned int* dev_words, int count) override {
    if (nranks == 1) return PBF_OK;
    PBF_NCCL(ctx, ncclAllReduce(dev_words, dev_words, (size_t)count, ncclUint32, ncclMax, comm, ctx->stream));
    return PBF_OK;
  }
  int reduce_status_host(pbf_ctx*, unsigned int*, int) override { return PBF_OK; }
  int allreduce_host(pbf_ctx* ctx, long long* words, int count, int op) override {
    if (nranks == 1) return PBF_OK;
    PBF_CUDA(ctx, ctx->slab.hist_dev.reserve((size_t)count));
    void* d = ctx->slab.hist_dev.p;
    PBF_CUDA(ctx, cudaMemcpyAsync(d, words, (size_t)count * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    PBF_NCCL(ctx, ncclAllReduce(d, d, (size_t)count, ncclInt64, op == 0 ? ncclMax : ncclSum, comm, ctx->stream));
    PBF_CUDA(ctx, cudaMemcpyAsync(words, d, (size_t)count * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PBF_OK;
  }
  bool capturable() const override { return true; }  // ncclSend/ncclRecv record into a CUDA graph
};

}  // namespace

// Several contexts of ONE process (same or different devices), one host thread per slab inside
// pbf_group_step.  Messages move with cudaMemcpyPeerAsync on the receiver's stream after waiting
// for the sender's "packed" event; a host barrier orders the publication of pointers and events.
struct pbf_group {
  std::vector<pbf_ctx*> ctxs;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t generation = 0;
  bool failed = false;
  std::vector<cudaEvent_t> packed;                 // per rank
  std::vector<float4*> send_ptr[2];                // [side][rank], published per exchange
  std::vector<std::vector<unsigned int>> words;    // status agreement
  std::vector<std::vector<long long>> wide;        // host all-reduces
  std::vector<float4*> window;                     // peer windows (direct-store transport)
  // global ids for upload / download
  size_t n_global = 0;

  // returns false if the group was aborted
  bool barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (failed) return false;
    const uint64_t gen = generation;
    if (++arrived == (int)ctxs.size()) {
      arrived = 0;
      ++generation;
      cv.notify_all();
      return true;
    }
    cv.wait(lk, [&] { return generation != gen || failed; });
    return !failed;
  }
  void abort() {
    std::lock_guard<std::mutex> lk(mu);
    failed = true;
    cv.notify_all();
  }
};

namespace {

struct LocalTransport : Transport {
  pbf_group* group = nullptr;
  int exchange(pbf_ctx* ctx, float4* const send[2], float4* const recv[2], size_t bytes) override {
    pbf_group* g = group;
    const int r = ctx->slab.rank, nr = ctx->slab.nranks;
    if (nr == 1) return PBF_OK;
    PBF_CUDA(ctx, cudaEventRecord(g->packed[r], ctx->stream));
    g->send_ptr[0][r] = send[0];
    g->send_ptr[1][r] = send[1];
    if (!g->barrier()) return fail(ctx, PBF_E_COMM, "slab group aborted by another rank");
    for (int side = 0; side < 2; ++side) {
      const int peer = side == 0 ? r - 1 : r + 1;
      if (peer < 0 || peer >= nr) continue;
      PBF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, g->packed[peer], 0));
      // the peer's message for me is the one it addressed to its other side
      PBF_CUDA(ctx, cudaMemcpyPeerAsync(recv[side], ctx->device, g->send_ptr[1 - side][peer], g->ctxs[peer]->device,
                                        bytes, ctx->stream));
    }
    if (!g->barrier()) return fail(ctx, PBF_E_COMM, "slab group aborted by another rank");
    return PBF_OK;
  }
  int reduce_status_device(pbf_ctx*, unsigned int*, int) override { return PBF_OK; }
  int reduce_status_host(pbf_ctx* ctx, unsigned int* host_words, int count) override {
    pbf_group* g = group;
    const int r = ctx->slab.rank, nr = ctx->slab.nranks;
    if (nr == 1) return PBF_OK;
    g->words[r].assign(host_words, host_words + count);
    if (!g->barrier()) return fail(ctx, PBF_E_COMM, "slab group aborted by another rank");
    for (int p = 0; p < nr; ++p)
      for (int k = 0; k < count; ++k) host_words[k] = std::max(host_words[k], g->words[p][k]);
    if (!g->barrier()) return fail(ctx, PBF_E_COMM, "slab group aborted by another rank");
    return PBF_OK;
  }
  int allreduce_host(pbf_ctx* ctx, long long* words, int count, int op) override {
    pbf_group* g = group;
    const int r = ctx->slab.rank, nr = ctx->slab.nranks;
    if (nr == 1) return PBF_OK;
    g->wide[r].assign(words, words + count);
    if (!g->barrier()) return fail(ctx, PBF_E_COMM, "slab group aborted by another rank");
    for (int k = 0; k < count; ++k) {
      long long v = g->wide[0][k];
      for (int p = 1; p < nr; ++p) v = op == 0 ? std::max(v, g->wide[p][k]) : v + g->wide[p][k];
      words[k] = v;
    }
    if (!g->barrier()) return fail(ctx, PBF_E_COMM, "slab group aborted by another rank");
    return PBF_OK;
  }
  void abort() override { group->abort(); }
};


// ---- direct peer stores + flags ------------------------------------------------------------------
// The fused transport: no message copy and no collective library on the substep
// path.  Every rank owns a "window" — a mailbox and three incoming-message buffers per side — that
// its x-neighbours map (cudaIpc between processes, plain pointers inside one process).  The pack
// kernels of kernels/slab.cu store boundary data STRAIGHT INTO THE NEIGHBOUR'S WINDOW over NVLink;
// an exchange is then one tiny kernel that publishes an epoch in both neighbours' mailboxes and
// spins until both neighbours have published theirs.  Everything is stream-ordered, so the whole
// substep, halo traffic included, replays as one CUDA graph.
//   * buffer of exchange i: i & 1, except the last exchange of a substep with an odd count, which
//     uses buffer 2 — consecutive exchanges never share a buffer, so a sender can never overwrite a
//     message its neighbour has not unpacked yet (the neighbour's signal for exchange i+1 is only
//     sent after it unpacked exchange i);
//   * epochs come from a device-side counter: replaying the graph needs no patched parameters.
constexpr size_t kMailboxElems = 16;  // float4 elements reserved for the mailbox (256 B)

// The exchange as a kernel of its own (in-process groups, where the consumer kernels of several
// slabs share one GPU and must not spin; and PBF_SLAB_FUSED=0).
__global__ void k_peer_signal_wait(PeerMailbox* mine, PeerMailbox* left, PeerMailbox* right, StatusBlock* st,
                                   unsigned long long timeout_ns) {
  pdl_wait();
  peer_signal_wait(mine, left, right, st, timeout_ns);
}

struct PeerTransport : Transport {
  Transport* inner = nullptr;     // owned: status agreement, rendezvous, abort
  pbf_group* group = nullptr;     // in-process rendezvous ...
  NcclTransport* nccl = nullptr;  // ... or IPC handles through the communicator
  float4* win = nullptr;
  size_t elems = 0;               // capacity of one message buffer
  bool dirty = true;              // neighbours do not know the current window yet
  float4* remote[2] = {nullptr, nullptr};
  bool ipc_open[2] = {false, false};
  void* stage = nullptr;          // 3 IPC handles: mine, from left, from right
  // The first batch after enabling runs through the inner (message) transport: it loads every
  // kernel and performs every allocation while no rank is spinning on a flag.  That matters when
  // several slabs share one GPU or one process (virtual ranks): a lazy module load or cudaMalloc
  // on behalf of one slab synchronises the device and would never return while another slab's
  // flag kernel waits for it.
  bool active = false;
  // how long a flag kernel waits for a neighbour before it fails the batch (PBF_SLAB_PEER_TIMEOUT_MS,
  // default 2000 ms; slab_step doubles it for every replay a time-out causes)
  unsigned long long timeout_ns = 2000ull * 1000000ull;
  PeerTransport() {
    if (const char* e = std::getenv("PBF_SLAB_PEER_TIMEOUT_MS")) {
      const long long ms = std::atoll(e);
      if (ms > 0) timeout_ns = (unsigned long long)ms * 1000000ull;
    }
  }
  void relax_timeout() override {
    if (timeout_ns < (1ull << 40)) timeout_ns *= 2;
  }
  // PBF_SLAB_FUSED=1: between processes the exchange runs inside the consumer kernel (pbf::peer_sync;
  // one launch less per exchange, 6 per substep).  MEASURED on 2 x B200 (fluid_million, substeps
  // 50-300, tools/gpu/r02n_fused.sh): 0.601 / 0.912 / 0.713 / 0.661 ms per substep fused against
  // 0.592 / 0.904 / 0.705 / 0.658 with the one-thread flag kernel — the launch it saves is paid back
  // by a whole grid (2 * gcap / 256 blocks) taking tickets and spinning, so it stays OFF.  Never
  // inside one process: several slabs may share a GPU there, and a grid of spinning blocks could
  // keep a neighbour's producer kernel from ever being scheduled.
  bool fuse(pbf_ctx* ctx, PeerSync* sync) override {
    (void)ctx;
    static const bool enabled = [] {
      const char* e = std::getenv("PBF_SLAB_FUSED");
      return e && e[0] == '1';
    }();
    if (!active || group || !enabled) return false;
    sync->mine = reinterpret_cast<PeerMailbox*>(win);
    sync->left = reinterpret_cast<PeerMailbox*>(remote[0]);
    sync->right = reinterpret_cast<PeerMailbox*>(remote[1]);
    sync->arrivals = sync->done = nullptr;  // per consumer kernel, filled by its launcher
    sync->timeout_ns = timeout_ns;
    return true;
  }

  ~PeerTransport() override {
    close_remote();
    if (win) cudaFree(win);
    if (stage) cudaFree(stage);
    delete inner;
  }
  void close_remote() {
    for (int side = 0; side < 2; ++side) {
      if (ipc_open[side] && remote[side]) cudaIpcCloseMemHandle(remote[side]);
      remote[side] = nullptr;
      ipc_open[side] = false;
    }
  }
  static int buffer_of(int index, int count) { return (index == count - 1 && (count & 1)) ? 2 : (index & 1); }
  float4* slot(float4* base, int buf, int side) const { return base + kMailboxElems + ((size_t)buf * 2 + (size_t)side) * elems; }

  int prepare(pbf_ctx* ctx, size_t msg_elems) override {
    const int r = ctx->slab.rank, nr = ctx->slab.nranks;
    int rc = inner->prepare(ctx, msg_elems);
    if (rc != PBF_OK) return rc;
    rc = prepare_window(ctx, msg_elems, r, nr);
    if (rc != PBF_OK) return rc;
    active = ctx->slab.warm;
    // all host-side preparation of every slab is done before any slab may start spinning
    if (group && !group->barrier()) return fail(ctx, PBF_E_COMM, "slab group aborted by another rank");
    return PBF_OK;
  }

  int prepare_window(pbf_ctx* ctx, size_t msg_elems, int r, int nr) {
    if (!win || msg_elems > elems) {
      // every rank grows at the same time: capacities are decided on max-reduced statistics
      PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      close_remote();
      if (win) cudaFree(win);
      win = nullptr;
      const size_t total = kMailboxElems + 6 * msg_elems;
      PBF_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&win), total * sizeof(float4)));
      PBF_CUDA(ctx, cudaMemset(win, 0, total * sizeof(float4)));
      elems = msg_elems;
      dirty = true;
      invalidate_graph(ctx);
      cudaFuncAttributes attr;  // forces the (lazily loaded) flag kernel into the context now
      PBF_CUDA(ctx, cudaFuncGetAttributes(&attr, k_peer_signal_wait));
    }
    if (!dirty) return PBF_OK;
    if (group) {
      group->window[(size_t)r] = win;
      if (!group->barrier()) return fail(ctx, PBF_E_COMM, "slab group aborted by another rank");
      remote[0] = r > 0 ? group->window[(size_t)r - 1] : nullptr;
      remote[1] = r + 1 < nr ? group->window[(size_t)r + 1] : nullptr;
      if (!group->barrier()) return fail(ctx, PBF_E_COMM, "slab group aborted by another rank");
    } else if (nccl && nr > 1) {
      cudaIpcMemHandle_t h[3];
      std::memset(h, 0, sizeof(h));
      PBF_CUDA(ctx, cudaIpcGetMemHandle(&h[0], win));
      if (!stage) PBF_CUDA(ctx, cudaMalloc(&stage, sizeof(h)));
      PBF_CUDA(ctx, cudaMemcpyAsync(stage, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
      char* d = static_cast<char*>(stage);
      const size_t hb = sizeof(cudaIpcMemHandle_t);
      PBF_NCCL(ctx, ncclGroupStart());
      if (r > 0) {
        PBF_NCCL(ctx, ncclSend(d, hb, ncclChar, r - 1, nccl->comm, ctx->stream));
        PBF_NCCL(ctx, ncclRecv(d + hb, hb, ncclChar, r - 1, nccl->comm, ctx->stream));
      }
      if (r + 1 < nr) {
        PBF_NCCL(ctx, ncclSend(d, hb, ncclChar, r + 1, nccl->comm, ctx->stream));
        PBF_NCCL(ctx, ncclRecv(d + 2 * hb, hb, ncclChar, r + 1, nccl->comm, ctx->stream));
      }
      PBF_NCCL(ctx, ncclGroupEnd());
      PBF_CUDA(ctx, cudaMemcpyAsync(h, stage, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
      PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      for (int side = 0; side < 2; ++side) {
        const bool has = side == 0 ? r > 0 : r + 1 < nr;
        if (!has) continue;
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h[1 + side], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
          return fail(ctx, PBF_E_COMM, std::string("cudaIpcOpenMemHandle (peer window): ") + cudaGetErrorString(e));
        remote[side] = static_cast<float4*>(ptr);
        ipc_open[side] = true;
      }
    }
    dirty = false;
    return PBF_OK;
  }

  bool bind(pbf_ctx* ctx, int index, int count, float4* send[2], float4* recv[2]) override {
    if (!active) return inner->bind(ctx, index, count, send, recv);
    const int buf = buffer_of(index, count);
    // my message to the left neighbour is what it receives "from the right" (side 1), and vice versa;
    // without a neighbour the pack kernels write into a local scratch buffer
    send[0] = remote[0] ? slot(remote[0], buf, 1) : ctx->slab.send[0][0].p;
    send[1] = remote[1] ? slot(remote[1], buf, 0) : ctx->slab.send[0][1].p;
    recv[0] = slot(win, buf, 0);
    recv[1] = slot(win, buf, 1);
    return true;
  }

  int exchange(pbf_ctx* ctx, float4* const send[2], float4* const recv[2], size_t bytes) override {
    if (!active) return inner->exchange(ctx, send, recv, bytes);
    PeerMailbox* mine = reinterpret_cast<PeerMailbox*>(win);
    PBF_LAUNCH(k_peer_signal_wait, 1, 1, ctx->stream, mine, reinterpret_cast<PeerMailbox*>(remote[0]),
                                                 reinterpret_cast<PeerMailbox*>(remote[1]), ctx->status.p,
                                                 timeout_ns);
    return PBF_OK;
  }
  int reduce_status_device(pbf_ctx* ctx, unsigned int* w, int n) override { return inner->reduce_status_device(ctx, w, n); }
  int reduce_status_host(pbf_ctx* ctx, unsigned int* w, int n) override { return inner->reduce_status_host(ctx, w, n); }
  int allreduce_host(pbf_ctx* ctx, long long* w, int n, int op) override { return inner->allreduce_host(ctx, w, n, op); }
  void abort() override { inner->abort(); }
  // Between processes the substep (flag kernels included) replays as a CUDA graph.  Inside one
  // process graph instantiation by one slab could stall behind another slab's spinning kernel.
  bool capturable() const override { return group == nullptr && (active || inner->capturable()); }
};

// ---- buffers -----------------------------------------------------------------------------------
int ensure_slab_buffers(pbf_ctx* ctx) {
  SlabState& sl = ctx->slab;
  if (!sl.counts.p) {
    PBF_CUDA(ctx, sl.counts.reserve(1));
    PBF_CUDA(ctx, cudaMemset(sl.counts.p, 0, sizeof(SlabCounts)));
    PBF_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&sl.counts_host), sizeof(SlabCounts)));
    std::memset(sl.counts_host, 0, sizeof(SlabCounts));
    PBF_CUDA(ctx, cudaStreamCreateWithFlags(&sl.side, cudaStreamNonBlocking));
    PBF_CUDA(ctx, cudaEventCreateWithFlags(&sl.ev_fork, cudaEventDisableTiming));
    PBF_CUDA(ctx, cudaEventCreateWithFlags(&sl.ev_join, cudaEventDisableTiming));
    if (const char* e = std::getenv("PBF_SLAB_OVERLAP")) sl.overlap = e[0] != '0';
  }
  PBF_CUDA(ctx, sl.holes.reserve(6 * (size_t)sl.mcap + 64));
  const size_t elems = 1 + 2 * (size_t)std::max(sl.mcap, sl.gcap);
  if (elems > sl.msg_elems) {
    for (int p = 0; p < 2; ++p)
      for (int side = 0; side < 2; ++side) {
        PBF_CUDA(ctx, sl.send[p][side].reserve(elems));
        PBF_CUDA(ctx, cudaMemset(sl.send[p][side].p, 0, elems * sizeof(float4)));
      }
    for (int side = 0; side < 2; ++side) {
      PBF_CUDA(ctx, sl.recv[side].reserve(elems));
      // a missing neighbour never writes its message: the zero header means "0 particles"
      PBF_CUDA(ctx, cudaMemset(sl.recv[side].p, 0, elems * sizeof(float4)));
    }
    sl.msg_elems = elems;
  }
  return PBF_OK;
}

void slab_fill(pbf_ctx* ctx, SlabBuffers& sb) {
  SlabState& sl = ctx->slab;
  sb.counts = sl.counts.p;
  sb.status = ctx->status.p;
  sb.gid_o = sl.gid_o.p;
  sb.holes = sl.holes.p;
  sb.send[0] = sb.send[1] = sb.recv[0] = sb.recv[1] = nullptr;  // slab_bind, per exchange
  sb.sync = PeerSync{nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  sb.cut_lo = sl.cut_lo;
  sb.cut_hi = sl.cut_hi;
  sb.cap = (int)sl.launch_own;                          // what the kernels cover, not what is allocated
  sb.tot_cap = (int)(sl.launch_own + sl.launch_ghost);
  sb.mcap = sl.mcap;
  sb.gcap = sl.gcap;
}

// Message buffers of exchange `index` (of `count` per substep), to be called before the kernels
// that pack it.  Message transports use the context's buffers: the send pair alternates between
// consecutive exchanges (a neighbour may still be copying the previous message out).
void slab_bind(pbf_ctx* ctx, SlabBuffers& sb, int index, int count) {
  SlabState& sl = ctx->slab;
  if (sl.transport->bind(ctx, index, count, sb.send, sb.recv)) return;
  sb.send[0] = sl.send[sl.parity][0].p;
  sb.send[1] = sl.send[sl.parity][1].p;
  sb.recv[0] = sl.recv[0].p;
  sb.recv[1] = sl.recv[1].p;
}

// One exchange with both neighbours.
int slab_exchange(pbf_ctx* ctx, SlabBuffers& sb, size_t elems) {
  SlabState& sl = ctx->slab;
  // fused: the kernel that consumes the messages (launched next, on this stream) runs the exchange
  sb.sync = PeerSync{nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  const bool side_stream = ctx->stream == sl.side;  // PBF_SLAB_OVERLAP: the consumer runs on another stream
  if (side_stream || !sl.transport->fuse(ctx, &sb.sync)) {
    sb.sync = PeerSync{nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    stage_mark(ctx, PBF_STAGE_EXCHANGE, 1);
    const int rc = sl.transport->exchange(ctx, sb.send, sb.recv, elems * sizeof(float4));
    stage_mark(ctx, PBF_STAGE_EXCHANGE, 0);
    if (rc != PBF_OK) return rc;
  }
  sl.exchanges++;
  sl.bytes_sent += ((sl.rank > 0) + (sl.rank + 1 < sl.nranks)) * elems * sizeof(float4);
  sl.parity ^= 1;
  return PBF_OK;
}

// ---- one substep of one slab ---------------------------------------------------------------------
// Same stages as enqueue_substep (pbf_capi.cu) with migration before the grid build, the ghost
// build after it and one halo refresh per solver iteration.  Returns launches (< 0: error code).
int slab_substep(pbf_ctx* ctx) {
  SlabState& sl = ctx->slab;
  cudaStream_t s = ctx->stream;
  const StepConsts& c = ctx->consts;
  const bool strict = ctx->mode == PBF_MODE_STRICT;
  StageTimer& t = ctx->timer;
  GridBuffers g{};
  fill_grid_buffers(ctx, g);
  SolveBuffers b{};
  fill_solve_buffers(ctx, b);
  SlabBuffers sb{};
  slab_fill(ctx, sb);
  NeighborList nl{ctx->nbr_idx.p, ctx->nbr_count.p, ctx->K};
  const NRef n_own = nref((int)sl.launch_own, &sl.counts.p->n_own);
  const NRef n_tot = nref((int)(sl.launch_own + sl.launch_ghost), &sl.counts.p->n_tot);
  int launches = 0, k, rc;
  // A replayed graph bakes the send-buffer pointers in: every substep must start on the same pair.
  // (Stream-ordered transports have no write-after-read hazard on the send buffers.)
  if (sl.transport->capturable()) sl.parity = 0;
  const int iters = ctx->params.solver_iterations;
  const bool tail_xsph = c.do_xsph != 0, tail_vort = c.do_vort != 0;
  const bool final_in_delta = !tail_xsph && !tail_vort;
  // exchanges of this substep: migration hops, ghost build, pred refreshes, post-XSPH velocities
  const int n_ex = sl.hops + 1 + (iters > 0 ? (final_in_delta ? iters - 1 : iters) : 0) +
                   ((iters > 0 && tail_xsph && tail_vort) ? 1 : 0);
  int xi = 0;

  stage_mark(ctx, PBF_STAGE_PREDICT, 1);
  k = launch_predict(ctx->pos_o.p, ctx->vel_o.p, ctx->pred_o.p, c, g, n_own, true, s);
  stage_mark(ctx, PBF_STAGE_PREDICT, 0);
  t.launches[PBF_STAGE_PREDICT] += k; launches += k;

  // migration: particles whose predicted x-cell left the slab move to the neighbour (hops > 1 only
  // after a batch found a particle more than one slab away)
  for (int hop = 0; hop < sl.hops; ++hop) {
    slab_bind(ctx, sb, xi++, n_ex);
    k = launch_slab_split(ctx->pos_o.p, ctx->pred_o.p, sb, c, s);
    if ((rc = slab_exchange(ctx, sb, 1 + 2 * (size_t)sl.mcap)) != PBF_OK) return rc;
    stage_mark(ctx, PBF_STAGE_EXCHANGE, 1);  // with a fused exchange the waiting happens in here
    k += launch_slab_merge(ctx->pos_o.p, ctx->pred_o.p, sb, c, hop == sl.hops - 1, s);
    stage_mark(ctx, PBF_STAGE_EXCHANGE, 0);
    t.launches[PBF_STAGE_EXCHANGE] += k; launches += k;
  }
  launches += launch_grid_finalize(g, 2, n_own, s);
  t.launches[PBF_STAGE_PREDICT] += 1;

  int out = 0;
  stage_mark(ctx, PBF_STAGE_SORT, 1);
  k = launch_sort(ctx->pred_o.p, c, g, n_own, &out, s);
  stage_mark(ctx, PBF_STAGE_SORT, 0);
  t.launches[PBF_STAGE_SORT] += k; launches += k;
  ctx->sorted_buf = out;

  stage_mark(ctx, PBF_STAGE_CELLS, 1);
  k = launch_cells_reorder(ctx->pred_o.p, ctx->pos_o.p, ctx->pred_a.p, ctx->pos_s.p, sl.gid_o.p, g, n_own, s);
  stage_mark(ctx, PBF_STAGE_CELLS, 0);
  t.launches[PBF_STAGE_CELLS] += k; launches += k;

  // ghost build: two boundary layers per side, appended after the owned slots
  slab_bind(ctx, sb, xi++, n_ex);
  k = launch_slab_ghost_pack(g.keys[out], ctx->pred_a.p, ctx->pos_s.p, g, sb, s);
  if ((rc = slab_exchange(ctx, sb, 1 + 2 * (size_t)sl.gcap)) != PBF_OK) return rc;
  stage_mark(ctx, PBF_STAGE_EXCHANGE, 1);
  k += launch_slab_ghost_unpack(ctx->pred_a.p, ctx->pos_s.p, g, sb, c, s);
  stage_mark(ctx, PBF_STAGE_EXCHANGE, 0);
  t.launches[PBF_STAGE_EXCHANGE] += k; launches += k;

  stage_mark(ctx, PBF_STAGE_NEIGHBORS, 1);
  k = launch_neighbors(ctx->pred_a.p, c, g, nl, n_tot, s);
  stage_mark(ctx, PBF_STAGE_NEIGHBORS, 0);
  t.launches[PBF_STAGE_NEIGHBORS] += k; launches += k;

  if (iters <= 0) {
    launches += launch_commit_only(b, c, n_own, strict, s);
    t.launches[PBF_STAGE_FINALIZE] += 1;
    return launches;
  }
  // Only the first cell layer next to a cut reads ghosts.  Optionally (PBF_SLAB_OVERLAP=1) the
  // lambda pass after iteration 0 is split: its interior part starts right after the delta pass
  // while the halo of that pass is exchanged, unpacked and followed by the boundary part on a side
  // stream.  MEASURED on B200: no gain at 2 slabs (1.671 vs 1.666 ms at 2 M particles per GPU) and
  // a loss at 8 (2.49 vs 2.00 ms: the boundary pass with ~0.5 M ghosts competes with the interior
  // pass for the same SMs and the fork/join adds two graph edges per iteration), so it is off.
  const bool overlap = sl.overlap && sl.nranks > 1 && !ctx->profile;
  const NRef n_interior = nref((int)sl.launch_own, &sl.counts.p->n_own);
  const NRef n_boundary = nref(4 * sl.gcap, &sl.counts.p->n_tot);
  int cur = 0;
  stage_mark(ctx, PBF_STAGE_LAMBDA, 1);
  launches += launch_lambda(b, nl, c, cur, n_tot, strict, s);  // owned + first-layer ghosts
  stage_mark(ctx, PBF_STAGE_LAMBDA, 0);
  t.launches[PBF_STAGE_LAMBDA] += 1;
  for (int it = 0; it < iters; ++it) {
    const bool last = it == iters - 1;
    // the delta pass stores its two boundary layers straight into the outgoing messages (no pack
    // kernel); after the very last pass nothing reads the ghosts any more
    const bool refresh = !(last && final_in_delta);
    b.halo = HaloOut{nullptr, {nullptr, nullptr}};
    if (refresh) {
      slab_bind(ctx, sb, xi++, n_ex);
      b.halo = HaloOut{sb.counts, {sb.send[0], sb.send[1]}};
    }
    stage_mark(ctx, PBF_STAGE_DELTA, 1);
    launches += launch_delta(b, nl, c, cur, last, last && final_in_delta, n_own, strict, s);
    stage_mark(ctx, PBF_STAGE_DELTA, 0);
    b.halo = HaloOut{nullptr, {nullptr, nullptr}};
    t.launches[PBF_STAGE_DELTA] += 1;
    cur ^= 1;
    if (!refresh) break;
    if (last || !overlap) {
      if ((rc = slab_exchange(ctx, sb, (size_t)sl.gcap)) != PBF_OK) return rc;
      stage_mark(ctx, PBF_STAGE_EXCHANGE, 1);
      k = launch_slab_halo_unpack(b.pred[cur], sb, s);
      stage_mark(ctx, PBF_STAGE_EXCHANGE, 0);
      t.launches[PBF_STAGE_EXCHANGE] += k; launches += k;
      if (!last) {
        stage_mark(ctx, PBF_STAGE_LAMBDA, 1);
        launches += launch_lambda(b, nl, c, cur, n_tot, strict, s);
        stage_mark(ctx, PBF_STAGE_LAMBDA, 0);
        t.launches[PBF_STAGE_LAMBDA] += 1;
      }
      continue;
    }
    // fork: [side] exchange -> unpack -> lambda(boundary + ghosts)   ||   [main] lambda(interior)
    if (cudaEventRecord(sl.ev_fork, s) != cudaSuccess || cudaStreamWaitEvent(sl.side, sl.ev_fork, 0) != cudaSuccess)
      return fail(ctx, PBF_E_CUDA, "slab substep: fork onto the halo stream failed");
    ctx->stream = sl.side;  // the transport enqueues on the context's stream
    rc = slab_exchange(ctx, sb, (size_t)sl.gcap);
    ctx->stream = s;
    if (rc != PBF_OK) return rc;
    k = launch_slab_halo_unpack(b.pred[cur], sb, sl.side);
    k += launch_lambda(b, nl, c, cur, n_boundary, strict, sl.side, Span{sl.counts.p, 1});
    k += launch_lambda(b, nl, c, cur, n_interior, strict, s, Span{sl.counts.p, 0});
    if (cudaEventRecord(sl.ev_join, sl.side) != cudaSuccess || cudaStreamWaitEvent(s, sl.ev_join, 0) != cudaSuccess)
      return fail(ctx, PBF_E_CUDA, "slab substep: join of the halo stream failed");
    t.launches[PBF_STAGE_LAMBDA] += 2;
    t.launches[PBF_STAGE_EXCHANGE] += 1;
    launches += k;
  }
  if (final_in_delta) return launches;

  float4* pos = b.pred[cur];
  launches += launch_slab_ghost_vel(pos, b.pos_s, b.rho, b.vel[0], xsph_record(b, c), sb, c, strict, s);
  t.launches[PBF_STAGE_EXCHANGE] += 1;
  int vcur = 0;
  if (tail_xsph) {
    if (tail_vort) {  // omega of the first-layer ghosts needs the post-XSPH velocity of both layers
      slab_bind(ctx, sb, xi++, n_ex);
      b.halo = HaloOut{sb.counts, {sb.send[0], sb.send[1]}};
    }
    stage_mark(ctx, PBF_STAGE_XSPH, 1);
    launches += launch_xsph(b, nl, c, pos, !tail_vort, n_own, strict, s);
    stage_mark(ctx, PBF_STAGE_XSPH, 0);
    b.halo = HaloOut{nullptr, {nullptr, nullptr}};
    t.launches[PBF_STAGE_XSPH] += 1;
    vcur = 1;
    if (tail_vort) {
      k = 0;
      if ((rc = slab_exchange(ctx, sb, (size_t)sl.gcap)) != PBF_OK) return rc;
      stage_mark(ctx, PBF_STAGE_EXCHANGE, 1);
      k += launch_slab_halo_unpack(b.vel[1], sb, s);
      stage_mark(ctx, PBF_STAGE_EXCHANGE, 0);
      t.launches[PBF_STAGE_EXCHANGE] += k; launches += k;
    }
  }
  if (tail_vort) {
    stage_mark(ctx, PBF_STAGE_VORT_OMEGA, 1);
    launches += launch_vort_omega(b, nl, c, pos, vcur, n_tot, strict, s);
    stage_mark(ctx, PBF_STAGE_VORT_OMEGA, 0);
    stage_mark(ctx, PBF_STAGE_VORT_APPLY, 1);
    launches += launch_vort_apply(b, nl, c, pos, vcur, n_own, strict, s);
    stage_mark(ctx, PBF_STAGE_VORT_APPLY, 0);
    t.launches[PBF_STAGE_VORT_OMEGA] += 1;
    t.launches[PBF_STAGE_VORT_APPLY] += 1;
  }
  return launches;
}

int set_device_count(pbf_ctx* ctx, int n_own) {
  SlabState& sl = ctx->slab;
  std::memset(sl.counts_host, 0, sizeof(SlabCounts));
  sl.counts_host->n_own = n_own;
  sl.counts_host->n_tot = n_own;
  sl.counts_host->cut_lo = sl.cut_lo;  // read by the migration / ghost kernels from device memory
  sl.counts_host->cut_hi = sl.cut_hi;
  PBF_CUDA(ctx, cudaMemcpyAsync(sl.counts.p, sl.counts_host, sizeof(SlabCounts), cudaMemcpyHostToDevice, ctx->stream));
  return PBF_OK;
}

}  // namespace

namespace {

// Substeps per slab batch (see slab_step): PBF_SLAB_CHUNK, default 25.
int slab_chunk() {
  static const int chunk = [] {
    const char* e = std::getenv("PBF_SLAB_CHUNK");
    const int v = e ? std::atoi(e) : 0;
    return v > 0 ? v : 25;
  }();
  return chunk;
}

// Automatic re-balancing.  Every rank sees every slab's owned count in the max-reduced status
// block; when the largest slab exceeds the mean by the threshold, all ranks together build the
// x-layer histogram of the whole scene (one sum all-reduce), plan equal-count cuts on it with the
// planner the upload uses, and install them.  Misplaced particles migrate during the next substep
// (hop count and message capacities grow on demand), so results do not depend on when this happens.
int maybe_rebalance(pbf_ctx* ctx, const StatusBlock& st) {
  SlabState& sl = ctx->slab;
  if (!(sl.rebalance_threshold > 0.0f) || sl.nranks < 2 || sl.nranks > kMaxSlabs) return PBF_OK;
  unsigned long long total = 0, largest = 0;
  for (int r = 0; r < sl.nranks; ++r) {
    total += st.own_by_rank[r];
    largest = std::max<unsigned long long>(largest, st.own_by_rank[r]);
  }
  // planned_ratio: max / mean owned the last plan itself produced (> 1 when cell layers are coarse
  // or the ghost weight gives the end slabs more) — only an imbalance beyond that is worth a re-plan
  if (total == 0 || (double)largest * sl.nranks <= (double)sl.rebalance_threshold * sl.planned_ratio * (double)total)
    return PBF_OK;
  SlabBuffers sb{};
  slab_fill(ctx, sb);
  // global x-cell range of the owned particles
  PBF_CUDA(ctx, sl.hist_dev.reserve(2));
  int* range_dev = reinterpret_cast<int*>(sl.hist_dev.p);
  const int init[2] = {INT_MAX, INT_MIN};
  PBF_CUDA(ctx, cudaMemcpyAsync(range_dev, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  // plan for the later part of the next batch (imbalance grows along it): positions advanced by 3/4
  // of a batch at constant velocity
  const float lookahead = 0.75f * (float)slab_chunk() * ctx->params.dt;
  ctx->launch_count += (uint64_t)launch_slab_xrange(ctx->pos_o.p, ctx->vel_o.p, sb, ctx->consts, lookahead, range_dev, ctx->stream);
  int range[2];
  PBF_CUDA(ctx, cudaMemcpyAsync(range, range_dev, sizeof(range), cudaMemcpyDeviceToHost, ctx->stream));
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  long long mm[2] = {ctx->n ? -(long long)range[0] : LLONG_MIN, ctx->n ? (long long)range[1] : LLONG_MIN};
  int rc = sl.transport->allreduce_host(ctx, mm, 2, 0);
  if (rc != PBF_OK) return rc;
  const long long x_min = -mm[0], x_max = mm[1];
  const long long layers = x_max - x_min + 1;
  if (layers < 2LL * sl.nranks || layers > (1LL << 16) || x_min < INT_MIN / 2 || x_max > INT_MAX / 2) return PBF_OK;
  // histogram of the whole scene
  PBF_CUDA(ctx, sl.hist_dev.reserve((size_t)layers));
  PBF_CUDA(ctx, cudaMemsetAsync(sl.hist_dev.p, 0, (size_t)layers * sizeof(unsigned long long), ctx->stream));
  ctx->launch_count += (uint64_t)launch_slab_xhist(ctx->pos_o.p, ctx->vel_o.p, sb, ctx->consts, lookahead, (int)x_min,
                                                   (int)layers, sl.hist_dev.p, ctx->stream);
  std::vector<long long> hist((size_t)layers);
  PBF_CUDA(ctx, cudaMemcpyAsync(hist.data(), sl.hist_dev.p, (size_t)layers * sizeof(long long), cudaMemcpyDeviceToHost,
                                ctx->stream));
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if ((rc = sl.transport->allreduce_host(ctx, hist.data(), (int)layers, 1)) != PBF_OK) return rc;
  std::vector<size_t> h((size_t)layers);
  for (size_t l = 0; l < h.size(); ++l) h[l] = (size_t)hist[l];
  std::vector<int> cuts;
  std::string err;
  if (plan_cuts_hist(h, (int)x_min, sl.nranks, default_ghost_weight(), cuts, err) != PBF_OK) return PBF_OK;  // can not do better: keep the cuts
  unsigned long long planned_max = 0;
  for (int r = 0; r < sl.nranks; ++r) {
    const long long a = r == 0 ? 0 : (long long)cuts[(size_t)r] - x_min;
    const long long b = r == sl.nranks - 1 ? layers : (long long)cuts[(size_t)r + 1] - x_min;
    unsigned long long own = 0;
    for (long long l = a; l < b; ++l) own += h[(size_t)l];
    planned_max = std::max(planned_max, own);
  }
  sl.planned_ratio = std::max(1.0, (double)planned_max * sl.nranks / (double)total);
  if (std::getenv("PBF_SLAB_DEBUG") && sl.rank == 0) {
    std::fprintf(stderr, "[replan] x layers %lld..%lld total %llu largest %llu planned_max %llu cuts:", x_min, x_max, total, largest,
                 planned_max);
    for (int r = 1; r < sl.nranks; ++r) std::fprintf(stderr, " %d", cuts[(size_t)r]);
    std::fprintf(stderr, " | own_by_rank:");
    for (int r = 0; r < sl.nranks; ++r) std::fprintf(stderr, " %u", st.own_by_rank[r]);
    std::fprintf(stderr, " | hist:");
    for (size_t l = 0; l < h.size(); ++l) std::fprintf(stderr, " %zu", h[l]);
    std::fprintf(stderr, "\n");
  }
  sl.rebalances++;  // the decision is global: every rank counts it, whether or not its own cuts move
  // The particles between an old and a new cut change owner during the next substep, whole cell
  // layers at once: size the migration messages for that now (identically on every rank: the largest
  // move of any cut), instead of letting the next batch overflow, grow and replay.
  {
    auto layer_sum = [&](long long a, long long b) {  // particles in the layers [a, b) of the histogram
      unsigned long long sum = 0;
      for (long long l = std::max(0LL, a); l < std::min(layers, b); ++l) sum += h[(size_t)l];
      return sum;
    };
    unsigned long long moved = 0;
    if (sl.rank > 0 && sl.cut_lo != INT_MIN) {
      const long long o = (long long)sl.cut_lo - x_min, n = (long long)cuts[(size_t)sl.rank] - x_min;
      moved = std::max(moved, layer_sum(std::min(o, n), std::max(o, n)));
    }
    if (sl.rank + 1 < sl.nranks && sl.cut_hi != INT_MAX) {
      const long long o = (long long)sl.cut_hi - x_min, n = (long long)cuts[(size_t)sl.rank + 1] - x_min;
      moved = std::max(moved, layer_sum(std::min(o, n), std::max(o, n)));
    }
    long long m = (long long)moved;
    if ((rc = sl.transport->allreduce_host(ctx, &m, 1, 0)) != PBF_OK) return rc;
    const long long want = m + m / 4 + 1024;
    if (want > (long long)sl.mcap && want < (1LL << 30)) {
      sl.mcap = (int)want;
      invalidate_graph(ctx);  // message capacities are kernel parameters and buffer sizes
    }
  }
  if (cuts[(size_t)sl.rank] == sl.cut_lo && cuts[(size_t)sl.rank + 1] == sl.cut_hi) return PBF_OK;
  sl.cut_lo = cuts[(size_t)sl.rank];
  sl.cut_hi = cuts[(size_t)sl.rank + 1];  // device data (SlabCounts): the captured substep graph stays valid
  return PBF_OK;
}

}  // namespace

namespace pbf {

void slab_release(pbf_ctx* ctx) {
  SlabState& sl = ctx->slab;
  if (sl.owns_transport && sl.transport) delete sl.transport;
  sl.transport = nullptr;
  sl.counts.release(); sl.gid_o.release(); sl.gid_bak.release(); sl.holes.release(); sl.hist_dev.release();
  for (int p = 0; p < 2; ++p)
    for (int side = 0; side < 2; ++side) sl.send[p][side].release();
  sl.recv[0].release(); sl.recv[1].release();
  if (sl.counts_host) cudaFreeHost(sl.counts_host);
  sl.counts_host = nullptr;
  if (sl.side) cudaStreamDestroy(sl.side);
  if (sl.ev_fork) cudaEventDestroy(sl.ev_fork);
  if (sl.ev_join) cudaEventDestroy(sl.ev_join);
  sl.side = nullptr;
  sl.ev_fork = sl.ev_join = nullptr;
}

// pbf_step of a slab context.  Every rank of the communicator (or every thread of the group)
// calls it with the same nsteps; the grow-and-replay decision is taken on max-reduced status
// words, so all ranks replay together.
static int slab_batch(pbf_ctx* ctx, int nsteps);

// A long pbf_step call is cut into batches of at most `chunk` substeps (PBF_SLAB_CHUNK, default 25):
// table capacities and the cuts are only re-decided BETWEEN batches, and a fluid that is spreading
// (fluid_million after its floor impact: the end slabs go from 130 k to 300 k particles within 100
// substeps) must not run that long on stale cuts — nor replay 100 substeps because a message buffer
// overflowed in the 90th.  Every rank makes the same cuts of the same nsteps, so the call stays
// collective.  Cost: one status reduction and host round trip per batch (~0.1 ms per 25 substeps).
// On failure the state is that of the last completed batch (time advanced accordingly).
int slab_step(pbf_ctx* ctx, int nsteps) {
  const int chunk = slab_chunk();
  for (int done = 0; done < nsteps;) {
    const int k = std::min(chunk, nsteps - done);
    const int rc = slab_batch(ctx, k);
    if (rc != PBF_OK) return rc;
    done += k;
  }
  return PBF_OK;
}

static int slab_batch(pbf_ctx* ctx, int nsteps) {
  SlabState& sl = ctx->slab;
  if (!sl.transport) return fail(ctx, PBF_E_COMM, "pbf_step: slab context without a communicator");
  int rc;
  const size_t n0 = ctx->n;
  sl.n_bak = n0;
  if (n0) {
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->pos_bak.p, ctx->pos_o.p, n0 * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->vel_bak.p, ctx->vel_o.p, n0 * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    PBF_CUDA(ctx, cudaMemcpyAsync(sl.gid_bak.p, sl.gid_o.p, n0 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  // State of the start of the batch back into place (device arrays and the owned count).
  auto restore = [&]() -> cudaError_t {
    cudaError_t e = cudaSuccess;
    if (n0) {
      auto cp = [&](void* d, const void* s, size_t bytes) {
        const cudaError_t r = cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = r;
      };
      cp(ctx->pos_o.p, ctx->pos_bak.p, n0 * sizeof(float4));
      cp(ctx->vel_o.p, ctx->vel_bak.p, n0 * sizeof(float4));
      cp(sl.gid_o.p, sl.gid_bak.p, n0 * sizeof(uint32_t));
    }
    ctx->n = n0;
    return e;
  };
  // Every failure after the batch has started leaves through here: the context keeps the state it
  // had before the call (like the single-GPU path) and the other ranks are released from whatever
  // exchange or reduction they are waiting in.
  auto bail = [&](int code) -> int {
    invalidate_graph(ctx);
    ctx->tables_dirty = true;
    restore();
    cudaStreamSynchronize(ctx->stream);
    sl.transport->abort();
    return code;
  };
  int timeouts = 0;
  for (int attempt = 0; attempt < 32; ++attempt) {
    if ((rc = ensure_slab_buffers(ctx)) != PBF_OK) return bail(rc);
    if ((rc = ensure_tables(ctx)) != PBF_OK) return bail(rc);
    {
      // launch bounds (rank-local: they are not part of the wire format).  Owned: half above the
      // current count (fluid_million's splash grows the end slabs of 8 by 40 % within 25 substeps;
      // outgrowing the bound replays the batch); ghosts: everything until a batch has shown how
      // many there are.
      const size_t need_own = std::min(ctx->cap, n0 + n0 / 2 + 16384);
      size_t own = sl.launch_own;
      // Re-decided with hysteresis — every change re-captures the substep graph (~2 ms): up when less
      // than an eighth of head-room is left, down only between batches and only by a factor of three
      // (a bound that was just grown because owned + ghosts overflowed it must survive the replay).
      const size_t low_water = std::min(ctx->cap, n0 + n0 / 8 + 4096);
      if (own < low_water || own > ctx->cap || (attempt == 0 && own > 3 * need_own)) own = need_own;
      size_t ghost = sl.launch_ghost;
      if (ghost == 0 || ghost > 2 * (size_t)sl.gcap) ghost = 2 * (size_t)sl.gcap;
      if (own != sl.launch_own || ghost != sl.launch_ghost) {
        sl.launch_own = own;
        sl.launch_ghost = ghost;
        invalidate_graph(ctx);
      }
    }
    if ((rc = sl.transport->prepare(ctx, sl.msg_elems)) != PBF_OK) return bail(rc);  // last: may end in a barrier
    if ((rc = reset_status(ctx)) != PBF_OK) return bail(rc);
    if ((rc = set_device_count(ctx, (int)n0)) != PBF_OK) return bail(rc);
    // Stream-ordered transports (NCCL) let the whole substep, exchanges included, replay as one
    // CUDA graph.  The first batch after joining the communicator runs un-captured so that NCCL
    // can set up its peer connections outside a capture.
    const bool graph = ctx->use_graph && !ctx->profile && sl.transport->capturable() && sl.warm;
    for (int sidx = 0; sidx < nsteps; ++sidx) {
      if (graph) {
        if (!ctx->graph_exec) {
          cudaGraph_t gr = nullptr;
          const uint64_t ex0 = sl.exchanges, by0 = sl.bytes_sent;
          uint64_t saved[PBF_STAGE_COUNT];
          std::memcpy(saved, ctx->timer.launches, sizeof(saved));
          PBF_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
          const int k = slab_substep(ctx);
          const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &gr);
          std::memcpy(ctx->timer.launches, saved, sizeof(saved));
          if (k < 0) {
            if (gr) cudaGraphDestroy(gr);
            return bail(k);
          }
          if (ce != cudaSuccess) return bail(fail(ctx, PBF_E_CUDA, std::string("slab batch: graph capture: ") + cudaGetErrorString(ce)));
          ctx->graph_kernels = k;
          sl.graph_exchanges = sl.exchanges - ex0;
          sl.graph_bytes = sl.bytes_sent - by0;
          sl.exchanges = ex0;
          sl.bytes_sent = by0;
          const cudaError_t ie = cudaGraphInstantiate(&ctx->graph_exec, gr, 0);
          cudaGraphDestroy(gr);
          if (ie != cudaSuccess) return bail(fail(ctx, PBF_E_CUDA, std::string("slab batch: graph instantiation: ") + cudaGetErrorString(ie)));
        }
        const cudaError_t le = cudaGraphLaunch(ctx->graph_exec, ctx->stream);
        if (le != cudaSuccess) return bail(fail(ctx, PBF_E_CUDA, std::string("slab batch: graph launch: ") + cudaGetErrorString(le)));
        ctx->launch_count += (uint64_t)ctx->graph_kernels;
        sl.exchanges += sl.graph_exchanges;
        sl.bytes_sent += sl.graph_bytes;
      } else {
        const int k = slab_substep(ctx);
        if (k < 0) return bail(k);
        ctx->launch_count += (uint64_t)k;
      }
    }
    {
      SlabBuffers sb{};
      slab_fill(ctx, sb);
      ctx->launch_count += (uint64_t)launch_slab_report(sb, sl.rank, ctx->stream);
    }
    {
      const cudaError_t ge = cudaGetLastError();
      if (ge != cudaSuccess) return bail(fail(ctx, PBF_E_CUDA, std::string("slab batch: ") + cudaGetErrorString(ge)));
    }
    unsigned int* dev_words = &ctx->status.p->max_neighbors;
    if ((rc = sl.transport->reduce_status_device(ctx, dev_words, kStatusShared)) != PBF_OK) return bail(rc);
    cudaError_t se = cudaMemcpyAsync(ctx->status_host, ctx->status.p, sizeof(StatusBlock), cudaMemcpyDeviceToHost, ctx->stream);
    if (se == cudaSuccess) se = cudaMemcpyAsync(&ctx->last_desc, ctx->desc.p, sizeof(GridDesc), cudaMemcpyDeviceToHost, ctx->stream);
    if (se == cudaSuccess) se = cudaMemcpyAsync(sl.counts_host, sl.counts.p, sizeof(SlabCounts), cudaMemcpyDeviceToHost, ctx->stream);
    if (se == cudaSuccess) se = cudaStreamSynchronize(ctx->stream);
    if (se != cudaSuccess) return bail(fail(ctx, PBF_E_CUDA, std::string("slab batch: ") + cudaGetErrorString(se)));
    if (ctx->profile) timer_resolve(ctx);
    if ((rc = sl.transport->reduce_status_host(ctx, &ctx->status_host->max_neighbors, kStatusShared)) != PBF_OK) return bail(rc);
    const StatusBlock st = *ctx->status_host;
    ctx->last_status = st;
    if (std::getenv("PBF_SLAB_DEBUG"))
      std::fprintf(stderr, "[slab %d/%d] attempt %d nsteps %d: n_own %d->%d ghosts %d+%d b %d+%d send %d+%d | grid %u nbr %u mig %u ghost %u own %u far %u peer %u | max_send %u max_ghost %u max_own %u | mcap %d gcap %d cap %zu hops %d K %d\n",
                   sl.rank, sl.nranks, attempt, nsteps, (int)n0, sl.counts_host->n_own, sl.counts_host->n_ghost[0],
                   sl.counts_host->n_ghost[1], sl.counts_host->b[0], sl.counts_host->b[1], sl.counts_host->n_send[0],
                   sl.counts_host->n_send[1], st.grid_overflow, st.nbr_overflow, st.mig_overflow, st.ghost_overflow,
                   st.own_overflow, st.far_migrant, st.peer_failed, st.max_send, st.max_ghost, st.max_own, sl.mcap,
                   sl.gcap, ctx->cap, sl.hops, ctx->K);
    const unsigned actionable = st.grid_overflow | st.nbr_overflow | st.mig_overflow | st.ghost_overflow |
                                st.own_overflow | st.far_migrant;
    if (!actionable && st.peer_failed)
      return bail(fail(ctx, PBF_E_COMM, "pbf_step: a neighbouring slab reported a failed batch but no rank knows why"));
    if (!actionable && st.peer_timeout) {
      // A neighbour's flag did not arrive in time and nothing else is wrong: a transient stall
      // (lazy module load, graph instantiation, profiler replay, a shared GPU).  The flag is
      // max-reduced, so every rank restores the batch and replays it with a longer limit.
      if (++timeouts > 3)
        return bail(fail(ctx, PBF_E_COMM, "pbf_step: a neighbouring slab did not answer within the peer time-out (4 attempts)"));
      sl.transport->relax_timeout();
      invalidate_graph(ctx);  // the limit is a kernel parameter of the captured substep
      ctx->batches_retried++;
      ctx->tables_dirty = true;
      if (restore() != cudaSuccess) return bail(fail(ctx, PBF_E_CUDA, "slab batch: restoring the backup failed"));
      continue;
    }
    if (!actionable) {
      ctx->n = (size_t)sl.counts_host->n_own;
      for (int s = 0; s < nsteps; ++s) ctx->time += ctx->params.dt;
      sl.warm = true;
      // Messages are sent at full capacity (their size is not known to the host), so capacities
      // follow the observed maxima down as well as up.  The maxima are max-reduced over the slabs:
      // every rank takes the same decision.  Shrinking re-allocates and re-captures the substep
      // graph (~1 ms) and, with peer windows, re-opens the IPC handles (~10 ms), so it needs a factor
      // of two, and growth (below) overshoots by half: a spreading fluid must not
      // re-capture every batch.  Before the re-plan: that one sizes the migration messages for the
      // layers it moves.
      // (Migration messages only grow: with direct peer stores their capacity costs memory, not
      // bandwidth, and a re-plan needs them large again a few batches later.)
      {  // ghost slots the kernels cover: both sides, half above the largest layer pair seen (max-reduced)
        const size_t cover = std::min<size_t>(2 * (size_t)sl.gcap, 2 * ((size_t)st.max_ghost + st.max_ghost / 2 + 1024));
        if (cover > sl.launch_ghost || 2 * cover < sl.launch_ghost) {
          sl.launch_ghost = cover;
          invalidate_graph(ctx);
        }
      }
      const int want_g = (int)(st.max_ghost + st.max_ghost / 2 + 1024);
      if (want_g < sl.gcap / 2) {
        sl.gcap = want_g;
        sl.tot_cap = ctx->cap + 2 * (size_t)sl.gcap;
        invalidate_graph(ctx);
      }
      if ((rc = maybe_rebalance(ctx, st)) != PBF_OK) {  // the batch itself is complete: keep its result
        sl.transport->abort();
        return rc;
      }
      return PBF_OK;
    }
    invalidate_graph(ctx);
    // Grow whatever overflowed (identically on every rank) and replay the batch from the backup.
    ctx->batches_retried++;
    ctx->tables_dirty = true;
    if (st.grid_overflow) {
      const unsigned long long max_cells = ((unsigned long long)st.max_cells_hi << 32) | st.max_cells_lo;
      if (max_cells > (1ull << 30))
        return bail(fail(ctx, PBF_E_CAPACITY, "pbf_step: bounding grid of a slab needs more than 2^30 cells (diverged or non-finite positions; state restored to the start of the batch)"));
      uint32_t cap = ctx->cell_cap;
      while ((unsigned long long)cap < max_cells + max_cells / 4) cap <<= 1;
      ctx->cell_cap = cap;
    }
    if (st.nbr_overflow) ctx->K = (int)((st.max_neighbors + st.max_neighbors / 2 + 16 + 7u) & ~7u);
    // the batch stopped at the first overflowing substep: the maxima are lower bounds of what the
    // rest of it needs, so overshoot
    if (st.mig_overflow) sl.mcap = std::max(8192, (int)(2 * st.max_send + 1024));
    if (st.ghost_overflow) sl.gcap = (int)(st.max_ghost + st.max_ghost / 2 + 1024);
    // far_migrant only counts when nothing else went wrong (a slab that stopped early leaves its
    // neighbours with incomplete data, which can look like a stray particle)
    const bool other = (st.grid_overflow | st.nbr_overflow | st.mig_overflow | st.ghost_overflow | st.own_overflow) != 0;
    if (st.far_migrant && !other) {
      if (sl.hops >= std::max(1, sl.nranks - 1))
        return bail(fail(ctx, PBF_E_COMM, "pbf_step: a particle is outside every reachable slab (non-finite position?)"));
      sl.hops++;
    }
    if (st.own_overflow) {
      // max_own = the largest owned (+ ghost) count any slab saw before it stopped: a lower bound of
      // what the batch needs.  First let the kernels cover more of what is already allocated ...
      const size_t need = (size_t)st.max_own + st.max_own / 2 + 1024;
      sl.launch_ghost = 2 * (size_t)sl.gcap;
      if (need <= ctx->cap) {
        sl.launch_own = std::max(sl.launch_own, std::min(ctx->cap, need));
      } else {  // ... and only then re-allocate (tot_cap = cap + 2 * gcap must cover owned + ghosts)
        ctx->cap = 0;  // force the resize
        if ((rc = ensure_particles(ctx, need, n0)) != PBF_OK) return bail(rc);
        sl.launch_own = ctx->cap;
      }
    }
    if (st.ghost_overflow) {
      const size_t keep_cap = ctx->cap;
      ctx->cap = 0;  // gcap changed: the sorted arrays (cap + 2 * gcap slots) are re-sized
      if ((rc = ensure_particles(ctx, keep_cap, n0)) != PBF_OK) return bail(rc);
      sl.launch_ghost = 2 * (size_t)sl.gcap;
    }
    if (restore() != cudaSuccess) return bail(fail(ctx, PBF_E_CUDA, "slab batch: restoring the backup failed"));
  }
  return bail(fail(ctx, PBF_E_CAPACITY, "pbf_step: slab tables kept overflowing after 32 growth attempts"));
}

}  // namespace pbf

namespace {

// Keeps the particles of this rank's slab out of a global set, in ascending global id.
int slab_take(pbf_ctx* ctx, size_t n_global, const float* px, const float* py, const float* pz, const float* vx,
              const float* vy, const float* vz, const std::vector<int>& cuts) {
  SlabState& sl = ctx->slab;
  if (n_global > 0xfffffff0ull) return fail(ctx, PBF_E_INVALID, "pbf_slab_upload: more than 2^32 particles");
  if (!(ctx->params.h > 0.0f)) return fail(ctx, PBF_E_INVALID, "pbf_slab_upload: set the parameters first (h is needed for the cuts)");
  cudaSetDevice(ctx->device);
  sl.cut_lo = cuts[(size_t)sl.rank];
  sl.cut_hi = cuts[(size_t)sl.rank + 1];
  const float inv_h = 1.0f / ctx->params.h;
  std::vector<uint32_t> gid;
  std::vector<float4> pos, vel;
  std::vector<size_t> per_rank((size_t)sl.nranks, 0);
  for (size_t i = 0; i < n_global; ++i) {
    const int c = host_cell(px[i], inv_h);
    size_t r = 0;
    while (r + 1 < (size_t)sl.nranks && c >= cuts[r + 1]) ++r;
    per_rank[r]++;
    if (c < sl.cut_lo || c >= sl.cut_hi) continue;
    gid.push_back((uint32_t)i);
    pos.push_back(make_float4(px[i], py[i], pz[i], 0.0f));
    vel.push_back(make_float4(vx[i], vy[i], vz[i], 0.0f));
  }
  const size_t n = gid.size();
  // Message capacities are part of the wire format (fixed-size messages, payload offsets), so
  // they are derived from the LARGEST slab: every rank computes the same numbers.
  const size_t n_max = *std::max_element(per_rank.begin(), per_rank.end());
  // Generous on purpose: growing a capacity later means re-allocating every per-particle array,
  // re-opening the peer windows and re-capturing the substep graph (tens of ms, measured on 8 GPUs
  // when fluid_million's splash drove the end slabs from 130 k to 180 k owned particles), while a
  // slot costs ~700 bytes of a 180 GB device.  A re-plan moves whole cell layers at once, and with
  // direct peer stores a message's capacity costs memory, not bandwidth.
  const size_t cap = 2 * n_max + 4096;
  if (sl.gcap == 0) sl.gcap = (int)std::max<size_t>(16384, cap / 2);
  if (sl.mcap == 0) sl.mcap = (int)std::max<size_t>(8192, cap / 2);
  int rc = ensure_particles(ctx, cap, 0);
  if (rc != PBF_OK) return rc;
  if ((rc = ensure_slab_buffers(ctx)) != PBF_OK) return rc;
  if (n) {
    PBF_CUDA(ctx, cudaMemcpy(ctx->pos_o.p, pos.data(), n * sizeof(float4), cudaMemcpyHostToDevice));
    PBF_CUDA(ctx, cudaMemcpy(ctx->vel_o.p, vel.data(), n * sizeof(float4), cudaMemcpyHostToDevice));
    PBF_CUDA(ctx, cudaMemcpy(sl.gid_o.p, gid.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  ctx->n = n;
  return PBF_OK;
}

int slab_enable(pbf_ctx* ctx, int rank, int nranks, Transport* tr, bool owns) {
  SlabState& sl = ctx->slab;
  if (sl.transport && sl.owns_transport) delete sl.transport;
  sl.enabled = true;
  sl.rank = rank;
  sl.nranks = nranks;
  sl.transport = tr;
  sl.owns_transport = owns;
  sl.hops = 1;
  invalidate_graph(ctx);
  ctx->cap = 0;  // the per-particle buffers are re-sized for the slab layout on the next upload
  return PBF_OK;
}

}  // namespace

// ================================================================== C ABI
extern "C" {

int pbf_slab_set_p2p(pbf_ctx* ctx, int enabled);

int pbf_slab_plan(size_t n, const float* px, float h, int nranks, int32_t* cuts) {
  if (!cuts || nranks < 1 || !(h > 0.0f) || (n > 0 && !px)) return fail(nullptr, PBF_E_INVALID, "pbf_slab_plan: bad arguments");
  std::vector<int> c;
  std::string err;
  const int rc = plan_cuts(n, px, h, nranks, c, err);
  if (rc != PBF_OK) return fail(nullptr, rc, err);
  for (int r = 0; r <= nranks; ++r) cuts[r] = c[(size_t)r];
  return PBF_OK;
}

int pbf_slab_plan_hist(const uint64_t* hist, int32_t nlayers, int32_t first_layer, int nranks, float ghost_weight,
                       int32_t* cuts) {
  if (!cuts || nranks < 1 || nlayers < 0 || (nlayers > 0 && !hist) || !(ghost_weight >= 0.0f && ghost_weight <= 1.0f))
    return fail(nullptr, PBF_E_INVALID, "pbf_slab_plan_hist: bad arguments");
  std::vector<size_t> h((size_t)nlayers);
  for (int32_t l = 0; l < nlayers; ++l) h[(size_t)l] = (size_t)hist[l];
  std::vector<int> c;
  std::string err;
  const int rc = plan_cuts_hist(h, first_layer, nranks, ghost_weight, c, err);
  if (rc != PBF_OK) return fail(nullptr, rc, err);
  for (int r = 0; r <= nranks; ++r) cuts[r] = c[(size_t)r];
  return PBF_OK;
}

int pbf_comm_unique_id(void* id_bytes) {
  if (!id_bytes) return PBF_E_INVALID;
  static_assert(sizeof(ncclUniqueId) <= PBF_COMM_ID_BYTES, "ncclUniqueId does not fit PBF_COMM_ID_BYTES");
  ncclUniqueId id;
  const ncclResult_t r = ncclGetUniqueId(&id);
  if (r != ncclSuccess) return fail(nullptr, PBF_E_COMM, std::string("ncclGetUniqueId: ") + ncclGetErrorString(r));
  std::memset(id_bytes, 0, PBF_COMM_ID_BYTES);
  std::memcpy(id_bytes, &id, sizeof(id));
  return PBF_OK;
}

int pbf_comm_init(pbf_ctx* ctx, int rank, int nranks, const void* id_bytes) {
  if (!ctx || !id_bytes || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, PBF_E_INVALID, "pbf_comm_init: bad arguments");
  cudaSetDevice(ctx->device);
  NcclTransport* tr = new NcclTransport();
  tr->rank = rank;
  tr->nranks = nranks;
  ncclUniqueId id;
  std::memcpy(&id, id_bytes, sizeof(id));
  const ncclResult_t r = ncclCommInitRank(&tr->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    tr->comm = nullptr;
    delete tr;
    return fail(ctx, PBF_E_COMM, std::string("ncclCommInitRank: ") + ncclGetErrorString(r));
  }
  const int rc = slab_enable(ctx, rank, nranks, tr, true);
  if (rc != PBF_OK) return rc;
  // Default between processes: direct peer stores + flags (NCCL then only carries the IPC handles
  // and the status agreement).  PBF_SLAB_P2P=0 keeps ncclSend/ncclRecv on the substep path.
  const char* e = std::getenv("PBF_SLAB_P2P");
  if (nranks > 1 && !(e && e[0] == '0')) return pbf_slab_set_p2p(ctx, 1);
  return PBF_OK;
}

int pbf_slab_upload(pbf_ctx* ctx, size_t n_global, const float* px, const float* py, const float* pz,
                    const float* vx, const float* vy, const float* vz) {
  if (!ctx || !ctx->slab.enabled) return fail(ctx, PBF_E_INVALID, "pbf_slab_upload: call pbf_comm_init (or pbf_group_create) first");
  if (n_global > 0 && (!px || !py || !pz || !vx || !vy || !vz)) return fail(ctx, PBF_E_INVALID, "pbf_slab_upload: null array");
  std::vector<int> cuts;
  std::string err;
  const int rc = plan_cuts(n_global, px, ctx->params.h, ctx->slab.nranks, cuts, err);
  if (rc != PBF_OK) return fail(ctx, rc, err);
  return slab_take(ctx, n_global, px, py, pz, vx, vy, vz, cuts);
}

// The per-rank analogue of pbf_upload: this rank's own particles (ascending global id, all inside
// its cuts) from host arrays; the cuts of the last pbf_slab_upload are kept.
int pbf_slab_upload_owned(pbf_ctx* ctx, size_t n, const int64_t* global_id, const float* px, const float* py,
                          const float* pz, const float* vx, const float* vy, const float* vz) {
  if (!ctx || !ctx->slab.enabled) return fail(ctx, PBF_E_INVALID, "pbf_slab_upload_owned: not a slab context");
  if (n > 0 && (!px || !py || !pz || !vx || !vy || !vz)) return fail(ctx, PBF_E_INVALID, "pbf_slab_upload_owned: null array");
  // global_id == NULL: the same particles in the same order as the slab holds them now (what the last
  // pbf_slab_download returned) — a caller that round-trips its particles through the host every
  // substep does not have to ship and convert the ids each time
  if (!global_id && n != ctx->n)
    return fail(ctx, PBF_E_INVALID, "pbf_slab_upload_owned: without ids the count must be that of the last pbf_slab_download");
  cudaSetDevice(ctx->device);
  int rc = ensure_particles(ctx, std::max(n, ctx->cap), ctx->n);
  if (rc != PBF_OK) return rc;
  std::vector<uint32_t>& gid = ctx->slab.gid_host;
  if (global_id) {
    gid.resize(n);
    for (size_t i = 0; i < n; ++i) {
      if (global_id[i] < 0 || global_id[i] > 0xfffffff0ll)
        return fail(ctx, PBF_E_INVALID, "pbf_slab_upload_owned: global id out of range");
      gid[i] = (uint32_t)global_id[i];
    }
  }
  // A particle outside the cuts is legal input: the next substep migrates it (one hop per slab).
  if (n) {
    const float* src[6] = {px, py, pz, vx, vy, vz};
    for (int a = 0; a < 6; ++a)
      PBF_CUDA(ctx, cudaMemcpyAsync(ctx->soa[a].p, src[a], n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (global_id)
      PBF_CUDA(ctx, cudaMemcpyAsync(ctx->slab.gid_o.p, gid.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    const float* dsoa[6] = {ctx->soa[0].p, ctx->soa[1].p, ctx->soa[2].p, ctx->soa[3].p, ctx->soa[4].p, ctx->soa[5].p};
    ctx->launch_count += launch_pack_state(dsoa, ctx->pos_o.p, ctx->vel_o.p, (int)n, ctx->stream);
    // the host arrays may be reused on return; with ids the staging vector as well
    PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->n = n;
  return PBF_OK;
}

size_t pbf_slab_owned(const pbf_ctx* ctx) { return ctx ? ctx->n : 0; }

int pbf_slab_set_p2p(pbf_ctx* ctx, int enabled) {
  if (!ctx || !ctx->slab.enabled || !ctx->slab.transport) return fail(ctx, PBF_E_INVALID, "pbf_slab_set_p2p: not a slab context");
  SlabState& sl = ctx->slab;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  PeerTransport* peer = dynamic_cast<PeerTransport*>(sl.transport);
  if (enabled && !peer) {
    PeerTransport* pt = new PeerTransport();
    pt->inner = sl.transport;
    pt->nccl = dynamic_cast<NcclTransport*>(sl.transport);
    if (LocalTransport* lt = dynamic_cast<LocalTransport*>(sl.transport)) pt->group = lt->group;
    if (!pt->nccl && !pt->group) {
      pt->inner = nullptr;
      delete pt;
      return fail(ctx, PBF_E_INVALID, "pbf_slab_set_p2p: unknown transport");
    }
    sl.transport = pt;
  } else if (!enabled && peer) {
    sl.transport = peer->inner;
    peer->inner = nullptr;
    delete peer;
  }
  sl.warm = false;
  invalidate_graph(ctx);
  return PBF_OK;
}

int pbf_slab_set_cuts(pbf_ctx* ctx, int32_t lo, int32_t hi) {
  if (!ctx || !ctx->slab.enabled) return fail(ctx, PBF_E_INVALID, "pbf_slab_set_cuts: not a slab context");
  const bool first = ctx->slab.rank == 0, last = ctx->slab.rank == ctx->slab.nranks - 1;
  if ((first && lo != INT_MIN) || (last && hi != INT_MAX) || (!first && lo == INT_MIN) || (!last && hi == INT_MAX) ||
      (long long)hi - (long long)lo < 2)
    return fail(ctx, PBF_E_INVALID, "pbf_slab_set_cuts: cuts must tile the x axis with >= 2 cell layers per slab");
  ctx->slab.cut_lo = lo;
  ctx->slab.cut_hi = hi;  // uploaded with the counts at the start of the next batch
  return PBF_OK;
}

int pbf_slab_set_rebalance(pbf_ctx* ctx, float threshold) {
  if (!ctx || !ctx->slab.enabled || !(threshold >= 0.0f)) return fail(ctx, PBF_E_INVALID, "pbf_slab_set_rebalance: bad arguments");
  ctx->slab.rebalance_threshold = threshold;
  return PBF_OK;
}

uint64_t pbf_slab_rebalance_count(const pbf_ctx* ctx) { return ctx ? ctx->slab.rebalances : 0; }

int pbf_slab_cuts(const pbf_ctx* ctx, int32_t* lo, int32_t* hi) {
  if (!ctx || !ctx->slab.enabled) return PBF_E_INVALID;
  if (lo) *lo = ctx->slab.cut_lo;
  if (hi) *hi = ctx->slab.cut_hi;
  return PBF_OK;
}

// Bytes of PAYLOAD this slab sent to its neighbours during the last substep of the last batch
// (messages travel at a fixed capacity with NCCL, so pbf_slab_stats' bytes_sent counts capacities;
// the direct-store transport moves exactly the payload): per neighbour side, `hops` migration
// messages of a header + (pos, pred) per migrant, the ghost build with a header + (pred, pos) of the
// two boundary layers, and one float4 per boundary particle for every later refresh.
int pbf_slab_payload(const pbf_ctx* ctx, uint64_t* bytes_last_substep) {
  if (!ctx || !ctx->slab.enabled || !bytes_last_substep) return PBF_E_INVALID;
  const SlabState& sl = ctx->slab;
  *bytes_last_substep = 0;
  if (!sl.counts_host) return PBF_OK;
  const StepConsts& c = ctx->consts;
  const int iters = ctx->params.solver_iterations;
  const bool final_in_delta = !c.do_xsph && !c.do_vort;
  const int refreshes = (iters > 0 ? (final_in_delta ? iters - 1 : iters) : 0) + ((iters > 0 && c.do_xsph && c.do_vort) ? 1 : 0);
  uint64_t elems = 0;
  for (int side = 0; side < 2; ++side) {
    const bool has = side == 0 ? sl.rank > 0 : sl.rank + 1 < sl.nranks;
    if (!has) continue;
    const uint64_t moved = (uint64_t)std::max(0, sl.counts_host->n_send[side]);
    const uint64_t layer = (uint64_t)std::max(0, sl.counts_host->b[side]);
    elems += (uint64_t)sl.hops * (1 + 2 * moved) + (1 + 2 * layer) + (uint64_t)refreshes * layer;
  }
  *bytes_last_substep = elems * sizeof(float4);
  return PBF_OK;
}

int pbf_slab_transport(const pbf_ctx* ctx) {
  if (!ctx || !ctx->slab.enabled || !ctx->slab.transport) return PBF_E_INVALID;
  if (const PeerTransport* pt = dynamic_cast<const PeerTransport*>(ctx->slab.transport))
    if (pt->active) return PBF_TRANSPORT_PEER_STORES;
  const Transport* base = ctx->slab.transport;
  if (const PeerTransport* pt = dynamic_cast<const PeerTransport*>(base)) base = pt->inner;
  return dynamic_cast<const NcclTransport*>(base) ? PBF_TRANSPORT_NCCL_MESSAGES : PBF_TRANSPORT_LOCAL_COPIES;
}

int pbf_slab_stats(const pbf_ctx* ctx, uint64_t* exchanges, uint64_t* bytes_sent, int32_t* ghosts, int32_t* hops) {
  if (!ctx || !ctx->slab.enabled) return PBF_E_INVALID;
  if (exchanges) *exchanges = ctx->slab.exchanges;
  if (bytes_sent) *bytes_sent = ctx->slab.bytes_sent;
  if (ghosts) *ghosts = ctx->slab.counts_host ? ctx->slab.counts_host->n_ghost[0] + ctx->slab.counts_host->n_ghost[1] : 0;
  if (hops) *hops = ctx->slab.hops;
  return PBF_OK;
}

int pbf_slab_download(pbf_ctx* ctx, int64_t* global_id, float* px, float* py, float* pz, float* vx, float* vy,
                      float* vz) {
  if (!ctx || !ctx->slab.enabled) return fail(ctx, PBF_E_INVALID, "pbf_slab_download: not a slab context");
  cudaSetDevice(ctx->device);
  const size_t n = ctx->n;
  if (n == 0) return PBF_OK;
  float* dst[6] = {px, py, pz, vx, vy, vz};
  float* dsoa[6];
  for (int a = 0; a < 6; ++a) dsoa[a] = dst[a] ? ctx->soa[a].p : nullptr;
  ctx->launch_count += launch_unpack_state(ctx->pos_o.p, ctx->vel_o.p, dsoa, (int)n, ctx->stream);
  for (int a = 0; a < 6; ++a)
    if (dst[a]) PBF_CUDA(ctx, cudaMemcpyAsync(dst[a], ctx->soa[a].p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<uint32_t>& gid = ctx->slab.gid_host;
  if (global_id) {
    gid.resize(n);
    PBF_CUDA(ctx, cudaMemcpyAsync(gid.data(), ctx->slab.gid_o.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  }
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (global_id)
    for (size_t i = 0; i < n; ++i) global_id[i] = (int64_t)gid[i];
  return PBF_OK;
}

// ---- in-process group (several slabs driven by one process) ------------------------------------
pbf_group* pbf_group_create(pbf_ctx** ctxs, int n) {
  if (!ctxs || n < 1) {
    fail(nullptr, PBF_E_INVALID, "pbf_group_create: bad arguments");
    return nullptr;
  }
  pbf_group* g = new pbf_group();
  g->ctxs.assign(ctxs, ctxs + n);
  g->packed.resize((size_t)n);
  g->send_ptr[0].assign((size_t)n, nullptr);
  g->send_ptr[1].assign((size_t)n, nullptr);
  g->words.resize((size_t)n);
  g->wide.resize((size_t)n);
  g->window.assign((size_t)n, nullptr);
  for (int r = 0; r < n; ++r) {
    cudaSetDevice(ctxs[r]->device);
    cudaEventCreateWithFlags(&g->packed[(size_t)r], cudaEventDisableTiming);
    for (int p = 0; p < n; ++p) {  // peer access between distinct devices (ignored when already on / unsupported)
      if (ctxs[p]->device != ctxs[r]->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[p]->device);
        if (can && cudaDeviceEnablePeerAccess(ctxs[p]->device, 0) != cudaSuccess) cudaGetLastError();
      }
    }
    LocalTransport* tr = new LocalTransport();
    tr->group = g;
    slab_enable(ctxs[r], r, n, tr, true);
  }
  return g;
}

void pbf_group_destroy(pbf_group* g) {
  if (!g) return;
  for (size_t r = 0; r < g->ctxs.size(); ++r) {
    cudaSetDevice(g->ctxs[r]->device);
    cudaStreamSynchronize(g->ctxs[r]->stream);
    cudaEventDestroy(g->packed[r]);
    SlabState& sl = g->ctxs[r]->slab;
    if (sl.owns_transport && sl.transport) delete sl.transport;
    sl.transport = nullptr;
    sl.enabled = false;
  }
  delete g;
}

int pbf_group_upload(pbf_group* g, size_t n_global, const float* px, const float* py, const float* pz,
                     const float* vx, const float* vy, const float* vz) {
  if (!g) return PBF_E_INVALID;
  g->n_global = n_global;
  for (pbf_ctx* ctx : g->ctxs) {
    const int rc = pbf_slab_upload(ctx, n_global, px, py, pz, vx, vy, vz);
    if (rc != PBF_OK) return rc;
  }
  return PBF_OK;
}

int pbf_group_step(pbf_group* g, int nsteps) {
  if (!g || nsteps < 0) return PBF_E_INVALID;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->failed = false;
    g->arrived = 0;
  }
  const size_t n = g->ctxs.size();
  std::vector<int> rcs(n, PBF_OK);
  std::vector<std::thread> threads;
  for (size_t r = 0; r < n; ++r)
    threads.emplace_back([g, r, nsteps, &rcs] {
      rcs[r] = pbf_step(g->ctxs[r], nsteps);
      if (rcs[r] != PBF_OK) g->abort();
    });
  for (auto& th : threads) th.join();
  for (size_t r = 0; r < n; ++r)
    if (rcs[r] != PBF_OK) return rcs[r];
  return PBF_OK;
}

// Gathers every slab into global arrays in ORIGINAL particle order (State order).
int pbf_group_download(pbf_group* g, float* px, float* py, float* pz, float* vx, float* vy, float* vz) {
  if (!g) return PBF_E_INVALID;
  size_t seen = 0;
  for (pbf_ctx* ctx : g->ctxs) {
    const size_t n = ctx->n;
    std::vector<int64_t> gid(n);
    std::vector<float> a[6];
    for (auto& v : a) v.resize(n);
    const int rc = pbf_slab_download(ctx, gid.data(), a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data(), a[5].data());
    if (rc != PBF_OK) return rc;
    float* dst[6] = {px, py, pz, vx, vy, vz};
    for (size_t i = 0; i < n; ++i) {
      if ((size_t)gid[i] >= g->n_global) return fail(ctx, PBF_E_INVALID, "pbf_group_download: global id out of range");
      for (int k = 0; k < 6; ++k)
        if (dst[k]) dst[k][gid[i]] = a[k][i];
    }
    seen += n;
  }
  if (seen != g->n_global) return fail(g->ctxs[0], PBF_E_COMM, "pbf_group_download: slabs do not add up to the global particle count");
  return PBF_OK;
}

size_t pbf_group_count(const pbf_group* g) { return g ? g->n_global : 0; }

// Test hook (not in pbf_b200.h): shrink the message capacities so that the overflow -> grow ->
// replay path of a slab batch can be exercised on small inputs.  Call before pbf_slab_upload.
int pbf_debug_set_slab_capacity(pbf_ctx* ctx, int mcap, int gcap) {
  if (!ctx || mcap < 1 || gcap < 1) return PBF_E_INVALID;
  ctx->slab.mcap = mcap;
  ctx->slab.gcap = gcap;
  return PBF_OK;
}

}  // extern "C"
