// pbf_context.h — host-side driver state behind one pbf_ctx (one GPU, one slab).
#pragma once

#include <cuda_runtime.h>

#include <climits>
#include <string>
#include <vector>

#include "kernels/pbf_kernels.h"
#include "pbf_b200.h"

struct pbf_ctx;

namespace pbf {

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;  // elements
  cudaError_t reserve(size_t count) {
    if (count <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  // grow, keeping the first `keep` elements (persistent state across a capacity change)
  cudaError_t grow_keep(size_t count, size_t keep) {
    if (count <= n) return cudaSuccess;
    T* q = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&q), count * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p && keep) e = cudaMemcpy(q, p, (keep < n ? keep : n) * sizeof(T), cudaMemcpyDeviceToDevice);
    if (p) cudaFree(p);
    p = q;
    n = count;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

// How one slab talks to its x-neighbours (pbf_slab.cu): NCCL between processes, or peer copies
// between contexts of one process.
struct Transport {
  virtual ~Transport() {}
  // send[0] -> left neighbour, send[1] -> right neighbour; recv[0] <- left, recv[1] <- right.
  // `bytes` is the same on every rank.  Enqueued on the context's stream.
  virtual int exchange(pbf_ctx* ctx, float4* const send[2], float4* const recv[2], size_t bytes) = 0;
  // max-reduce the shared words of the device status block over all slabs (stream-ordered) ...
  virtual int reduce_status_device(pbf_ctx* ctx, unsigned int* dev_words, int count) = 0;
  // ... or of its host copy after the batch's synchronisation (whichever the transport supports)
  virtual int reduce_status_host(pbf_ctx* ctx, unsigned int* host_words, int count) = 0;
  // all-reduce of a small host array over all slabs (op 0 = max, 1 = sum); blocking, collective
  virtual int allreduce_host(pbf_ctx* ctx, long long* words, int count, int op) = 0;
  virtual void abort() {}
  // a batch was replayed because a neighbour's flag timed out: wait longer next time
  virtual void relax_timeout() {}
  // Direct-store transports between processes: instead of a flag kernel of its own, the exchange
  // runs inside the kernel that consumes the messages (pbf::peer_sync).  Returns false when the
  // transport has to run exchange() itself.
  virtual bool fuse(pbf_ctx* ctx, pbf::PeerSync* sync) { (void)ctx; (void)sync; return false; }
  // true when exchange() is purely stream-ordered, i.e. may be recorded into a CUDA graph
  virtual bool capturable() const { return false; }
  // Called at the start of every slab batch attempt, before anything is enqueued: (re)establish
  // what the transport needs for messages of `msg_elems` float4.  Collective when it has to talk
  // to the neighbours (every rank reaches it with the same msg_elems).
  virtual int prepare(pbf_ctx* ctx, size_t msg_elems) { (void)ctx; (void)msg_elems; return PBF_OK; }
  // Buffers of exchange number `index` out of `count` per substep: where the pack kernels write
  // the two outgoing messages and where the incoming ones will be found.  Returns false when the
  // transport has no opinion (the context's own send/recv buffers are used).
  virtual bool bind(pbf_ctx* ctx, int index, int count, float4* send[2], float4* recv[2]) {
    (void)ctx; (void)index; (void)count; (void)send; (void)recv;
    return false;
  }
};

struct SlabState {
  bool enabled = false;
  int rank = 0, nranks = 1;
  int cut_lo = INT_MIN, cut_hi = INT_MAX;  // owned x-cells [cut_lo, cut_hi)
  int mcap = 0, gcap = 0;                  // message capacities (particles)
  int hops = 1;                            // migration hops per substep
  size_t tot_cap = 0;                      // capacity of the sorted arrays (owned + ghosts)
  // What the captured kernels COVER (grid sizes, overflow checks): <= the allocated capacities.  The
  // allocation is generous (growing it re-allocates everything); the launch bound follows the
  // slab's current size, so that a 16 M-particle scene does not launch twice the blocks it needs,
  // and growing it only re-captures the graph.
  size_t launch_own = 0;                   // owned slots covered (<= ctx->cap)
  size_t launch_ghost = 0;                 // ghost slots covered (<= 2 * gcap); 0 = not decided yet
  size_t msg_elems = 0;                    // float4 elements per message buffer
  int parity = 0;                          // which send buffer pair the next exchange uses
  size_t n_bak = 0;                        // owned count at the start of the batch
  uint64_t exchanges = 0;
  uint64_t bytes_sent = 0;
  uint64_t graph_exchanges = 0, graph_bytes = 0;  // per replay of the captured substep
  bool warm = false;                       // a batch has completed since the communicator was joined
  float rebalance_threshold = 1.1f;         // re-plan the cuts when max / mean owned exceeds it (0 = never)
  double planned_ratio = 1.0;              // max / mean owned the last automatic plan produced by itself
  uint64_t rebalances = 0;
  DevBuf<unsigned long long> hist_dev;     // x-layer histogram / staging of host all-reduces
  Transport* transport = nullptr;          // not owned when it belongs to a group
  bool owns_transport = false;
  DevBuf<SlabCounts> counts;
  DevBuf<uint32_t> gid_o, gid_bak;
  DevBuf<uint32_t> holes;                  // 6 * mcap: slots vacated by migrants + two scratch lists
  DevBuf<float4> send[2][2], recv[2];      // [parity][side], [side]
  SlabCounts* counts_host = nullptr;       // pinned
  cudaStream_t side = nullptr;             // halo of iteration k in flight while lambda runs on the interior
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap = false;                    // PBF_SLAB_OVERLAP=1: split lambda pass (measured slower at 8 x 2 M, see pbf_slab.cu)
  std::vector<uint32_t> gid_host;          // staging of global ids for upload / download
};

struct StageTimer {
  std::vector<cudaEvent_t> begin, end;  // recorded pairs awaiting resolution
  std::vector<int> stage;
  double total_ms[PBF_STAGE_COUNT] = {0};
  uint64_t launches[PBF_STAGE_COUNT] = {0};
  std::vector<cudaEvent_t> pool;
};

}  // namespace pbf

struct pbf_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string error;

  pbf_params params{};
  std::vector<float4> planes_host;
  pbf::StepConsts consts{};
  int mode = PBF_MODE_STRICT;
  bool use_graph = true;
  bool debug = false;
  bool profile = false;

  size_t n = 0;         // particles
  size_t cap = 0;       // particle capacity of the per-particle buffers
  float time = 0.0f;

  // persistent state (original order) + batch backup
  pbf::DevBuf<float4> pos_o, vel_o, pos_bak, vel_bak, pred_o;
  // sorted-order work arrays
  pbf::DevBuf<float4> pred_a, pred_b, pos_s, vel_a, vel_b, omega;
  pbf::DevBuf<pbf::PosVel> pv;  // (pos, vel, m/rho) records gathered by XSPH
  pbf::DevBuf<float> rho;
  pbf::DevBuf<float4> planes_dev;
  // grid
  pbf::DevBuf<pbf::GridDesc> desc;
  pbf::DevBuf<pbf::StatusBlock> status;
  pbf::DevBuf<uint32_t> keys0, keys1, vals0, vals1, chunk_total;
  pbf::DevBuf<int2> cell_range;
  pbf::DevBuf<uint32_t> cell_count, cell_excl, slot_id;
  pbf::DevBuf<unsigned long long> cell_key;  // sparse cell table (hash slots), all-ones when empty
  uint32_t cell_cap = 1u << 22;   // a power of two (the sparse table masks with cell_cap - 1)
  bool tables_dirty = true;       // cell counters / sparse keys / descriptor need a reset before the next batch
  int sorted_buf = 0;  // which keys/vals buffer holds the last substep's sorted order
  // neighbour list
  pbf::DevBuf<uint32_t> nbr_idx, nbr_count;
  int K = 96;
  // brick path (kernels/brick.cu): one CTA per brick of grid cells, neighbourhood staged in shared
  // memory, 16-bit tile-relative neighbour entries.  `brick_want` is the user's / environment's
  // choice (PBF_BRICK=0 disables it), `brick_on` what the current batch attempt uses: a batch whose
  // tiles do not fit (or that needs the sparse cell table) is replayed with the global-gather family
  // and the brick path is tried again `brick_retry` batches later.
  pbf::DevBuf<pbf::BrickRec> bricks;
  pbf::DevBuf<unsigned int> brick_ctl;  // {brick ticket, finished CTAs} of the persistent brick kernels
  int brick_cap = 0;
  bool brick_want = false, brick_on = false;
  bool brick_persist = true;      // which brick driver: persistent CTAs (PBF_BRICK_PERSISTENT) or one CTA per brick
  int brick_retry = 0;
  uint64_t brick_fallbacks = 0;
  bool last_brick = false;        // the last completed batch ran on the brick path (debug surface)
  // SoA staging for upload / download
  pbf::DevBuf<float> soa[6];
  // debug scratch (sorted order)
  pbf::DevBuf<float> dbg_lambda, dbg_rho;
  pbf::DevBuf<float4> dbg_delta, dbg_dv, dbg_eta;

  pbf::StatusBlock* status_host = nullptr;  // pinned
  pbf::StatusBlock last_status{};           // of the last completed batch
  pbf::GridDesc last_desc{};

  // CUDA graph of one substep (re-captured when the configuration changes)
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_key = 0;
  int graph_kernels = 0;
  // the cuda_step contract (pbf_step_host) as one graph: copies in, substep, copies out (pbf_capi.cu)
  cudaGraphExec_t host_graph = nullptr;
  const void* host_graph_ptr[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int host_graph_kernels = 0;
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

  // Frame snapshots (pbf_snapshot_begin / _wait): positions are unpacked into a device
  // staging copy on the compute stream and travel to library-owned pinned host buffers on a copy
  // stream of their own, so the download runs under the next batch of substeps.  Two slots.
  struct Snapshot {
    pbf::DevBuf<float> dev[3];
    float* host[3] = {nullptr, nullptr, nullptr};  // pinned
    size_t host_cap = 0;
    size_t n = 0;
    float time = 0.0f;
    bool pending = false;
    cudaEvent_t ready = nullptr, done = nullptr;
  } snap[2];
  cudaStream_t copy_stream = nullptr;

  pbf::SlabState slab;

  pbf::StageTimer timer;
  uint64_t launch_count = 0;
  uint64_t batches_retried = 0;
};

// ---- internals shared by pbf_capi.cu and pbf_slab.cu ---------------------------------------------
#define PBF_CUDA(ctx, expr)                                                                       \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return pbf::fail(ctx, PBF_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
  } while (0)

namespace pbf {
int fail(pbf_ctx* ctx, int code, const std::string& msg);
void invalidate_graph(pbf_ctx* ctx);
int ensure_particles(pbf_ctx* ctx, size_t n, size_t keep);
int ensure_tables(pbf_ctx* ctx);
int reset_status(pbf_ctx* ctx);
void stage_mark(void* user, int stage, int begin);
void timer_resolve(pbf_ctx* ctx);
void fill_grid_buffers(pbf_ctx* ctx, GridBuffers& g);
void fill_neighbor_list(pbf_ctx* ctx, NeighborList& nl);
void fill_solve_buffers(pbf_ctx* ctx, SolveBuffers& b);
// pbf_slab.cu
int slab_step(pbf_ctx* ctx, int nsteps);
void slab_release(pbf_ctx* ctx);
}  // namespace pbf
