// comm.cu -- communicator over NVSwitch peer memory (no NCCL on this path) and
// the op-surface AlltoallvN built on it.
//
// One process per GPU.  Every rank cudaMalloc's one symmetric allocation
//   [ control block | data window ]
// and exports it as a CUDA IPC handle inside a 128-byte token (the reference
// broadcasts a 128-byte NCCL id the same way, distribute/collective.py:108-115).
// hbCommConnect maps every peer's allocation (cudaIpcOpenMemHandle => NVLink P2P
// through NVSwitch).  All cross-GPU signalling is done by kernels with
// st.release.sys / ld.acquire.sys on epoch-valued flags living in the control
// block of the RECEIVER, so a waiter spins on local HBM only.
//
// AlltoallvN (HbNcclAlltoallvN, nccl_alltoallv.cc:359-580):
//   phase 1  hbAlltoallvNSizes: all-gather of the N x W send-size vectors through
//            peer mailboxes (replaces the NCCL AlltoallN size pre-exchange,
//            nccl_collective.cc:153-199); builds, on device, the segment tables
//            for phase 2 in a FIFO snapshot slot; recv sizes go to the caller
//            (device + pinned host, the reference blocks the host here too).
//   phase 2  hbAlltoallvN: every CTA streams its share of the N x W segments from
//            the local input straight into the peers' windows with 128-bit stores
//            (push), the last CTA publishes an epoch flag to every peer; a second
//            kernel waits for all peers' flags and copies window -> output
//            (outputs are caller-allocated and not peer-mapped).
// The window is used in two halves alternating per call, which is what makes a
// trailing barrier unnecessary (see DESIGN.md "window reuse").
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>

#include "comm.cuh"

namespace hb {

// ---- barrier -----------------------------------------------------------------
// Two kernels (arrive, then wait): see comm_submit -- a kernel that publishes
// never waits, so an in-process group can issue all arrivals before any wait.
__global__ void barrier_arrive_kernel(PeerPtrs peers, int me, int world, uint32_t epoch) {
  const int q = threadIdx.x;
  if (q < world) {
    __threadfence_system();
    Control* remote = reinterpret_cast<Control*>(peers.p[q]);
    st_release_sys_u32(&remote->barrier_flags[me], epoch);
  }
}
__global__ void barrier_wait_kernel(PeerPtrs peers, int me, int world, uint32_t epoch,
                                    int32_t* status) {
  const int q = threadIdx.x;
  if (q < world) {
    Control* mine = reinterpret_cast<Control*>(peers.p[me]);
    if (!wait_flag(&mine->barrier_flags[q], epoch)) raise_status(status, HB_STATUS_PEER_TIMEOUT);
  }
}

// ---- alltoallv phase 1: sizes all-gather + segment tables ---------------------
struct SizesParams {
  const int32_t* send_sizes[kMaxA2aTensors];
  int32_t* recv_sizes[kMaxA2aTensors];
};

__global__ void __launch_bounds__(256)
a2a_sizes_publish_kernel(const __grid_constant__ SizesParams P, PeerPtrs peers, int me, int world,
                         int n, uint32_t call) {
  const int parity = call & 1;
  // publish my N x W send sizes into every peer's mailbox row `me`
  for (int i = threadIdx.x; i < n * world; i += blockDim.x) {
    const int k = i / world, r = i % world;
    const int32_t v = P.send_sizes[k][r];
    for (int q = 0; q < world; ++q) {
      Control* remote = reinterpret_cast<Control*>(peers.p[q]);
      remote->mailbox[parity][(me * kMaxA2aTensors + k) * kMaxWorld + r] = v;
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    Control* remote = reinterpret_cast<Control*>(peers.p[threadIdx.x]);
    st_release_sys_u32(&remote->sizes_flags[parity][me], call);
  }
}

__global__ void __launch_bounds__(256)
a2a_sizes_collect_kernel(const __grid_constant__ SizesParams P, PeerPtrs peers, int me, int world,
                         int n, uint32_t call, int slot, int32_t* status) {
  const int parity = call & 1;
  Control* mine = reinterpret_cast<Control*>(peers.p[me]);
  if ((int)threadIdx.x < world) {
    if (!wait_flag(&mine->sizes_flags[parity][threadIdx.x], call))
      raise_status(status, HB_STATUS_PEER_TIMEOUT);
  }
  __syncthreads();
  // snapshot the matrix
  Snapshot* S = &mine->snap[slot];
  for (int i = threadIdx.x; i < world * n * world; i += blockDim.x) {
    const int q = i / (n * world), k = (i / world) % n, r = i % world;
    const int idx = (q * kMaxA2aTensors + k) * kMaxWorld + r;
    S->matrix[idx] = *reinterpret_cast<volatile int32_t*>(&mine->mailbox[parity][idx]);
  }
  __syncthreads();
  // recv sizes for the caller
  for (int i = threadIdx.x; i < n * world; i += blockDim.x) {
    const int k = i / world, q = i % world;
    const int32_t v = S->matrix[(q * kMaxA2aTensors + k) * kMaxWorld + me];
    S->recv_sizes[k * world + q] = v;
    if (P.recv_sizes[k] != nullptr) P.recv_sizes[k][q] = v;
  }
}

// Build the segment tables once row sizes (bytes per element row) are known:
// done at the start of phase 2 by one CTA.
struct TableParams {
  uint64_t row_bytes[kMaxA2aTensors];
};

__global__ void __launch_bounds__(256)
a2a_tables_kernel(const __grid_constant__ TableParams P, PeerPtrs peers, int me, int world, int n,
                  int slot, uint64_t half_bytes, int32_t* status) {
  Control* mine = reinterpret_cast<Control*>(peers.p[me]);
  Snapshot* S = &mine->snap[slot];
  __shared__ uint64_t s_region[kMaxWorld][2];  // running window offset per destination r
  auto M = [&](int q, int k, int r) -> uint64_t {
    return (uint64_t)(uint32_t)S->matrix[(q * kMaxA2aTensors + k) * kMaxWorld + r];
  };
  // thread r (< world): walk tensors in order, computing for destination r the
  // window offset of every (k, q) segment -> the push entry [k][r] of q == me.
  const int r = threadIdx.x;
  if (r < world) {
    uint64_t off = 0;
    for (int k = 0; k < n; ++k) {
      for (int q = 0; q < world; ++q) {
        const uint64_t bytes = M(q, k, r) * P.row_bytes[k];
        if (q == me) {
          SegEntry& e = S->push[k * world + r];
          uint64_t src = 0;
          for (int rr = 0; rr < r; ++rr) src += M(me, k, rr) * P.row_bytes[k];
          e.src_off = src;
          e.dst_off = off;
          e.bytes = bytes;
        }
        if (r == me) {
          SegEntry& e = S->pull[k * world + q];
          uint64_t dst = 0;
          for (int qq = 0; qq < q; ++qq) dst += M(qq, k, me) * P.row_bytes[k];
          e.src_off = off;
          e.dst_off = dst;
          e.bytes = bytes;
        }
        off += (bytes + 15) & ~(uint64_t)15;
      }
    }
    s_region[r][0] = off;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t c = 0, ovf = 0;
    for (int i = 0; i < n * world; ++i) {
      S->push[i].chunk_begin = c;
      c += (S->push[i].bytes + kChunkBytes - 1) / kChunkBytes;
    }
    S->push_chunks = c;
    c = 0;
    for (int i = 0; i < n * world; ++i) {
      S->pull[i].chunk_begin = c;
      c += (S->pull[i].bytes + kChunkBytes - 1) / kChunkBytes;
    }
    S->pull_chunks = c;
    for (int rr = 0; rr < world; ++rr)
      if (s_region[rr][0] > half_bytes) ovf = 1;
    S->overflow = ovf;
    if (ovf) raise_status(status, HB_STATUS_WINDOW_OVERFLOW);
  }
}

// Generic chunked copy of one byte range with the widest safe vector width.
__device__ __forceinline__ void copy_chunk(const unsigned char* src, unsigned char* dst,
                                           uint64_t bytes) {
  const uintptr_t a = (uintptr_t)src | (uintptr_t)dst | (uintptr_t)bytes;
  if ((a & 15) == 0) {
    const int4* s = reinterpret_cast<const int4*>(src);
    int4* d = reinterpret_cast<int4*>(dst);
    const uint64_t n = bytes >> 4;
    uint64_t i = threadIdx.x;
    for (; i + 3 * blockDim.x < n; i += 4 * blockDim.x) {
      int4 v0 = s[i], v1 = s[i + blockDim.x], v2 = s[i + 2 * blockDim.x], v3 = s[i + 3 * blockDim.x];
      d[i] = v0; d[i + blockDim.x] = v1; d[i + 2 * blockDim.x] = v2; d[i + 3 * blockDim.x] = v3;
    }
    for (; i < n; i += blockDim.x) d[i] = s[i];
  } else if ((a & 7) == 0) {
    const uint64_t n = bytes >> 3;
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x)
      reinterpret_cast<uint64_t*>(dst)[i] = reinterpret_cast<const uint64_t*>(src)[i];
  } else if ((a & 3) == 0) {
    const uint64_t n = bytes >> 2;
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x)
      reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  } else {
    for (uint64_t i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
  }
}

__device__ __forceinline__ int find_chunk_seg(const SegEntry* t, int nseg, uint64_t chunk) {
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {  // last entry with chunk_begin <= chunk and non-empty beyond
    const int mid = (lo + hi + 1) >> 1;
    if (t[mid].chunk_begin <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

struct PushParams {
  const unsigned char* inputs[kMaxA2aTensors];
};

// last-CTA-done: the CTA that finishes last publishes `epoch` into
// data_flags[half][me] of every peer
__device__ __forceinline__ void publish_window(PeerPtrs& peers, Control* mine, int me, int world,
                                               int half, uint32_t epoch) {
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned done = atomicAdd(&mine->done_counter[half], 1u);
    s_last = (done == gridDim.x - 1);
    if (s_last) mine->done_counter[half] = 0;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < world) {
    __threadfence_system();
    Control* remote = reinterpret_cast<Control*>(peers.p[threadIdx.x]);
    st_release_sys_u32(&remote->data_flags[half][me], epoch);
  }
}

__global__ void __launch_bounds__(256)
a2a_push_kernel(const __grid_constant__ PushParams P, PeerPtrs peers, int me, int world, int n,
                int slot, int half, uint64_t window_off, uint64_t half_bytes, uint32_t epoch) {
  Control* mine = reinterpret_cast<Control*>(peers.p[me]);
  const Snapshot* S = &mine->snap[slot];
  const int nseg = n * world;
  if (!S->overflow) {
    const uint64_t total = S->push_chunks;
    for (uint64_t c = blockIdx.x; c < total; c += gridDim.x) {
      const int si = find_chunk_seg(S->push, nseg, c);
      const SegEntry e = S->push[si];
      const int k = si / world, r = si % world;
      const uint64_t o = (c - e.chunk_begin) * kChunkBytes;
      const uint64_t len = (e.bytes - o < kChunkBytes) ? e.bytes - o : kChunkBytes;
      copy_chunk(P.inputs[k] + e.src_off + o,
                 peers.p[r] + window_off + (uint64_t)half * half_bytes + e.dst_off + o, len);
    }
  }
  publish_window(peers, mine, me, world, half, epoch);
}

struct PullParams {
  unsigned char* outputs[kMaxA2aTensors];
};

__global__ void __launch_bounds__(256)
a2a_copyout_kernel(const __grid_constant__ PullParams P, PeerPtrs peers, int me, int world, int n,
                   int slot, int half, uint64_t window_off, uint64_t half_bytes, uint32_t epoch,
                   int32_t* status) {
  Control* mine = reinterpret_cast<Control*>(peers.p[me]);
  __shared__ int s_timeout;
  if (threadIdx.x == 0) s_timeout = 0;
  __syncthreads();
  if ((int)threadIdx.x < world) {
    if (!wait_flag(&mine->data_flags[half][threadIdx.x], epoch)) {
      s_timeout = 1;
      raise_status(status, HB_STATUS_PEER_TIMEOUT);
    }
  }
  __syncthreads();
  const Snapshot* S = &mine->snap[slot];
  if (S->overflow || s_timeout) return;  // never copy out data that did not arrive
  const int nseg = n * world;
  const uint64_t total = S->pull_chunks;
  const unsigned char* win = peers.p[me] + window_off + (uint64_t)half * half_bytes;
  for (uint64_t c = blockIdx.x; c < total; c += gridDim.x) {
    const int si = find_chunk_seg(S->pull, nseg, c);
    const SegEntry e = S->pull[si];
    const int k = si / world;
    const uint64_t o = (c - e.chunk_begin) * kChunkBytes;
    const uint64_t len = (e.bytes - o < kChunkBytes) ? e.bytes - o : kChunkBytes;
    copy_chunk(win + e.src_off + o, P.outputs[k] + e.dst_off + o, len);
  }
}

// ---- small dense all-reduce (sum, fp32) over the peer windows ---------------------
// Replaces Collective.allreduce for the dense gradients of replicated small
// tables (training/gradient.py:157-160): every rank stores its vector into slot
// `me` of every peer's window half (an all-gather over NVSwitch), then reduces the
// W slots locally in RANK ORDER -- the same summation order on every rank, so
// replicas stay bit-identical -- and scales by `scale` (the reference's 1/W mean,
// gradient.py:77-97).  W x the bytes of a ring all-reduce: meant for the small
// replicated tables of the embedding path, not for dense model gradients.
__global__ void __launch_bounds__(256)
ar_push_kernel(const float* __restrict__ in, int64_t count, uint64_t slot_bytes, PeerPtrs peers,
               int me, int world, int half, uint64_t window_off, uint64_t half_bytes,
               uint32_t epoch) {
  Control* mine = reinterpret_cast<Control*>(peers.p[me]);
  const uint64_t bytes = (uint64_t)count * 4;
  const uint64_t chunks = (bytes + kChunkBytes - 1) / kChunkBytes;
  for (uint64_t w = blockIdx.x; w < chunks * world; w += gridDim.x) {
    const int r = (int)(w / chunks);
    const uint64_t o = (w % chunks) * kChunkBytes;
    const uint64_t len = (bytes - o < kChunkBytes) ? bytes - o : kChunkBytes;
    copy_chunk(reinterpret_cast<const unsigned char*>(in) + o,
               peers.p[r] + window_off + (uint64_t)half * half_bytes + (uint64_t)me * slot_bytes + o,
               len);
  }
  publish_window(peers, mine, me, world, half, epoch);
}

__global__ void __launch_bounds__(256)
ar_reduce_kernel(float* __restrict__ out, int64_t count, uint64_t slot_bytes, float scale,
                 PeerPtrs peers, int me, int world, int half, uint64_t window_off,
                 uint64_t half_bytes, uint32_t epoch, int32_t* status) {
  Control* mine = reinterpret_cast<Control*>(peers.p[me]);
  __shared__ int s_timeout;
  if (threadIdx.x == 0) s_timeout = 0;
  __syncthreads();
  if ((int)threadIdx.x < world) {
    if (!wait_flag(&mine->data_flags[half][threadIdx.x], epoch)) {
      s_timeout = 1;
      raise_status(status, HB_STATUS_PEER_TIMEOUT);
    }
  }
  __syncthreads();
  if (s_timeout) return;
  const unsigned char* win = peers.p[me] + window_off + (uint64_t)half * half_bytes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    float acc = reinterpret_cast<const float*>(win)[i];  // plain loads: written by peers
    for (int q = 1; q < world; ++q)
      acc = __fadd_rn(acc, reinterpret_cast<const float*>(win + (uint64_t)q * slot_bytes)[i]);
    out[i] = __fmul_rn(acc, scale);
  }
}

// ---- in-process group ---------------------------------------------------------------
// W communicators on ONE device whose "peer" pointers are each other's
// allocations.  One host thread per rank drives its communicator exactly as a
// process would; collective entry points rendezvous here (comm_submit).
struct LocalGroup {
  std::mutex mu;
  std::condition_variable cv;
  int world = 0;
  int device = 0;
  int refs = 0;
  hbComm* members[kMaxWorld] = {};
  // rendezvous of the current op
  uint64_t seq = 0;
  int arrived = 0;
  struct Pending {
    const PhaseFn* run;
    cudaStream_t stream;
    int op_code, nphases;
  } pend[kMaxWorld] = {};
  int result = HB_OK;
  std::string error;
  cudaEvent_t ev[kMaxWorld] = {};
  bool have_events = false;
};

static int group_flush(LocalGroup* g) {
  const int W = g->world;
  int nph = g->pend[0].nphases;
  bool same_stream = true;
  for (int r = 0; r < W; ++r) {
    if (g->pend[r].op_code != g->pend[0].op_code || g->pend[r].nphases != nph) {
      set_last_error("in-process group: ranks submitted different collectives (op %d/%d phases vs op %d/%d phases on rank %d)",
                     g->pend[0].op_code, nph, g->pend[r].op_code, g->pend[r].nphases, r);
      return HB_ERR_COMM;
    }
    if (g->pend[r].stream != g->pend[0].stream) same_stream = false;
  }
  if (!same_stream && !g->have_events) {
    for (int r = 0; r < W; ++r) HB_CUDA_OK(cudaEventCreateWithFlags(&g->ev[r], cudaEventDisableTiming));
    g->have_events = true;
  }
  for (int p = 0; p < nph; ++p) {
    for (int r = 0; r < W; ++r) {
      const int rc = (*g->pend[r].run)(p);
      if (rc != HB_OK) return rc;
    }
    if (!same_stream && p + 1 < nph) {
      // phase fence: every rank's next phase is ordered behind all ranks' phase p
      for (int r = 0; r < W; ++r) HB_CUDA_OK(cudaEventRecord(g->ev[r], g->pend[r].stream));
      for (int r = 0; r < W; ++r)
        for (int q = 0; q < W; ++q)
          if (q != r) HB_CUDA_OK(cudaStreamWaitEvent(g->pend[r].stream, g->ev[q], 0));
    }
  }
  return HB_OK;
}

int comm_submit(hbComm* c, int op_code, int nphases, cudaStream_t stream, const PhaseFn& run) {
  LocalGroup* g = c->group;
  if (g == nullptr) {
    for (int p = 0; p < nphases; ++p) {
      const int rc = run(p);
      if (rc != HB_OK) return rc;
    }
    return HB_OK;
  }
  std::unique_lock<std::mutex> lk(g->mu);
  const uint64_t my_seq = g->seq;
  g->pend[c->rank] = LocalGroup::Pending{&run, stream, op_code, nphases};
  g->arrived++;
  if (g->arrived == g->world) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != g->device) cudaSetDevice(g->device);
    g->result = group_flush(g);
    if (g->result != HB_OK) g->error = get_last_error();
    if (dev != g->device) cudaSetDevice(dev);
    g->arrived = 0;
    g->seq++;
    g->cv.notify_all();
    return g->result;
  }
  // the submitting thread stays here until the last rank has issued everything:
  // `run` (and what it captures by reference) must stay alive, and the entry
  // point's contract is that the work is enqueued when it returns
  static const int timeout_s = [] {
    const char* e = getenv("HB_LOCAL_GROUP_TIMEOUT_S");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : 120;
  }();
  const bool ok = g->cv.wait_for(lk, std::chrono::seconds(timeout_s), [&] { return g->seq != my_seq; });
  if (!ok) {
    g->arrived--;
    set_last_error("in-process group: rank %d waited %d s for the other ranks to submit op %d "
                   "(every rank must issue the same collective sequence, one thread per rank)",
                   c->rank, timeout_s, op_code);
    return HB_ERR_COMM;
  }
  if (g->result != HB_OK) set_last_error("%s", g->error.c_str());
  return g->result;
}

}  // namespace hb

// ---------------------------------------------------------------------------------
extern "C" {

int hbCommCreate(int rank, int world_size, int local_size, size_t window_bytes, hbComm** comm,
                 unsigned char token_out[HB_COMM_TOKEN_BYTES]) {
  using namespace hb;
  HB_REQUIRE(comm && token_out, "hbCommCreate: null argument");
  HB_REQUIRE(world_size >= 1 && world_size <= kMaxWorld, "hbCommCreate: world_size %d not in [1,%d]",
             world_size, kMaxWorld);
  HB_REQUIRE(rank >= 0 && rank < world_size, "hbCommCreate: bad rank %d", rank);
  HB_REQUIRE(local_size >= 1, "hbCommCreate: bad local_size %d", local_size);
  hbComm* c = new hbComm();
  memset(c, 0, sizeof(*c));
  c->rank = rank; c->world = world_size; c->local = local_size;
  c->window_bytes = align_up(window_bytes, 4096);
  c->alloc_bytes = control_bytes() + c->window_bytes;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->base), c->alloc_bytes);
  if (e != cudaSuccess) {
    set_last_error("hbCommCreate: cudaMalloc(%zu) failed: %s", c->alloc_bytes, cudaGetErrorString(e));
    delete c;
    return HB_ERR_CUDA;
  }
  e = cudaMemset(c->base, 0, control_bytes());
  if (e == cudaSuccess && world_size > 1) e = cudaIpcGetMemHandle(&c->handle, c->base);
  if (e != cudaSuccess) {
    set_last_error("hbCommCreate: IPC export failed: %s", cudaGetErrorString(e));
    cudaFree(c->base);
    delete c;
    return HB_ERR_COMM;
  }
  cudaDeviceSynchronize();
  c->peer[rank] = c->base;
  c->connected = (world_size == 1);
  memset(token_out, 0, HB_COMM_TOKEN_BYTES);
  uint32_t hdr[4] = {kTokenMagic, (uint32_t)rank, (uint32_t)world_size, (uint32_t)getpid()};
  memcpy(token_out, hdr, sizeof(hdr));
  uint64_t sz = c->alloc_bytes;
  memcpy(token_out + 16, &sz, 8);
  memcpy(token_out + 32, &c->handle, sizeof(cudaIpcMemHandle_t));
  *comm = c;
  return HB_OK;
}

int hbCommConnect(hbComm* c, const unsigned char* all_tokens) {
  using namespace hb;
  HB_REQUIRE(c && all_tokens, "hbCommConnect: null argument");
  if (c->connected) return HB_OK;
  for (int q = 0; q < c->world; ++q) {
    const unsigned char* t = all_tokens + (size_t)q * HB_COMM_TOKEN_BYTES;
    uint32_t hdr[4];
    memcpy(hdr, t, sizeof(hdr));
    uint64_t sz;
    memcpy(&sz, t + 16, 8);
    if (hdr[0] != kTokenMagic || (int)hdr[1] != q || (int)hdr[2] != c->world || sz != c->alloc_bytes) {
      set_last_error("hbCommConnect: token %d is malformed or from a differently-sized communicator", q);
      return HB_ERR_COMM;
    }
    if (q == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, t + 32, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_last_error("hbCommConnect: cudaIpcOpenMemHandle(rank %d) failed: %s", q, cudaGetErrorString(e));
      return HB_ERR_COMM;
    }
    c->peer[q] = reinterpret_cast<unsigned char*>(p);
    c->opened[q] = true;
  }
  c->connected = true;
  return HB_OK;
}

int hbCommCreateLocalGroup(int world_size, size_t window_bytes, hbComm** comms) {
  using namespace hb;
  HB_REQUIRE(comms, "hbCommCreateLocalGroup: null argument");
  HB_REQUIRE(world_size >= 1 && world_size <= kMaxWorld,
             "hbCommCreateLocalGroup: world_size %d not in [1,%d]", world_size, kMaxWorld);
  LocalGroup* g = new LocalGroup();
  g->world = world_size;
  HB_CUDA_OK(cudaGetDevice(&g->device));
  for (int r = 0; r < world_size; ++r) {
    hbComm* c = new hbComm();
    memset(c, 0, sizeof(*c));
    c->rank = r; c->world = world_size; c->local = world_size;
    c->window_bytes = align_up(window_bytes, 4096);
    c->alloc_bytes = control_bytes() + c->window_bytes;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->base), c->alloc_bytes);
    if (e == cudaSuccess) e = cudaMemset(c->base, 0, control_bytes());
    if (e != cudaSuccess) {
      set_last_error("hbCommCreateLocalGroup: allocating rank %d (%zu B) failed: %s", r, c->alloc_bytes,
                     cudaGetErrorString(e));
      if (c->base) cudaFree(c->base);
      delete c;
      for (int q = 0; q < r; ++q) { cudaFree(g->members[q]->base); delete g->members[q]; }
      delete g;
      return HB_ERR_CUDA;
    }
    c->group = g;
    g->members[r] = c;
  }
  for (int r = 0; r < world_size; ++r) {
    for (int q = 0; q < world_size; ++q) g->members[r]->peer[q] = g->members[q]->base;
    g->members[r]->connected = true;
    comms[r] = g->members[r];
  }
  g->refs = world_size;
  cudaDeviceSynchronize();
  return HB_OK;
}

int hbCommDestroy(hbComm* c) {
  if (!c) return HB_OK;
  cudaDeviceSynchronize();
  for (int q = 0; q < c->world; ++q)
    if (c->opened[q]) cudaIpcCloseMemHandle(c->peer[q]);
  if (c->base) cudaFree(c->base);
  if (c->group != nullptr) {
    hb::LocalGroup* g = c->group;
    bool last;
    {
      std::lock_guard<std::mutex> lk(g->mu);
      last = (--g->refs == 0);
    }
    if (last) {
      if (g->have_events)
        for (int r = 0; r < g->world; ++r) cudaEventDestroy(g->ev[r]);
      delete g;
    }
  }
  delete c;
  return HB_OK;
}

int hbCommRank(const hbComm* c) { return c ? c->rank : -1; }
int hbCommWorldSize(const hbComm* c) { return c ? c->world : -1; }
void* hbCommWindow(hbComm* c) { return c ? c->base + hb::control_bytes() : nullptr; }
size_t hbCommWindowBytes(const hbComm* c) { return c ? c->window_bytes : 0; }

int hbCommSetStatusWord(hbComm* c, int32_t* d_status) {
  HB_REQUIRE(c, "hbCommSetStatusWord: null communicator");
  c->d_status = d_status;
  return HB_OK;
}

int hbCommBarrier(hbComm* c, hbStream stream_) {
  using namespace hb;
  cudaStream_t stream = (cudaStream_t)stream_;
  HB_REQUIRE(c && c->connected, "hbCommBarrier: communicator not connected");
  const uint32_t epoch = ++c->barrier_epoch;
  return comm_submit(c, kOpBarrier, 2, stream, [&](int phase) -> int {
    KernelScope ks(HB_K_BARRIER, stream);
    if (phase == 0) barrier_arrive_kernel<<<1, 32, 0, stream>>>(peer_ptrs(c), c->rank, c->world, epoch);
    else barrier_wait_kernel<<<1, 32, 0, stream>>>(peer_ptrs(c), c->rank, c->world, epoch, c->d_status);
    HB_CUDA_OK(cudaGetLastError());
    return HB_OK;
  });
}

int hbAlltoallvNSizes(hbComm* c, int n, const int32_t* const* d_send_sizes,
                      int32_t* const* d_recv_sizes, int32_t* h_recv_sizes, hbStream stream_) {
  using namespace hb;
  cudaStream_t stream = (cudaStream_t)stream_;
  HB_REQUIRE(c && c->connected, "hbAlltoallvNSizes: communicator not connected");
  HB_REQUIRE(n >= 1 && n <= kMaxA2aTensors, "hbAlltoallvNSizes: N=%d not in [1,%d]", n, kMaxA2aTensors);
  HB_REQUIRE(d_send_sizes, "hbAlltoallvNSizes: null send sizes");
  HB_REQUIRE(c->sizes_calls - c->data_calls < (uint32_t)kSnapSlots,
             "hbAlltoallvNSizes: more than %d size exchanges without their hbAlltoallvN", kSnapSlots);
  SizesParams P;
  for (int k = 0; k < n; ++k) {
    HB_REQUIRE(d_send_sizes[k], "hbAlltoallvNSizes: null send sizes for tensor %d", k);
    P.send_sizes[k] = d_send_sizes[k];
    P.recv_sizes[k] = d_recv_sizes ? d_recv_sizes[k] : nullptr;
  }
  const uint32_t call = ++c->sizes_calls;
  const int slot = call % kSnapSlots;
  c->n_of_call[slot] = n;
  return comm_submit(c, kOpA2aSizes, 2, stream, [&](int phase) -> int {
    if (phase == 0) {
      KernelScope ks(HB_K_A2A_SIZES, stream);
      a2a_sizes_publish_kernel<<<1, 256, 0, stream>>>(P, peer_ptrs(c), c->rank, c->world, n, call);
      HB_CUDA_OK(cudaGetLastError());
      return HB_OK;
    }
    {
      KernelScope ks(HB_K_A2A_SIZES, stream);
      a2a_sizes_collect_kernel<<<1, 256, 0, stream>>>(P, peer_ptrs(c), c->rank, c->world, n, call, slot,
                                                       c->d_status);
    }
    HB_CUDA_OK(cudaGetLastError());
    if (h_recv_sizes != nullptr) {
      Control* ctl = reinterpret_cast<Control*>(c->base);
      HB_CUDA_OK(cudaMemcpyAsync(h_recv_sizes, ctl->snap[slot].recv_sizes,
                                 sizeof(int32_t) * (size_t)n * c->world, cudaMemcpyDeviceToHost, stream));
    }
    return HB_OK;
  });
}

int hbAlltoallvN(hbComm* c, int n, const void* const* d_inputs, const int64_t* common_sizes,
                 const int32_t* elem_bytes, void* const* d_outputs, int32_t* d_status,
                 hbStream stream_) {
  using namespace hb;
  cudaStream_t stream = (cudaStream_t)stream_;
  HB_REQUIRE(c && c->connected, "hbAlltoallvN: communicator not connected");
  HB_REQUIRE(c->data_calls < c->sizes_calls, "hbAlltoallvN: no pending hbAlltoallvNSizes to pair with");
  HB_REQUIRE(d_inputs && d_outputs && common_sizes && elem_bytes, "hbAlltoallvN: null argument");
  const uint32_t call = c->data_calls + 1;
  const int slot = call % kSnapSlots;
  HB_REQUIRE(c->n_of_call[slot] == n, "hbAlltoallvN: N=%d differs from the paired size exchange (N=%d)",
             n, c->n_of_call[slot]);
  TableParams T;
  PushParams PP;
  PullParams QP;
  for (int k = 0; k < n; ++k) {
    HB_REQUIRE(common_sizes[k] >= 0 && elem_bytes[k] >= 1, "hbAlltoallvN: bad common size / element size for tensor %d", k);
    T.row_bytes[k] = (uint64_t)common_sizes[k] * (uint64_t)elem_bytes[k];
    PP.inputs[k] = reinterpret_cast<const unsigned char*>(d_inputs[k]);
    QP.outputs[k] = reinterpret_cast<unsigned char*>(d_outputs[k]);
  }
  HB_REQUIRE(c->window_bytes > c->reserved_bytes, "hbAlltoallvN: no window space left beside the sharded plan");
  c->data_calls = call;
  const uint32_t epoch = ++c->win_seq;
  const int half = epoch & 1;
  const uint64_t half_bytes = ((c->window_bytes - c->reserved_bytes) / 2) & ~(uint64_t)255;
  const uint64_t woff = control_bytes() + c->reserved_bytes;
  if (d_status == nullptr) d_status = c->d_status;
  const int grid = device_sm_count() * 2;
  return comm_submit(c, kOpA2aData, 2, stream, [&](int phase) -> int {
    PeerPtrs pp = peer_ptrs(c);
    if (phase == 0) {
      {
        KernelScope ks(HB_K_A2A_TABLES, stream);
        a2a_tables_kernel<<<1, 256, 0, stream>>>(T, pp, c->rank, c->world, n, slot, half_bytes, d_status);
      }
      HB_CUDA_OK(cudaGetLastError());
      KernelScope ks(HB_K_A2A_PUSH, stream);
      a2a_push_kernel<<<grid, 256, 0, stream>>>(PP, pp, c->rank, c->world, n, slot, half, woff, half_bytes, epoch);
      HB_CUDA_OK(cudaGetLastError());
      return HB_OK;
    }
    KernelScope ks(HB_K_A2A_COPYOUT, stream);
    a2a_copyout_kernel<<<grid, 256, 0, stream>>>(QP, pp, c->rank, c->world, n, slot, half, woff, half_bytes,
                                                  epoch, d_status);
    HB_CUDA_OK(cudaGetLastError());
    return HB_OK;
  });
}

int hbAllreduceSumF32(hbComm* c, const float* d_in, float* d_out, int64_t count, float scale,
                      int32_t* d_status, hbStream stream_) {
  using namespace hb;
  cudaStream_t stream = (cudaStream_t)stream_;
  HB_REQUIRE(c && c->connected, "hbAllreduceSumF32: communicator not connected");
  HB_REQUIRE(count >= 0 && (count == 0 || (d_in && d_out)), "hbAllreduceSumF32: bad arguments");
  HB_REQUIRE(c->window_bytes > c->reserved_bytes, "hbAllreduceSumF32: no window space left beside the sharded plan");
  const uint64_t half_bytes = ((c->window_bytes - c->reserved_bytes) / 2) & ~(uint64_t)255;
  const uint64_t slot_bytes = align_up((size_t)count * 4, 256);
  if (slot_bytes * (uint64_t)c->world > half_bytes) {
    set_last_error("hbAllreduceSumF32: %lld floats x %d ranks need %llu B, window half is %llu B",
                   (long long)count, c->world, (unsigned long long)(slot_bytes * c->world),
                   (unsigned long long)half_bytes);
    return HB_ERR_WORKSPACE;
  }
  const uint32_t epoch = ++c->win_seq;
  const int half = epoch & 1;
  const uint64_t woff = control_bytes() + c->reserved_bytes;
  if (d_status == nullptr) d_status = c->d_status;
  const uint64_t chunks = ((uint64_t)count * 4 + kChunkBytes - 1) / kChunkBytes * c->world;
  const int maxg = device_sm_count() * 2;
  const int grid_push = chunks < 1 ? 1 : (chunks < (uint64_t)maxg ? (int)chunks : maxg);
  const int64_t rb = (count + 255) / 256;
  const int grid_red = rb < 1 ? 1 : (rb < maxg ? (int)rb : maxg);
  return comm_submit(c, kOpAllreduce, 2, stream, [&](int phase) -> int {
    PeerPtrs pp = peer_ptrs(c);
    KernelScope ks(phase == 0 ? HB_K_AR_PUSH : HB_K_AR_REDUCE, stream);
    if (phase == 0)
      ar_push_kernel<<<grid_push, 256, 0, stream>>>(d_in, count, slot_bytes, pp, c->rank, c->world, half, woff,
                                                    half_bytes, epoch);
    else
      ar_reduce_kernel<<<grid_red, 256, 0, stream>>>(d_out, count, slot_bytes, scale, pp, c->rank, c->world,
                                                     half, woff, half_bytes, epoch, d_status);
    HB_CUDA_OK(cudaGetLastError());
    return HB_OK;
  });
}

}  // extern "C"

// ---- bootstrap by a broadcast 128-byte id (the reference's protocol) -----------------------
// The reference creates its communicator from ONE 128-byte id that rank 0 generates
// (HbGetNcclId, nccl_get_id.cc:35-70) and the Python side broadcasts
// (distribute/collective.py:108-115); HbCreateNcclCollective(handle, id) then runs on every
// rank (nccl_create.cc:45-62).  Our tokens are per rank (an IPC handle each), so the id
// names a rendezvous directory on the node (one NVSwitch domain = one node): every rank
// writes its token file there and reads the others'.
static std::string rendezvous_dir(const unsigned char* id) {
  const char* root = getenv("HB_B200_RENDEZVOUS_DIR");
  std::string d = (root != nullptr && root[0]) ? root : "/dev/shm";
  char tag[40];
  uint64_t a, b;
  memcpy(&a, id + 8, 8);
  memcpy(&b, id + 16, 8);
  snprintf(tag, sizeof(tag), "/hb_b200_%016llx%016llx", (unsigned long long)a, (unsigned long long)b);
  return d + tag;
}

extern "C" int hbGetUniqueId(unsigned char id_out[HB_COMM_TOKEN_BYTES]) {
  using namespace hb;
  HB_REQUIRE(id_out, "hbGetUniqueId: null argument");
  memset(id_out, 0, HB_COMM_TOKEN_BYTES);
  const uint32_t magic = kTokenMagic ^ 0x1d1d1d1du;
  memcpy(id_out, &magic, 4);
  uint64_t r[2] = {0, 0};
  FILE* f = fopen("/dev/urandom", "rb");
  if (f != nullptr) {
    if (fread(r, 1, sizeof(r), f) != sizeof(r)) r[0] = 0;
    fclose(f);
  }
  if (r[0] == 0) {
    r[0] = (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count();
    r[1] = ((uint64_t)getpid() << 32) ^ (uint64_t)(uintptr_t)id_out;
  }
  memcpy(id_out + 8, r, 16);
  return HB_OK;
}

extern "C" int hbCommCreateFromId(const unsigned char id[HB_COMM_TOKEN_BYTES], int rank, int world_size,
                                  int local_size, size_t window_bytes, hbComm** comm) {
  using namespace hb;
  HB_REQUIRE(id && comm, "hbCommCreateFromId: null argument");
  uint32_t magic;
  memcpy(&magic, id, 4);
  HB_REQUIRE(magic == (kTokenMagic ^ 0x1d1d1d1du), "hbCommCreateFromId: not an id made by hbGetUniqueId");
  unsigned char token[HB_COMM_TOKEN_BYTES];
  int rc = hbCommCreate(rank, world_size, local_size, window_bytes, comm, token);
  if (rc != HB_OK) return rc;
  if (world_size == 1) return HB_OK;
  const std::string dir = rendezvous_dir(id);
  if (mkdir(dir.c_str(), 0700) != 0 && errno != EEXIST) {
    set_last_error("hbCommCreateFromId: cannot create %s: %s", dir.c_str(), strerror(errno));
    hbCommDestroy(*comm);
    *comm = nullptr;
    return HB_ERR_COMM;
  }
  auto path = [&](int q, const char* suffix) { return dir + "/" + std::to_string(q) + suffix; };
  {  // write-then-rename: a reader never sees a partial token
    const std::string tmp = path(rank, ".tmp"), fin = path(rank, ".tok");
    FILE* f = fopen(tmp.c_str(), "wb");
    const bool ok = f != nullptr && fwrite(token, 1, HB_COMM_TOKEN_BYTES, f) == HB_COMM_TOKEN_BYTES;
    if (f != nullptr) fclose(f);
    if (!ok || rename(tmp.c_str(), fin.c_str()) != 0) {
      set_last_error("hbCommCreateFromId: cannot publish the token in %s: %s", dir.c_str(), strerror(errno));
      hbCommDestroy(*comm);
      *comm = nullptr;
      return HB_ERR_COMM;
    }
  }
  std::string all((size_t)world_size * HB_COMM_TOKEN_BYTES, '\0');
  const auto t0 = std::chrono::steady_clock::now();
  for (int q = 0; q < world_size; ++q) {
    while (true) {
      FILE* f = fopen(path(q, ".tok").c_str(), "rb");
      if (f != nullptr) {
        const size_t got = fread(&all[(size_t)q * HB_COMM_TOKEN_BYTES], 1, HB_COMM_TOKEN_BYTES, f);
        fclose(f);
        if (got == HB_COMM_TOKEN_BYTES) break;
      }
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) {
        set_last_error("hbCommCreateFromId: rank %d never published its token in %s (120 s)", q, dir.c_str());
        hbCommDestroy(*comm);
        *comm = nullptr;
        return HB_ERR_COMM;
      }
      usleep(2000);
    }
  }
  rc = hbCommConnect(*comm, reinterpret_cast<const unsigned char*>(all.data()));
  if (rc != HB_OK) {
    hbCommDestroy(*comm);
    *comm = nullptr;
    return rc;
  }
  // every rank has mapped every window once all ".ok" files exist; the last one cleans up
  { FILE* f = fopen(path(rank, ".ok").c_str(), "wb"); if (f != nullptr) fclose(f); }
  bool all_ok = true;
  for (int q = 0; q < world_size; ++q) all_ok = all_ok && access(path(q, ".ok").c_str(), F_OK) == 0;
  if (all_ok) {
    for (int q = 0; q < world_size; ++q) { unlink(path(q, ".tok").c_str()); unlink(path(q, ".ok").c_str()); }
    rmdir(dir.c_str());
  }
  return HB_OK;
}
