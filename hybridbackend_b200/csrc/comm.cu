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
#include <string.h>
#include <unistd.h>

#include "comm.cuh"

namespace hb {

// ---- barrier -----------------------------------------------------------------
__global__ void barrier_kernel(PeerPtrs peers, int me, int world, uint32_t epoch) {
  const int q = threadIdx.x;
  if (q < world) {
    __threadfence_system();
    Control* remote = reinterpret_cast<Control*>(peers.p[q]);
    st_release_sys_u32(&remote->barrier_flags[me], epoch);
    Control* mine = reinterpret_cast<Control*>(peers.p[me]);
    wait_flag(&mine->barrier_flags[q], epoch);
  }
}

// ---- alltoallv phase 1: sizes all-gather + segment tables ---------------------
struct SizesParams {
  const int32_t* send_sizes[kMaxA2aTensors];
  int32_t* recv_sizes[kMaxA2aTensors];
};

__global__ void __launch_bounds__(256)
a2a_sizes_kernel(const __grid_constant__ SizesParams P, PeerPtrs peers, int me, int world, int n,
                 uint32_t call, int slot, uint64_t half_bytes) {
  const int parity = call & 1;
  Control* mine = reinterpret_cast<Control*>(peers.p[me]);
  // 1. publish my N x W send sizes into every peer's mailbox row `me`
  for (int i = threadIdx.x; i < n * world; i += blockDim.x) {
    const int k = i / world, r = i % world;
    const int32_t v = P.send_sizes[k][r];
    for (int q = 0; q < world; ++q) {
      Control* remote = reinterpret_cast<Control*>(peers.p[q]);
      remote->mailbox[parity][(me * kMaxA2aTensors + k) * kMaxWorld + r] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    Control* remote = reinterpret_cast<Control*>(peers.p[threadIdx.x]);
    st_release_sys_u32(&remote->sizes_flags[parity][me], call);
    wait_flag(&mine->sizes_flags[parity][threadIdx.x], call);
  }
  __syncthreads();
  // 2. snapshot the matrix
  Snapshot* S = &mine->snap[slot];
  for (int i = threadIdx.x; i < world * n * world; i += blockDim.x) {
    const int q = i / (n * world), k = (i / world) % n, r = i % world;
    const int idx = (q * kMaxA2aTensors + k) * kMaxWorld + r;
    S->matrix[idx] = *reinterpret_cast<volatile int32_t*>(&mine->mailbox[parity][idx]);
  }
  __syncthreads();
  // 3. recv sizes for the caller
  for (int i = threadIdx.x; i < n * world; i += blockDim.x) {
    const int k = i / world, q = i % world;
    const int32_t v = S->matrix[(q * kMaxA2aTensors + k) * kMaxWorld + me];
    S->recv_sizes[k * world + q] = v;
    if (P.recv_sizes[k] != nullptr) P.recv_sizes[k][q] = v;
  }
  (void)half_bytes;
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

__global__ void __launch_bounds__(256)
a2a_push_kernel(const __grid_constant__ PushParams P, PeerPtrs peers, int me, int world, int n,
                int slot, int half, uint64_t window_off, uint64_t half_bytes, uint32_t call) {
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
  // publish: the last CTA to finish releases the epoch flag on every peer
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
    st_release_sys_u32(&remote->data_flags[half][me], call);
  }
}

struct PullParams {
  unsigned char* outputs[kMaxA2aTensors];
};

__global__ void __launch_bounds__(256)
a2a_copyout_kernel(const __grid_constant__ PullParams P, PeerPtrs peers, int me, int world, int n,
                   int slot, int half, uint64_t window_off, uint64_t half_bytes, uint32_t call) {
  Control* mine = reinterpret_cast<Control*>(peers.p[me]);
  if ((int)threadIdx.x < world) {
    wait_flag(&mine->data_flags[half][threadIdx.x], call);
  }
  __syncthreads();
  const Snapshot* S = &mine->snap[slot];
  if (S->overflow) return;
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

int hbCommDestroy(hbComm* c) {
  if (!c) return HB_OK;
  cudaDeviceSynchronize();
  for (int q = 0; q < c->world; ++q)
    if (c->opened[q]) cudaIpcCloseMemHandle(c->peer[q]);
  if (c->base) cudaFree(c->base);
  delete c;
  return HB_OK;
}

int hbCommRank(const hbComm* c) { return c ? c->rank : -1; }
int hbCommWorldSize(const hbComm* c) { return c ? c->world : -1; }
void* hbCommWindow(hbComm* c) { return c ? c->base + hb::control_bytes() : nullptr; }
size_t hbCommWindowBytes(const hbComm* c) { return c ? c->window_bytes : 0; }

int hbCommBarrier(hbComm* c, hbStream stream) {
  using namespace hb;
  HB_REQUIRE(c && c->connected, "hbCommBarrier: communicator not connected");
  c->barrier_epoch++;
  KernelScope ks(HB_K_BARRIER, (cudaStream_t)stream);
  barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(peer_ptrs(c), c->rank, c->world, c->barrier_epoch);
  HB_CUDA_OK(cudaGetLastError());
  return HB_OK;
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
  {
    KernelScope ks(HB_K_A2A_SIZES, stream);
    a2a_sizes_kernel<<<1, 256, 0, stream>>>(P, peer_ptrs(c), c->rank, c->world, n, call, slot,
                                             c->window_bytes / 2);
  }
  HB_CUDA_OK(cudaGetLastError());
  if (h_recv_sizes != nullptr) {
    Control* ctl = reinterpret_cast<Control*>(c->base);
    HB_CUDA_OK(cudaMemcpyAsync(h_recv_sizes, ctl->snap[slot].recv_sizes,
                               sizeof(int32_t) * (size_t)n * c->world, cudaMemcpyDeviceToHost, stream));
  }
  return HB_OK;
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
  c->data_calls = call;
  const int half = call & 1;
  HB_REQUIRE(c->window_bytes > c->reserved_bytes, "hbAlltoallvN: no window space left beside the sharded plan");
  const uint64_t half_bytes = ((c->window_bytes - c->reserved_bytes) / 2) & ~(uint64_t)255;
  const uint64_t woff = control_bytes() + c->reserved_bytes;
  PeerPtrs pp = peer_ptrs(c);
  {
    KernelScope ks(HB_K_A2A_TABLES, stream);
    a2a_tables_kernel<<<1, 256, 0, stream>>>(T, pp, c->rank, c->world, n, slot, half_bytes, d_status);
  }
  HB_CUDA_OK(cudaGetLastError());
  const int grid = device_sm_count() * 2;
  {
    KernelScope ks(HB_K_A2A_PUSH, stream);
    a2a_push_kernel<<<grid, 256, 0, stream>>>(PP, pp, c->rank, c->world, n, slot, half, woff, half_bytes, call);
  }
  HB_CUDA_OK(cudaGetLastError());
  {
    KernelScope ks(HB_K_A2A_COPYOUT, stream);
    a2a_copyout_kernel<<<grid, 256, 0, stream>>>(QP, pp, c->rank, c->world, n, slot, half, woff, half_bytes, call);
  }
  HB_CUDA_OK(cudaGetLastError());
  return HB_OK;
}

}  // extern "C"
