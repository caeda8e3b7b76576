// nts_exchange: the data plane of the device-resident partition-boundary exchange, in C++ behind the C ABI.
//
// Replaces the GPU drivers of the reference engine - Graph::sync_compute_decoupled (forward, core/graph.hpp:3639-3719)
// and Graph::compute_sync_decoupled (backward, :3455-3622) - and the host-staged NtsGraphCommunicator they drive
// (comm/network.cpp:159-844) for the peer-memory ("p2p") transport:
//
//   forward   publish my rows in a CUDA-IPC window -> aggregate the local chunk while the receiver side PULLS the rows
//             it needs out of every peer's window over NVLink (nts_gather_rows on mapped peer pointers, side stream)
//             -> ONE launch over the merged CSC of all remote chunks.
//   backward  ONE launch computes the partial gradients of the active sources of all remote chunks straight into my
//             window -> publish -> local chunk while the slices the peers computed for me are copied out of their
//             windows -> one scatter-add.
//
// Cross-GPU ordering: epoch-numbered flags in peer memory (published[rank], consumed[rank][peer]) written / awaited by
// tiny kernels with release / acquire semantics at system scope; everything else is stream order + two events.
// The CONTROL plane (exchanging row lists and IPC handles between ranks) stays with the caller - torch.distributed in
// this repo, MPI in the reference's host code - so this file has no dependency on either.
#include <vector>

#include "nts_common.cuh"

struct nts_exchange {
  nts_exchange_desc d;
  std::vector<uint32_t> need_count, send_count, recv_offs, peer_bwd_offset;
  std::vector<const uint32_t *> need;
  float *window = nullptr;
  size_t capacity_floats = 0;
  uint32_t *flags = nullptr;                 // [1 + P]: published, consumed[peer]
  std::vector<float *> peer_window;          // opened IPC mappings (own entry = window)
  std::vector<uint32_t *> peer_flags;
  uint32_t **d_peer_flags = nullptr;         // device copy of peer_flags for the signalling kernel
  bool peers_open = false;
  uint32_t epoch = 0;
  float *recv = nullptr;                     // receive staging (forward) / pulled slices (backward)
  size_t recv_cap = 0;
  cudaStream_t comm = nullptr;
  cudaEvent_t ev_main = nullptr, ev_comm = nullptr;
};

namespace nts {

__global__ void wait_all_consumed_kernel(const uint32_t *flags, int P, int p, uint32_t value) {
  const int j = threadIdx.x;
  if (j >= P || j == p)
    return;
  uint32_t v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + 1 + j) : "memory");
    if (v < value)
      __nanosleep(200);
  } while (v < value);
}

__global__ void signal_consumed_kernel(uint32_t *const *peer_flags, int P, int p, uint32_t value) {
  const int j = threadIdx.x;
  if (j >= P || j == p)
    return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[j] + 1 + p), "r"(value) : "memory");
}

static int grow(float **buf, size_t *cap, size_t floats) {
  if (floats <= *cap)
    return 0;
  if (*buf)
    NTS_CUDA_OK(cudaFree(*buf));
  *buf = nullptr;
  NTS_CUDA_OK(cudaMalloc(reinterpret_cast<void **>(buf), floats * sizeof(float)));
  *cap = floats;
  return 0;
}

} // namespace nts

using namespace nts;

extern "C" {

nts_exchange *nts_exchange_create(const nts_exchange_desc *desc) {
  if (!desc || desc->partitions < 1 || desc->rank < 0 || desc->rank >= desc->partitions) {
    fail(-1, "bad exchange descriptor", __FILE__, __LINE__);
    return nullptr;
  }
  nts_exchange *ex = new nts_exchange();
  ex->d = *desc;
  const int P = desc->partitions;
  ex->need_count.assign(desc->need_count, desc->need_count + P);
  ex->send_count.assign(desc->send_count, desc->send_count + P);
  ex->peer_bwd_offset.assign(desc->peer_bwd_offset, desc->peer_bwd_offset + P);
  ex->need.assign(desc->need, desc->need + P);
  ex->recv_offs.assign(P + 1, 0);
  for (int i = 0; i < P; i++)
    ex->recv_offs[i + 1] = ex->recv_offs[i] + (i == desc->rank ? 0u : ex->need_count[i]);
  ex->peer_window.assign(P, nullptr);
  ex->peer_flags.assign(P, nullptr);
  bool ok = cudaMalloc(reinterpret_cast<void **>(&ex->flags), sizeof(uint32_t) * (1 + P)) == cudaSuccess &&
            cudaMemset(ex->flags, 0, sizeof(uint32_t) * (1 + P)) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void **>(&ex->d_peer_flags), sizeof(uint32_t *) * P) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ex->comm, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&ex->ev_main, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ex->ev_comm, cudaEventDisableTiming) == cudaSuccess &&
            cudaDeviceSynchronize() == cudaSuccess;
  if (!ok) {
    fail(-1, "exchange resource allocation failed", __FILE__, __LINE__);
    delete ex;
    return nullptr;
  }
  return ex;
}

int nts_exchange_destroy(nts_exchange *ex) {
  if (!ex)
    return 0;
  cudaDeviceSynchronize();
  if (ex->peers_open)
    for (int j = 0; j < ex->d.partitions; j++)
      if (j != ex->d.rank) {
        cudaIpcCloseMemHandle(ex->peer_window[j]);
        cudaIpcCloseMemHandle(ex->peer_flags[j]);
      }
  cudaFree(ex->window);
  cudaFree(ex->flags);
  cudaFree(ex->d_peer_flags);
  cudaFree(ex->recv);
  cudaStreamDestroy(ex->comm);
  cudaEventDestroy(ex->ev_main);
  cudaEventDestroy(ex->ev_comm);
  delete ex;
  return 0;
}

// Rows the window must hold for feature width F: my own rows (forward) or the partials for all peers (backward).
uint64_t nts_exchange_required_floats(const nts_exchange *ex, nts_vid_t feature_size) {
  uint64_t rows = ex->d.owned_vertices;
  if (ex->d.recv_total > rows)
    rows = ex->d.recv_total;
  if (rows == 0)
    rows = 1;
  return rows * (uint64_t)feature_size;
}

// (Re)allocate the exported window.  COLLECTIVE in effect: after it returns 1 on any rank, every rank must exchange
// the new handles and call nts_exchange_open_peers again before the next forward/backward.
int nts_exchange_reserve(nts_exchange *ex, uint64_t floats, int *reallocated) {
  NTS_ARG_CHECK(ex && reallocated, "null argument");
  *reallocated = 0;
  if (floats <= ex->capacity_floats)
    return 0;
  NTS_CUDA_OK(cudaDeviceSynchronize());
  if (ex->peers_open) {
    for (int j = 0; j < ex->d.partitions; j++)
      if (j != ex->d.rank) {
        NTS_CUDA_OK(cudaIpcCloseMemHandle(ex->peer_window[j]));
        NTS_CUDA_OK(cudaIpcCloseMemHandle(ex->peer_flags[j]));
      }
    ex->peers_open = false;
  }
  if (ex->window)
    NTS_CUDA_OK(cudaFree(ex->window));
  ex->window = nullptr;
  NTS_CUDA_OK(cudaMalloc(reinterpret_cast<void **>(&ex->window), floats * sizeof(float)));
  ex->capacity_floats = floats;
  *reallocated = 1;
  return 0;
}

int nts_exchange_handles(nts_exchange *ex, unsigned char window_handle[NTS_IPC_HANDLE_BYTES],
                         unsigned char flags_handle[NTS_IPC_HANDLE_BYTES]) {
  NTS_ARG_CHECK(ex && ex->window && ex->flags, "window not allocated");
  cudaIpcMemHandle_t h;
  NTS_CUDA_OK(cudaIpcGetMemHandle(&h, ex->window));
  memcpy(window_handle, &h, sizeof(h));
  NTS_CUDA_OK(cudaIpcGetMemHandle(&h, ex->flags));
  memcpy(flags_handle, &h, sizeof(h));
  return 0;
}

// handles: P consecutive 64-byte window handles and P consecutive flag handles (own entries ignored)
int nts_exchange_open_peers(nts_exchange *ex, const unsigned char *window_handles, const unsigned char *flag_handles) {
  NTS_ARG_CHECK(ex && window_handles && flag_handles, "null argument");
  const int P = ex->d.partitions, p = ex->d.rank;
  for (int j = 0; j < P; j++) {
    if (j == p) {
      ex->peer_window[j] = ex->window;
      ex->peer_flags[j] = ex->flags;
      continue;
    }
    cudaIpcMemHandle_t h;
    void *ptr = nullptr;
    memcpy(&h, window_handles + (size_t)j * NTS_IPC_HANDLE_BYTES, sizeof(h));
    NTS_CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    ex->peer_window[j] = static_cast<float *>(ptr);
    memcpy(&h, flag_handles + (size_t)j * NTS_IPC_HANDLE_BYTES, sizeof(h));
    NTS_CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    ex->peer_flags[j] = static_cast<uint32_t *>(ptr);
  }
  NTS_CUDA_OK(cudaMemcpy(ex->d_peer_flags, ex->peer_flags.data(), sizeof(uint32_t *) * P, cudaMemcpyHostToDevice));
  ex->peers_open = true;
  return 0;
}

static int begin_epoch(nts_exchange *ex, cudaStream_t st, uint32_t *epoch) {
  ex->epoch += 1;
  *epoch = ex->epoch;
  if (ex->epoch > 1 && ex->d.partitions > 1) { // every peer is done reading what I published last time
    wait_all_consumed_kernel<<<1, 32 * ((ex->d.partitions + 31) / 32), 0, st>>>(ex->flags, ex->d.partitions, ex->d.rank,
                                                                               ex->epoch - 1);
    NTS_LAUNCH_CHECK();
  }
  return 0;
}

static int signal_consumed(nts_exchange *ex, uint32_t epoch, cudaStream_t st) {
  signal_consumed_kernel<<<1, 32 * ((ex->d.partitions + 31) / 32), 0, st>>>(ex->d_peer_flags, ex->d.partitions,
                                                                           ex->d.rank, epoch);
  NTS_LAUNCH_CHECK();
  return 0;
}

#define NTS_TRY(expr)                                                                                               \
  do {                                                                                                              \
    int nts_rc_ = (expr);                                                                                           \
    if (nts_rc_ != 0)                                                                                               \
      return nts_rc_;                                                                                               \
  } while (0)

// Y_p += sum_i A_{p<-i} X_i.  `y` must be zeroed by the caller (accumulate semantics, like every aggregation entry).
int nts_exchange_forward(nts_exchange *ex, const float *x, float *y, nts_vid_t F, void *stream) {
  NTS_ARG_CHECK(ex != nullptr, "null engine");
  const nts_exchange_desc &d = ex->d;
  // a rank that owns no vertices (the 1024-aligned partitioner leaves such ranks on small graphs) has no rows to
  // publish or produce, but still takes part in the flag protocol below
  NTS_ARG_CHECK(d.owned_vertices == 0 || (x && y), "null feature pointer");
  cudaStream_t st = as_stream(stream);
  const int P = d.partitions, p = d.rank;
  if (P == 1)
    return nts_gather_by_dst_from_src(x, y, d.local_weight_forward, d.local_row_indices, d.local_column_offset,
                                      d.dst_start, d.dst_start + d.owned_vertices, d.dst_start,
                                      d.dst_start + d.owned_vertices, d.local_edges, d.owned_vertices, F, 1, stream);
  NTS_ARG_CHECK(ex->peers_open && nts_exchange_required_floats(ex, F) <= ex->capacity_floats,
                "exchange window not reserved / peers not opened for this feature width");
  uint32_t epoch = 0;
  NTS_TRY(begin_epoch(ex, st, &epoch));
  NTS_CUDA_OK(cudaMemcpyAsync(ex->window, x, (size_t)d.owned_vertices * F * sizeof(float), cudaMemcpyDeviceToDevice, st));
  NTS_TRY(nts_signal_set(ex->flags, epoch, st));                     // published
  NTS_CUDA_OK(cudaEventRecord(ex->ev_main, st));
  NTS_CUDA_OK(cudaStreamWaitEvent(ex->comm, ex->ev_main, 0));
  NTS_TRY(grow(&ex->recv, &ex->recv_cap, (size_t)(d.recv_total ? d.recv_total : 1) * F));
  for (int s = 1; s < P; s++) {                                      // the reference's ring order (p+1, p+2, ...)
    const int i = (p + s) % P;
    const uint32_t n = ex->need_count[i];
    if (!n)
      continue;
    NTS_TRY(nts_signal_wait_geq(ex->peer_flags[i], epoch, ex->comm));
    NTS_TRY(nts_gather_rows(ex->recv + (size_t)ex->recv_offs[i] * F, ex->peer_window[i], ex->need[i], n, F, ex->comm));
  }
  NTS_TRY(signal_consumed(ex, epoch, ex->comm));
  NTS_CUDA_OK(cudaEventRecord(ex->ev_comm, ex->comm));
  // local chunk overlaps with the pulls
  NTS_TRY(nts_gather_by_dst_from_src(x, y, d.local_weight_forward, d.local_row_indices, d.local_column_offset, d.dst_start,
                                     d.dst_start + d.owned_vertices, d.dst_start, d.dst_start + d.owned_vertices,
                                     d.local_edges, d.owned_vertices, F, 1, st));
  NTS_CUDA_OK(cudaStreamWaitEvent(st, ex->ev_comm, 0));
  if (d.remote_edges)
    NTS_TRY(nts_segment_gather_sum(ex->recv, y, d.remote_weight, d.remote_slots, d.remote_column_offset, 0,
                                   d.owned_vertices, d.remote_edges, F, st));
  return 0;
}

// dX_p += sum_j A_{j<-p}^T dY_j.  `dx` must be zeroed by the caller.
int nts_exchange_backward(nts_exchange *ex, const float *g, float *dx, nts_vid_t F, void *stream) {
  NTS_ARG_CHECK(ex != nullptr, "null engine");
  const nts_exchange_desc &d = ex->d;
  NTS_ARG_CHECK(d.owned_vertices == 0 || (g && dx), "null gradient pointer");
  cudaStream_t st = as_stream(stream);
  const int P = d.partitions, p = d.rank;
  if (P == 1)
    return nts_gather_by_src_from_dst(g, dx, d.local_weight_backward, d.local_row_offset, d.local_column_indices,
                                      d.dst_start, d.dst_start + d.owned_vertices, d.dst_start,
                                      d.dst_start + d.owned_vertices, d.local_edges, d.owned_vertices, F, 1, stream);
  NTS_ARG_CHECK(ex->peers_open && nts_exchange_required_floats(ex, F) <= ex->capacity_floats,
                "exchange window not reserved / peers not opened for this feature width");
  uint32_t epoch = 0;
  NTS_TRY(begin_epoch(ex, st, &epoch));
  if (d.recv_total) {
    NTS_CUDA_OK(cudaMemsetAsync(ex->window, 0, (size_t)d.recv_total * F * sizeof(float), st));
    if (d.remote_edges)
      NTS_TRY(nts_segment_gather_sum(g, ex->window, d.backward_weight, d.backward_indices, d.backward_offsets,
                                     d.dst_start, d.recv_total, d.remote_edges, F, st));
  }
  NTS_TRY(nts_signal_set(ex->flags, epoch, st));                     // published
  NTS_CUDA_OK(cudaEventRecord(ex->ev_main, st));
  NTS_CUDA_OK(cudaStreamWaitEvent(ex->comm, ex->ev_main, 0));
  NTS_TRY(grow(&ex->recv, &ex->recv_cap, (size_t)(d.send_total ? d.send_total : 1) * F));
  size_t pos = 0;
  for (int j = 0; j < P; j++) {                                      // staging order = order of send_rows_all
    const uint32_t n = ex->send_count[j];
    if (j == p || !n)
      continue;
    NTS_TRY(nts_signal_wait_geq(ex->peer_flags[j], epoch, ex->comm));
    NTS_CUDA_OK(cudaMemcpyAsync(ex->recv + pos * F, ex->peer_window[j] + (size_t)ex->peer_bwd_offset[j] * F,
                                (size_t)n * F * sizeof(float), cudaMemcpyDeviceToDevice, ex->comm));
    pos += n;
  }
  NTS_TRY(signal_consumed(ex, epoch, ex->comm));
  NTS_CUDA_OK(cudaEventRecord(ex->ev_comm, ex->comm));
  NTS_TRY(nts_gather_by_src_from_dst(g, dx, d.local_weight_backward, d.local_row_offset, d.local_column_indices,
                                     d.dst_start, d.dst_start + d.owned_vertices, d.dst_start,
                                     d.dst_start + d.owned_vertices, d.local_edges, d.owned_vertices, F, 1, st));
  NTS_CUDA_OK(cudaStreamWaitEvent(st, ex->ev_comm, 0));
  if (d.send_total)
    NTS_TRY(nts_scatter_add_rows_atomic(dx, ex->recv, d.send_rows_all, d.send_total, F, st));
  return 0;
}

} // extern "C"
