// Host-side builder of the exchange plan (the index structures nts_exchange_create() consumes) from the reference's
// per-source-partition chunks (`CSC_segment_pinned`, core/GraphSegment.h:52-139, built by
// PartitionedGraph::PartitionToChunks, core/PartitionedGraph.hpp:324-420).  C++ twin of
// neutronstarlite_b200/exchange.py::ExchangePlan, so that the reference's C++ host code can drive the peer-memory
// exchange without Python: include/nts_dropin/core/ntsDistGPUFusedGraphOp.hpp is the caller.
//
// The builder is pure host code until nts_exchange_create_from_plan() uploads the arrays.  The control plane stays
// with the caller: it moves every rank's packed need lists (nts_exchange_plan_pack_needs) to every other rank with
// whatever transport it has (MPI in the reference) and feeds them back through nts_exchange_plan_set_peer_needs.
#include "nts_common.cuh"

#include <algorithm>
#include <vector>

struct nts_exchange_plan {
  int P = 0, p = 0;
  std::vector<nts_host_chunk> chunks;
  std::vector<std::vector<uint32_t>> need;            // [P] rows of partition i with at least one edge into mine
  std::vector<std::vector<uint32_t>> send_rows;       // [P] rows of mine that rank j reads
  std::vector<std::vector<uint32_t>> peer_need_count; // [P][P] every rank's need counts (own entries 0)
  std::vector<char> have_peer;
  bool finalized = false;
  // merged arrays (host)
  std::vector<uint32_t> need_count, send_count, peer_bwd_offset, bwd_push_offset, recv_offs;
  // per remote chunk: row_indices as ranks in the need list, row_offset restricted to the active sources
  std::vector<std::vector<uint32_t>> chunk_slots, chunk_offc;
  std::vector<nts_exchange_chunk> dev_chunks;
  std::vector<uint32_t> remote_col_offset, remote_slots, bwd_offsets, bwd_indices, send_rows_all;
  std::vector<float> remote_w, bwd_w;
  uint64_t remote_edges = 0;
  uint32_t recv_total = 0, send_total = 0;
  // device copies (owned)
  std::vector<void *> dev_allocs;
  std::vector<const uint32_t *> need_dev;
};

namespace {

using nts::fail;

template <class T> int upload(nts_exchange_plan *pl, const std::vector<T> &h, const T **out) {
  *out = nullptr;
  if (h.empty())
    return 0;
  void *d = nullptr;
  NTS_CUDA_OK(cudaMalloc(&d, h.size() * sizeof(T)));
  pl->dev_allocs.push_back(d);
  NTS_CUDA_OK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = static_cast<const T *>(d);
  return 0;
}

} // namespace

extern "C" {

nts_exchange_plan *nts_exchange_plan_create(const nts_host_chunk *chunks, int partitions, int rank) {
  if (!chunks || partitions < 1 || rank < 0 || rank >= partitions) {
    fail(-1, "bad arguments to nts_exchange_plan_create", __FILE__, __LINE__);
    return nullptr;
  }
  for (int i = 0; i < partitions; i++) {
    const nts_host_chunk &c = chunks[i];
    const bool ok = c.src_end >= c.src_start && c.dst_end >= c.dst_start && c.edges < 0xffffffffull &&
                    c.dst_start == chunks[rank].dst_start && c.dst_end == chunks[rank].dst_end &&
                    (c.edges == 0 || (c.column_offset && c.row_indices && c.row_offset && c.column_indices &&
                                      c.edge_weight_forward && c.edge_weight_backward)) &&
                    (c.src_end == c.src_start || c.row_offset) && (c.dst_end == c.dst_start || c.column_offset);
    if (!ok) {
      fail(-1, "inconsistent chunk passed to nts_exchange_plan_create", __FILE__, __LINE__);
      return nullptr;
    }
  }
  nts_exchange_plan *pl = new nts_exchange_plan();
  pl->P = partitions;
  pl->p = rank;
  pl->chunks.assign(chunks, chunks + partitions);
  pl->need.resize(partitions);
  pl->send_rows.resize(partitions);
  pl->peer_need_count.assign(partitions, std::vector<uint32_t>(partitions, 0u));
  pl->have_peer.assign(partitions, 0);
  for (int i = 0; i < partitions; i++) {
    const nts_host_chunk &c = pl->chunks[i];
    const uint32_t rows = c.src_end - c.src_start;
    std::vector<uint32_t> &n = pl->need[i];
    for (uint32_t r = 0; r < rows; r++)
      if (c.row_offset[r + 1] > c.row_offset[r])
        n.push_back(r);
    pl->peer_need_count[rank][i] = i == rank ? 0u : (uint32_t)n.size();
  }
  pl->have_peer[rank] = 1;
  return pl;
}

void nts_exchange_plan_destroy(nts_exchange_plan *pl) {
  if (!pl)
    return;
  for (void *d : pl->dev_allocs)
    cudaFree(d);
  delete pl;
}

const nts_vid_t *nts_exchange_plan_need(const nts_exchange_plan *pl, int i, nts_vid_t *count) {
  if (!pl || i < 0 || i >= pl->P) {
    if (count)
      *count = 0;
    return nullptr;
  }
  if (count)
    *count = (nts_vid_t)pl->need[i].size();
  return pl->need[i].data();
}

uint64_t nts_exchange_plan_packed_rows(const nts_exchange_plan *pl) {
  uint64_t n = 0;
  if (pl)
    for (int i = 0; i < pl->P; i++)
      if (i != pl->p)
        n += pl->need[i].size();
  return n;
}

int nts_exchange_plan_pack_needs(const nts_exchange_plan *pl, nts_vid_t *need_counts, nts_vid_t *need_rows) {
  NTS_ARG_CHECK(pl && need_counts, "null argument");
  uint64_t pos = 0;
  for (int i = 0; i < pl->P; i++) {
    const uint32_t n = i == pl->p ? 0u : (uint32_t)pl->need[i].size();
    need_counts[i] = n;
    if (n) {
      NTS_ARG_CHECK(need_rows != nullptr, "need_rows is null");
      std::copy(pl->need[i].begin(), pl->need[i].end(), need_rows + pos);
    }
    pos += n;
  }
  return 0;
}

int nts_exchange_plan_set_peer_needs(nts_exchange_plan *pl, int j, const nts_vid_t *need_counts,
                                     const nts_vid_t *need_rows) {
  NTS_ARG_CHECK(pl && need_counts && j >= 0 && j < pl->P, "bad argument");
  NTS_ARG_CHECK(!pl->finalized, "plan already finalized");
  if (j == pl->p)
    return 0; // own lists are known
  uint64_t pos = 0;
  for (int i = 0; i < pl->P; i++) {
    const uint32_t n = i == j ? 0u : need_counts[i];
    pl->peer_need_count[j][i] = n;
    if (i == pl->p) {
      NTS_ARG_CHECK(n == 0 || need_rows != nullptr, "need_rows is null");
      const uint32_t mine = pl->chunks[pl->p].src_end - pl->chunks[pl->p].src_start;
      pl->send_rows[j].assign(need_rows + pos, need_rows + pos + n);
      for (uint32_t r : pl->send_rows[j])
        NTS_ARG_CHECK(r < mine, "peer asks for a row outside this partition");
    }
    pos += n;
  }
  pl->have_peer[j] = 1;
  return 0;
}

int nts_exchange_plan_finalize(nts_exchange_plan *pl) {
  NTS_ARG_CHECK(pl != nullptr, "null plan");
  if (pl->finalized)
    return 0;
  const int P = pl->P, p = pl->p;
  for (int j = 0; j < P; j++)
    NTS_ARG_CHECK(pl->have_peer[j], "nts_exchange_plan_set_peer_needs was not called for every peer");
  const uint32_t Vp = pl->chunks[p].dst_end - pl->chunks[p].dst_start;
  pl->need_count.assign(P, 0);
  pl->send_count.assign(P, 0);
  pl->peer_bwd_offset.assign(P, 0);
  pl->bwd_push_offset.assign(P, 0);
  pl->recv_offs.assign(P + 1, 0);
  for (int i = 0; i < P; i++) {
    pl->need_count[i] = i == p ? 0u : (uint32_t)pl->need[i].size();
    pl->send_count[i] = i == p ? 0u : (uint32_t)pl->send_rows[i].size();
    pl->recv_offs[i + 1] = pl->recv_offs[i] + pl->need_count[i];
    uint32_t before = 0; // rows peer i computes for the partitions before mine = start of my slice in its window
    for (int q = 0; q < p; q++)
      before += pl->peer_need_count[i][q];
    pl->peer_bwd_offset[i] = before;
    uint32_t grads_before = 0; // partial-gradient rows the ranks before me return to rank i = start of mine in its staging
    for (int q = 0; q < p; q++)
      if (q != i)
        grads_before += pl->peer_need_count[q][i];
    pl->bwd_push_offset[i] = grads_before;
  }
  // ---- per remote chunk: slots (rank of the source in the need list) and compact row offsets
  pl->chunk_slots.assign(P, {});
  pl->chunk_offc.assign(P, {});
  for (int i = 0; i < P; i++) {
    if (i == p)
      continue;
    const nts_host_chunk &c = pl->chunks[i];
    const std::vector<uint32_t> &need = pl->need[i];
    std::vector<uint32_t> rank_of(c.src_end - c.src_start, 0u);
    for (size_t k = 0; k < need.size(); k++)
      rank_of[need[k]] = (uint32_t)k;
    pl->chunk_slots[i].resize(c.edges);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < (int64_t)c.edges; e++)
      pl->chunk_slots[i][e] = rank_of[c.row_indices[e] - c.src_start];
    pl->chunk_offc[i].resize(need.size() + 1);
    for (size_t k = 0; k < need.size(); k++)
      pl->chunk_offc[i][k] = c.row_offset[need[k]];
    pl->chunk_offc[i][need.size()] = (uint32_t)c.edges;
  }
  pl->recv_total = pl->recv_offs[P];
  pl->send_total = 0;
  pl->send_rows_all.clear();
  for (int j = 0; j < P; j++)
    if (j != p) {
      pl->send_total += pl->send_count[j];
      pl->send_rows_all.insert(pl->send_rows_all.end(), pl->send_rows[j].begin(), pl->send_rows[j].end());
    }
  // ---- merged remote CSC: destination-major over ALL remote chunks; within a destination the chunks appear in
  // partition order and each chunk's edges keep their order (== stable sort by destination of the concatenation)
  uint64_t total = 0;
  for (int i = 0; i < P; i++)
    if (i != p)
      total += pl->chunks[i].edges;
  NTS_ARG_CHECK(total < 0xffffffffull, "remote edges of one rank must fit uint32 offsets");
  pl->remote_edges = total;
  pl->remote_col_offset.clear();
  pl->remote_slots.clear();
  pl->remote_w.clear();
  if (total) {
    pl->remote_col_offset.assign((size_t)Vp + 1, 0u);
    for (int i = 0; i < P; i++) {
      if (i == p || pl->chunks[i].edges == 0)
        continue;
      const uint32_t *co = pl->chunks[i].column_offset;
#pragma omp parallel for schedule(static)
      for (int64_t d = 0; d < (int64_t)Vp; d++)
        pl->remote_col_offset[d + 1] += co[d + 1] - co[d];
    }
    for (uint32_t d = 0; d < Vp; d++)
      pl->remote_col_offset[d + 1] += pl->remote_col_offset[d];
    pl->remote_slots.resize(total);
    pl->remote_w.resize(total);
    std::vector<uint32_t> cursor(pl->remote_col_offset.begin(), pl->remote_col_offset.end() - 1);
    for (int i = 0; i < P; i++) {
      const nts_host_chunk &c = pl->chunks[i];
      if (i == p || c.edges == 0)
        continue;
      std::vector<uint32_t> slot_of(c.src_end - c.src_start, 0u);
      for (size_t k = 0; k < pl->need[i].size(); k++)
        slot_of[pl->need[i][k]] = (uint32_t)k + pl->recv_offs[i];
#pragma omp parallel for schedule(dynamic, 1024)
      for (int64_t d = 0; d < (int64_t)Vp; d++) {
        uint32_t pos = cursor[d];
        for (uint32_t e = c.column_offset[d]; e < c.column_offset[d + 1]; e++, pos++) {
          pl->remote_slots[pos] = slot_of[c.row_indices[e] - c.src_start];
          pl->remote_w[pos] = c.edge_weight_forward[e];
        }
        cursor[d] = pos;
      }
    }
  }
  // ---- compact CSR over the active sources of all remote chunks (rows = the send-staging layout of the backward)
  pl->bwd_offsets.clear();
  pl->bwd_indices.clear();
  pl->bwd_w.clear();
  pl->bwd_indices.reserve(total);
  pl->bwd_w.reserve(total);
  uint32_t edge_base = 0;
  for (int i = 0; i < P; i++) {
    if (i == p)
      continue;
    const nts_host_chunk &c = pl->chunks[i];
    for (uint32_t r : pl->need[i])
      pl->bwd_offsets.push_back(c.row_offset[r] + edge_base);
    if (c.edges) {
      pl->bwd_indices.insert(pl->bwd_indices.end(), c.column_indices, c.column_indices + c.edges);
      pl->bwd_w.insert(pl->bwd_w.end(), c.edge_weight_backward, c.edge_weight_backward + c.edges);
    }
    edge_base += (uint32_t)c.edges;
  }
  pl->bwd_offsets.push_back(edge_base);
  pl->finalized = true;
  return 0;
}

int nts_exchange_plan_get_view(const nts_exchange_plan *pl, nts_exchange_plan_view *v) {
  NTS_ARG_CHECK(pl && v, "null argument");
  NTS_ARG_CHECK(pl->finalized, "plan not finalized");
  v->partitions = pl->P;
  v->rank = pl->p;
  v->owned_vertices = pl->chunks[pl->p].dst_end - pl->chunks[pl->p].dst_start;
  v->recv_total = pl->recv_total;
  v->send_total = pl->send_total;
  v->remote_edges = pl->remote_edges;
  v->need_count = pl->need_count.data();
  v->send_count = pl->send_count.data();
  v->peer_bwd_offset = pl->peer_bwd_offset.data();
  v->fwd_push_offset = pl->peer_bwd_offset.data(); // same number: where partition `rank` starts in rank j's staging
  v->bwd_push_offset = pl->bwd_push_offset.data();
  v->remote_column_offset = pl->remote_col_offset.empty() ? nullptr : pl->remote_col_offset.data();
  v->remote_slots = pl->remote_slots.empty() ? nullptr : pl->remote_slots.data();
  v->remote_weight = pl->remote_w.empty() ? nullptr : pl->remote_w.data();
  v->backward_offsets = pl->bwd_offsets.data();
  v->backward_rows = (nts_vid_t)pl->bwd_offsets.size() - 1;
  v->backward_indices = pl->bwd_indices.empty() ? nullptr : pl->bwd_indices.data();
  v->backward_weight = pl->bwd_w.empty() ? nullptr : pl->bwd_w.data();
  v->send_rows_all = pl->send_rows_all.empty() ? nullptr : pl->send_rows_all.data();
  return 0;
}

int nts_exchange_plan_chunk(const nts_exchange_plan *pl, int i, const nts_vid_t **slots,
                            const nts_vid_t **row_offset_compact) {
  NTS_ARG_CHECK(pl && pl->finalized && i >= 0 && i < pl->P && i != pl->p, "bad argument");
  if (slots)
    *slots = pl->chunk_slots[i].empty() ? nullptr : pl->chunk_slots[i].data();
  if (row_offset_compact)
    *row_offset_compact = pl->chunk_offc[i].data();
  return 0;
}

nts_exchange *nts_exchange_create_from_plan(nts_exchange_plan *pl, const nts_device_chunk *device_chunks) {
  if (!pl || !pl->finalized || !device_chunks) {
    fail(-1, "nts_exchange_create_from_plan needs a finalized plan and the device arrays of every chunk", __FILE__,
         __LINE__);
    return nullptr;
  }
  const int P = pl->P, p = pl->p;
  nts_exchange_desc d = {};
  d.partitions = P;
  d.rank = p;
  d.owned_vertices = pl->chunks[p].dst_end - pl->chunks[p].dst_start;
  d.dst_start = pl->chunks[p].dst_start;
  const nts_device_chunk &mine = device_chunks[p];
  d.local_column_offset = mine.column_offset;
  d.local_row_indices = mine.row_indices;
  d.local_row_offset = mine.row_offset;
  d.local_column_indices = mine.column_indices;
  d.local_weight_forward = mine.edge_weight_forward;
  d.local_weight_backward = mine.edge_weight_backward;
  d.local_edges = (nts_vid_t)pl->chunks[p].edges;
  if (d.local_edges && !(mine.column_offset && mine.row_indices && mine.row_offset && mine.column_indices &&
                         mine.edge_weight_forward && mine.edge_weight_backward)) {
    fail(-1, "device arrays of the local chunk are missing", __FILE__, __LINE__);
    return nullptr;
  }
  pl->dev_chunks.assign(P, nts_exchange_chunk{});
  int rc = 0;
  rc |= upload(pl, pl->send_rows_all, &d.send_rows_all);
  for (int i = 0; i < P && !rc; i++) {
    if (i == p)
      continue;
    nts_exchange_chunk &c = pl->dev_chunks[i];
    c.edges = pl->chunks[i].edges;
    if (c.edges && !(device_chunks[i].column_offset && device_chunks[i].column_indices &&
                     device_chunks[i].edge_weight_forward && device_chunks[i].edge_weight_backward)) {
      fail(-1, "device arrays of a remote chunk are missing", __FILE__, __LINE__);
      return nullptr;
    }
    c.column_offset = device_chunks[i].column_offset;
    c.column_indices = device_chunks[i].column_indices;
    c.weight_forward = device_chunks[i].edge_weight_forward;
    c.weight_backward = device_chunks[i].edge_weight_backward;
    rc |= upload(pl, pl->chunk_slots[i], &c.slots);
    rc |= upload(pl, pl->chunk_offc[i], &c.row_offset_compact);
  }
  if (rc)
    return nullptr;
  pl->need_dev.assign(P, nullptr);
  if (upload(pl, pl->need[p], &pl->need_dev[p]))
    return nullptr;
  d.local_need = pl->need_dev[p];
  d.local_need_count = (nts_vid_t)pl->need[p].size();
  d.chunks = pl->dev_chunks.data();
  d.need_count = pl->need_count.data();
  d.send_count = pl->send_count.data();
  d.fwd_push_offset = pl->peer_bwd_offset.data();
  d.bwd_push_offset = pl->bwd_push_offset.data();
  return nts_exchange_create(&d);
}

} // extern "C"
