/* nts_b200.h - C ABI of libnts_b200.so: NeutronStar's sparse neighbour-aggregation hot path,
 * hand-written for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  Every entry point replaces one member of the reference's device
 * interface `cuda/ntsCUDA.hpp` (free functions :25-47, `deviceCSC` :49-95, `Cuda_Stream` :97-217,
 * implemented by `cuda/ntsCUDAGraphOP.cu`); the reference interface it stands in for is cited at
 * each declaration.  The C++ surface of that header (class `Cuda_Stream`, ...) is provided on top
 * of this ABI by `include/nts_cuda_compat.hpp` so the reference's host code links unchanged - see
 * INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes; all device pointers are CUDA device (or mapped/peer) addresses,
 *    `stream` is a `cudaStream_t` passed as `void*` (NULL = the legacy default stream);
 *  - feature matrices are row-major contiguous float32 [rows, feature_size], borrowed, never
 *    re-laid-out (the reference borrows torch storage the same way, core/NtsScheduler.hpp:505-515);
 *  - vertex ids / offsets are uint32 (`VertexId_CUDA`, cuda/cuda_type.h:21); all address
 *    arithmetic is 64-bit (the reference's 32-bit `feature_size*batch_size` products overflow for
 *    V*F >= 2^32, cuda/ntsCUDAFuseKernel.cuh:280,299);
 *  - aggregation kernels ACCUMULATE into `output` (caller zeroes), exactly like the reference
 *    (cuda/ntsCUDAFuseKernel.cuh:272-309; tensors come zero-filled from NtsScheduler::NewKeyTensor);
 *  - every call is asynchronous on `stream` unless stated; launch errors are checked;
 *  - return value: 0 on success, otherwise the cudaError_t (or -1 for argument errors).  The
 *    message is available from nts_last_error().  With NTS_B200_ABORT_ON_ERROR=1 in the
 *    environment the library prints the message and exit(1)s instead - the reference's convention
 *    (cuda/ntsCUDAGraphOP.cu:13-19).
 *  - there is no CPU fallback anywhere in this library.
 */
#ifndef NTS_B200_H
#define NTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint32_t nts_vid_t; /* VertexId_CUDA, cuda/cuda_type.h:21 */

/* ---- library / device ----------------------------------------------------------------------- */
int nts_version(void);                 /* ABI version, currently 1 */
const char *nts_last_error(void);      /* thread-local text of the last failure */
int nts_device_count(void);
int nts_set_device(int device);        /* the reference never calls cudaSetDevice (always device 0) */
int nts_device_sm_count(int *sm_count);
int nts_device_synchronize(void);      /* ::CUDA_DEVICE_SYNCHRONIZE(), ntsCUDA.hpp:46 */
int nts_device_reset(void);            /* ::ResetDevice(), ntsCUDA.hpp:47 */

/* ---- memory (ntsCUDA.hpp:25-45) ---------------------------------------------------------------- */
void *nts_malloc_device(size_t bytes);            /* ::cudaMallocGPU / allocate_gpu_buffer / allocate_gpu_edge */
int nts_free_device(void *ptr);                   /* ::FreeBuffer / ::FreeEdge */
void *nts_malloc_pinned(size_t bytes);            /* ::cudaMallocPinned (cudaHostAllocMapped) */
int nts_free_pinned(void *ptr);                   /* ::ntsFreeHost */
void *nts_pinned_device_pointer(void *host_ptr);  /* ::getDevicePointer */
int nts_memcpy_h2d(void *dst_device, const void *src_host, size_t bytes, void *stream, int sync);
                                                  /* ::move_data_in / move_edge_in / move_bytes_in */
int nts_memcpy_d2h(void *dst_host, const void *src_device, size_t bytes, void *stream, int sync);
                                                  /* ::move_result_out */
int nts_memcpy_d2d(void *dst_device, const void *src_device, size_t bytes, void *stream);
int nts_zero(void *device_ptr, size_t bytes, void *stream); /* ::zero_buffer */

/* ---- streams / events (Cuda_Stream ctor, destory_Stream, CUDA_DEVICE_SYNCHRONIZE; ntsCUDA.hpp:97-103) */
void *nts_stream_create(int non_blocking);
int nts_stream_destroy(void *stream);
int nts_stream_synchronize(void *stream);
void *nts_event_create(int with_timing);
int nts_event_destroy(void *event);
int nts_event_record(void *event, void *stream);
int nts_stream_wait_event(void *stream, void *event);
int nts_event_elapsed_ms(void *start, void *stop, float *ms);

/* ---- the hot path: segmented weighted gather-sum (SpMM-like) --------------------------------------
 * output[r, :] += sum_{e in [offsets[r], offsets[r+1])} input[indices[e] - index_base, :] * (weight ? weight[e] : 1)
 *
 * Forward  = Cuda_Stream::Gather_By_Dst_From_Src[_Optim] (ntsCUDA.hpp:125-138, ntsCUDAGraphOP.cu:157-211):
 *            offsets = column_offset[Vdst+1], indices = row_indices (GLOBAL source ids), index_base = src_start.
 * Backward = Cuda_Stream::Gather_By_Src_From_Dst[_Optim] (ntsCUDA.hpp:139-152, ntsCUDAGraphOP.cu:213-281):
 *            offsets = row_offset[Vsrc+1], indices = column_indices (GLOBAL destination ids), index_base = dst_start.
 * The two named wrappers keep the reference's parameter list (minus the unused tensor_weight flag).
 */
int nts_segment_gather_sum(const float *input, float *output, const float *weight,
                           const nts_vid_t *indices, const nts_vid_t *offsets, nts_vid_t index_base,
                           nts_vid_t n_rows, uint64_t n_edges, nts_vid_t feature_size, void *stream);

int nts_gather_by_dst_from_src(const float *input, float *output, const float *weight_forward,
                               const nts_vid_t *row_indices, const nts_vid_t *column_offset,
                               nts_vid_t src_start, nts_vid_t src_end, nts_vid_t dst_start,
                               nts_vid_t dst_end, nts_vid_t edges, nts_vid_t batch_size,
                               nts_vid_t feature_size, int with_weight, void *stream);

int nts_gather_by_src_from_dst(const float *input, float *output, const float *weight_backward,
                               const nts_vid_t *row_offset, const nts_vid_t *column_indices,
                               nts_vid_t src_start, nts_vid_t src_end, nts_vid_t dst_start,
                               nts_vid_t dst_end, nts_vid_t edges, nts_vid_t batch_size,
                               nts_vid_t feature_size, int with_weight, void *stream);

/* Row-range launch of the same contraction: `offsets` points at the first row of the range (offsets[0] ==
 * edge_begin, offsets[n_rows] == edge_end), `output` at its first output row; indices / weight stay the whole
 * arrays (addressed by absolute edge position).  Used by the exchange engine to aggregate the remote chunks in
 * pipeline stages (core/graph.hpp:3678-3719 processes one chunk per ring step). */
int nts_segment_gather_sum_range(const float *input, float *output, const float *weight, const nts_vid_t *indices,
                                 const nts_vid_t *offsets, const nts_vid_t *slot_of, nts_vid_t index_base,
                                 nts_vid_t n_rows, uint64_t edge_begin, uint64_t edge_end, nts_vid_t feature_size,
                                 void *stream);

/* ---- preprocessed aggregation: nts_gather_plan --------------------------------------------------------------------
 * The arrays of one chunk direction (CSC: column_offset / row_indices / edge_weight_forward; CSR: row_offset /
 * column_indices / edge_weight_backward; core/GraphSegment.h:52-139) regrouped ONCE on the device for repeated
 * aggregation: edges bucketed by (slab of the gathered row, output row) with a stable sort so that one slab of the
 * feature matrix stays L2-resident per launch, (row, weight) stored as interleaved pairs, gathers from 16-byte
 * aligned (padded) rows.  The reference has no counterpart (its chunks are consumed as built); results equal
 * nts_segment_gather_sum up to fp32 re-association across slabs.  `gather_rows` = rows of the gathered matrix
 * (every mapped index must be < gather_rows); n_slabs = 0/1 disables the bucketing (pairs only).
 * The plan owns copies of everything it needs: the input arrays may be released after create returns
 * (create synchronises `stream`). */
typedef struct nts_gather_plan nts_gather_plan;
int nts_gather_plan_pick_slabs(nts_vid_t gather_rows, uint64_t n_edges, nts_vid_t n_rows, nts_vid_t feature_size,
                               uint64_t l2_budget_bytes /* 0 = default */);
nts_gather_plan *nts_gather_plan_create(const nts_vid_t *offsets, const nts_vid_t *indices, const float *weight,
                                        const nts_vid_t *slot_of, nts_vid_t index_base, nts_vid_t n_rows,
                                        uint64_t n_edges, nts_vid_t gather_rows, int n_slabs, void *stream);
/* slab count chosen by MEASUREMENT on the real arrays at this feature width (candidates 1, 2, 4, ... built and timed
 * once; hub-dominated graphs prefer no bucketing, uniform ones 8-16 slabs) */
nts_gather_plan *nts_gather_plan_create_tuned(const nts_vid_t *offsets, const nts_vid_t *indices, const float *weight,
                                              const nts_vid_t *slot_of, nts_vid_t index_base, nts_vid_t n_rows,
                                              uint64_t n_edges, nts_vid_t gather_rows, nts_vid_t feature_size,
                                              void *stream);
/* Several chunks merged into ONE plan: part-local output row r becomes row_add + r, a mapped index g becomes index_add
 * + g; inside a (slab, row) segment the parts follow each other in the order given.  n_slabs = 0 measures the slab
 * count for feature_size.  Used by the exchange engine to aggregate all remote chunks of a rank in one launch. */
typedef struct nts_plan_part {
  const nts_vid_t *offsets, *indices;   /* [n_rows+1], [n_edges] of this part */
  const float *weight;                  /* [n_edges] or NULL */
  const nts_vid_t *slot_of;             /* optional slot table applied to the indices */
  nts_vid_t index_base, index_add, n_rows, row_add;
  uint64_t n_edges;
} nts_plan_part;
nts_gather_plan *nts_gather_plan_create_parts(const nts_plan_part *parts, int n_parts, nts_vid_t n_rows,
                                              nts_vid_t gather_rows, int n_slabs, nts_vid_t feature_size, void *stream);
float nts_gather_plan_tuned_ms(const nts_gather_plan *plan); /* time of the winning candidate of a measured plan */
int nts_gather_plan_destroy(nts_gather_plan *plan);
int nts_gather_plan_slabs(const nts_gather_plan *plan);
uint64_t nts_gather_plan_bytes(const nts_gather_plan *plan);
/* output[r,:] += sum_e input[row(e),:] * w(e)   (accumulates; one launch per non-empty slab, in stream order) */
int nts_gather_plan_run(nts_gather_plan *plan, const float *input, float *output, nts_vid_t feature_size,
                        void *stream);
int nts_gather_plan_last_launch(const nts_gather_plan *plan, int *launches, int *grid, int *k, int *u, int *outv);
int nts_gather_plan_set_tuning(int u, int min_blocks, int edges_per_warp); /* measurement hook, 0 = default */
/* 0 = gathered rows through registers (default); 1 = rows staged in shared memory by per-row cp.async.bulk (TMA) into
 * a per-warp ring, U of set_tuning = ring depth - the north star's "feature tiles via TMA", kept for measurement */
int nts_gather_plan_set_variant(int variant);

/* Same contraction with the source row taken through a slot table instead of `index - base`:
 * row = slot_of[indices[e]].  Used with MirrorIndex (core/PartitionedGraph.hpp:295-305) for the fused
 * GAT aggregation, DistAggregateDstFuseWeight::forward (core/ntsDistCPUGraphOp.hpp:516-546), and with
 * the compact receive-staging slots of the multi-GPU exchange. */
int nts_segment_gather_sum_slots(const float *input, float *output, const float *weight,
                                 const nts_vid_t *indices, const nts_vid_t *offsets,
                                 const nts_vid_t *slot_of, nts_vid_t n_rows, uint64_t n_edges,
                                 nts_vid_t feature_size, void *stream);

/* Multi-head fused GAT aggregation: weight is [n_edges, heads] row-major, feature_size = heads * D and head h scales
 * columns [h*D, (h+1)*D):  output[r, hD+c] += sum_e weight[e,h] * input[row(e), hD+c],  row(e) = slot_of ?
 * slot_of[indices[e]] : indices[e] - index_base.  heads = 1 is the single-head operator above.  (The reference's
 * DistAggregateDstFuseWeight has one head; 8 heads is config D of BASELINE.json.) */
int nts_segment_gather_sum_heads(const float *input, float *output, const float *weight,
                                 const nts_vid_t *indices, const nts_vid_t *offsets, const nts_vid_t *slot_of,
                                 nts_vid_t index_base, nts_vid_t n_rows, uint64_t n_edges,
                                 nts_vid_t feature_size, nts_vid_t heads, void *stream);

/* Tuning / introspection of the aggregation kernel (does not change results beyond fp32
 * re-association): variant 0 = auto, see DESIGN.md "kernel variants". */
int nts_aggregate_set_variant(int variant, int edges_per_warp);
int nts_aggregate_last_launch(int *grid, int *block, int *smem_bytes, int *variant);
uint64_t nts_kernel_launch_count(void); /* kernels launched by this library since load */

/* ---- edge-granular operators (GAT building blocks) -------------------------------------------------
 * All take the whole-partition CSC of core/PartitionedGraph.hpp:105-143 (column_offset[Vp+1], row_indices[Ep]
 * global source ids) as uploaded by `deviceCSC` (ntsCUDA.hpp:49-95). `batch_size` = Vp. */
/* msg[e,:] = mirror[mirror_index[row_indices[e]],:]      Cuda_Stream::Scatter_Src_Mirror_to_Msg (ntsCUDA.hpp:154) */
int nts_scatter_src_mirror_to_msg(float *message, const float *src_mirror_feature,
                                  const nts_vid_t *row_indices, const nts_vid_t *column_offset,
                                  const nts_vid_t *mirror_index, nts_vid_t batch_size,
                                  nts_vid_t feature_size, void *stream);
/* mirror_grad[mirror_index[row_indices[e]],:] += msg[e,:]  Cuda_Stream::Gather_Msg_To_Src_Mirror (ntsCUDA.hpp:159) */
int nts_gather_msg_to_src_mirror(float *src_mirror_feature, const float *message,
                                 const nts_vid_t *row_indices, const nts_vid_t *column_offset,
                                 const nts_vid_t *mirror_index, nts_vid_t batch_size,
                                 nts_vid_t feature_size, void *stream);
/* msg[e,:] = dst_feature[dst(e),:]                         Cuda_Stream::Scatter_Dst_to_Msg (ntsCUDA.hpp:164) */
int nts_scatter_dst_to_msg(float *message, const float *dst_feature, const nts_vid_t *row_indices,
                           const nts_vid_t *column_offset, nts_vid_t batch_size,
                           nts_vid_t feature_size, void *stream);
/* dst_feature[d,:] += sum_{e->d} msg[e,:]                  Cuda_Stream::Gather_Msg_to_Dst (ntsCUDA.hpp:168) */
int nts_gather_msg_to_dst(float *dst_feature, const float *message, const nts_vid_t *row_indices,
                          const nts_vid_t *column_offset, nts_vid_t batch_size,
                          nts_vid_t feature_size, void *stream);
/* a[seg,h] = softmax_seg(m[seg,h]) per destination segment and column h (max-subtracted; oracle =
 * DistEdgeSoftMax::forward core/ntsDistCPUGraphOp.hpp:449-470); msg_cached receives a copy.
 * Cuda_Stream::Edge_Softmax_Forward_Block (ntsCUDA.hpp:172) - the reference kernel handles 1 column only. */
int nts_edge_softmax_forward(float *msg_output, const float *msg_input, float *msg_cached,
                             const nts_vid_t *row_indices, const nts_vid_t *column_offset,
                             nts_vid_t batch_size, nts_vid_t feature_size, void *stream);
/* g_in[e,h] = a[e,h]*g[e,h] - a[e,h]*sum_seg(g*a)            Cuda_Stream::Edge_Softmax_Backward_Block (ntsCUDA.hpp:177) */
int nts_edge_softmax_backward(float *msg_input_grad, const float *msg_output_grad,
                              const float *msg_cached, const nts_vid_t *row_indices,
                              const nts_vid_t *column_offset, nts_vid_t batch_size,
                              nts_vid_t feature_size, void *stream);
/* message_grad[e,:] += input[dst(e),:]                     Cuda_Stream::Scatter_Grad_Back_To_Message (ntsCUDA.hpp:194) */
int nts_scatter_grad_back_to_message(const float *input, float *message_grad,
                                     const nts_vid_t *row_indices, const nts_vid_t *column_offset,
                                     nts_vid_t batch_size, nts_vid_t feature_size, void *stream);

/* Fused GAT aggregation backward, DistAggregateDstFuseWeight::backward (core/ntsDistCPUGraphOp.hpp:548-589)
 * WITHOUT the reference's spurious extra unweighted add (:572):
 *   mirror_grad[slot(e),:] += a[e] * g[dst(e),:]        (needs mirror_grad zeroed by the caller)
 *   a_grad[e]               = < mirror[slot(e),:], g[dst(e),:] >  */
int nts_aggregate_dst_fuse_weight_backward(float *mirror_grad, float *edge_weight_grad,
                                           const float *mirror, const float *edge_weight,
                                           const float *dst_grad, const nts_vid_t *row_indices,
                                           const nts_vid_t *column_offset,
                                           const nts_vid_t *mirror_index, nts_vid_t batch_size,
                                           nts_vid_t feature_size, void *stream);
/* multi-head form: edge_weight / edge_weight_grad are [E, heads], head h owns columns [h*D, (h+1)*D) */
int nts_aggregate_dst_fuse_weight_backward_heads(float *mirror_grad, float *edge_weight_grad,
                                                 const float *mirror, const float *edge_weight,
                                                 const float *dst_grad, const nts_vid_t *row_indices,
                                                 const nts_vid_t *column_offset,
                                                 const nts_vid_t *mirror_index, nts_vid_t batch_size,
                                                 nts_vid_t feature_size, nts_vid_t heads, void *stream);

/* ---- fully fused GAT layer (K7): edge logits / attention are never materialised ---------------------------------
 * The flow of toolkits/GAT_CPU_DIST_OPTM.hpp:196-241 (per-vertex scores -> leaky_relu -> edge softmax ->
 * DistAggregateDstFuseWeight) with a[e,h] = softmax_seg(leaky_relu(src_score[slot(e),h] + dst_score[dst(e),h]))
 * recomputed inside the kernels; slot(e) = mirror_index[row_indices[e]] (or row_indices[e] itself when mirror_index
 * is NULL, i.e. the caller has applied the lookup once), head h owns columns [h*D,(h+1)*D), D <= 512.
 *   stats    : seg_max[d,h], seg_sum[d,h]                                   ([batch_size, heads] each)
 *   forward  : output[d, hD+c] += sum_e a[e,h] * mirror[slot(e), hD+c]
 *   backward : mirror_grad[slot,hD+c] += a*g[d,hD+c];  src_score_grad[slot,h] += dpre;  dst_score_grad[d,h] += dpre
 *              with dpre = a*(<mirror[slot,h],g[d,h]> - out_dot_grad[d,h]) * leaky_relu'(pre) and
 *              out_dot_grad[d,h] = <output[d,h], g[d,h]> supplied by the caller (a per-vertex dot product).
 *   The three gradient outputs must be zeroed by the caller. */
int nts_gat_softmax_stats(float *seg_max, float *seg_sum, const float *src_score, const float *dst_score,
                          const nts_vid_t *row_indices, const nts_vid_t *column_offset,
                          const nts_vid_t *mirror_index, nts_vid_t batch_size, nts_vid_t heads,
                          float negative_slope, void *stream);
int nts_gat_fused_aggregate_forward(const float *mirror, float *output, const float *src_score,
                                    const float *dst_score, const float *seg_max, const float *seg_sum,
                                    const nts_vid_t *row_indices, const nts_vid_t *column_offset,
                                    const nts_vid_t *mirror_index, nts_vid_t batch_size, uint64_t n_edges,
                                    nts_vid_t feature_size, nts_vid_t heads, float negative_slope, void *stream);
int nts_gat_fused_aggregate_backward(float *mirror_grad, float *src_score_grad, float *dst_score_grad,
                                     const float *mirror, const float *src_score, const float *dst_score,
                                     const float *seg_max, const float *seg_sum, const float *out_dot_grad,
                                     const float *dst_grad, const nts_vid_t *row_indices,
                                     const nts_vid_t *column_offset, const nts_vid_t *mirror_index,
                                     nts_vid_t batch_size, nts_vid_t feature_size, nts_vid_t heads,
                                     float negative_slope, void *stream);
/* Same results without per-edge atomics: a destination-major pass over the CSC (dst_score_grad) and a source-major
 * pass over the CSR of the same edges keyed by MIRROR SLOT (mirror_grad, src_score_grad), both with register
 * accumulators.  slot_row_offset[mirror_size+1] / slot_column_indices[E] (local destination ids) list the out-edges
 * of every mirror slot; dst_pack is a 16-byte-aligned workspace of batch_size*heads*4 floats.  Shapes the passes do
 * not cover (heads > 1 with a non-power-of-two head width in vectors, rows wider than 128 vectors) fall back to
 * nts_gat_fused_aggregate_backward.  The three gradient outputs must be zeroed by the caller. */
int nts_gat_fused_aggregate_backward_two_pass(float *mirror_grad, float *src_score_grad, float *dst_score_grad,
                                              float *dst_pack, const float *mirror, const float *src_score,
                                              const float *dst_score, const float *seg_max, const float *seg_sum,
                                              const float *out_dot_grad, const float *dst_grad,
                                              const nts_vid_t *row_indices, const nts_vid_t *column_offset,
                                              const nts_vid_t *mirror_index, const nts_vid_t *slot_row_offset,
                                              const nts_vid_t *slot_column_indices, nts_vid_t batch_size,
                                              nts_vid_t mirror_size, nts_vid_t feature_size, nts_vid_t heads,
                                              float negative_slope, void *stream);

/* ---- (vid,row) message records: the reference's host-staged exchange format (comm/network.h:143-149) ---
 * record k = { uint32 vid; float row[feature_size]; }, read through mapped pinned host memory. */
/* mirror[vid - partition_start,:] = record.row if vid in [partition_start, partition_end)
 * Cuda_Stream::deSerializeToGPU (ntsCUDA.hpp:113, ntsCUDATransferKernel.cuh:70-93) */
int nts_deserialize_records(float *mirror, const float *records, nts_vid_t n_records,
                            nts_vid_t feature_size, nts_vid_t partition_start,
                            nts_vid_t partition_end, void *stream);
/* Y[vid - partition_start,:] += record.row
 * Cuda_Stream::aggregate_comm_result_debug (ntsCUDA.hpp:117, ntsCUDATransferKernel.cuh:49-68) */
int nts_aggregate_records(float *aggregate, const float *records, nts_vid_t n_records,
                          nts_vid_t feature_size, nts_vid_t partition_start,
                          nts_vid_t partition_end, void *stream);

/* ---- dense-row exchange helpers of the B200 engine (replace the record format on NVLink) -------------- */
/* dst[k,:] = src[rows[k],:]   (sender-side compaction of mirror rows / pull from a peer's mapped buffer) */
int nts_gather_rows(float *dst, const float *src, const nts_vid_t *rows, nts_vid_t n_rows,
                    nts_vid_t feature_size, void *stream);
/* dst[rows[k],:] += src[k,:]  (receiver-side add of partial gradients; rows must be unique) */
int nts_scatter_add_rows(float *dst, const float *src, const nts_vid_t *rows, nts_vid_t n_rows,
                         nts_vid_t feature_size, void *stream);
/* same with vector red.global.add: rows may repeat (partials of several senders merged into one launch) */
int nts_scatter_add_rows_atomic(float *dst, const float *src, const nts_vid_t *rows, nts_vid_t n_rows,
                                nts_vid_t feature_size, void *stream);

/* ---- parameter update ------------------------------------------------------------------------------------
 * Parameter::learnC2G_with_decay_Adam (core/NtsScheduler.hpp:774-781) as ONE kernel, in place on W / M / V:
 *   W_g = W*weight_decay + grad;  M = beta1*M + (1-beta1)*W_g;  V = beta2*V + (1-beta2)*W_g*W_g;
 *   W  -= alpha * M / (sqrt(V) + epsilon)
 * alpha / beta1 / beta2 are the CURRENT values of the reference's schedule (Parameter::next(), :727-736).  `grad` is the
 * all-reduced gradient (Parameter::all_reduce_to_gradient, :719-722 - ncclAllReduce here instead of .cpu() + MPI). */
int nts_adam_update(float *W, float *M, float *V, const float *grad, uint64_t n, float weight_decay, float beta1,
                    float beta2, float alpha, float epsilon, void *stream);

/* ---- peer memory (CUDA IPC) for the NVLink exchange ------------------------------------------------------ */
#define NTS_IPC_HANDLE_BYTES 64
int nts_ipc_get_handle(void *device_ptr, unsigned char handle[NTS_IPC_HANDLE_BYTES]);
void *nts_ipc_open_handle(const unsigned char handle[NTS_IPC_HANDLE_BYTES]);
int nts_ipc_close_handle(void *peer_ptr);
/* cross-GPU flag signalling on IPC-mapped uint32 flags (release/acquire at system scope) */
int nts_signal_set(uint32_t *flag, uint32_t value, void *stream);
int nts_signal_wait_geq(const uint32_t *flag, uint32_t value, void *stream);

/* ---- the exchange engine: data plane of the distributed fused aggregation (peer-memory transport) ---------------------
 * Replaces Graph::sync_compute_decoupled / compute_sync_decoupled (core/graph.hpp:3455-3719) and the host-staged
 * NtsGraphCommunicator (comm/network.cpp:159-844).  PUSH model over CUDA-IPC windows, one pipeline stage per source
 * partition in the reference's ring order (core/graph.hpp:3678-3683): the owner of a row stores it straight into the
 * reader's receive window over NVLink and raises an epoch flag; the reader aggregates chunk (p+s) as soon as the rows
 * of partition (p+s) have landed (csrc/nts_exchange.cu documents the protocol).
 * The caller owns the CONTROL plane: it builds the plan arrays (who needs which rows;
 * neutronstarlite_b200/exchange.py::ExchangePlan or nts_exchange_plan_* below) and moves the two 64-byte IPC handles
 * per rank between processes (torch.distributed here, MPI in the reference's host code); the engine owns windows,
 * flags, streams, events and the launch sequence.  All pointers are device pointers that must outlive the engine;
 * per-partition arrays have `partitions` entries (own entry ignored).  Row layouts ("partition order"): the receive
 * staging of rank r holds the rows it reads from partition 0, 1, ... (r skipped), need_count[i] rows each; the
 * gradient staging holds what ranks 0, 1, ... (r skipped) return, send_count[j] rows each = the order of
 * send_rows_all. */
typedef struct nts_exchange nts_exchange;
typedef struct nts_exchange_chunk {          /* remote chunk i: sources in partition i -> my destinations */
  const nts_vid_t *column_offset;            /* [V_p+1]  graph_chunks[i]->column_offset_gpu */
  const nts_vid_t *slots;                    /* [E_i]    row_indices remapped to 0..need_count[i]-1 (rank in need list) */
  const float *weight_forward;               /* [E_i]    graph_chunks[i]->edge_weight_forward_gpu */
  const nts_vid_t *row_offset_compact;       /* [need_count[i]+1] row_offset restricted to the active sources */
  const nts_vid_t *column_indices;           /* [E_i]    graph_chunks[i]->column_indices_gpu (global destination ids) */
  const float *weight_backward;              /* [E_i]    graph_chunks[i]->edge_weight_backward_gpu */
  uint64_t edges;
} nts_exchange_chunk;
typedef struct nts_exchange_desc {
  int partitions, rank;
  nts_vid_t owned_vertices, dst_start;      /* V_p and partition_offset[rank] */
  /* local chunk (sources in this partition): CSC + CSR of CSC_segment_pinned */
  const nts_vid_t *local_column_offset, *local_row_indices, *local_row_offset, *local_column_indices;
  const float *local_weight_forward, *local_weight_backward;
  nts_vid_t local_edges;
  const nts_exchange_chunk *chunks;         /* [P] remote chunks (host array of device pointers) */
  const nts_vid_t *need_count;              /* [P] rows of partition i that I read (= active sources of chunk i) */
  const nts_vid_t *send_count;              /* [P] rows of mine that rank j reads */
  const nts_vid_t *send_rows_all;           /* device: concatenation over j != rank (ascending) of those local row ids */
  const nts_vid_t *fwd_push_offset;         /* [P] first row of MY rows inside rank j's receive staging */
  const nts_vid_t *bwd_push_offset;         /* [P] first row of MY partials inside rank i's gradient staging */
  /* rows of MY partition that are sources of my own in-edges (active sources of the local chunk, ascending); only
   * the mirror fetch below needs them */
  const nts_vid_t *local_need;              /* device */
  nts_vid_t local_need_count;
} nts_exchange_desc;

nts_exchange *nts_exchange_create(const nts_exchange_desc *desc);
int nts_exchange_destroy(nts_exchange *ex);
/* floats ONE epoch buffer of the receive window needs at this width; take the MAX over ranks before reserving */
uint64_t nts_exchange_required_floats(const nts_exchange *ex, nts_vid_t feature_size);
uint64_t nts_exchange_capacity_floats(const nts_exchange *ex);
/* Replacing the exported window is COLLECTIVE and must not race with peers that still map or write it.  On every
 * rank: nts_exchange_release_peers (drains this rank's device work, closes its mappings of the peers' windows) ->
 * barrier -> nts_exchange_reserve (frees / allocates; n_buffers = 2 lets a rank run one exchange ahead of a slow
 * peer, 1 halves the memory) -> nts_exchange_handles -> all-gather of the handles -> nts_exchange_open_peers ->
 * barrier.  Reserve once for the widest layer to keep cudaMalloc out of the epoch loop. */
int nts_exchange_release_peers(nts_exchange *ex);
int nts_exchange_reserve(nts_exchange *ex, uint64_t floats_per_buffer, int n_buffers);
int nts_exchange_handles(nts_exchange *ex, unsigned char window_handle[64], unsigned char flags_handle[64]);
int nts_exchange_open_peers(nts_exchange *ex, const unsigned char *window_handles, const unsigned char *flag_handles);
/* Y_p += sum_i A_{p<-i} X_i  (ForwardGPUfuseOp::forward, core/ntsDistGPUFusedGraphOp.hpp:56-73); y zeroed by caller.
 * Every rank must issue the same sequence of forward / backward calls (SPMD, like the reference's ring). */
int nts_exchange_forward(nts_exchange *ex, const float *x, float *y, nts_vid_t feature_size, void *stream);
/* dX_p += sum_j A_{j<-p}^T dY_j (ForwardGPUfuseOp::backward, :75-90); dx zeroed by caller */
int nts_exchange_backward(nts_exchange *ex, const float *g, float *dx, nts_vid_t feature_size, void *stream);
/* per-phase device timeline of the last forward (measurement): [0] push kernel, [1] local chunk, then per ring step
 * s: [2s] wait for the rows of partition (p+s), [2s+1] aggregation of chunk (p+s); [2P] whole call.  2P+1 floats (ms);
 * nts_exchange_last_timeline synchronises the device */
int nts_exchange_set_trace(nts_exchange *ex, int enable);
int nts_exchange_last_timeline(nts_exchange *ex, float *ms, int capacity);
/* DistGPUGetDepNbrOp (core/ntsDistGPUGraphOp.hpp:48-143) on the same windows - the reference moves the whole feature
 * matrix GPU -> host -> MPI -> host -> GPU.  forward: mirror[MirrorIndex[s], :] = X[s, :] for every source s of a local
 * in-edge ([owned_mirrors, F], partition order); backward: dx[v, :] += every partition's mirror gradient of my vertex
 * v (dx zeroed by the caller). */
int nts_exchange_fetch_mirrors(nts_exchange *ex, const float *x, float *mirror, nts_vid_t feature_size, void *stream);
int nts_exchange_return_mirror_grads(nts_exchange *ex, const float *mirror_grad, float *dx, nts_vid_t feature_size,
                                     void *stream);

/* ---- exchange plan builder (host C++): the reference's chunks -> the arrays of nts_exchange_desc ------------------------
 * C++ twin of neutronstarlite_b200/exchange.py::ExchangePlan for hosts without Python (the reference's own host
 * code: include/nts_dropin/core/ntsDistGPUFusedGraphOp.hpp).  One nts_host_chunk per source partition i describes
 * CSC_segment_pinned `graph_chunks[i]` of this rank (core/GraphSegment.h:52-139) through its HOST arrays:
 * column_offset[V_p+1] / row_indices[E_i] (global source ids) / edge_weight_forward, row_offset[V_i+1] /
 * column_indices[E_i] (global destination ids) / edge_weight_backward, src_range, dst_range, edge_size.
 * Sequence:  create -> pack_needs -> (caller moves every rank's pack to every rank) -> set_peer_needs for each
 * peer -> finalize -> create_from_plan (uploads; the plan owns the device copies and must outlive the engine).
 * The merged arrays of the view (one CSC / one compact CSR over all remote chunks) serve transports that aggregate
 * all remote chunks in one launch (the NCCL all-to-all path); the peer-memory engine works per chunk. */
typedef struct nts_host_chunk {
  const nts_vid_t *column_offset, *row_indices, *row_offset, *column_indices;
  const float *edge_weight_forward, *edge_weight_backward;
  nts_vid_t src_start, src_end, dst_start, dst_end;
  uint64_t edges;
} nts_host_chunk;
typedef struct nts_exchange_plan nts_exchange_plan;
nts_exchange_plan *nts_exchange_plan_create(const nts_host_chunk *chunks, int partitions, int rank);
void nts_exchange_plan_destroy(nts_exchange_plan *plan);
/* rows of partition i (local ids, ascending) that have an edge into this rank's partition = what it reads from i */
const nts_vid_t *nts_exchange_plan_need(const nts_exchange_plan *plan, int i, nts_vid_t *count);
/* this rank's lists in wire form: need_counts[P] (own entry 0) and the lists concatenated in partition order */
uint64_t nts_exchange_plan_packed_rows(const nts_exchange_plan *plan);
int nts_exchange_plan_pack_needs(const nts_exchange_plan *plan, nts_vid_t *need_counts, nts_vid_t *need_rows);
/* rank j's wire form, as received */
int nts_exchange_plan_set_peer_needs(nts_exchange_plan *plan, int j, const nts_vid_t *need_counts,
                                     const nts_vid_t *need_rows);
int nts_exchange_plan_finalize(nts_exchange_plan *plan);
/* host view of the finalized plan (tests, other transports); pointers live as long as the plan */
typedef struct nts_exchange_plan_view {
  int partitions, rank;
  nts_vid_t owned_vertices, recv_total, send_total, backward_rows;
  uint64_t remote_edges;
  const nts_vid_t *need_count, *send_count, *peer_bwd_offset;   /* [P] each */
  const nts_vid_t *fwd_push_offset, *bwd_push_offset;           /* [P] each, see nts_exchange_desc */
  const nts_vid_t *remote_column_offset, *remote_slots;         /* [V_p+1], [remote_edges] */
  const float *remote_weight;
  const nts_vid_t *backward_offsets, *backward_indices;         /* [backward_rows+1], [remote_edges] */
  const float *backward_weight;
  const nts_vid_t *send_rows_all;                               /* [send_total] */
} nts_exchange_plan_view;
int nts_exchange_plan_get_view(const nts_exchange_plan *plan, nts_exchange_plan_view *view);
/* per remote chunk i of the finalized plan (host arrays, live as long as the plan): row_indices as ranks in the need
 * list [E_i], row_offset restricted to the active sources [need_count[i]+1] */
int nts_exchange_plan_chunk(const nts_exchange_plan *plan, int i, const nts_vid_t **slots,
                            const nts_vid_t **row_offset_compact);
/* DEVICE arrays of CSC_segment_pinned graph_chunks[i] as uploaded by CopyGraphToDevice (core/GraphSegment.cpp:178-220),
 * one entry per source partition; the plan uploads the derived arrays itself and must outlive the engine */
typedef struct nts_device_chunk {
  const nts_vid_t *column_offset, *row_indices, *row_offset, *column_indices;
  const float *edge_weight_forward, *edge_weight_backward;
} nts_device_chunk;
nts_exchange *nts_exchange_create_from_plan(nts_exchange_plan *plan, const nts_device_chunk *device_chunks);

/* ---- host-side graph preparation (C++ with OpenMP; no device involved) -----------------------------------
 * Restates the layout contract of core/graph.hpp:1185-1211 (partitioner), :4396-4401 (degree clamp),
 * core/ntsBaseOp.hpp:194-197 (edge weight) and core/PartitionedGraph.hpp:324-420 (per-source-partition chunks). */
/* degrees with multiplicity over packed {u32 src,u32 dst} edges, clamped to >= 1 */
int nts_host_degrees(const nts_vid_t *edges_src_dst, uint64_t n_edges, nts_vid_t n_vertices,
                     nts_vid_t *out_degree, nts_vid_t *in_degree);
/* partition_offset[P+1] */
int nts_host_partition_offsets(const nts_vid_t *edges_src_dst, uint64_t n_edges, nts_vid_t n_vertices,
                               int partitions, nts_vid_t *partition_offset);
/* number of edges of chunk (src partition i -> dst partition p) for every i: counts[P] */
int nts_host_chunk_edge_counts(const nts_vid_t *edges_src_dst, uint64_t n_edges,
                               const nts_vid_t *partition_offset, int partitions, int rank,
                               uint64_t *counts);
/* build chunk i of rank p into caller-allocated arrays:
 * column_offset[Vp+1], row_indices[Ei], edge_weight_forward[Ei], row_offset[Vi+1], column_indices[Ei],
 * edge_weight_backward[Ei], source_active[Vi] (bytes) */
int nts_host_build_chunk(const nts_vid_t *edges_src_dst, uint64_t n_edges, nts_vid_t n_vertices,
                         const nts_vid_t *partition_offset, int partitions, int rank, int src_partition,
                         const nts_vid_t *out_degree, const nts_vid_t *in_degree,
                         nts_vid_t *column_offset, nts_vid_t *row_indices, float *edge_weight_forward,
                         nts_vid_t *row_offset, nts_vid_t *column_indices, float *edge_weight_backward,
                         unsigned char *source_active);
/* MirrorIndex[V+1] of rank p (core/PartitionedGraph.hpp:295-305); returns owned_mirrors through *owned */
int nts_host_mirror_index(const nts_vid_t *edges_src_dst, uint64_t n_edges, nts_vid_t n_vertices,
                          const nts_vid_t *partition_offset, int rank, nts_vid_t *mirror_index,
                          nts_vid_t *owned);

/* ---- feature / label / mask tables (GNNDatum, core/ntsDataloador.hpp) ----------------------------------------------------
 * Text tables exactly as GNNDatum::readFeature_Label_Mask (:156-221) reads them - "id f0 .. fF-1", "id label",
 * "id train|val|eval|test", the k-th label / mask record belongs to the k-th feature record - parsed in parallel;
 * rows with id in [v_begin, v_end) land at id - v_begin (mask: train 0, val/eval 1, test 2, other 3).  label_path /
 * mask_path (and their outputs) may be NULL.  Returns 0, or a negative code (-2/-3/-4 unreadable file, -5 malformed). */
int nts_host_read_feature_label_mask(const char *feature_path, const char *label_path, const char *mask_path,
                                     nts_vid_t feature_size, nts_vid_t v_begin, nts_vid_t v_end, float *features,
                                     int64_t *labels, int32_t *masks);
/* rows [v_begin, v_end) of a packed float32 [V, feature_size] table (the twin of the packed binary edge file) */
int nts_host_read_feature_binary(const char *path, nts_vid_t feature_size, nts_vid_t v_begin, nts_vid_t v_end,
                                 float *features);

#ifdef __cplusplus
}
#endif
#endif /* NTS_B200_H */
