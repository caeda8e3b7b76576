// Additive C++ operator: `nts::op::DistGPUFusedGATOp` - the fully fused multi-head GAT layer (K7 of DESIGN.md) for the
// reference's C++ host code.  The flow of toolkits/GAT_CPU_DIST_OPTM.hpp:196-241 / toolkits/GAT_GPU_DIST.hpp:187-219
// (per-vertex scores -> leaky_relu -> edge softmax -> weighted aggregation of the mirror rows) in two kernels forward
// and two edge passes backward, with NO edge-sized tensor (the reference keeps four [E,F] messages on the GPU path).
// Python twin: neutronstarlite_b200/ops.py::DistGPUFusedGATOp (that one is what the GPU tests and the bench drive).
//
//   forward(mirror [M, H*D], src_score [M, H], dst_score [V, H]) -> out [V, H*D]
//       a[e,h] = softmax over the in-edges of dst(e) of leaky_relu(src_score[slot(e),h] + dst_score[dst(e),h])
//       out[d, hD:(h+1)D] = sum_e a[e,h] * mirror[slot(e), hD:(h+1)D],     slot(e) = MirrorIndex[row_indices[e]]
//   backward(grad_out) -> d_mirror;  get_src_score_grad() / get_dst_score_grad() afterwards
//
// Topology comes from the reference's own PartitionedGraph (whole-partition CSC column_offset / row_indices,
// MirrorIndex, and the mirror-keyed CSR compressed_row_offset / column_indices of GenerateWholeGraphTopo,
// core/PartitionedGraph.hpp:105-143), uploaded ONCE per PartitionedGraph and cached - the reference's edge operators
// re-upload a deviceCSC per operator instance (core/ntsDistGPUGraphOp.hpp:150-159).
// The reference's tape (NtsContext::runGraphOp) has one- and two-input entries only, so a toolkit calls this class
// directly and chains the three gradients itself.
#ifndef NTS_B200_DIST_GPU_FUSED_GAT_OP_HPP
#define NTS_B200_DIST_GPU_FUSED_GAT_OP_HPP

#if CUDA_ENABLE
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "nts_b200.h"

namespace nts {
namespace op {
namespace b200 {

struct GatTopology {
  nts_vid_t *column_offset = nullptr, *row_indices = nullptr, *mirror_index = nullptr;
  nts_vid_t *slot_row_offset = nullptr, *slot_column_indices = nullptr;
  nts_vid_t V = 0, E = 0, M = 0;

  static GatTopology &of(PartitionedGraph *pg) {
    static std::map<PartitionedGraph *, GatTopology *> all;
    GatTopology *&t = all[pg];
    if (!t)
      t = new GatTopology(pg);
    return *t;
  }

private:
  static nts_vid_t *upload(const nts_vid_t *host, size_t n) {
    nts_vid_t *d = static_cast<nts_vid_t *>(nts_malloc_device((n ? n : 1) * sizeof(nts_vid_t)));
    if (!d || (n && nts_memcpy_h2d(d, host, n * sizeof(nts_vid_t), nullptr, 1))) {
      std::fprintf(stderr, "nts_b200 GAT topology upload failed: %s\n", nts_last_error());
      std::exit(1);
    }
    return d;
  }
  explicit GatTopology(PartitionedGraph *pg) {
    V = pg->owned_vertices, E = pg->owned_edges, M = pg->owned_mirrors;
    column_offset = upload(pg->column_offset, (size_t)V + 1);
    row_indices = upload(pg->row_indices, E);
    mirror_index = upload(pg->MirrorIndex, (size_t)pg->global_vertices + 1);
    slot_row_offset = upload(pg->compressed_row_offset, (size_t)M + 1);
    std::vector<nts_vid_t> local(E); // the mirror-keyed CSR lists GLOBAL destination ids; the kernels want local ones
    const nts_vid_t v0 = pg->graph_->gnnctx->p_v_s;
    for (size_t e = 0; e < (size_t)E; e++)
      local[e] = pg->column_indices[e] - v0;
    slot_column_indices = upload(local.data(), E);
  }
};

} // namespace b200

class DistGPUFusedGATOp : public ntsGraphOp {
public:
  float negative_slope;

  DistGPUFusedGATOp(PartitionedGraph *partitioned_graph, VertexSubset *active, float negative_slope_ = 0.2f)
      : ntsGraphOp(partitioned_graph, active), negative_slope(negative_slope_) {}

  NtsVar forward(NtsVar &) {
    std::fprintf(stderr, "DistGPUFusedGATOp::forward takes (mirror, src_score, dst_score)\n");
    std::exit(1);
    return NtsVar();
  }

  NtsVar forward(NtsVar &mirror_, NtsVar &src_score_, NtsVar &dst_score_) {
    b200::GatTopology &t = b200::GatTopology::of(partitioned_graph_);
    mirror = mirror_.contiguous(), src_score = src_score_.contiguous(), dst_score = dst_score_.contiguous();
    const long F = mirror.size(1), H = src_score.size(1);
    auto opt = mirror.options();
    seg_max = torch::empty({(long)t.V, H}, opt);
    seg_sum = torch::empty({(long)t.V, H}, opt);
    out = torch::zeros({(long)t.V, F}, opt);
    if (nts_gat_softmax_stats(seg_max.data_ptr<float>(), seg_sum.data_ptr<float>(), src_score.data_ptr<float>(),
                              dst_score.data_ptr<float>(), t.row_indices, t.column_offset, t.mirror_index, t.V,
                              (nts_vid_t)H, negative_slope, nullptr) ||
        nts_gat_fused_aggregate_forward(mirror.data_ptr<float>(), out.data_ptr<float>(), src_score.data_ptr<float>(),
                                        dst_score.data_ptr<float>(), seg_max.data_ptr<float>(),
                                        seg_sum.data_ptr<float>(), t.row_indices, t.column_offset, t.mirror_index, t.V,
                                        t.E, (nts_vid_t)F, (nts_vid_t)H, negative_slope, nullptr))
      die("fused GAT forward");
    return out;
  }

  // d_mirror; the two score gradients are kept for get_src_score_grad() / get_dst_score_grad()
  NtsVar backward(NtsVar &grad_out) {
    b200::GatTopology &t = b200::GatTopology::of(partitioned_graph_);
    NtsVar g = grad_out.contiguous();
    const long F = mirror.size(1), H = src_score.size(1), D = F / H;
    // sum_e a[e,h] * <mirror[slot(e),h], g[d,h]> == <out[d,h], g[d,h]>: the softmax backward needs no edge pass
    NtsVar out_dot_g = (out.detach() * g).view({(long)t.V, H, D}).sum(-1).contiguous();
    NtsVar d_mirror = torch::zeros_like(mirror);
    src_score_grad = torch::zeros_like(src_score);
    dst_score_grad = torch::zeros_like(dst_score);
    NtsVar pack = torch::empty({(long)t.V, H, 4}, mirror.options());
    if (nts_gat_fused_aggregate_backward_two_pass(
            d_mirror.data_ptr<float>(), src_score_grad.data_ptr<float>(), dst_score_grad.data_ptr<float>(),
            pack.data_ptr<float>(), mirror.data_ptr<float>(), src_score.data_ptr<float>(), dst_score.data_ptr<float>(),
            seg_max.data_ptr<float>(), seg_sum.data_ptr<float>(), out_dot_g.data_ptr<float>(), g.data_ptr<float>(),
            t.row_indices, t.column_offset, t.mirror_index, t.slot_row_offset, t.slot_column_indices, t.V, t.M,
            (nts_vid_t)F, (nts_vid_t)H, negative_slope, nullptr))
      die("fused GAT backward");
    return d_mirror;
  }

  NtsVar get_additional_grad() { return src_score_grad; }
  NtsVar get_src_score_grad() { return src_score_grad; }
  NtsVar get_dst_score_grad() { return dst_score_grad; }

private:
  static void die(const char *what) {
    std::fprintf(stderr, "nts_b200 %s: %s\n", what, nts_last_error());
    std::exit(1); // the reference's convention for device errors (cuda/ntsCUDAGraphOP.cu:13-19)
  }
  NtsVar mirror, src_score, dst_score, seg_max, seg_sum, out, src_score_grad, dst_score_grad;
};

} // namespace op
} // namespace nts
#endif // CUDA_ENABLE
#endif
