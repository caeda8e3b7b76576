// Drop-in for `nts::op::DistGPUGetDepNbrOp` of core/ntsDistGPUGraphOp.hpp:48-143 (the mirror fetch every GAT layer of
// toolkits/GAT_GPU_DIST.hpp:187-219 starts with): same class name, constructor, forward / backward and tensor shapes,
// but device-resident.  The original copies the feature matrix to the host, pushes (vid,row) records through
// NtsGraphCommunicator / MPI and copies the mirror matrix back (f_input_.cpu() ... f_output.cuda()); here the owners
// store the rows straight into the readers' CUDA-IPC windows over NVLink (nts_exchange_fetch_mirrors /
// nts_exchange_return_mirror_grads, the same engine and windows as ForwardGPUfuseOp).
//
// Wiring (oracle/Makefile target dropin_dist, no change to the reference tree): dist_fused_prelude.hpp includes
// core/neutronstar.hpp with the original class renamed out of the way by a macro, then this file.
#ifndef NTS_B200_DIST_GPU_GET_DEP_NBR_OP_HPP
#define NTS_B200_DIST_GPU_GET_DEP_NBR_OP_HPP

#if CUDA_ENABLE
#include "nts_dropin/core/ntsDistGPUFusedGraphOp.hpp"

namespace nts {
namespace op {

class DistGPUGetDepNbrOp : public ntsGraphOp {
public:
  std::vector<CSC_segment_pinned *> subgraphs;

  DistGPUGetDepNbrOp(PartitionedGraph *partitioned_graph, VertexSubset *active)
      : ntsGraphOp(partitioned_graph, active) {
    subgraphs = partitioned_graph->graph_chunks;
  }

  // mirror[MirrorIndex[s], :] = X[s, :] for every source s of a local in-edge: [owned_mirrors, F] on the device
  NtsVar forward(NtsVar &f_input_) {
    const int feature_size = f_input_.size(1);
    NtsVar x = f_input_.cuda().contiguous();
    NtsVar f_output = graph_->Nts->NewKeyTensor({(long)partitioned_graph_->owned_mirrors, (long)feature_size},
                                                torch::DeviceType::CUDA);
    b200::DistExchange &ex = b200::DistExchange::of(partitioned_graph_);
    ex.reserve(feature_size);
    if (nts_exchange_fetch_mirrors(ex.engine(), x.data_ptr<float>(), f_output.data_ptr<float>(),
                                   (nts_vid_t)feature_size, nullptr))
      b200::die("nts_exchange_fetch_mirrors");
    return f_output;
  }

  // every partition's mirror gradients go back to the owner of the source vertex, who sums them
  NtsVar backward(NtsVar &f_output_grad_) {
    const int feature_size = f_output_grad_.size(1);
    NtsVar g = f_output_grad_.cuda().contiguous();
    NtsVar f_input_grad = graph_->Nts->NewLeafTensor({(long)graph_->gnnctx->l_v_num, (long)feature_size},
                                                     torch::DeviceType::CUDA);
    b200::DistExchange &ex = b200::DistExchange::of(partitioned_graph_);
    ex.reserve(feature_size);
    if (nts_exchange_return_mirror_grads(ex.engine(), g.data_ptr<float>(), f_input_grad.data_ptr<float>(),
                                         (nts_vid_t)feature_size, nullptr))
      b200::die("nts_exchange_return_mirror_grads");
    return f_input_grad;
  }
};

} // namespace op
} // namespace nts
#endif // CUDA_ENABLE
#endif
