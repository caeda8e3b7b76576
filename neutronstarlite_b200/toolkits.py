"""Callers of the hot path, mirrored from the reference so the path can be driven and timed end to end:

  * `Parameter`  <-> core/NtsScheduler.hpp:639-791 (weight, L2-regularised Adam with the reference's bias-correction
                     folded into alpha by `next()`, gradient SUM-allreduce - NCCL instead of MPI on host copies)
  * `GCNImpl`    <-> toolkits/GCN.hpp (2-layer GCN: aggregate -> X.W -> relu / log_softmax, nll loss on the train
                     mask, tape backward, Adam); the single-GPU op of toolkits/GCN_EAGER_single.hpp when P = 1.
  * `GCNEagerImpl` <-> toolkits/GCN_EAGER_single.hpp / GCN_EAGER.hpp order (X.W first, aggregate the narrow result).
  * `GATImpl`    <-> the flow of toolkits/GAT_CPU_DIST_OPTM.hpp on the fused multi-head aggregation (K7).

Dense NN work (mm, relu, log_softmax, nll_loss, Adam element-wise) stays on torch/cuBLAS exactly as in the
reference (libtorch); the aggregation goes through libnts_b200."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist

from .context import NtsContext
from . import ops


class Parameter:
    def __init__(self, w, h, alpha, beta1, beta2, epsilon, weight_decay, device=None, generator=None):
        scale = math.sqrt(6.0 / (w + h))
        W = (2 * scale) * torch.rand((w, h), dtype=torch.float32, generator=generator) - scale
        self.W = W.to(device).requires_grad_(True)
        self.M = torch.zeros((w, h), dtype=torch.float32, device=device)
        self.V = torch.zeros((w, h), dtype=torch.float32, device=device)
        self.W_gradient = None
        # the reference keeps every hyper-parameter in `ValueType` = float and does the schedule arithmetic in float
        # (1 - 0.999f != 0.001: a 1.3e-5 relative difference in V that the golden vectors of tests/test_adam.py see)
        f32 = np.float32
        self.alpha = f32(alpha)
        self.beta1, self.beta2, self.epsilon = f32(beta1), f32(beta2), f32(epsilon)
        self.alpha_t, self.beta1_t, self.beta2_t = f32(alpha), f32(beta1), f32(beta2)
        self.weight_decay = f32(weight_decay)
        self.curr_epoch = 0
        self.decay_rate, self.decay_epoch = 1, -1

    def init_parameter(self):
        """Network_simple::broadcast from rank 0 (comm/network.h:205-211)."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(self.W.data, src=0)

    def set_decay(self, decay_rate, decay_epoch):
        # the reference stores both in `int` members (NtsScheduler.hpp:663-664): 0.97 truncates to 0
        self.decay_rate, self.decay_epoch = int(decay_rate), int(decay_epoch)

    def all_reduce_to_gradient(self, grad):
        """SUM (not mean) over ranks, NtsScheduler.hpp:719-722 -> comm/network.h:198-203."""
        self.W_gradient = grad.detach().clone().contiguous()
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.W_gradient, op=dist.ReduceOp.SUM)

    def next(self):
        """NtsScheduler.hpp:727-736."""
        if self.decay_epoch != -1 and self.curr_epoch != 0 and self.curr_epoch % self.decay_epoch == 0:
            self.alpha_t *= self.decay_rate
        one = np.float32(1)
        self.alpha_t = np.float32(self.alpha_t)
        self.alpha = np.float32(self.alpha_t * np.sqrt(one - self.beta2) / (one - self.beta1))
        self.beta1 = np.float32(self.beta1 * self.beta1_t)
        self.beta2 = np.float32(self.beta2 * self.beta2_t)
        self.curr_epoch += 1

    def forward(self, x):
        return x.mm(self.W)

    def learn_with_decay_Adam(self):
        """learnC2G_with_decay_Adam, NtsScheduler.hpp:774-781.  On the GPU: ONE fused kernel (nts_adam_update)
        instead of the reference's six element-wise libtorch ops; the torch expression below is the host mirror
        used by the CPU tests (same arithmetic, pinned to the reference's golden vectors in tests/test_adam.py)."""
        with torch.no_grad():
            if self.W.is_cuda:
                from . import _lib
                _lib.call("nts_adam_update", self.W.data_ptr(), self.M.data_ptr(), self.V.data_ptr(),
                          self.W_gradient.data_ptr(), self.W.numel(), float(self.weight_decay), float(self.beta1),
                          float(self.beta2), float(self.alpha), float(self.epsilon),
                          torch.cuda.current_stream().cuda_stream)
                return
            one = np.float32(1)
            W_g = self.W * float(self.weight_decay) + self.W_gradient
            self.M = float(self.beta1) * self.M + float(one - self.beta1) * W_g
            self.V = float(self.beta2) * self.V + float(one - self.beta2) * W_g * W_g
            self.W -= float(self.alpha) * self.M / (torch.sqrt(self.V) + float(self.epsilon))

    def zero_grad(self):
        self.W.grad = None


class GCNImpl:
    """toolkits/GCN.hpp:33-354.  `layers` = LAYERS of the cfg, e.g. [602, 128, 41]."""

    def __init__(self, partitioned_graph, layers, features, labels, mask, learn_rate=0.01, weight_decay=0.0001,
                 decay_rate=0.97, decay_epoch=100, drop_rate=0.5, op_class=None, op_kwargs=None, seed=0):
        self.pg = partitioned_graph
        self.layers = list(layers)
        self.device = features.device
        self.drop_rate = drop_rate
        self.ctx = NtsContext()
        gen = torch.Generator().manual_seed(seed)
        self.P = []
        for i in range(len(self.layers) - 1):
            p = Parameter(self.layers[i], self.layers[i + 1], learn_rate, 0.9, 0.999, 1e-9, weight_decay,
                          device=self.device, generator=gen)
            p.init_parameter()
            p.set_decay(decay_rate, decay_epoch)
            self.P.append(p)
        self.L_GT = labels.to(self.device)
        self.MASK = mask.to(self.device)
        self.train_rows = (self.MASK == 0).nonzero().view(-1)
        self.X = [None] * len(self.layers)
        self.X[0] = features.requires_grad_(True)
        if op_class is None:
            op_class = ops.ForwardSingleGPUfuseOp if partitioned_graph.partitions == 1 else ops.ForwardGPUfuseOp
        self.op_class = op_class
        self.op_kwargs = op_kwargs or {}
        self.loss = None
        self.epoch = 0

    def vertexForward(self, a, x, layer):
        """GCN.hpp:183-196."""
        if layer < len(self.layers) - 2:
            return torch.relu(self.P[layer].forward(a))
        return self.P[layer].forward(a).log_softmax(1)

    def Forward(self):
        """GCN.hpp:217-235."""
        for i in range(len(self.layers) - 1):
            x_i = self.X[i]
            if i != 0 and self.drop_rate > 0:
                # the reference drops in place (GCN.hpp:221-223); out of place + chaining onto the NN segment
                # keeps torch's autograd version check happy and is numerically the same operation
                dropped = torch.nn.functional.dropout(x_i, self.drop_rate, training=True)
                self.ctx.appendNNOp(x_i, dropped)
                x_i = dropped
            x_i = x_i.contiguous()
            y_i = self.ctx.runGraphOp(self.op_class, self.pg, None, x_i, **self.op_kwargs)
            self.X[i + 1] = self.ctx.runVertexForward(lambda n, v, _l=i: self.vertexForward(n, v, _l), y_i, x_i)

    def Loss(self):
        """GCN.hpp:198-207: nll_loss over the local train rows (mean)."""
        a = self.X[-1]
        self.loss = torch.nn.functional.nll_loss(a.index_select(0, self.train_rows),
                                                 self.L_GT.index_select(0, self.train_rows))
        self.ctx.appendNNOp(a, self.loss)

    def Update(self):
        """GCN.hpp:209-215."""
        for p in self.P:
            p.all_reduce_to_gradient(p.W.grad)
            p.learn_with_decay_Adam()
            p.next()

    def Test(self, s):
        """GCN.hpp:150-181: accuracy over mask == s, summed over ranks."""
        sel = self.MASK == s
        correct = (self.X[-1].argmax(1) == self.L_GT)[sel].sum()
        total = sel.sum()
        pair = torch.stack([correct, total]).to(torch.int64)
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(pair)
        return pair

    def run_epoch(self, test=False):
        """One iteration of GCN.hpp:244-262."""
        if self.epoch != 0:
            for p in self.P:
                p.zero_grad()
        self.Forward()
        acc = [self.Test(s) for s in (0, 1, 2)] if test else None
        self.Loss()
        self.ctx.self_backward(True)
        self.Update()
        self.epoch += 1
        return self.loss, acc


class GCNEagerImpl(GCNImpl):
    """Transform-then-aggregate GCN: toolkits/GCN_EAGER_single.hpp:184-232 (P = 1) and toolkits/GCN_EAGER.hpp (P > 1).
    Each layer runs the weight GEMM first and aggregates the NARROW result (widths LAYERS[1:], e.g. 128 and 41
    instead of 602 and 128 - 4.7x fewer gathered bytes on config B), then `log_softmax` + `nll_loss` on the last
    aggregate.  The tape is [NNOP, GRAPHOP, NNOP, GRAPHOP, NNOP(loss)], so every aggregation has a backward
    (L forward + L backward calls per epoch)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        # the input features are the input of the first NN op here; nobody reads their gradient, so do not make
        # autograd compute dX0 = dY0 W0^T (a [V, LAYERS[0]] tensor) every epoch
        self.X[0] = self.X[0].detach()

    def vertexForward(self, a, x, layer):
        """GCN_EAGER_single.hpp:201-212: layer 0 = W0 x; deeper layers = W_l relu(dropout(a))."""
        if layer == 0:
            return self.P[layer].forward(a)
        if self.drop_rate > 0:
            a = torch.nn.functional.dropout(a, self.drop_rate, training=True)
        return self.P[layer].forward(torch.relu(a))

    def Forward(self):
        """GCN_EAGER_single.hpp:214-227."""
        for i in range(len(self.layers) - 1):
            y_i = self.ctx.runVertexForward(lambda n, v, _l=i: self.vertexForward(n, v, _l), self.X[i], self.X[i])
            self.X[i + 1] = self.ctx.runGraphOp(self.op_class, self.pg, None, y_i.contiguous(), **self.op_kwargs)

    def Loss(self):
        """GCN_EAGER_single.hpp:185-194."""
        a = self.X[-1].log_softmax(1)
        self.loss = torch.nn.functional.nll_loss(a.index_select(0, self.train_rows),
                                                 self.L_GT.index_select(0, self.train_rows))
        self.ctx.appendNNOp(self.X[-1], self.loss)


class GATImpl:
    """Multi-head GAT on the fused aggregation path - the flow of toolkits/GAT_CPU_DIST_OPTM.hpp:196-241 (per-vertex
    attention scores -> [E, H] edge logits -> edge softmax -> DistAggregateDstFuseWeight) on the GPU operators, with
    H heads (the reference has one).  `layers` are total widths, e.g. [602, 64, 64, 41] with heads=8 gives hidden
    layers of 8 heads x 8 and a single-head output layer (config D of BASELINE.json).  Never materialises an [E, F]
    message; the only edge-sized tensors are [E, H]."""

    def __init__(self, partitioned_graph, layers, features, labels, mask, heads=8, learn_rate=0.01,
                 weight_decay=0.0001, exchange=None, seed=0, sum_fanout_grads=True, fused_kernel=False,
                 two_pass_backward=True):
        self.pg = partitioned_graph
        self.fused_kernel = fused_kernel  # True: K7 (ops.DistGPUFusedGATOp), no edge-sized tensors at all
        self.two_pass_backward = two_pass_backward
        self.layers = list(layers)
        self.device = features.device
        self.heads = [heads] * (len(self.layers) - 2) + [1]
        self.ctx = NtsContext(sum_fanout_grads=sum_fanout_grads)
        gen = torch.Generator().manual_seed(seed)
        self.P, self.al, self.ar = [], [], []
        for i in range(len(self.layers) - 1):
            H = self.heads[i]
            D = self.layers[i + 1] // H
            assert H * D == self.layers[i + 1], "layer width must be a multiple of heads"
            mk = lambda a, b: Parameter(a, b, learn_rate, 0.9, 0.999, 1e-9, weight_decay, device=self.device,
                                        generator=gen)
            for lst, shape in ((self.P, (self.layers[i], H * D)), (self.al, (H, D)), (self.ar, (H, D))):
                prm = mk(*shape)
                prm.init_parameter()
                lst.append(prm)
        if exchange is None:
            from .exchange import GpuExchange
            exchange = GpuExchange(partitioned_graph)
        self.exchange = exchange
        self.L_GT = labels.to(self.device)
        self.MASK = mask.to(self.device)
        self.train_rows = (self.MASK == 0).nonzero().view(-1)
        self.X = [None] * len(self.layers)
        self.X[0] = features.requires_grad_(True)
        self.loss = None
        self.epoch = 0

    def params(self):
        return self.P + self.al + self.ar

    def Forward(self):
        ctx, pg = self.ctx, self.pg
        for i in range(len(self.layers) - 1):
            H = self.heads[i]
            D = self.layers[i + 1] // H
            last = i == len(self.layers) - 2
            X_trans = ctx.runVertexForward(lambda x, _i=i: self.P[_i].forward(x), self.X[i])
            mirror = ctx.runGraphOp(ops.DistGPUGetDepNbrOp, pg, None, X_trans.contiguous(), exchange=self.exchange)
            src_att = ctx.runVertexForward(
                lambda m, _i=i: (m.view(-1, H, D) * self.al[_i].W).sum(-1).contiguous(), mirror)
            dst_att = ctx.runVertexForward(
                lambda x, _i=i: (x.view(-1, H, D) * self.ar[_i].W).sum(-1).contiguous(), X_trans)
            if self.fused_kernel:
                nbr = ctx.runGraphOpN(ops.DistGPUFusedGATOp, pg, None, [mirror, src_att, dst_att],
                                      two_pass_backward=self.two_pass_backward)
            else:
                e_src = ctx.runGraphOp(ops.DistGPUScatterSrc, pg, None, src_att)
                e_dst = ctx.runGraphOp(ops.DistGPUScatterDst, pg, None, dst_att)
                e_msg = e_src + e_dst  # (the reference concatenates the two [E,1] columns and sums them, :218-224)
                m = ctx.runEdgeForward(lambda t: torch.nn.functional.leaky_relu(t, 0.2), e_msg)
                a = ctx.runGraphOp(ops.DistGPUEdgeSoftMax, pg, None, m)
                nbr = ctx.runGraphOp(ops.DistGPUAggregateDstFuseWeight, pg, None, mirror, a)
            if last:
                self.X[i + 1] = ctx.runVertexForward(lambda t: t.log_softmax(1), nbr)
            else:
                self.X[i + 1] = ctx.runVertexForward(lambda t: torch.relu(t), nbr)

    def Loss(self):
        a = self.X[-1]
        self.loss = torch.nn.functional.nll_loss(a.index_select(0, self.train_rows),
                                                 self.L_GT.index_select(0, self.train_rows))
        self.ctx.appendNNOp(a, self.loss)

    def Update(self):
        for p in self.params():
            if p.W.grad is None:
                continue
            p.all_reduce_to_gradient(p.W.grad)
            p.learn_with_decay_Adam()
            p.next()

    def run_epoch(self):
        if self.epoch != 0:
            for p in self.params():
                p.zero_grad()
        self.Forward()
        self.Loss()
        self.ctx.self_backward(True)
        self.Update()
        self.epoch += 1
        return self.loss
