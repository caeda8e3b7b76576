"""`NtsContext`: the reference's hand-rolled autograd tape (core/ntsContext.hpp:96-409), host side only.

Graph operators run hand-written kernels outside torch's autograd; NN segments between them are ordinary torch
autograd graphs.  The tape records, in execution order, GRAPHOP / BIGRAPHOP entries (operator object kept for its
`backward`) and NNOP entries (only the boundary tensors; consecutive NN ops are chained into one entry,
ntsContext.hpp:228-251).  `self_backward` pops the tape exactly like the reference (:276-359), including its
quirk that the loop stops when a single GRAPHOP is left: the FIRST graph operator of a model that starts with one
(GCN.hpp) is never back-propagated (`while (count > 1 || (count == 1 && NNOP == op.top()))`, :283)."""
from __future__ import annotations

import torch

GRAPHOP, BIGRAPHOP, NGRAPHOP, NNOP = "GRAPHOP", "BIGRAPHOP", "NGRAPHOP", "NNOP"


class _Entry:
    __slots__ = ("kind", "op", "input", "output", "o_id", "i_id1", "i_id2", "i_ids", "grad")

    def __init__(self, kind, op, inp, out, i_id2=None, i_ids=None):
        self.kind, self.op, self.input, self.output = kind, op, inp, out
        self.o_id = out.data_ptr()
        self.i_id1 = inp.data_ptr()
        self.i_id2 = i_id2
        self.i_ids = i_ids
        self.grad = None


class NtsContext:
    def __init__(self, sum_fanout_grads=False):
        self.tape = []
        self.training = True
        # The reference overwrites: when a graph-op output feeds BOTH another graph op (gradient assigned through the
        # tape) and an NN segment (gradient accumulated in tensor.grad by torch), only the assigned one survives
        # (`if (output_grad[top].dim() < 2) output_grad[top] = output.top().grad()`, ntsContext.hpp:289-291) - e.g. the
        # attention-score path mirror -> src_att of GAT_CPU_DIST_OPTM.hpp is dropped.  sum_fanout_grads=True adds the
        # two contributions instead (the mathematically complete gradient); default is the reference behaviour.
        self.sum_fanout_grads = sum_fanout_grads

    @property
    def count(self):
        return len(self.tape)

    def train(self):
        self.training = True

    def eval(self):
        self.training = False

    # ---- forward recording -------------------------------------------------------------------------------
    def runGraphOp(self, op_class, partitioned_graph, active, f_input, f_input2=None, **op_kwargs):
        """ntsContext.hpp:108-149 (one- and two-input graph operators)."""
        op = op_class(partitioned_graph, active, **op_kwargs)
        if f_input2 is None:
            out = op.forward(f_input)
            # NewKeyTensor: outputs are leaves that collect gradients.  The one exception: a graph op recorded as
            # the FIRST tape entry is never back-propagated (self_backward stops at it, ntsContext.hpp:283), so
            # nobody ever reads its output gradient - do not make torch compute it (saves the dY = dH.W^T GEMM of
            # the input layer; parameter gradients are unaffected).
            out.requires_grad_(bool(self.tape) or not self.training)
            if self.training:
                self.tape.append(_Entry(GRAPHOP, op, f_input, out))
        else:
            out = op.forward(f_input, f_input2)
            out.requires_grad_(True)
            self.tape.append(_Entry(BIGRAPHOP, op, f_input, out, f_input2.data_ptr()))
        return out

    def runGraphOpN(self, op_class, partitioned_graph, active, inputs, **op_kwargs):
        """Additive generalisation of the two-input form (ntsContext.hpp:130-149) to N tensor inputs: the operator's
        `backward(grad)` returns one gradient per input, each routed to the tape entry that produced that input."""
        op = op_class(partitioned_graph, active, **op_kwargs)
        out = op.forward(*inputs)
        out.requires_grad_(True)
        if self.training:
            self.tape.append(_Entry(NGRAPHOP, op, inputs[0], out, i_ids=[t.data_ptr() for t in inputs]))
        return out

    def runVertexForward(self, vertexforward, nbr_input, vtx_input=None):
        """ntsContext.hpp:198-216."""
        out = vertexforward(nbr_input) if vtx_input is None else vertexforward(nbr_input, vtx_input)
        if self.training:
            self.appendNNOp(nbr_input, out)
        return out

    def runEdgeForward(self, edgeforward, edge_input):
        """ntsContext.hpp:218-226."""
        out = edgeforward(edge_input)
        if self.training:
            self.appendNNOp(edge_input, out)
        return out

    def appendNNOp(self, input_t, output_t):
        """ntsContext.hpp:228-251: chain onto the previous NNOP when this op consumes its output."""
        assert self.training
        if self.tape and self.tape[-1].kind == NNOP and input_t.data_ptr() == self.tape[-1].o_id:
            self.tape[-1].output = output_t
            self.tape[-1].o_id = output_t.data_ptr()
        else:
            self.tape.append(_Entry(NNOP, None, input_t, output_t))

    # ---- backward --------------------------------------------------------------------------------------------
    def _producer_of(self, data_ptr, upto):
        for k in range(upto, -1, -1):
            if self.tape[k].o_id == data_ptr:
                return k
        return -1

    def self_backward(self, retain_graph=True):
        """ntsContext.hpp:276-359."""
        assert self.training and self.tape
        top = self.tape[-1]
        top.output.backward(torch.ones_like(top.output), retain_graph=retain_graph)
        if len(self.tape) >= 2 and not self.sum_fanout_grads:
            # (in sum mode the entry fetches the same tensor through output.grad below; assigning it here as well
            #  would count it twice)
            self.tape[-2].grad = top.input.grad
        self.tape.pop()
        while len(self.tape) > 1 or (len(self.tape) == 1 and self.tape[-1].kind == NNOP):
            e = self.tape[-1]
            idx = len(self.tape) - 1
            if e.grad is None or e.grad.dim() < 2:
                e.grad = e.output.grad
            elif self.sum_fanout_grads and e.kind != NNOP and e.output.grad is not None:
                e.grad = e.grad + e.output.grad
            if e.kind == NGRAPHOP:
                grads = e.op.backward(e.grad)
                for ptr, g_in in zip(e.i_ids, grads):
                    k = self._producer_of(ptr, idx)
                    if k >= 0:
                        prev = self.tape[k].grad
                        self.tape[k].grad = g_in if prev is None or prev.dim() < 2 else prev + g_in
            elif e.kind in (GRAPHOP, BIGRAPHOP):
                g_in = e.op.backward(e.grad)
                k = self._producer_of(e.i_id1, idx)
                if k >= 0:
                    self.tape[k].grad = g_in
                if e.kind == BIGRAPHOP:
                    k2 = self._producer_of(e.i_id2, idx)
                    if k2 >= 0:
                        self.tape[k2].grad = e.op.get_additional_grad()
            else:  # NNOP: torch autograd carries the gradient through the NN segment
                if e.grad is not None and e.grad.dim() > 1:
                    assert e.grad.shape == e.output.shape
                    e.output.backward(e.grad, retain_graph=retain_graph)
            self.tape.pop()
        self.reset()

    def reset(self):
        assert len(self.tape) <= 1
        self.tape = []
