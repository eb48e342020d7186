"""NeuralNet: theanet's network object (reference: theanet/neuralnet.py:59-333) on hand-written
sm_100a kernels.

Same constructor (``NeuralNet(layers, training_params, allwts=None)``), same ``get_trin_model`` /
``get_test_model`` / ``get_data_test_model`` callables, same ``.prms`` vocabulary and ``.pkl``
schema, so the reference's ``train.py`` loop runs unchanged against it.  What differs is below
the API: there is no symbolic graph and no autodiff.  The constructor lays all parameters,
velocities and gradients out in three flat device buffers, allocates one activation tensor per
layer, and a training step is a fixed sequence of C-ABI kernel launches (forward, softmax-NLL,
hand-derived backward, one optional NCCL all-reduce of the flat gradient buffer, one fused
optimiser launch) that is captured once into a CUDA graph and replayed per minibatch.  Per-step
scalars (step counter, corpus row, learning rate) live in a small device control block that the
graph refreshes from pinned host memory.

Beyond the hot path the same engine runs ColorLayer, MeanLayer, strided
convolutions, the auxiliary-input layers, ExpLossLayer, HingeLayer and the nllsq / truncated-NLL
losses; CenteredOutLayer is not implemented (it cannot be trained in the reference either).

PyTorch is used for device memory, streams, CUDA graphs and torch.distributed only.
"""
import ctypes
import os
import weakref
from contextlib import contextmanager
from types import SimpleNamespace
from functools import reduce
from operator import mul

import numpy as np
import torch

from . import _C
from . import layer
from .dist import DistContext
from .layer import (InputLayer, ElasticLayer, ColorLayer, ConvLayer, PoolLayer, MeanLayer,
                    DropOutLayer, HiddenLayer, SoftmaxLayer, ExpLossLayer, HingeLayer, OutputLayer,
                    AuxConcatLayer, SoftAuxLayer, OUT_KINDS)

# ########################### Helper Functions #################################


def get_layers_info(layers):
    lines = []
    for name, args in layers:
        lines.append('\n{} : '.format(name))
        lines.extend('\n\t{} : \t{}'.format(k, args[k]) for k in args)
    return ''.join(lines)


def get_wts_info(wts, detailed=False):
    out, n_wts = [], 0
    for l, ww in enumerate(wts):
        out.append("\nLayer {}:".format(l))
        for w in ww:
            n_ww = reduce(mul, w.shape)
            n_wts += n_ww
            out.append('\n\t {} {} ❲{}❳'.format(w.shape, w.dtype, n_ww))
            if detailed:
                out.append(" ❲{:.2e}, {:.2e}, {:.2e}❳".format(w.min(), w.mean(), w.max()))
    out.append('\n\nTotal Number of Weights : {:,}'.format(n_wts))
    return ''.join(out)


def get_training_params_info(training_params):
    return "Training Parameters:" + ''.join(
        '\n\t{} : \t{}'.format(k, training_params[k]) for k in sorted(training_params.keys()))


class _Scalar:
    """Stand-in for the reference's shared learning-rate scalar (neuralnet.py:110)."""

    def __init__(self, v=0.0):
        self.v = np.float32(v)

    def set_value(self, v):
        self.v = np.float32(v)

    def get_value(self):
        return self.v


def _as_numpy(data):
    if hasattr(data, 'get_value'):
        data = data.get_value()
    if isinstance(data, torch.Tensor):
        return data
    return np.asarray(data)


###############################################################################
#                           The Neural Network
###############################################################################


class NeuralNet():
    def __init__(self, layers, training_params, allwts=None, test_x=None, device=None, dist=None,
                 use_graph=True, fuse_conv=True, fuse_head=True):
        if allwts is None:
            self.rand_gen = np.random.RandomState(training_params['SEED'])
        else:
            self.rand_gen = None
        self.tr_prms = training_params
        self.layers = layers
        self.allwts = allwts
        self.tr_layers = []
        self.te_layers = []
        self.dist = dist or DistContext()
        self.batch_sz = training_params['BATCH_SZ']          # global minibatch
        assert self.batch_sz % self.dist.world == 0, "BATCH_SZ must divide by the world size"
        self.local_bsz = self.batch_sz // self.dist.world
        self.num_layers = 0
        if device is None:
            device = 'cuda' if torch.cuda.is_available() else 'cpu'
        self.device = torch.device(device)
        self.use_graph = use_graph and self.device.type == 'cuda'
        self.fuse_conv = fuse_conv
        self.fuse_head = fuse_head
        # data parallel: the collective is captured into the step's CUDA graph and bucketed -- the
        # gradients of the trailing dense layers (99% of the bytes in the shipped networks) are
        # all-reduced on a side stream while the conv layers below are still back-propagating;
        # TN_GRAPH_NCCL=0: graph -> one eager all-reduce -> graph (round 1)
        self.nccl_in_graph = os.environ.get('TN_GRAPH_NCCL', '1') == '1'
        self._comm = None
        self.overlap_wgrad = os.environ.get('TN_OVERLAP_WGRAD', '1') == '1'
        self.sm_count = (torch.cuda.get_device_properties(self.device).multi_processor_count
                         if self.device.type == 'cuda' else 148)
        self._side = None

        # Input Layer
        input_layer_type = getattr(layer, layers[0][0])
        assert input_layer_type in (InputLayer, ElasticLayer, ColorLayer), \
            "First layer needs to be Input or Elastic or Color Layer"
        self.tr_layers.append(input_layer_type(None, rand_gen=self.rand_gen, **layers[0][1]))
        self.te_layers.append(self.tr_layers[0].TestVersion(None))
        self.num_layers += 1
        while self.num_layers < len(layers):
            self.append_next_layer()
        assert isinstance(self.tr_layers[-1], OutputLayer), "last layer must be an output layer"
        # Handle Auxiliary input (neuralnet.py:99-104)
        self.aux_index = None
        for li, tr_layer in enumerate(self.tr_layers):
            if type(tr_layer) in (AuxConcatLayer, SoftAuxLayer):
                assert self.aux_index is None, "Multiple Aux Inputs"
                self.aux_index = li
                self.aux_inpt_tr = tr_layer.aux_inpt
                self.aux_inpt_te = self.te_layers[li].aux_inpt
        self._aux_corpus = None

        if 'CUR_EPOCH' not in training_params:
            training_params['CUR_EPOCH'] = 0
        self.cur_learn_rate = _Scalar(0.0)
        self.set_rate()

        self.step_count = 0
        self._early_reduced = None
        self.inject = {}            # (layer index, 'noise'|'u'|'flip'|'mask') -> device tensor
        self.keep_conv_out = False  # also write the un-pooled output of fused conv+pool layers when training
        self.debug_elastic = False
        self._allocate()

    # ------------------------------------------------------------------------------------------
    def append_next_layer(self):
        layer_type, layer_args = self.layers[self.num_layers]
        prev_tr_layer = self.tr_layers[self.num_layers - 1]
        prev_te_layer = self.te_layers[self.num_layers - 1]
        wts = self.allwts[self.num_layers] if self.allwts else None
        tr_inpt, te_inpt = prev_tr_layer.output, prev_te_layer.output
        curr_layer_type = getattr(layer, layer_type, None)

        if curr_layer_type in (ElasticLayer, ColorLayer, ConvLayer, PoolLayer, MeanLayer):
            # a DropOutLayer carries no map geometry: look through it (neuralnet.py:123-130)
            use = self.tr_layers[self.num_layers - 2] if type(prev_tr_layer) is DropOutLayer \
                else prev_tr_layer
            num_prev_maps, prev_out_sz = use.num_maps, use.out_sz

        if curr_layer_type in (ElasticLayer, ColorLayer):     # neuralnet.py:132-142
            layer_args.pop("num_maps", None)
            layer_args.pop("img_sz", None)
            curr_layer = curr_layer_type(tr_inpt, num_maps=num_prev_maps, img_sz=prev_out_sz,
                                         rand_gen=self.rand_gen, **layer_args)
        elif curr_layer_type is ConvLayer:
            curr_layer = ConvLayer(tr_inpt, wts, self.rand_gen, self.batch_sz, num_prev_maps,
                                   prev_out_sz, **layer_args)
        elif curr_layer_type is MeanLayer:
            curr_layer = MeanLayer(tr_inpt, num_maps=num_prev_maps, in_sz=prev_out_sz,
                                   **layer_args)
        elif curr_layer_type is PoolLayer:
            curr_layer = PoolLayer(tr_inpt, num_maps=num_prev_maps, in_sz=prev_out_sz,
                                   **layer_args)
        elif curr_layer_type is DropOutLayer:
            curr_layer = DropOutLayer(tr_inpt, self.rand_gen, prev_tr_layer.n_out, **layer_args)
        elif curr_layer_type in (AuxConcatLayer, HiddenLayer, SoftmaxLayer, SoftAuxLayer,
                                 ExpLossLayer, HingeLayer):
            te_inpt = te_inpt.flatten(2)
            curr_layer = curr_layer_type(tr_inpt.flatten(2), wts, self.rand_gen,
                                         prev_tr_layer.n_out, **layer_args)
        else:
            raise NotImplementedError("Unknown Layer Type" + layer_type)

        self.tr_layers.append(curr_layer)
        self.te_layers.append(curr_layer.TestVersion(te_inpt))
        self.num_layers += 1

    # ------------------------------------------------------------------------------------------
    # memory layout
    # ------------------------------------------------------------------------------------------
    def _allocate(self):
        dev, B = self.device, self.local_bsz
        f32 = torch.float32
        # flat parameter / velocity / gradient buffers; every tensor starts on a 16-byte boundary
        params = []
        for lyr in self.tr_layers:
            for p in lyr.params:
                if p not in params:
                    params.append(p)
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.size + 3) // 4 * 4
        self.n_flat = total
        self.theta = torch.zeros(total, dtype=f32, device=dev)
        self.vel = torch.zeros(total, dtype=f32, device=dev)
        # gradient buffer + 4 trailing floats: [nll partial sum, pad]; one all-reduce covers both
        # data parallel: either one NCCL all-reduce of `grad` per step, or (TN_DP_FUSED=1) the
        # all-reduce folded into the optimiser kernel over CUDA-IPC mapped peer buffers, with the
        # gradient buffer double-buffered by step parity (update.cu)
        # default (hybrid): the big early bucket goes through NCCL inside the graph, overlapped with
        # the conv backward pass; the few late conv gradients are summed over peer memory by the
        # update kernel itself (no second collective, no launch gap before the update)
        self.dp_hybrid = (self.dist.world > 1 and self.nccl_in_graph and
                          os.environ.get('TN_DP_HYBRID', '1') == '1' and self.device.type == 'cuda')
        self.dp_fused = self.dist.world > 1 and (os.environ.get('TN_DP_FUSED', '0') == '1' or
                                                 self.dp_hybrid)
        self.peer_ar = os.environ.get('TN_PEER_AR', '1') == '1'    # 0: NCCL for the big bucket
        if self.dp_fused:
            from .dist import PeerBuffers
            self.peers = PeerBuffers(self.dist, total + 4, dev)
            self.grads2 = self.peers.grad
        else:
            self.grads2 = [torch.zeros(total + 4, dtype=f32, device=dev)]
        self.params = params
        self.param_offset = dict()
        self._grad_views = []
        for gbuf in self.grads2:
            self._grad_views.append([gbuf[o:o + p.size].view(p.shape) for p, o in zip(params, offs)])
        for p, o in zip(params, offs):
            shp = p.shape
            p.bind(self.theta[o:o + p.size].view(shp), self.vel[o:o + p.size].view(shp), None)
            self.param_offset[id(p)] = o
        self._bind_grads(0)
        # optimiser segment table (theanet/layer/layer.py:70-107)
        segs = []
        for lyr in self.tr_layers:
            for p, reg in lyr.update_segments():
                s = _C.ParamSeg()
                s.offset, s.size, s.ndim = self.param_offset[id(p)], p.size, p.ndim
                if p.ndim == 2:
                    s.rows, s.cols = p.shape
                elif p.ndim == 4:
                    s.rows, s.cols = p.shape[0], p.size // p.shape[0]
                else:
                    s.rows, s.cols = 1, p.size
                s.momentum, s.rate = reg['momentum'], reg['rate'] or 0
                s.maxnorm, s.l1, s.l2 = reg['maxnorm'] or 0, reg['L1'], reg['L2']
                segs.append(s)
        self.segs = (_C.ParamSeg * max(1, len(segs)))(*segs)
        self.n_segs = len(segs)
        trainable = [bool(getattr(l, 'reg', None) and l.reg['rate'] and l.params)
                     for l in self.tr_layers]
        self.trainable = trainable
        # need_below[li]: some trainable layer sits strictly below li
        self.need_below = [any(trainable[:li]) for li in range(len(self.tr_layers))]

        # activations (shared by the train and test twins), gradient buffers
        self.out, self.dbuf = [], []
        for li, lyr in enumerate(self.tr_layers):
            shp = (B,) + lyr.output.shape
            t = torch.empty(shp, dtype=f32, device=dev)
            self.out.append(t)
            lyr.output.tensor = t
            self.te_layers[li].output.tensor = t
            need = li + 1 < len(self.tr_layers) and self.need_below[li + 1]
            self.dbuf.append(torch.empty(shp, dtype=f32, device=dev) if need else None)
        last = self.tr_layers[-1]
        n_out = last.n_out
        self.z = self.out[-1]                                   # pre-softmax scores
        # log-probabilities and the cost share one buffer: a call that returns them to the host
        # (the reference's training function does, neuralnet.py:236-241) needs ONE device-to-host copy
        self._res = torch.zeros(B * n_out + 4, dtype=f32, device=dev)
        self.logprob = self._res[:B * n_out].view(B, n_out)
        # output stage (outlayers.py): kind of layer, loss; SoftmaxLayer + 'nll' is the hot path
        self.out_kind = OUT_KINDS[last.kind]
        self.loss_code, self.log_thr = last.cost()
        self.plain_nll = self.out_kind == _C.OUT_SOFTMAX and self.loss_code == _C.LOSS_NLL
        # ExpLossLayer's features are the centred scores, not its log-probabilities
        self.feat = torch.empty((B, n_out), dtype=f32, device=dev) \
            if self.out_kind == _C.OUT_EXPLOSS else self.logprob
        self.gsoft = torch.empty((B, n_out), dtype=f32, device=dev)
        self.rowloss = torch.empty(B, dtype=f32, device=dev)
        self.cost = self._res[B * n_out:B * n_out + 1]
        self.stats = torch.zeros(2 + 2 * B, dtype=f32, device=dev)
        self.preds = torch.zeros(B, dtype=torch.int64, device=dev)
        for l in (last, self.te_layers[-1]):
            l.logprob.tensor = self.logprob
            l.features.tensor = self.feat
            l.y_preds.tensor = self.preds
        # control block: per-step scalars, written by a tiny kernel launched ahead of the captured
        # graph (tn_set_ctl: the values are launch arguments, so a host running ahead of the device
        # cannot overwrite what a queued step has yet to read).  Index vectors go through a ring
        # of pinned buffers, each guarded by an event before it is reused.
        pin = self.device.type == 'cuda'
        self.ctl = torch.zeros(_C.CTL_WORDS, dtype=torch.int32, device=dev)
        self._ctl_np = np.zeros(_C.CTL_WORDS, dtype=np.int32)      # host mirror of the last write
        self._ctl_ptr = self.ctl.data_ptr()
        self._lr_cached = (None, 0)
        self._ctl_issued = None
        self._idx_ring = [torch.zeros(B, dtype=torch.int32, pin_memory=pin) for _ in range(8)]
        self._idx_ev = [None] * len(self._idx_ring)
        self._idx_k = 0
        self.idx = torch.zeros(B, dtype=torch.int32, device=dev)
        # elastic scratch
        # the (single) distorting ElasticLayer: normally layer 0, possibly behind a Color / Input layer
        els = [i for i, l in enumerate(self.tr_layers) if isinstance(l, ElasticLayer) and not l.identity]
        if len(els) > 1:
            raise NotImplementedError("more than one distorting ElasticLayer")
        self.el_index = els[0] if els else None
        self.idx_iota = torch.arange(B, dtype=torch.int32, device=dev)   # rows of an in-flight batch
        l0 = self.tr_layers[self.el_index] if els else None
        if l0 is not None:
            h = l0.img_sz
            self.el_noise = torch.zeros(2 * h * h, dtype=f32, device=dev)
            # two sets of sampling grids: under CUDA graphs the field of step s+1 is computed on a
            # side branch while step s runs (it depends on (seed, step) only), see _train_launches
            self.el_grids = [(torch.zeros(h * h, dtype=torch.int32, device=dev),
                              torch.zeros(2 * h * h, dtype=f32, device=dev)) for _ in range(2)]
            self.el_gidx, self.el_gfrac = self.el_grids[0]
            self._field_step = None          # step whose field sits in el_grids[step & 1]
            self.el_target = torch.zeros(2 * h * h, dtype=torch.float64, device=dev)
            self.el_tyx = torch.zeros(2 * h * h, dtype=torch.float64, device=dev)
            self.el_filt = torch.from_numpy(l0.filt).to(dev) if l0.filt is not None else None
            prm = _C.ElasticPrm()
            prm.h, prm.sigma = h, int(l0.sigma)
            prm.translation, prm.magnitude = float(l0.translation), float(l0.magnitude)
            prm.zoom_on = int(l0.zoom != 1)
            prm.log_zoom = float(np.float32(np.log(l0.zoom)))
            prm.angle_rad = float(np.float32(l0.angle * np.pi / 180))
            prm.nearest = int(bool(l0.nearest))
            prm.clip_hi = h - 1 - .001
            prm.step_offset = 0
            self.el_prm = prm
            nxt = _C.ElasticPrm.from_buffer_copy(prm)
            nxt.step_offset = 1
            self.el_prm_next = nxt
        self.field_prefetch = (os.environ.get('TN_FIELD_PREFETCH', '1') == '1' and self.use_graph
                               and l0 is not None and l0.has_grid)
        # workspaces; conv_fused[li]: ConvLayer li (+ the PoolLayer right above it) runs on the
        # fused small-channel kernels (conv_fused.cu)
        self.ws = {}
        self.conv_fused = {}
        self.conv_small = {}
        self.conv_full = {}
        self._plan_conv_tc()
        for li, lyr in enumerate(self.tr_layers):
            if isinstance(lyr, ConvLayer) and li in self.conv_tc:
                continue
            if isinstance(lyr, ConvLayer):
                nb = _C.lib.tn_conv2d_wgrad_workspace_bytes(B, lyr.num_prev_maps, lyr.in_sz,
                                                            lyr.num_maps, lyr.filter_sz)
                nxt = self.tr_layers[li + 1] if li + 1 < len(self.tr_layers) else None
                if lyr.stride != 1:      # stride-1 activations and gradients at full resolution
                    full = (B, lyr.num_maps, lyr.full_sz, lyr.full_sz)
                    self.conv_full[li] = (torch.empty(full, dtype=f32, device=dev),
                                          torch.empty(full, dtype=f32, device=dev))
                if self.fuse_conv and lyr.stride == 1 and isinstance(nxt, PoolLayer):
                    # second-generation kernels (conv_small.cu): groups of images per CTA, one
                    # backward launch per layer; else the first-generation fused kernels
                    geom = (lyr.num_prev_maps, lyr.in_sz, lyr.num_maps, lyr.filter_sz, lyr.pad_lo,
                            lyr.out_sz, lyr.act.code, nxt.pool_sz, nxt.out_sz)
                    small = bool(_C.lib.tn_convpool_small_supported(*geom))
                    gen1 = self._fused_conv_ok(lyr)
                    if small or gen1:
                        self.conv_fused[li] = nxt
                    if gen1:
                        nb = max(nb, _C.lib.tn_convpool_bwd_weights_workspace_bytes(
                            B, lyr.num_prev_maps, lyr.num_maps, lyr.filter_sz))
                    if small:
                        nbs = _C.lib.tn_convpool_bwd_workspace_bytes(B, *geom,
                                                                     int(self.need_below[li]))
                        # zero-filled: the kernel's ticket counters start (and are left) at zero;
                        # tie: which window elements equal the pooled maximum, written by the
                        # training forward, read by the backward kernel instead of out[li]
                        self.conv_small[li] = SimpleNamespace(
                            ws=torch.zeros((nbs + 3) // 4, dtype=f32, device=dev),
                            tie=torch.zeros(B * lyr.num_maps * nxt.out_sz ** 2, dtype=torch.uint8,
                                            device=dev),
                            tie_valid=False)
                self.ws[li] = torch.empty((nb + 3) // 4, dtype=f32, device=dev)
        # auxiliary-input scratch (auxiliary.py): the gathered (B,2,2) rows, the mixed input, the
        # two LocationInfo activations and, for SoftAuxLayer, their gradients and the cross term
        if self.aux_index is not None:
            ai = self.tr_layers[self.aux_index].aux_info
            z = lambda *shp: torch.zeros(shp, dtype=f32, device=dev)   # noqa: E731
            self.aux_batch, self.aux_loc = z(B, 4), z(B, 2)
            self.aux_hid, self.aux_out = z(B, ai.n_hid), z(B, ai.n_out)
            self.aux_ghid, self.aux_gout = z(B, ai.n_hid), z(B, ai.n_out)
            self.aux_cross = z(B, n_out)
        # classifier head (narrow SoftmaxLayer) on the fused kernels of head.cu
        self.head = bool(self.fuse_head and self.trainable[-1] and self.plain_nll and
                         not isinstance(last, SoftAuxLayer) and
                         _C.lib.tn_softmax_head_supported(last.n_in, last.n_out))
        if self.head:
            nb = _C.lib.tn_softmax_head_workspace_bytes(B, last.n_in, last.n_out)
            self.ws_head = torch.zeros((nb + 3) // 4, dtype=f32, device=dev)   # tickets start at 0
        nb = _C.lib.tn_update_workspace_bytes(max(1, self.n_segs), total)
        self.ws_update = torch.zeros((nb + 3) // 4, dtype=f32, device=dev)    # ticket starts at 0
        self._graphs = {}
        self.launches = {}          # 'train' / 'test' -> kernels of this library per step
        if self.dist.world > 1:     # replicas must start identical (rank 0 wins)
            torch.distributed.broadcast(self.theta, src=0, group=self.dist.group)

    def _bind_grads(self, parity):
        """Point every parameter's .grad (and the NLL slot) at the gradient buffer of this parity."""
        k = parity % len(self.grads2)
        self.grad = self.grads2[k]
        self.nll_sum = self.grad[self.n_flat:self.n_flat + 1]
        for p, v in zip(self.params, self._grad_views[k]):
            p.grad = v
        self._grad_parity = k

    # ------------------------------------------------------------------------------------------
    # mixed-precision conv stack (training_params['CONV_DTYPE'] = 'bfloat16'; config C4)
    # ------------------------------------------------------------------------------------------
    def _conv_tc_kind(self, li):
        """'tc' (wide layer, direct), 'im2col' (first weighted layer with C*f*f <= 64) or None."""
        lyr = self.tr_layers[li]
        nxt = self.tr_layers[li + 1] if li + 1 < len(self.tr_layers) else None
        if lyr.mode != 'same' or lyr.act.code not in (_C.ACT_LINEAR, _C.ACT_RELU, _C.ACT_LEAKY):
            return None
        if lyr.stride != 1:
            return None
        if isinstance(nxt, MeanLayer):                # float32 path: its backward fuses act' itself
            return None
        if isinstance(nxt, PoolLayer) and (nxt.pool_sz != 2 or lyr.out_sz % 2 or nxt.ignore_border):
            return None
        C_, S, M, f, O = lyr.num_prev_maps, lyr.in_sz, lyr.num_maps, lyr.filter_sz, lyr.out_sz
        if _C.lib.tn_conv2d_tc_supported(C_, S, M, f, O):
            return 'tc'
        if (not self.need_below[li] and f == 3 and C_ <= 7 and
                _C.lib.tn_conv2d_tc_supported(64, S, M, 1, O)):
            return 'im2col'
        return None

    def _plan_conv_tc(self):
        """conv_tc[li]: buffers of ConvLayer li (and the PoolLayer above it) on the tcgen05 bf16
        implicit-GEMM kernels (conv_tc.cu).  Activations inside the stack are NHWC bfloat16; they
        are converted from / to the float32 NCHW tensors of the neighbouring layers at its ends."""
        self.conv_tc = {}
        if str(self.tr_prms.get('CONV_DTYPE', 'float32')).lower() not in ('bf16', 'bfloat16'):
            return
        dev, B, bf = self.device, self.local_bsz, torch.bfloat16
        L = self.tr_layers
        for li, lyr in enumerate(L):
            kind = self._conv_tc_kind(li) if isinstance(lyr, ConvLayer) else None
            if kind is None:
                continue
            nxt = L[li + 1] if li + 1 < len(L) else None
            st = SimpleNamespace(pool=nxt if isinstance(nxt, PoolLayer) else None,
                                 im2col=kind == 'im2col')
            S, O, Ci, M, f = lyr.in_sz, lyr.out_sz, lyr.num_prev_maps, lyr.num_maps, lyr.filter_sz
            st.fuse_pool = st.pool is not None and O <= 16     # else a separate NHWC pool kernel
            # input: the bf16 output of a tensor-core block right below, a converted copy, or the
            # 64-wide im2col of a narrow first layer
            src = None
            if not st.im2col:
                if (li - 1) in self.conv_tc and self.conv_tc[li - 1].pool is None:
                    src = self.conv_tc[li - 1].a
                elif (li - 2) in self.conv_tc and self.conv_tc[li - 2].pool is L[li - 1]:
                    src = self.conv_tc[li - 2].pooled
            st.convert_in = src is None and not st.im2col
            Cx = 64 if st.im2col else Ci
            st.xin = src if src is not None else torch.empty((B, S, S, Cx), dtype=bf, device=dev)
            st.a = torch.empty((B, O, O, M), dtype=bf, device=dev)
            st.pooled = torch.empty((B, O // 2, O // 2, M), dtype=bf, device=dev) if st.pool else None
            st.gz = torch.empty((B, O, O, M), dtype=bf, device=dev)
            st.dx = torch.empty((B, S, S, Ci), dtype=bf, device=dev) if self.need_below[li] else None
            nW = M * 64 if st.im2col else M * f * f * Ci
            st.Wp = torch.empty(nW, dtype=bf, device=dev)
            st.Wpd = torch.empty(nW, dtype=bf, device=dev)
            st.dWcol = torch.empty(M * 64, dtype=torch.float32, device=dev) if st.im2col else None
            nb = _C.lib.tn_conv2d_tc_wgrad_workspace_bytes(B, Cx, M, 1 if st.im2col else f, O)
            st.ws = torch.empty((nb + 3) // 4, dtype=torch.float32, device=dev)
            st.out_idx = li + 1 if st.pool else li
            st.convert_out = True                 # cleared below when a tensor-core conv consumes it
            self.conv_tc[li] = st
            if src is not None:
                below = self.conv_tc[li - 1] if (li - 1) in self.conv_tc else self.conv_tc[li - 2]
                below.convert_out = False

    @staticmethod
    def _fused_conv_ok(lyr):
        """Limits of the fused conv(+pool) kernels (conv_fused.cu: fill_geom and shared memory)."""
        f, C_, M, S, O = lyr.filter_sz, lyr.num_prev_maps, lyr.num_maps, lyr.in_sz, lyr.out_sz
        G, CG = (M + 3) // 4, (C_ + 3) // 4
        if f not in (3, 5) or G * C_ * f > 256 or M * O * O >= 65536 or C_ * S * S >= 65536:
            return False
        strips_o, strips_s = (O + 3) // 4, (S + 3) // 4
        ng = max(1, min(256 // max(1, CG * S * strips_s), M))
        smem = 4 * max(
            C_ * f * f * 4 * G + C_ * (O + f - 1) * (strips_o * 4 + f - 1) + M * O * O,
            C_ * (O + f - 1) ** 2 + 4 * G * O * O,
            M * f * f * 4 * CG + M * (S + f - 1) * (strips_s * 4 + f - 1) + ng * 4 * CG * S * S)
        return smem <= 96 * 1024

    # ------------------------------------------------------------------------------------------
    # launch helpers
    # ------------------------------------------------------------------------------------------
    def _stream(self):
        if self.device.type != 'cuda':
            raise RuntimeError("theanet_b200 has no CPU execution path: a B200 (sm_100a) device "
                               "is required to run the network")
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _inj(self, li, kind):
        t = self.inject.get((li, kind))
        return _C.ptr(t)

    def _forward(self, layers, train, corpus, idx, labels):
        """Launch the forward pass of ``layers`` (the train or the test twins)."""
        st = self._stream()
        B = self.local_bsz
        ctl = _C.ptr(self.ctl)
        idxp = _C.ptr(idx)
        for li, lyr in enumerate(layers):
            if train and self.head and li == len(layers) - 1:
                break                                 # scores are computed by the fused head
            out = self.out[li]
            x = self.out[li - 1] if li else None
            if isinstance(lyr, ColorLayer):
                C_, h = lyr.num_maps, lyr.out_sz
                if li == 0:       # batch loader first (tn_elastic_warp mode 0 = gather the rows)
                    _C.call('tn_elastic_warp', _C.ptr(corpus), idxp, ctl, B, C_, h, 0, 0, None,
                            None, 0.0, None, 0, _C.ptr(out), st)
                    x = out
                if train and not lyr.identity:
                    _C.call('tn_color_jitter', _C.ptr(x), _C.ptr(out), B, C_, h, lyr.log_balance,
                            lyr.log_gamma, float(lyr.maxval), lyr.seed, ctl,
                            self._inj(li, 'color'), st)
                elif li:                                  # identity twin: plain copy
                    _C.call('tn_dropout_apply', _C.ptr(x), _C.ptr(out), B, C_ * h * h, 1.0, 0, ctl,
                            None, 1.0, st)
            elif isinstance(lyr, (InputLayer, ElasticLayer)):
                if li and isinstance(lyr, InputLayer):
                    raise NotImplementedError("InputLayer past position 0")
                # layer 0 gathers its rows from the corpus; further up the source is the batch
                # the layer below produced (rows 0..B-1)
                src, rows = (_C.ptr(corpus), idxp) if li == 0 else (_C.ptr(x), _C.ptr(self.idx_iota))
                # geometry from the TRAIN twin: the reference's Elastic test twin does not forward
                # num_maps (inlayers.py:157-163), harmless there because it is elementwise
                C_, h = self.tr_layers[li].num_maps, lyr.out_sz
                invert = int(getattr(lyr, 'invert', False))
                mode, gidx, gfrac, pflip, seed = 0, None, None, 0.0, 0
                if train and isinstance(lyr, ElasticLayer) and not lyr.identity:
                    seed = lyr.seed
                    if lyr.has_grid:
                        if self._prefetching():           # computed during the previous step
                            g_i, g_f = self.el_grids[self.step_count & 1]
                        else:
                            g_i, g_f = self.el_gidx, self.el_gfrac
                            self._launch_field(self.el_prm, g_i, g_f, st)
                        mode = 1 if lyr.nearest else 2
                        gidx, gfrac = _C.ptr(g_i), _C.ptr(g_f)
                    pflip = float(lyr.pflip)
                _C.call('tn_elastic_warp', src, rows, ctl, B, C_, h, invert, mode, gidx,
                        gfrac, pflip, self._inj(li, 'flip') if train else None, seed, _C.ptr(out),
                        st)
            elif isinstance(lyr, ConvLayer) and li in self.conv_tc:
                t = self.conv_tc[li]
                S_, O_, Ci, M_, f_ = lyr.in_sz, lyr.out_sz, lyr.num_prev_maps, lyr.num_maps, lyr.filter_sz
                fp = _C.ptr(t.pooled) if t.fuse_pool else None
                if t.im2col:
                    _C.call('tn_im2col_bf16', _C.ptr(x), _C.ptr(t.xin), B, Ci, S_, f_, lyr.pad_lo, st)
                    _C.call('tn_conv2d_tc_pack_weights_im2col', _C.ptr(lyr.W.tensor), _C.ptr(t.Wp),
                            M_, Ci, f_, st)
                    _C.call('tn_conv2d_tc_fprop', _C.ptr(t.xin), _C.ptr(t.Wp), _C.ptr(lyr.b.tensor),
                            _C.ptr(t.a), fp, B, 64, S_, M_, 1, 0, O_, lyr.act.code, lyr.act.nn, st)
                else:
                    if t.convert_in:
                        _C.call('tn_nchw_f32_to_nhwc_bf16', _C.ptr(x), _C.ptr(t.xin), B, Ci, S_, S_, st)
                    _C.call('tn_conv2d_tc_pack_weights', _C.ptr(lyr.W.tensor), _C.ptr(t.Wp), M_, Ci,
                            f_, 0, st)
                    if train and self.need_below[li]:
                        _C.call('tn_conv2d_tc_pack_weights', _C.ptr(lyr.W.tensor), _C.ptr(t.Wpd), M_,
                                Ci, f_, 1, st)
                    _C.call('tn_conv2d_tc_fprop', _C.ptr(t.xin), _C.ptr(t.Wp), _C.ptr(lyr.b.tensor),
                            _C.ptr(t.a), fp, B, Ci, S_, M_, f_, lyr.pad_lo, O_, lyr.act.code,
                            lyr.act.nn, st)
                if t.pool is not None and not t.fuse_pool:
                    _C.call('tn_maxpool2_nhwc_bf16', _C.ptr(t.a), _C.ptr(t.pooled), B, O_, M_, st)
                if t.convert_out:
                    src = t.pooled if t.pool else t.a
                    Po = O_ // 2 if t.pool else O_
                    _C.call('tn_nhwc_bf16_to_nchw_f32', _C.ptr(src), _C.ptr(self.out[t.out_idx]), B,
                            M_, Po, Po, st)
            elif isinstance(lyr, ConvLayer) and li in self.conv_small and train:
                # training forward: the tie pattern goes to the backward kernel; the un-pooled
                # activations are only written on request (keep_conv_out) -- nothing reads them
                pl, sm = self.conv_fused[li], self.conv_small[li]
                _C.call('tn_convpool_fprop_train', _C.ptr(x), _C.ptr(lyr.W.tensor),
                        _C.ptr(lyr.b.tensor), _C.ptr(out) if self.keep_conv_out else None,
                        _C.ptr(self.out[li + 1]), _C.ptr(sm.tie), B, lyr.num_prev_maps, lyr.in_sz,
                        lyr.num_maps, lyr.filter_sz, lyr.pad_lo, lyr.out_sz, lyr.act.code,
                        lyr.act.nn, pl.pool_sz, pl.out_sz, st)
            elif isinstance(lyr, ConvLayer) and li in self.conv_fused:
                pl = self.conv_fused[li]
                _C.call('tn_convpool_fprop', _C.ptr(x), _C.ptr(lyr.W.tensor),
                        _C.ptr(lyr.b.tensor), _C.ptr(out), _C.ptr(self.out[li + 1]), B,
                        lyr.num_prev_maps, lyr.in_sz, lyr.num_maps, lyr.filter_sz, lyr.pad_lo,
                        lyr.out_sz, lyr.act.code, lyr.act.nn, pl.pool_sz, pl.out_sz, st)
            elif isinstance(lyr, ConvLayer):
                dst = self.conv_full[li][0] if lyr.stride != 1 else out
                _C.call('tn_conv2d_fprop', _C.ptr(x), _C.ptr(lyr.W.tensor), _C.ptr(lyr.b.tensor),
                        _C.ptr(dst), B, lyr.num_prev_maps, lyr.in_sz, lyr.num_maps, lyr.filter_sz,
                        lyr.pad_lo, lyr.full_sz, lyr.act.code, lyr.act.nn, st)
                if lyr.stride != 1:                   # conv2d(subsample=(s, s)), convpool.py:54-56
                    _C.call('tn_subsample2d', _C.ptr(dst), _C.ptr(out), B * lyr.num_maps,
                            lyr.full_sz, lyr.stride, lyr.out_sz, st)
            elif isinstance(lyr, PoolLayer) and ((li - 1) in self.conv_fused or
                                                 ((li - 1) in self.conv_tc and
                                                  self.conv_tc[li - 1].pool is not None)):
                pass                                  # produced by the fused conv kernel below it
            elif isinstance(lyr, PoolLayer):
                _C.call('tn_maxpool_fwd', _C.ptr(x), _C.ptr(out), B * lyr.num_maps, lyr.in_sz,
                        lyr.pool_sz, lyr.out_sz, st)
            elif isinstance(lyr, MeanLayer):
                _C.call('tn_meanpool_fwd', _C.ptr(x), _C.ptr(out), B * lyr.num_maps, lyr.in_sz, st)
            elif isinstance(lyr, DropOutLayer):
                n = x[0].numel()
                if train and lyr.pdrop:
                    _C.call('tn_dropout_apply', _C.ptr(x), _C.ptr(out), B, n, 1. - lyr.pdrop,
                            lyr.seed, ctl, self._inj(li, 'mask'), 1.0, st)
                else:
                    _C.call('tn_dropout_apply', _C.ptr(x), _C.ptr(out), B, n, 1.0, 0, ctl, None,
                            float(lyr.test_scale), st)
            elif isinstance(lyr, AuxConcatLayer):     # auxiliary.py:80
                self._aux_forward(li, lyr.aux_info, train, idxp, st)
                _C.call('tn_concat_cols', _C.ptr(x), lyr.n_in, _C.ptr(self.aux_out),
                        lyr.aux_info.n_out, _C.ptr(out), B, st)
            elif isinstance(lyr, HiddenLayer):        # incl. the output layers (scores only)
                pkeep = 1. - lyr.pdrop if (train and lyr.pdrop) else 1.0
                _C.call('tn_dense_fwd', _C.ptr(x), _C.ptr(lyr.w.tensor), _C.ptr(lyr.b.tensor),
                        _C.ptr(out), B, lyr.n_in, lyr.n_out, lyr.act.code, lyr.act.nn, pkeep,
                        lyr.seed or 0, ctl, self._inj(li, 'mask') if train else None,
                        float(lyr.test_scale), st)
                if isinstance(lyr, SoftAuxLayer):     # + cross_b + aux . cross_w (auxiliary.py:133-134)
                    ai = lyr.aux_info
                    self._aux_forward(li, ai, train, idxp, st)
                    _C.call('tn_dense_fwd', _C.ptr(self.aux_out), _C.ptr(lyr.cross_w.tensor),
                            _C.ptr(lyr.cross_b.tensor), _C.ptr(self.aux_cross), B, ai.n_out,
                            lyr.n_out, _C.ACT_LINEAR, 0, 1.0, 0, ctl, None, 1.0, st)
                    _C.call('tn_add_inplace', _C.ptr(out), _C.ptr(self.aux_cross),
                            B * lyr.n_out, st)
            else:
                raise NotImplementedError(type(lyr).__name__)

    def _aux_forward(self, li, ai, train, idxp, st):
        """LocationInfo (auxiliary.py:14-57): gather this batch's (2,2) rows, mix them, two small
        dense layers -> self.aux_out."""
        B, ctl = self.local_bsz, _C.ptr(self.ctl)
        assert self._aux_corpus is not None, "Auxillary data not supplied"
        _C.call('tn_elastic_warp', _C.ptr(self._aux_corpus), idxp, ctl, B, 1, 2, 0, 0, None, None,
                0.0, None, 0, _C.ptr(self.aux_batch), st)              # row gather, 4 floats each
        _C.call('tn_aux_location_mix', _C.ptr(self.aux_batch), _C.ptr(self.aux_loc), B,
                float(ai.boost), int(train), ai.seed or 0, ctl,
                self._inj(li, 'auxu') if train else None, st)
        _C.call('tn_dense_fwd', _C.ptr(self.aux_loc), _C.ptr(ai.w1.tensor), _C.ptr(ai.b1.tensor),
                _C.ptr(self.aux_hid), B, 2, ai.n_hid, ai.act1.code, ai.act1.nn, 1.0, 0, ctl, None,
                1.0, st)
        _C.call('tn_dense_fwd', _C.ptr(self.aux_hid), _C.ptr(ai.w2.tensor), _C.ptr(ai.b2.tensor),
                _C.ptr(self.aux_out), B, ai.n_hid, ai.n_out, ai.act2.code, ai.act2.nn, 1.0, 0, ctl,
                None, 1.0, st)

    def _aux_backward(self, lyr, g, st):
        """SoftAuxLayer's extra parameters (auxiliary.py:108-141): the cross term and, through
        it, the two LocationInfo layers.  g = dL/d(scores)."""
        B, ctl, ai = self.local_bsz, _C.ptr(self.ctl), lyr.aux_info
        _C.call('tn_dense_bwd_weights', _C.ptr(self.aux_out), _C.ptr(g), _C.ptr(lyr.cross_w.grad),
                _C.ptr(lyr.cross_b.grad), B, ai.n_out, lyr.n_out, st)
        _C.call('tn_dense_bwd_data', _C.ptr(g), _C.ptr(lyr.cross_w.tensor), _C.ptr(self.aux_gout),
                B, ai.n_out, lyr.n_out, _C.ptr(self.aux_out), ai.act2.code, ai.act2.nn, 1.0, 0, ctl,
                None, st)
        _C.call('tn_dense_bwd_weights', _C.ptr(self.aux_hid), _C.ptr(self.aux_gout),
                _C.ptr(ai.w2.grad), _C.ptr(ai.b2.grad), B, ai.n_hid, ai.n_out, st)
        _C.call('tn_dense_bwd_data', _C.ptr(self.aux_gout), _C.ptr(ai.w2.tensor),
                _C.ptr(self.aux_ghid), B, ai.n_hid, ai.n_out, _C.ptr(self.aux_hid), ai.act1.code,
                ai.act1.nn, 1.0, 0, ctl, None, st)
        _C.call('tn_dense_bwd_weights', _C.ptr(self.aux_loc), _C.ptr(self.aux_ghid),
                _C.ptr(ai.w1.grad), _C.ptr(ai.b1.grad), B, 2, ai.n_hid, st)

    def _prefetching(self):
        return (self.field_prefetch and self.use_graph and not self.inject
                and not self.debug_elastic)

    def _launch_field(self, prm, g_i, g_f, st):
        li = self.el_index
        lyr = self.tr_layers[li]
        dbg = self.debug_elastic
        _C.call('tn_elastic_field', ctypes.byref(prm), self._inj(li, 'noise'), self._inj(li, 'u'),
                _C.ptr(self.el_filt), lyr.seed, _C.ptr(self.ctl),
                _C.ptr(self.el_target) if dbg else None, _C.ptr(self.el_tyx) if dbg else None,
                _C.ptr(g_i), _C.ptr(g_f), st)

    def _fuse_info(self, li):
        """How a consumer turns dL/d(out[li]) into dL/dz of layer li inside its own epilogue:
        (prev_out, act code, nn, pkeep, seed, injected mask) or None when layer li has no
        activation of its own."""
        lyr = self.tr_layers[li]
        if isinstance(lyr, OutputLayer):
            return None
        if isinstance(lyr, HiddenLayer):
            pk = 1. - lyr.pdrop if lyr.pdrop else 1.0
            return (self.out[li], lyr.act.code, lyr.act.nn, pk, lyr.seed or 0, self._inj(li, 'mask'))
        if isinstance(lyr, ConvLayer):
            if li in self.conv_tc:
                # the bf16 tensor-core branch of _backward applies act' itself
                # (tn_poolbwd_nhwc_bf16 gets the layer's activation): a consumer that fused it as
                # well would square the negative-side slope of a leaky reluNN
                return None
            return (self.out[li], lyr.act.code, lyr.act.nn, 1.0, 0, None)
        return None

    @contextmanager
    def _wgrad_stream(self):
        """Weight-gradient kernels do not feed the rest of the backward pass: they are forked
        onto side streams (graph branches once captured) and joined before the update, so that
        e.g. the dW GEMM overlaps the dX GEMM and the conv wgrad overlaps the conv dgrad."""
        if not self.overlap_wgrad:
            yield self._stream()
            return
        if self._side is None:
            self._side = [torch.cuda.Stream(self.device) for _ in range(2)]
            self._side_rr = 0
        side = self._side[self._side_rr]
        self._side_rr ^= 1
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            yield ctypes.c_void_p(side.cuda_stream)

    def _join_wgrad(self):
        main = torch.cuda.current_stream(self.device)
        if self.overlap_wgrad and self._side is not None:
            for side in self._side:
                main.wait_stream(side)
        if self._early_reduced:
            main.wait_stream(self._comm)

    def _bucket_split(self):
        """(index of the lowest layer of the trailing run of dense layers, its offset in the flat
        gradient buffer) when the data-parallel all-reduce can be split there, else None: the early
        bucket [offset, end + NLL slot) is complete once that layer's weight gradient is, long
        before the conv layers below have finished."""
        if not (self.dist.world > 1 and self.nccl_in_graph and self.head and
                (self.dp_hybrid or not self.dp_fused)):
            return None
        L = self.tr_layers
        li = len(L) - 1
        while li > 0 and (isinstance(L[li], DropOutLayer) or
                          (isinstance(L[li], HiddenLayer) and not isinstance(L[li], SoftAuxLayer))):
            li -= 1
        first = li + 1
        while first < len(L) and not L[first].params:
            first += 1
        if first >= len(L) or not any(self.trainable[j] for j in range(1, first)):
            return None
        return first, self.param_offset[id(L[first].params[0])]

    def _reduce_early_bucket(self, offset):
        """All-reduce grad[offset:] (dense-layer gradients + the NLL slot) on the communication
        stream, behind every weight-gradient kernel enqueued so far."""
        main = torch.cuda.current_stream(self.device)
        if self._comm is None:
            self._comm = torch.cuda.Stream(self.device)
        self._comm.wait_stream(main)
        if self.overlap_wgrad and self._side is not None:
            for side in self._side:
                self._comm.wait_stream(side)
        with torch.cuda.stream(self._comm):
            if self.dp_hybrid and self.peer_ar:
                # two-shot reduce-scatter + all-gather over peer memory (tn_peer_allreduce): at 8
                # ranks the NCCL ring takes 56 us for this 1.4 MB bucket and outlasts the conv
                # backward pass it is supposed to hide behind
                n = self.n_flat + 4 - offset
                _C.call('tn_peer_allreduce', self.peers.grad_ptrs[self._grad_parity],
                        self.peers.flag_ptrs, self.dist.world, self.dist.rank, offset, n,
                        ctypes.c_void_p(self._comm.cuda_stream))
            else:
                self.dist.all_reduce_sum(self.grad[offset:])
        self._early_reduced = offset

    def _backward(self):
        st = self._stream()
        B = self.local_bsz
        ctl = _C.ptr(self.ctl)
        L = self.tr_layers
        g = self.gsoft                       # dL/dz of the layer being visited
        g_bf16 = None                        # ... or, inside the bf16 conv stack, an NHWC bf16 tensor
        bucket = self._bucket_split()
        self._early_reduced = None
        for li in range(len(L) - 1, 0, -1):
            if bucket and li == bucket[0] - 1:       # every dense layer above has been enqueued
                self._reduce_early_bucket(bucket[1])
            lyr = L[li]
            x = self.out[li - 1]
            below = self.need_below[li]
            dx = self.dbuf[li - 1]
            fuse = self._fuse_info(li - 1) if below else None
            if self.head and li == len(L) - 1:
                # dL/dz of the layer below was already produced by tn_softmax_head_fwd_bwd
                with self._wgrad_stream() as sw:
                    _C.call('tn_softmax_head_bwd_weights', _C.ptr(x), _C.ptr(g),
                            _C.ptr(lyr.w.grad), _C.ptr(lyr.b.grad), _C.ptr(self.ws_head), B,
                            lyr.n_in, lyr.n_out, _C.ptr(self.rowloss), _C.ptr(self.nll_sum), sw)
            elif isinstance(lyr, AuxConcatLayer):
                if below:          # gradient of the concatenation wrt its first n_in columns
                    _C.call('tn_slice_cols', _C.ptr(g), lyr.n_out, 0, lyr.n_in, _C.ptr(dx), B, st)
                    if fuse:
                        po, ac, nn, pkp, sd, mi = fuse
                        if pkp < 1.0 or mi is not None:
                            _C.call('tn_dropout_apply', _C.ptr(dx), _C.ptr(dx), B, lyr.n_in, pkp, sd,
                                    ctl, mi, 1.0, st)
                        _C.call('tn_act_bwd', _C.ptr(dx), _C.ptr(po), _C.ptr(dx), B * lyr.n_in, ac,
                                nn, st)
            elif isinstance(lyr, HiddenLayer):
                if self.trainable[li] and isinstance(lyr, SoftAuxLayer):
                    self._aux_backward(lyr, g, st)
                # the two gradients are independent: when both are needed (and may overlap) they get
                # 1/3 and 2/3 of the SMs so that they run side by side instead of taking turns
                share = bool(self.trainable[li] and below and self.overlap_wgrad)
                if self.trainable[li]:
                    with self._wgrad_stream() as sw:
                        _C.call('tn_dense_bwd_weights_sm', _C.ptr(x), _C.ptr(g), _C.ptr(lyr.w.grad),
                                _C.ptr(lyr.b.grad), B, lyr.n_in, lyr.n_out,
                                self.sm_count // 3 if share else 0, sw)
                if below:
                    po, ac, nn, pk, sd, mi = fuse or (None, 0, 0, 1.0, 0, None)
                    _C.call('tn_dense_bwd_data_sm', _C.ptr(g), _C.ptr(lyr.w.tensor), _C.ptr(dx), B,
                            lyr.n_in, lyr.n_out, _C.ptr(po), ac, nn, pk, sd, ctl, mi,
                            self.sm_count - self.sm_count // 3 if share else 0, st)
            elif isinstance(lyr, ConvLayer) and li in self.conv_tc:
                t = self.conv_tc[li]
                S_, O_, Ci, M_, f_ = lyr.in_sz, lyr.out_sz, lyr.num_prev_maps, lyr.num_maps, lyr.filter_sz
                dtop, fmt = (g_bf16, 0) if g_bf16 is not None else (g, 1)
                _C.call('tn_poolbwd_nhwc_bf16', _C.ptr(t.a), _C.ptr(t.pooled), _C.ptr(dtop), fmt,
                        _C.ptr(t.gz), B, O_, M_, lyr.act.code, lyr.act.nn, st)
                if self.trainable[li]:
                    with self._wgrad_stream() as sw:
                        if t.im2col:
                            _C.call('tn_conv2d_tc_wgrad', _C.ptr(t.xin), _C.ptr(t.gz),
                                    _C.ptr(t.dWcol), _C.ptr(lyr.b.grad), _C.ptr(t.ws), B, 64, S_, M_,
                                    1, 0, O_, sw)
                            _C.call('tn_conv2d_tc_unpack_wgrad_im2col', _C.ptr(t.dWcol),
                                    _C.ptr(lyr.W.grad), M_, Ci, f_, sw)
                        else:
                            _C.call('tn_conv2d_tc_wgrad', _C.ptr(t.xin), _C.ptr(t.gz),
                                    _C.ptr(lyr.W.grad), _C.ptr(lyr.b.grad), _C.ptr(t.ws), B, Ci, S_,
                                    M_, f_, lyr.pad_lo, O_, sw)
                g_bf16 = None
                if below:
                    _C.call('tn_conv2d_tc_dgrad', _C.ptr(t.gz), _C.ptr(t.Wpd), _C.ptr(t.dx), B, Ci,
                            S_, M_, f_, lyr.pad_lo, O_, st)
                    if t.convert_in:
                        _C.call('tn_nhwc_bf16_to_nchw_f32', _C.ptr(t.dx), _C.ptr(dx), B, Ci, S_, S_, st)
                        if fuse:                      # a float32 conv directly below: its act'
                            if fuse[3] < 1.0:
                                raise NotImplementedError("dropout-masked output feeding a conv")
                            _C.call('tn_act_bwd', _C.ptr(dx), _C.ptr(fuse[0]), _C.ptr(dx),
                                    B * Ci * S_ * S_, fuse[1], fuse[2], st)
                    else:
                        g_bf16 = t.dx
            elif isinstance(lyr, ConvLayer) and li in self.conv_fused:
                # g is dL/d(pooled output); the pool backward is folded into both kernels
                pl = self.conv_fused[li]
                geom = (B, lyr.num_prev_maps, lyr.in_sz, lyr.num_maps, lyr.filter_sz, lyr.pad_lo,
                        lyr.out_sz, lyr.act.code, lyr.act.nn, pl.pool_sz, pl.out_sz)
                if fuse and fuse[3] < 1.0:
                    raise NotImplementedError("dropout-masked dense output feeding a conv")
                if (li in self.conv_small and
                        (not fuse or fuse[1] in (_C.ACT_LINEAR, _C.ACT_RELU, _C.ACT_LEAKY))):
                    # dW, db and dx from one staging of dL/dz (tn_convpool_bwd); the tie pattern
                    # comes from this step's forward kernel
                    po, ac, nn = (fuse[0], fuse[1], fuse[2]) if fuse else (None, 0, 0)
                    sm = self.conv_small[li]
                    _C.call('tn_convpool_bwd', _C.ptr(x), None, _C.ptr(sm.tie),
                            _C.ptr(self.out[li + 1]), _C.ptr(g), _C.ptr(lyr.W.tensor),
                            _C.ptr(lyr.W.grad), _C.ptr(lyr.b.grad),
                            _C.ptr(dx) if below else None, _C.ptr(po), _C.ptr(sm.ws), *geom, ac, nn,
                            st)
                    if not below:
                        break
                    g = dx
                    continue
                if self.trainable[li]:
                    with self._wgrad_stream() as sw:
                        _C.call('tn_convpool_bwd_weights', _C.ptr(x), _C.ptr(self.out[li]),
                                _C.ptr(self.out[li + 1]), _C.ptr(g), _C.ptr(lyr.W.grad),
                                _C.ptr(lyr.b.grad), _C.ptr(self.ws[li]), *geom, sw)
                if below:
                    if fuse and fuse[3] < 1.0:
                        raise NotImplementedError("dropout-masked dense output feeding a conv")
                    po, ac, nn = (fuse[0], fuse[1], fuse[2]) if fuse else (None, 0, 0)
                    _C.call('tn_convpool_bwd_data', _C.ptr(self.out[li]), _C.ptr(self.out[li + 1]),
                            _C.ptr(g), _C.ptr(lyr.W.tensor), _C.ptr(dx), _C.ptr(po), *geom, ac, nn,
                            st)
            elif isinstance(lyr, ConvLayer):
                if lyr.stride != 1:      # dL/dz back on the sampling lattice, zeros elsewhere
                    gfull = self.conv_full[li][1]
                    _C.call('tn_upsample2d_zero', _C.ptr(g), _C.ptr(gfull), B * lyr.num_maps,
                            lyr.full_sz, lyr.stride, lyr.out_sz, st)
                    g = gfull
                if self.trainable[li]:
                    with self._wgrad_stream() as sw:
                        _C.call('tn_conv2d_wgrad', _C.ptr(x), _C.ptr(g), _C.ptr(lyr.W.grad),
                                _C.ptr(lyr.b.grad), _C.ptr(self.ws[li]), B, lyr.num_prev_maps,
                                lyr.in_sz, lyr.num_maps, lyr.filter_sz, lyr.pad_lo, lyr.full_sz, sw)
                if below:
                    po, ac, nn = (fuse[0], fuse[1], fuse[2]) if fuse else (None, 0, 0)
                    _C.call('tn_conv2d_dgrad', _C.ptr(g), _C.ptr(lyr.W.tensor), _C.ptr(dx),
                            _C.ptr(po), B, lyr.num_prev_maps, lyr.in_sz, lyr.num_maps,
                            lyr.filter_sz, lyr.pad_lo, lyr.full_sz, ac, nn, st)
                    if fuse and fuse[3] < 1.0:
                        raise NotImplementedError("dropout-masked dense output feeding a conv")
            elif isinstance(lyr, PoolLayer) and ((li - 1) in self.conv_fused or
                                                 ((li - 1) in self.conv_tc and
                                                  self.conv_tc[li - 1].pool is not None)):
                continue                              # g stays dL/d(pooled): see the conv branch
            elif isinstance(lyr, PoolLayer):
                if below:
                    ac, nn = (fuse[1], fuse[2]) if fuse else (_C.ACT_LINEAR, 0)
                    _C.call('tn_maxpool_bwd', _C.ptr(g), _C.ptr(x), _C.ptr(self.out[li]),
                            _C.ptr(dx), B * lyr.num_maps, lyr.in_sz, lyr.pool_sz, lyr.out_sz, ac,
                            nn, st)
                    if fuse and fuse[3] < 1.0:
                        raise NotImplementedError("dropout-masked dense output feeding a pool")
            elif isinstance(lyr, MeanLayer):
                if below:
                    po, ac, nn = (fuse[0], fuse[1], fuse[2]) if fuse else (None, _C.ACT_LINEAR, 0)
                    _C.call('tn_meanpool_bwd', _C.ptr(g), _C.ptr(po), _C.ptr(dx), B * lyr.num_maps,
                            lyr.in_sz, ac, nn, st)
                    if fuse and fuse[3] < 1.0:
                        raise NotImplementedError("dropout-masked dense output feeding a MeanLayer")
            elif isinstance(lyr, (ColorLayer, ElasticLayer)):
                if below:
                    raise NotImplementedError("{} above a trainable layer (no gradient through "
                                              "it is implemented)".format(type(lyr).__name__))
            elif isinstance(lyr, DropOutLayer):
                if below:
                    n = x[0].numel()
                    pk = 1. - lyr.pdrop if lyr.pdrop else 1.0
                    _C.call('tn_dropout_apply', _C.ptr(g), _C.ptr(dx), B, n, pk, lyr.seed or 0, ctl,
                            self._inj(li, 'mask'), 1.0, st)
                    if fuse:
                        po, ac, nn, pkp, sd, mi = fuse
                        if pkp < 1.0 or mi is not None:
                            _C.call('tn_dropout_apply', _C.ptr(dx), _C.ptr(dx), B, n, pkp, sd, ctl,
                                    mi, 1.0, st)
                        _C.call('tn_act_bwd', _C.ptr(dx), _C.ptr(po), _C.ptr(dx), B * n, ac, nn, st)
            else:
                raise NotImplementedError(type(lyr).__name__)
            if not below:
                break
            g = dx

    def _set_ctl(self, row0):
        """Per-step scalars -> the device control block, as the launch arguments of tn_set_ctl
        (eager, stream-ordered ahead of the step's graph).  Kept lean: it sits on the host's critical
        path between two steps of the synchronous API -- and usually finds that `_preissue_ctl`
        already enqueued exactly these values behind the previous step."""
        lr = self.cur_learn_rate.get_value()
        if lr != self._lr_cached[0]:
            self._lr_cached = (lr, int(np.float32(lr).view(np.int32)))
        step, s0 = self.step_count & 0x7fffffff, self.dist.rank * self.local_bsz
        want = (step, s0, int(row0), self._lr_cached[1])
        c = self._ctl_np                      # host mirror (debugging, CPU-side tests)
        c[_C.CTL_STEP], c[_C.CTL_SAMPLE0], c[_C.CTL_ROW0], c[_C.CTL_LR_BITS] = want
        if self.device.type == 'cuda':
            st = self._stream()
            if self._ctl_issued == (want, st.value):
                return                        # already in the stream, behind the previous step
            _C.call('tn_set_ctl', self._ctl_ptr, want[0], want[1], want[2], want[3], st)
            self._ctl_issued = (want, st.value)

    def _preissue_ctl(self, row0_next):
        """Right after a training step has been enqueued (step_count already advanced): enqueue the
        NEXT step's scalars behind it.  The tiny kernel then runs as soon as this step's graph has
        finished, not when the host comes back with the next call -- 8 us less between two steps of
        the synchronous API.  A call that needs different values (another batch, a new learning
        rate, a test pass in between) simply issues its own."""
        if self.device.type == 'cuda' and not self.inject:
            self._set_ctl(row0_next)

    def _upload_idx(self, ids):
        """Index vector of this step -> self.idx, through the next pinned ring slot (eager copy,
        stream-ordered ahead of the step's graph; a slot is reused only after its copy ran)."""
        k = self._idx_k
        self._idx_k = (k + 1) % len(self._idx_ring)
        if self._idx_ev[k] is not None:
            self._idx_ev[k].synchronize()
        self._idx_ring[k].copy_(torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int32)))
        self.idx.copy_(self._idx_ring[k], non_blocking=True)
        if self.device.type == 'cuda':
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._idx_ev[k] = ev

    def _train_launches(self, corpus, idx, labels):
        """Everything one training step enqueues (this is what the CUDA graph captures)."""
        st = self._stream()
        B = self.local_bsz
        self._forward(self.tr_layers, True, corpus, idx, labels)
        if self._prefetching():       # next step's sampling grid, off the critical path
            g_i, g_f = self.el_grids[(self.step_count + 1) & 1]
            with self._wgrad_stream() as sw:
                self._launch_field(self.el_prm_next, g_i, g_f, sw)
        last = self.tr_layers[-1]
        n_out = last.n_out
        if self.head:
            li = len(self.tr_layers) - 1
            below = self.need_below[li]
            fuse = self._fuse_info(li - 1) if below else None
            po, ac, nn, pk, sd, mi = fuse or (None, 0, 0, 1.0, 0, None)
            _C.call('tn_softmax_head_fwd_bwd', _C.ptr(self.out[li - 1]), _C.ptr(last.w.tensor),
                    _C.ptr(last.b.tensor), _C.ptr(labels), _C.ptr(idx), _C.ptr(self.ctl), B,
                    last.n_in, n_out, 1.0 / self.batch_sz, _C.ptr(self.logprob),
                    _C.ptr(self.gsoft), _C.ptr(self.rowloss),
                    _C.ptr(self.dbuf[li - 1]) if below else None, int(fuse is not None), ac, nn,
                    pk, sd, mi, st)
        elif self.plain_nll:
            _C.call('tn_softmax_nll_fwd_bwd', _C.ptr(self.z), _C.ptr(labels), _C.ptr(idx),
                    _C.ptr(self.ctl), B, n_out, 1.0 / self.batch_sz, _C.ptr(self.logprob),
                    _C.ptr(self.gsoft), _C.ptr(self.rowloss), st)
        else:                                        # nllsq / nllNN / ExpLossLayer / HingeLayer
            _C.call('tn_output_loss_fwd_bwd', _C.ptr(self.z), _C.ptr(labels), _C.ptr(idx),
                    _C.ptr(self.ctl), B, n_out, self.out_kind, self.loss_code, self.log_thr,
                    1.0 / self.batch_sz, _C.ptr(self.feat),
                    _C.ptr(self.logprob) if self.feat is not self.logprob else None,
                    _C.ptr(self.gsoft), _C.ptr(self.rowloss), st)
        self._backward()
        self._join_wgrad()
        if not self.head:                      # the fused head already reduced the row losses
            _C.call('tn_reduce_rowloss', _C.ptr(self.rowloss), B, _C.ptr(self.nll_sum), st)
        if self.dp_fused:                            # all-reduce inside the optimiser kernel
            _C.call('tn_allreduce_sgd_update', _C.ptr(self.theta), _C.ptr(self.vel),
                    self.peers.grad_ptrs[self._grad_parity], self.peers.flag_ptrs, self.dist.world,
                    self.dist.rank, self.segs, self.n_segs, self.n_flat,
                    int(self._early_reduced or 0), _C.ptr(self.ctl), 1.0,
                    1.0 / self.batch_sz, _C.ptr(self.cost), _C.ptr(self.ws_update), st)
            return
        if self.dist.world > 1 and not self.nccl_in_graph:
            return                                   # the caller reduces, then _update_launches
        if self._early_reduced:                      # the small rest: conv gradients
            self.dist.all_reduce_sum(self.grad[:self._early_reduced])
        else:
            self.dist.all_reduce_sum(self.grad)      # the one collective of the step
        self._update_launches()

    def _update_launches(self):
        _C.call('tn_sgd_momentum_maxnorm_update', _C.ptr(self.theta), _C.ptr(self.vel),
                _C.ptr(self.grad), self.segs, self.n_segs, self.n_flat, _C.ptr(self.ctl), 1.0,
                _C.ptr(self.nll_sum), 1.0 / self.batch_sz, _C.ptr(self.cost),
                _C.ptr(self.ws_update), self._stream())

    def _train_step(self, key, corpus, idx, labels):
        """One training step.  Single GPU: one CUDA graph.  Data parallel: graph (forward +
        backward) -> NCCL all-reduce of the flat gradient buffer -> graph (update); set
        TN_GRAPH_NCCL=1 to capture the collective inside a single graph instead."""
        prefetch = self._prefetching()
        if prefetch:
            # the graph of step s reads the sampling grid from el_grids[s & 1] and, on a side
            # branch, fills el_grids[(s+1) & 1] for the next step: one graph per parity
            if self._field_step != self.step_count:      # first step, or the sequence was broken
                g_i, g_f = self.el_grids[self.step_count & 1]
                self._launch_field(self.el_prm, g_i, g_f, self._stream())
            key = key + ('f%d' % (self.step_count & 1),)
            self._field_step = self.step_count + 1
        if self.dp_fused:
            parity = self.step_count & 1              # graphs bake pointers: one per parity
            self._bind_grads(parity)
            self._run(key + ('p%d' % parity,), self._train_launches, (corpus, idx, labels),
                      restore=(self.theta, self.vel))
        elif self.dist.world > 1 and not self.nccl_in_graph:
            self._run(('train_pre',) + key[1:], self._train_launches, (corpus, idx, labels))
            self.dist.all_reduce_sum(self.grad)
            self._run(('update',), self._update_launches, (), restore=(self.theta, self.vel))
            self.launches['train'] = self.launches.get('train_pre', 0) + self.launches.get('update', 0)
        else:
            self._run(key, self._train_launches, (corpus, idx, labels),
                      restore=(self.theta, self.vel))

    def _test_launches(self, corpus, idx, labels):
        st = self._stream()
        B = self.local_bsz
        self._forward(self.te_layers, False, corpus, idx, labels)
        if self.out_kind == _C.OUT_SOFTMAX:
            _C.call('tn_softmax_test_stats', _C.ptr(self.z), _C.ptr(labels), _C.ptr(idx),
                    _C.ptr(self.ctl), B, self.tr_layers[-1].n_out, _C.ptr(self.logprob),
                    _C.ptr(self.preds), _C.ptr(self.stats), st)
        else:
            _C.call('tn_output_test_stats', _C.ptr(self.z), _C.ptr(labels), _C.ptr(idx),
                    _C.ptr(self.ctl), B, self.tr_layers[-1].n_out, self.out_kind,
                    _C.ptr(self.feat),
                    _C.ptr(self.logprob) if self.feat is not self.logprob else None,
                    _C.ptr(self.preds), _C.ptr(self.stats), st)
        if self.dist.world > 1:
            self.dist.all_reduce_sum(self.stats[:2])

    def _run(self, key, fn, args, restore=()):
        """Run ``fn(*args)`` eagerly, or capture it once into a CUDA graph and replay it.
        ``restore`` lists the tensors whose contents the warm-up execution must not change."""
        if not self.use_graph or self.inject:
            n0 = _C.lib.tn_launch_count()
            fn(*args)
            self.launches[key[0]] = _C.lib.tn_launch_count() - n0
            return
        g = self._graphs.get(key)
        if g is None:
            snap = [t.clone() for t in restore]
            n0 = _C.lib.tn_launch_count()
            fn(*args)                                # warm-up (module loading, NCCL setup)
            self.launches[key[0]] = _C.lib.tn_launch_count() - n0   # kernels per replay
            torch.cuda.synchronize(self.device)
            if self.dp_fused and self.dist.world > 1:
                # the replay below recomputes the SAME parity of the peer-visible gradient buffers
                # the warm-up just published: every rank must have finished reading them first
                self.dist.barrier()
            for t, s in zip(restore, snap):
                t.copy_(s)
            g = torch.cuda.CUDAGraph()
            # thread_local: other threads (the NCCL watchdog polling its events) may keep issuing
            # CUDA calls while this thread captures; the default 'global' mode fails the capture
            with torch.cuda.graph(g, capture_error_mode='thread_local'):
                fn(*args)
            self._graphs[key] = g
        g.replay()

    # ------------------------------------------------------------------------------------------
    # data staging
    # ------------------------------------------------------------------------------------------
    def _stage(self, x_data, y_data, resident):
        x = _as_numpy(x_data)
        y = _as_numpy(y_data)
        l0 = self.tr_layers[0]
        shp = (-1, l0.num_maps, l0.out_sz, l0.out_sz)
        if isinstance(x, torch.Tensor):
            x = x.to(torch.float32).reshape(shp).contiguous()
        else:
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32).reshape(shp))
        if isinstance(y, torch.Tensor):
            y = y.to(torch.int32).contiguous()
        else:
            y = torch.from_numpy(np.ascontiguousarray(y, dtype=np.int32))
        if resident:
            return x.to(self.device), y.to(self.device), None
        if self.device.type == 'cuda':
            if not x.is_cuda and not x.is_pinned():
                x = x.pin_memory()
            if not y.is_cuda and not y.is_pinned():
                y = y.pin_memory()
        B = self.local_bsz
        xb = torch.empty((B,) + tuple(x.shape[1:]), dtype=torch.float32, device=self.device)
        yb = torch.empty(B, dtype=torch.int32, device=self.device)
        return xb, yb, (x, y)

    def _stage_aux(self, aux_data, resident):
        """The auxiliary corpus (N, 2, 2) -> device (N, 4); None when the net takes none."""
        if self.aux_index is None:
            return None
        assert aux_data is not None, "Auxillary data not supplied"       # neuralnet.py:216,267
        assert resident, "auxiliary inputs need the HBM-resident corpus (resident=True)"
        a = _as_numpy(aux_data)
        if isinstance(a, torch.Tensor):
            a = a.to(torch.float32).reshape(-1, 4).contiguous()
        else:
            a = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4))
        return a.to(self.device)

    # ------------------------------------------------------------------------------------------
    # public API (theanet/neuralnet.py:203-296)
    # ------------------------------------------------------------------------------------------
    def get_trin_model(self, x_data, y_data, aux_data=None, take_index_list=False, resident=True,
                       lazy=False):
        """Returns ``f(batch_index) -> [cost, features, logprob]`` (or ``f(index_vector)`` with
        take_index_list), the analogue of the reference's compiled training function.

        resident=True keeps the corpus in HBM (the reference keeps it in a Theano shared
        variable); resident=False leaves it in pinned host memory and copies each minibatch
        host->device inside the call.  lazy=True returns device tensors without synchronising.
        """
        xd, yd, host = self._stage(x_data, y_data, resident)
        auxd = self._stage_aux(aux_data, resident)
        B, Bl, rank = self.batch_sz, self.local_bsz, self.dist.rank
        # captured graphs bake the raw device pointers of THIS closure's staged corpus: key them by a
        # serial number (id(xd) can be recycled once a closure is collected) and drop them with it
        key = ('train', self._new_model_key(), take_index_list)
        n_rows = int(host[0].shape[0] if host is not None else xd.shape[0])
        pin = self.device.type == 'cuda'
        n_lp = self.logprob.numel()
        h_res = torch.zeros(self._res.shape, dtype=torch.float32, pin_memory=pin)
        h_res_np = h_res.numpy()
        lp_shape = tuple(self.logprob.shape)
        two = self.feat is not self.logprob          # ExpLossLayer: features != logprob
        h_ft = torch.zeros(lp_shape, dtype=torch.float32, pin_memory=pin) if two else None

        def results():
            """[cost, features, logprob] on the host (neuralnet.py:236-241): one device-to-host
            copy of [logprob | cost], one synchronisation."""
            h_res.copy_(self._res, non_blocking=True)
            if two:
                h_ft.copy_(self.feat, non_blocking=True)
            if pin:
                torch.cuda.current_stream(self.device).synchronize()
            lp = h_res_np[:n_lp].reshape(lp_shape).copy()
            return [h_res_np[n_lp].copy(), h_ft.numpy().copy() if two else lp, lp]

        # host-resident corpus: double-buffered minibatch staging.  While step i runs, the H2D copy
        # of batch i+1 (the reference driver walks the batches in order, train.py:210) proceeds on a
        # copy stream into the other device buffer; a call with the predicted index finds its data
        # already on the device.  Any other index is simply copied on demand.
        pre = {'index': None, 'slot': 0}
        if host is not None and not take_index_list and self.device.type == 'cuda':
            bufs = [(xd, yd), (torch.empty_like(xd), torch.empty_like(yd))]
            copy_stream = torch.cuda.Stream(self.device)
            ready = [torch.cuda.Event(), torch.cuda.Event()]
            done = [None, None]                   # last step that read each buffer
            n_batches = max(1, host[0].shape[0] // B)

            def fetch(i, slot):
                lo_ = int(i) * B + rank * Bl
                if done[slot] is not None:
                    copy_stream.wait_event(done[slot])
                with torch.cuda.stream(copy_stream):
                    bufs[slot][0].copy_(host[0][lo_:lo_ + Bl], non_blocking=True)
                    bufs[slot][1].copy_(host[1][lo_:lo_ + Bl], non_blocking=True)
                    ready[slot].record(copy_stream)
        else:
            bufs = None

        def check_batch(i):
            # the reference slices a shared variable and Theano raises on a short batch
            if not (0 <= int(i) and (int(i) + 1) * B <= n_rows):
                raise IndexError("minibatch {} of {} images is outside the corpus of {} rows".format(
                    int(i), B, n_rows))

        def training_fn(indx):
            self._aux_corpus = auxd
            if not take_index_list:
                check_batch(indx)
            if bufs is not None:
                if pre['index'] == int(indx):
                    slot = pre['slot']
                else:
                    slot = pre['slot'] ^ 1 if pre['index'] is not None else 0
                    fetch(indx, slot)
                torch.cuda.current_stream(self.device).wait_event(ready[slot])
                self._set_ctl(0)
                self._train_step(key + (slot,), bufs[slot][0], None, bufs[slot][1])
                if done[slot] is None:
                    done[slot] = torch.cuda.Event()
                done[slot].record(torch.cuda.current_stream(self.device))
                self.step_count += 1
                self._preissue_ctl(0)
                nxt = (int(indx) + 1) % n_batches
                fetch(nxt, slot ^ 1)              # overlaps the step that was just enqueued
                pre['index'], pre['slot'] = nxt, slot ^ 1
                if lazy:
                    return self.cost, self.feat, self.logprob
                return results()
            if take_index_list:
                ids = np.asarray(indx, dtype=np.int32)[rank * Bl:(rank + 1) * Bl]
                if ids.size != Bl or ids.min() < 0 or ids.max() >= n_rows:
                    raise IndexError("index list must hold BATCH_SZ row numbers in [0, {})".format(n_rows))
                if host is None:
                    self._upload_idx(ids)
                    idx, row0 = self.idx, 0
                else:
                    sel = torch.from_numpy(ids.astype(np.int64))
                    xd.copy_(host[0][sel], non_blocking=True)
                    yd.copy_(host[1][sel], non_blocking=True)
                    idx, row0 = None, 0
            else:
                lo = int(indx) * B + rank * Bl
                if host is None:
                    idx, row0 = None, lo
                else:
                    xd.copy_(host[0][lo:lo + Bl], non_blocking=True)
                    yd.copy_(host[1][lo:lo + Bl], non_blocking=True)
                    idx, row0 = None, 0
            self._set_ctl(row0)
            self._train_step(key, xd, idx, yd)
            self.step_count += 1
            if host is None and not take_index_list and not lazy:
                # the reference's driver walks the batches in order (train.py:210)
                self._preissue_ctl(((int(indx) + 1) % max(1, n_rows // B)) * B + rank * Bl)
            if lazy:
                return self.cost, self.feat, self.logprob
            return results()

        weakref.finalize(training_fn, self._drop_graphs, key[1])
        return training_fn

    def _new_model_key(self):
        self._model_serial = getattr(self, '_model_serial', 0) + 1
        return self._model_serial

    def _drop_graphs(self, serial):
        for k in [k for k in self._graphs if len(k) > 1 and k[1] == serial]:
            del self._graphs[k]

    def reset_accumulated_gradients(self):
        self.vel.zero_()

    def get_test_model(self, x_data, y_data, aux_data=None, preds_feats=False, resident=True):
        """``f(batch_index) -> (sym_err_rate, mean p[y]) [+ (features, y_preds)]``."""
        xd, yd, host = self._stage(x_data, y_data, resident)
        auxd = self._stage_aux(aux_data, resident)
        B, Bl, rank = self.batch_sz, self.local_bsz, self.dist.rank
        key = ('test', self._new_model_key())
        n_rows = int(host[0].shape[0] if host is not None else xd.shape[0])

        def test_fn(indx):
            self._aux_corpus = auxd
            if not (0 <= int(indx) and (int(indx) + 1) * B <= n_rows):
                raise IndexError("minibatch {} of {} images is outside the corpus of {} rows".format(
                    int(indx), B, n_rows))
            lo = int(indx) * B + rank * Bl
            if host is None:
                row0 = lo
            else:
                xd.copy_(host[0][lo:lo + Bl], non_blocking=True)
                yd.copy_(host[1][lo:lo + Bl], non_blocking=True)
                row0 = 0
            self._set_ctl(row0)
            self._run(key, self._test_launches, (xd, None, yd))
            st = self.stats[:2].cpu().numpy() / self.dist.world
            outs = [np.float32(st[0]), np.float32(st[1])]
            if preds_feats:                       # features_and_predictions, outlayers.py:66-67
                outs += [self.feat.cpu().numpy(), self.preds.cpu().numpy()]
            return outs

        weakref.finalize(test_fn, self._drop_graphs, key[1])
        return test_fn

    def takes_aux(self):
        return self.aux_index is not None

    def get_data_test_model(self, get_output_of_layers=()):
        """``f(x_batch) -> [features, y_preds, *layer outputs]`` on raw input batches of exactly
        the (local) batch size (neuralnet.py:282-296)."""
        Bl = self.local_bsz
        yd = torch.zeros(Bl, dtype=torch.int32, device=self.device)

        def data_test_fn(x, aux=None):
            xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(self.device)
            assert xd.shape[0] == Bl, "expected a batch of {} images".format(Bl)
            self._aux_corpus = self._stage_aux(aux, True)      # inputs += [aux] (neuralnet.py:288-290)
            self._set_ctl(0)
            self._test_launches(xd.reshape((Bl,) + self.tr_layers[0].output.shape), None, yd)
            outs = [self.feat.cpu().numpy(), self.preds.cpu().numpy()]
            outs += [self.out[i].cpu().numpy() for i in get_output_of_layers]
            return outs

        return data_test_fn

    def get_init_params(self):
        return {"layers": self.layers,
                "training_params": self.tr_prms,
                "allwts": [l.get_wts() for l in self.tr_layers]}

    # -- exact resume (the reference's .pkl drops momentum and RNG state) ----------
    def get_resume_state(self):
        """What ``get_init_params()`` does not carry: the momentum buffers (layer.py:78-80 keeps
        them only inside the compiled graph), the step counter and the per-layer stream seeds that
        key the Philox draws.  ``set_resume_state`` on a net built from the same .pkl continues the
        run bit for bit; a .pkl without it resumes like the reference does (fresh momentum, new
        random streams)."""
        aux = [getattr(getattr(l, 'aux_info', None), 'seed', None) for l in self.tr_layers]
        return {"velocities": self.get_velocities(), "step_count": int(self.step_count),
                "seeds": [getattr(l, 'seed', None) for l in self.tr_layers], "aux_seeds": aux}

    def set_resume_state(self, state):
        assert len(state["velocities"]) == len(self.tr_layers) == len(state["seeds"])
        for lyr, vels, seed in zip(self.tr_layers, state["velocities"], state["seeds"]):
            assert len(vels) == len(lyr.params)
            for p, v in zip(lyr.params, vels):
                p.vel.copy_(torch.as_tensor(np.asarray(v, np.float32)).reshape(p.vel.shape))
            if seed is not None:
                lyr.seed = seed
        for lyr, seed in zip(self.tr_layers, state.get("aux_seeds", ())):
            if seed is not None:
                lyr.aux_info.seed = seed
        self.step_count = int(state["step_count"])
        self._field_step = None
        self._graphs = {}            # captured graphs bake the seeds in as kernel arguments

    def set_rate(self):
        self.cur_learn_rate.set_value(
            self.tr_prms['INIT_LEARNING_RATE'] /
            (1 + self.tr_prms['CUR_EPOCH'] / self.tr_prms['EPOCHS_TO_HALF_RATE']))

    def inc_epoch_set_rate(self):
        self.tr_prms['CUR_EPOCH'] += 1
        self.set_rate()

    def get_epoch(self):
        return self.tr_prms['CUR_EPOCH']

    def __str__(self):
        prmstr = '; '.join(', '.join(str(p) for p in lyr.params) for lyr in self.tr_layers)
        return ('\nTrain Layers\n\t' + '\n\t'.join(str(l) for l in self.tr_layers) +
                '\nTest Layers\n\t' + '\n\t'.join(str(l) for l in self.te_layers) +
                '\nParams ' + prmstr)

    def get_layers_info(self):
        return get_layers_info(self.layers)

    def get_wts_info(self, detailed=False):
        return get_wts_info((l.get_wts() for l in self.tr_layers), detailed)

    def get_training_params_info(self):
        return get_training_params_info(self.tr_prms)

    # ------------------------------------------------------------------------------------------
    # extras used by tests / debugging
    # ------------------------------------------------------------------------------------------
    def elastic_debugout(self):
        """[displacement (2,h,h) float64, clipped coordinates (2,h,h)] of the last training step
        (needs ``debug_elastic = True``), cf. ElasticLayer.debugout (inlayers.py:145-155)."""
        h = self.tr_layers[self.el_index].img_sz
        tgt = self.el_target.cpu().numpy().reshape(2, h, h)
        return [tgt - np.indices((h, h)), self.el_tyx.cpu().numpy().reshape(2, h, h)]

    def get_velocities(self):
        return [[p.vel.detach().cpu().numpy().copy() for p in l.params] for l in self.tr_layers]

    def get_gradients(self):
        return [[p.grad.detach().cpu().numpy().copy() for p in l.params] for l in self.tr_layers]
