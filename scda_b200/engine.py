"""One SCDA training iteration on one rank — the body of `train()` in the reference driver
(tools/faster_rcnn_train_val.py:507-770), kept as a function so the driver, bench.py and
the tests run the same code.

Four networks, four Adam optimisers stepping in sequence, as the reference does:
  detector forward on the source AND target image            (:526)
  crops around the cluster centres, decoder                   (:528-560)
  (1) image discriminator update                              (:567-616)
  (2) patch (feature) discriminator update                    (:623-635)
  (3) decoder update                                          (:642-704)
  (4) detector update                                         (:716-750)

Mechanism differences (results are the same):
  * each network's parameters, gradients and Adam moments live in ONE flat fp32 buffer
    (FlatAdam): one NCCL all-reduce and one fused optimiser launch per network instead of
    one collective and ~10 elementwise kernels per parameter tensor;
  * every backward names the parameters it is for (`inputs=`), so autograd does not
    compute the stray gradients the reference accumulates into the other networks and
    then discards with the next zero_grad();
  * the "Max Grad" logging loops (:607-611, 695-699, 741-745: one host sync per parameter
    tensor) are off unless asked for;
  * the phases that do not depend on each other run side by side on separate CUDA streams
    (`overlap`): the target-image branch of the detector forward beside the source branch, the
    anchor targets beside the backbone, the detector's backward + Adam (phase 4) beside phases
    1-3 — the cluster features are detached (functions/mask.py:234), so phase 4's gradient never
    depended on them — and phase 2 beside phase 1.  The whole iteration, all streams included,
    is replayed from one CUDA graph (world 1) or from eight graphs cut at the four gradient
    all-reduces (world > 1).  `overlap=False` gives the reference's single-stream order; the
    tests compare the two.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

from . import gan_ops
from ._lib import check, load, stream_ptr
from .gan_ops import run_pair
from .loss_ops import bce_sigmoid_rows
from .utils.distributed_utils import FlatGradBucket


class FlatAdam(object):
    """Adam(lr, betas, eps, weight_decay) of torch 0.4.1 over one network, flat storage,
    stepped by libscda_b200's scda_adam_step.

    Parameters, gradients and both moments live in flat buffers with one 64-element-aligned
    segment per parameter.  channels_last=True stores 4-D weights [O][kh][kw][I] (the
    parameter keeps its [O,I,kh,kw] shape with channels_last strides); tensor_core=True adds
    a bf16 shadow of the parameters in the same layout, refreshed by the optimiser kernel
    itself: the shadow segment of a conv weight IS the KRSC operand of the tcgen05 kernels.

    The bias-corrected step size lives in a device scalar (`lr_t_dev`) that `begin_step`
    refreshes, so `step_dev` is a pure kernel launch and can sit inside a CUDA graph."""

    def __init__(self, module, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, tensor_core=False,
                 channels_last=None):
        from .utils.distributed_utils import flat_layout, flat_view
        channels_last = tensor_core if channels_last is None else channels_last
        self.params = [p for p in module.parameters() if p.requires_grad]
        assert self.params and all(p.is_cuda and p.dtype == torch.float32 for p in self.params)
        dev = self.params[0].device
        self.offsets, n = flat_layout(self.params, 64)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, self.offsets):
            view = flat_view(self.flat, off, p) if channels_last else self.flat[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        self.n = n
        self.bucket = FlatGradBucket(self.params, self.offsets, n, channels_last=channels_last)
        module._scda_bucket = self.bucket
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.shadow = None
        if tensor_core:
            self.shadow = self.flat.to(torch.bfloat16)
            for p, off in zip(self.params, self.offsets):
                seg = self.shadow[off:off + p.numel()]
                if p.dim() == 4:
                    o, i, kh, kw = p.shape
                    p._scda_shadow = seg.view(o, kh, kw, i)
                else:
                    p._scda_shadow = seg.view(p.shape)
                p._scda_shadow_version = p._version
                p._scda_direct_grad = True
                p.register_hook(self._stale_guard(p))
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.t = 0
        self.lr_t_dev = torch.zeros(1, dtype=torch.float32, device=dev)

    @staticmethod
    def _stale_guard(p):
        """FlatGradBucket.zero() leaves the big direct-written gradients un-zeroed and only marks them
        fresh (the tensor-core sinks overwrite them).  If a gradient reaches such a parameter through
        ordinary autograd instead (`_fp32_graph`, a sink's fall-through), it would be ADDED onto last
        step's values: this hook runs before that accumulation and zeroes the stale contents first."""
        def hook(grad):
            if getattr(p, "_scda_grad_fresh", False):
                if p.grad is not None:
                    p.grad.zero_()
                p._scda_grad_fresh = False
            return grad
        return hook

    def zero_grad(self):
        self.bucket.zero()

    def all_reduce(self, lo=0, hi=None, group=None):
        self.bucket.rebind()
        self.bucket.settle(lo, hi)
        self.bucket.all_reduce(lo=lo, hi=hi, group=group)

    def offset_of(self, param):
        """start of `param`'s segment in the flat buffers (bucket boundaries)"""
        for p, off in zip(self.params, self.offsets):
            if p is param:
                return off
        raise KeyError("parameter is not owned by this optimiser")

    def begin_step(self, lr=None):
        """host side of one optimiser step: count it and publish the step size"""
        self.t += 1
        lr = float(self.lr if lr is None else lr)
        bc1, bc2 = 1.0 - self.betas[0] ** self.t, 1.0 - self.betas[1] ** self.t
        # a fill kernel carries the value in its launch arguments; an async copy from a reused
        # pinned scalar could be overtaken by the next step's value (the host runs ahead)
        self.lr_t_dev.fill_(lr * (bc2 ** 0.5) / bc1)

    def step_dev(self, grad_scale=1.0, lo=0, hi=None):
        """device side: one fused kernel over the flat buffers, or over their slice [lo, hi) — a bucket
        of whole parameter segments (graph-capturable)"""
        hi = self.n if hi is None else hi
        assert 0 <= lo < hi <= self.n and lo % 64 == 0
        self.bucket.rebind()
        self.bucket.settle(lo, hi)
        with torch.cuda.device(self.flat.device):
            check(load().scda_adam_step(
                self.flat.data_ptr() + 4 * lo, self.bucket.flat.data_ptr() + 4 * lo, self.exp_avg.data_ptr() + 4 * lo,
                self.exp_avg_sq.data_ptr() + 4 * lo,
                self.shadow.data_ptr() + 2 * lo if self.shadow is not None else None,
                hi - lo, max(self.t, 1), float(self.lr), self.betas[0], self.betas[1], self.eps,
                self.weight_decay, float(grad_scale), self.lr_t_dev.data_ptr(),
                stream_ptr(self.flat.device)), "scda_adam_step")
        # the kernel wrote the parameters through raw pointers: stamp what is (not) in sync
        for p, off in zip(self.params, self.offsets):
            if off < lo or off >= hi:
                continue
            p._scda_epoch = getattr(p, "_scda_epoch", 0) + 1
            if self.shadow is not None:
                p._scda_shadow_version = p._version
            elif hasattr(p, "_scda_shadow_version"):
                p._scda_shadow_version = -1

    def step(self, lr=None, grad_scale=1.0):
        self.begin_step(lr)
        self.step_dev(grad_scale)


def get_corner_from_center(center, recon_size, new_w, new_h):
    """(cluster_num, 2) centres -> recon_size windows shifted to stay inside new_w x new_h
    (tools/faster_rcnn_train_val.py:411-438)."""
    half = recon_size // 2
    corner = []
    for cx, cy in center:
        x_1 = max(int(cx) - half, 0)
        y_1 = max(int(cy) - half, 0)
        if x_1 == 0:
            x_2 = recon_size
        else:
            x_2 = min(int(cx) + half, new_w)
            if x_2 == new_w:
                x_1 = new_w - recon_size
        if y_1 == 0:
            y_2 = recon_size
        else:
            y_2 = min(int(cy) + half, new_h)
            if y_2 == new_h:
                y_1 = new_h - recon_size
        corner.append([x_1, y_1, x_2, y_2])
    return corner


def _crops(image, corners, recon_size):
    out = []
    for x1, y1, x2, y2 in corners:
        assert x2 - x1 == recon_size and y2 - y1 == recon_size, "crop size does not match recon_size"
        out.append(image[:, :, y1:y2, x1:x2])
    return torch.cat(out, 0)


def soft_label(flag, like, generator=None):
    """U(0.8, 1) for 1, U(0, 0.3) for 0 (:440-448), drawn on the device."""
    u = torch.rand(like.shape, device=like.device, generator=generator)
    return 0.8 + 0.2 * u if flag == 1 else 0.3 * u


def _bce_rows(logits, label_row):
    """per-cluster vector of F.binary_cross_entropy(sigmoid(logits[c:c+1]), label_row)
    (tools/faster_rcnn_train_val.py:577-600): one fused kernel each way (csrc/loss_ops.cu).
    CUDA only — there is no CPU path (loss_ops raises on host tensors)."""
    return bce_sigmoid_rows(logits, label_row)


class SCDATrainer(object):
    """The reference's four networks + optimisers for one rank.

    Nothing in the iteration synchronises with the host and every shape is fixed: after one
    eager warm-up iteration it is captured into ONE CUDA graph (two streams inside: detector
    backward + Adam beside the reconstruction / discriminator updates; with world > 1 the four
    NCCL all-reduces are captured too) and replayed — ~2800 kernel launches per iteration
    leave the Python/driver launch path.  use_graphs=False keeps it eager."""

    def __init__(self, model, dec_model, dis_model, dis_model_patch, lr, cluster_num=4,
                 threshold=128, recon_size=256, new_w=1024, new_h=512, world_size=1,
                 weight_decay=1e-4, use_graphs=True, overlap=True, graph_collectives=None, force_cut=False):
        self.model, self.dec_model = model, dec_model
        self.dis_model, self.dis_model_patch = dis_model, dis_model_patch
        self.opt = FlatAdam(model, lr, weight_decay=weight_decay, tensor_core=True)
        # channels_last weights: cuDNN's tensor-core convolutions then run NHWC end to end
        # (the decoder's 64-multiple 3x3 convolutions run on the tensor-core kernels: bf16 shadows)
        self.opt_dec = FlatAdam(dec_model, lr, weight_decay=weight_decay, tensor_core=gan_ops.TC_GAN,
                                channels_last=True)
        # (the discriminators' stride-2 convolutions run on the tensor-core kernels too: bf16 shadows, direct sinks)
        self.opt_dis = FlatAdam(dis_model, lr, weight_decay=weight_decay, tensor_core=gan_ops.TC_GAN,
                                channels_last=True)
        self.opt_dis_patch = FlatAdam(dis_model_patch, lr, weight_decay=weight_decay, tensor_core=gan_ops.TC_GAN,
                                      channels_last=True)
        self.cluster_num, self.threshold, self.recon_size = cluster_num, threshold, recon_size
        self.new_w, self.new_h, self.world_size = new_w, new_h, world_size
        self.use_graphs = use_graphs
        # overlap: the three reconstruction / discriminator updates run on a second stream
        # beside the detector's backward and optimiser step (nothing flows between them: the
        # cluster features are detached, functions/mask.py:234 of the reference).
        # graph_collectives (world > 1): capture the NCCL all-reduces inside the one graph;
        # None = the SCDA_GRAPH_COLLECTIVES environment variable, default ON.  (With ONE communicator the
        # captured form hung on 2 x B200, profiles/r1_ddp2_check.txt: two forked branches of the capture
        # issued collectives with no order between them.  With a communicator per issuing stream
        # (`_group_of`) it replays on 2 and on 8 GPUs: 8 x B200 987 images/s vs 958 for the cut plan,
        # profiles/r2_scale8.txt.)  SCDA_GRAPH_COLLECTIVES=0 / graph_collectives=False select the cut plan.
        self.overlap = overlap
        self.pair_streams = os.environ.get("SCDA_PAIR_STREAMS", "1") != "0"
        if graph_collectives is None:
            graph_collectives = os.environ.get("SCDA_GRAPH_COLLECTIVES", "1") != "0"
        self.graph_collectives = graph_collectives
        self.force_cut = force_cut          # tests: the cut (world > 1) replay plan on one GPU
        self.split_detector = os.environ.get("SCDA_SPLIT_DETECTOR", "1") != "0"
        self.wgrad_side = os.environ.get("SCDA_WGRAD_SIDE", "0") != "0"     # measured: 7.69 ms with, 7.59 without
        # the backbone's weight / bias gradients on their own stream beside its data-gradient chain
        self.body_wgrad_side = os.environ.get("SCDA_BODY_WGRAD_SIDE", "1") != "0"
        # the same for the decoder's convolution nodes in phase 3 (gan_ops.WGRAD_SIDE)
        self.gan_wgrad_side = os.environ.get("SCDA_GAN_WGRAD_SIDE", "1") != "0"
        # third gradient bucket of the detector (conv4_1 .. RPN head), see _det_chain
        self.mid_bucket = os.environ.get("SCDA_MID_BUCKET", "1") != "0"
        self._mid_off, self._mid_conv = None, None
        self._det_head_lo, self._oside = None, None
        self._comm_per_stream = os.environ.get("SCDA_ONE_COMM", "0") == "0"
        self._groups = None
        self._side = None
        self._tside = None
        self._aside = None
        self._pside = None
        self._ksides, self._gan_go, self._bwside = None, None, None
        self._copy_stream, self._stage, self._prefetched = None, {}, None
        if overlap and hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            # gradients are produced on whichever stream ran the forward of their branch and are
            # accumulated into the flat buffers there; the mismatch torch warns about is intended
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        self._static, self._st = None, {}
        self._graphs = None
        self._by_shape = {}
        # parity-test hooks (tests/test_iteration_parity_gpu.py); all None in production:
        #   rng  {'anchor': rng, 'proposal': rng}  prescribed sampling keys (functions/_sampling.ArrayRng)
        #   soft {'score_1', 'score_0' [1, M]; 'score_0_patch', 'score_1_patch' [K, P]}  soft-label rows
        #   taps {}  receives the detector's intermediate decisions (forces eager execution)
        self.rng, self.soft, self.taps = None, None, None

    def nets(self):
        return (self.model, self.dec_model, self.dis_model, self.dis_model_patch)

    def release_graphs(self):
        """Drop the captured iteration graphs (the next iteration re-captures).  With world > 1 and the
        all-reduces captured, call this before dist.destroy_process_group(): NCCL will not destroy a
        communicator while a live CUDA graph still holds kernels captured on it — the teardown waits forever."""
        import gc
        torch.cuda.synchronize()
        for ent in self._by_shape.values():
            ent['graphs'] = None
        self._graphs = None
        gc.collect()
        torch.cuda.synchronize()

    def close(self):
        """release the graphs and the per-stream communicators (world > 1)"""
        self.release_graphs()
        if self._groups is not None:
            import torch.distributed as dist
            for g in self._groups.values():
                dist.destroy_process_group(g)
            self._groups = None

    def train_mode(self):
        for n in self.nets():
            n.train()

    # ------------------------------------------------------------------ phases 1-3 (+ 4 forward)
    def _seg_patch_fwd(self):
        """forwards of the patch discriminator on both cluster-feature blocks (used by phases 1 and 2)"""
        b, st = self._static, self._st
        st['t_patch_pro'] = self.dis_model_patch(b['xt'])
        st['t_patch_mean'] = torch.mean(st['t_patch_pro'], 1)
        st['s_patch_pro'] = self.dis_model_patch(b['xs'])

    def _seg_dis(self, patch_stream=None, reduce=None, patch_done=False):
        """(1) image discriminator: forward + backward (tools/faster_rcnn_train_val.py:567-611).
        With `patch_stream` the patch discriminator's two forwards AND its whole update (phase 2:
        loss, backward, all-reduce, Adam) run on that stream beside this phase: phase 2 needs
        nothing from phase 1, and phase 1 needs only the VALUE mean(patch(x_target)) from it."""
        b, st, ws = self._static, self._st, float(self.world_size)
        xs, xt, cs, ct = b['xs'], b['xt'], b['cs'], b['ct']
        cur = torch.cuda.current_stream()
        if patch_stream is not None:
            patch_stream.wait_stream(cur)
            with torch.cuda.stream(patch_stream):
                st['t_patch_pro'] = self.dis_model_patch(xt)
                t_patch_mean = torch.mean(st['t_patch_pro'], 1)
                st['s_patch_pro'] = self.dis_model_patch(xs)
                have_mean = torch.cuda.Event()
                have_mean.record(patch_stream)
                self._patch_update()
                reduce(self.opt_dis_patch)
                self.opt_dis_patch.step_dev()
        st['recon'] = self.dec_model(xs, xt)
        x_source_recon, x_target_recon = st['recon']
        self.opt_dis.zero_grad()
        on_recon, on_real = run_pair(lambda: self.dis_model(x_source_recon, x_target_recon),
                                     lambda: self.dis_model(cs, ct))
        (s_dis, t_dis), (s_real, t_real) = on_recon, on_real        # logits: the sigmoid lives in _bce_rows
        score_1 = self._soft('score_1', 1, s_real[:1])
        score_0 = self._soft('score_0', 0, s_dis[:1])
        adloss_source = (_bce_rows(s_dis, score_1) + _bce_rows(s_real, score_0)).sum()
        if patch_done:                       # cut plan: an earlier stretch on the patch stream ran them
            t_patch_mean = st['t_patch_mean']
        elif patch_stream is None:
            st['t_patch_pro'] = self.dis_model_patch(xt)
            t_patch_mean = torch.mean(st['t_patch_pro'], 1)
            st['s_patch_pro'] = self.dis_model_patch(xs)
        else:
            cur.wait_event(have_mean)
            t_patch_mean.record_stream(cur)
        adloss_target = (t_patch_mean * _bce_rows(t_dis, score_0) + _bce_rows(t_real, score_1)).sum()
        adloss = (adloss_source + adloss_target) / ws
        from . import disc_ops
        with disc_ops.frozen_inputs():          # parameters of the discriminator only: no gradient to the decoder here
            adloss.backward(retain_graph=True, inputs=self.opt_dis.params)
        gan_ops.join_wgrad_streams()
        st['dis_loss'] = adloss.detach()

    def _soft(self, name, flag, like):
        if self.soft is not None and name in self.soft:
            return self.soft[name].to(like.device, like.dtype).view_as(like)
        return soft_label(flag, like)

    def _patch_update(self):
        """(2) patch discriminator: loss + backward (:616-630)"""
        st, ws = self._st, float(self.world_size)
        self.opt_dis_patch.zero_grad()
        score_0_patch = self._soft('score_0_patch', 0, st['t_patch_pro'])
        score_1_patch = self._soft('score_1_patch', 1, st['s_patch_pro'])
        dis_patch_loss = (F.binary_cross_entropy(st['s_patch_pro'], score_1_patch)
                          + F.binary_cross_entropy(st['t_patch_pro'], score_0_patch)) / ws
        dis_patch_loss.backward(retain_graph=True, inputs=self.opt_dis_patch.params)
        st['dis_patch_loss'] = dis_patch_loss.detach()

    def _seg_dis_patch(self):
        """(1) step; (2) patch discriminator update, in the reference's order"""
        self.opt_dis.step_dev()
        self._patch_update()

    def _seg_dec(self):
        """(2) step; (3) decoder: loss + backward (:635-699)"""
        self.opt_dis_patch.step_dev()
        self._dec_update()

    def _dec_update(self):
        b, st, ws = self._static, self._st, float(self.world_size)
        self.opt_dec.zero_grad()
        x_source_recon, x_target_recon = st['recon']
        (s_dis2, t_dis2), on_real = run_pair(lambda: self.dis_model(x_source_recon, x_target_recon),
                                             lambda: self.dis_model(b['cs'], b['ct']))
        s_real2, t_real2 = on_real                                  # logits, as above
        st['t_patch_mean2'] = torch.mean(self.dis_model_patch(b['xt']), 1).detach()
        t_patch_mean2 = st['t_patch_mean2']
        ones, zeros = torch.ones_like(t_dis2[:1]), torch.zeros_like(t_dis2[:1])
        fake_loss1_target = (t_patch_mean2 * (_bce_rows(t_dis2, ones) + _bce_rows(t_real2, zeros))).sum()
        fake_loss1_source = (_bce_rows(s_dis2, ones) + _bce_rows(s_real2, zeros)).sum()
        recon_loss = (fake_loss1_source + fake_loss1_target) / ws
        from . import disc_ops
        with disc_ops.frozen_params():          # through the discriminators to the decoder only
            recon_loss.backward(inputs=self.opt_dec.params)
        gan_ops.join_wgrad_streams()
        st['dec_loss'] = recon_loss.detach()
        st.pop('recon'), st.pop('t_patch_pro'), st.pop('s_patch_pro'), st.pop('t_patch_mean', None)

    def _seg_fake(self):
        """(3) step; forward of (4): decoders swapped, values only (:704-732).  The cluster
        features are detached (functions/mask.py:234), so these two terms carry no gradient
        to the detector: no autograd graph is built for them."""
        b, st = self._static, self._st
        self.opt_dec.step_dev()
        with torch.no_grad():
            x_source_recon2, x_target_recon2 = self.dec_model(b['xt'], b['xs'])
            s_dis3, t_dis3 = self.dis_model(x_source_recon2, x_target_recon2)
            # F.binary_cross_entropy over the whole [K, M] block = the mean of the K row means
            st['fake_loss_source'] = _bce_rows(t_dis3, torch.ones_like(t_dis3[:1])).mean()
            st['fake_loss_target'] = (st['t_patch_mean2']
                                      * _bce_rows(s_dis3, torch.ones_like(s_dis3[:1]))).sum()

    def _seg_forward(self):
        """detector forward on both images, crops around the cluster centres"""
        b, st = self._static, self._st
        x = {'cfg': b['cfg'], 'image': b['image'], 'image_info': b['info'], 'ground_truth_bboxes': b['gts'],
             'ignore_regions': None, 'cluster_num': self.cluster_num, 'threshold': self.threshold,
             'device_clusters': True}
        if self.overlap:
            x['target_stream'] = self._target_stream()
            x['aux_stream'] = self._aux_stream()
            x['side_streams'] = self._kmeans_streams()
            x['on_clusters'] = self._on_clusters
        if self.rng is not None:
            x['rng'] = self.rng
        if self.taps is not None:
            x['taps'] = self.taps
        self._gan_go = None
        outputs = self.model(x, b['target'])
        st['det_losses'] = outputs['losses']
        st['feat'] = outputs['feature_map']
        st['acc'] = outputs['accuracy']
        if self._gan_go is None:
            self._on_clusters(outputs['cluster_centers'], outputs['cluster_features'], mark=False)

    def _on_clusters(self, centers, features, mark=True):
        """crops around the cluster centres (tools/faster_rcnn_train_val.py:411-438,528-557).  Called by the
        detector's forward as soon as the clusters are known — before its loss kernels — and marks that point
        of the stream: everything phases 1-3 read (crops, cluster features) exists there, so the
        reconstruction chain is forked from the mark instead of from the end of the forward."""
        b = self._static
        centers_source, centers_target = centers
        b['cs'] = crops_device(b['image'], centers_source, self.recon_size, self.new_w, self.new_h)
        b['ct'] = crops_device(b['target'], centers_target, self.recon_size, self.new_w, self.new_h)
        b['xs'], b['xt'] = features
        if mark:
            self._gan_go = torch.cuda.Event()
            self._gan_go.record(torch.cuda.current_stream())

    def _seg_det_backward(self):
        """(4) detector backward (:736-745).  The two "fake" terms of the reference's loss carry
        no gradient to the detector (quirk above), so the backward runs on the four detection
        losses and does not wait for the reconstruction networks."""
        st, ws = self._st, float(self.world_size)
        rpn_cls_loss, rpn_loc_loss, rcnn_cls_loss, rcnn_loc_loss = st['det_losses']
        st['det_loss_sum'] = rpn_cls_loss + rpn_loc_loss + rcnn_cls_loss + rcnn_loc_loss
        self.opt.zero_grad()
        (st['det_loss_sum'] / ws).backward(inputs=self.opt.params)

    def _seg_step(self):
        self.opt.step_dev()

    # The same backward CUT AT THE FEATURE MAP into two buckets (">= 2 buckets, head
    # first"): the RCNN head's gradients (fc6 + fc7 + cls / loc = 478 of the detector's 547 MB) are complete
    # ~0.6 ms into the backward; their all-reduce and their Adam sweep (HBM bound) then run on a side stream
    # beside the backbone's backward (tensor-core bound) instead of behind it.
    def _head_lo(self):
        if self._det_head_lo is None:
            self._det_head_lo = self.opt.offset_of(self.model.classifier[0].weight)
        return self._det_head_lo

    def _mid_lo(self):
        """start of the third bucket = conv4_1's weight (the 8th convolution of the VGG16 stack); None when the
        backbone is not that stack"""
        if self._mid_off is None:
            convs = [m for m in self.model.features if isinstance(m, torch.nn.Conv2d)]
            if len(convs) != 13:
                self._mid_off = -1
            else:
                self._mid_off, self._mid_conv = self.opt.offset_of(convs[7].weight), 7
        return self._mid_off if self._mid_off > 0 else None

    def _seg_det_backward_head(self):
        st, ws = self._st, float(self.world_size)
        rpn_cls_loss, rpn_loc_loss, rcnn_cls_loss, rcnn_loc_loss = st['det_losses']
        st['det_loss_sum'] = rpn_cls_loss + rpn_loc_loss + rcnn_cls_loss + rcnn_loc_loss
        self.opt.zero_grad()
        # every head parameter takes its gradient through the direct sinks of tc_detector (written into
        # the flat buffer as a side effect), so differentiating w.r.t. the feature map alone runs the
        # whole head backward
        (st['g_feat'],) = torch.autograd.grad((rcnn_cls_loss + rcnn_loc_loss) / ws, [st['feat']])

    def _seg_det_backward_body(self):
        st, ws = self._st, float(self.world_size)
        rpn_cls_loss, rpn_loc_loss = st['det_losses'][:2]
        lo = self._head_lo()
        front = [p for p, off in zip(self.opt.params, self.opt.offsets) if off < lo]
        from . import tc_detector
        tc_detector.BODY_WGRAD_STREAM = self._body_wgrad_stream() if (self.overlap and self.body_wgrad_side) else None
        try:
            torch.autograd.backward([(rpn_cls_loss + rpn_loc_loss) / ws, st['feat']], [None, st.pop('g_feat')],
                                    inputs=front)
        finally:
            tc_detector.BODY_WGRAD_STREAM = None
        st.pop('feat')

    def _seg_step_head(self):
        self.opt.step_dev(lo=self._head_lo())

    def _seg_step_body(self):
        self.opt.step_dev(lo=0, hi=self._head_lo())

    def _opt_stream(self):
        if self._oside is None:
            self._oside = torch.cuda.Stream(device=self.opt.flat.device)
        return self._oside

    def _seg_outputs(self):
        st = self._st
        rpn_cls_loss, rpn_loc_loss, rcnn_cls_loss, rcnn_loc_loss = st.pop('det_losses')
        loss = st.pop('det_loss_sum').detach() + 0.1 * (st['fake_loss_source'] + st['fake_loss_target'])
        st['out'] = {'loss': loss, 'rpn_cls': rpn_cls_loss.detach(),
                     'rpn_loc': rpn_loc_loss.detach(), 'rcnn_cls': rcnn_cls_loss.detach(),
                     'rcnn_loc': rcnn_loc_loss.detach(), 'fake_loss': st['fake_loss_target'],
                     'fake_loss_source': st['fake_loss_source'],
                     'dec_loss': st['dec_loss'], 'dis_loss': st['dis_loss'],
                     'dis_patch_loss': st['dis_patch_loss'],
                     'rpn_acc': st['acc'][0], 'rcnn_acc': st['acc'][1]}

    # ------------------------------------------------------------------ the iteration
    def _gan_chain(self, reduce):
        if self.overlap and self.pair_streams and self._whole_graph():
            # phase 2 beside phase 1 (see _seg_dis); the cut plan of world > 1 keeps the phases in
            # the reference's order, one stretch each
            cur = torch.cuda.current_stream()
            helper = self._patch_stream()
            self._seg_dis(patch_stream=helper, reduce=reduce)
            reduce(self.opt_dis)
            self.opt_dis.step_dev()
            cur.wait_stream(helper)
            from . import timestamps as ts
            ts.mark("phase 1+2 done")
            self._dec_update()
        else:
            self._seg_dis()
            reduce(self.opt_dis)
            self._seg_dis_patch()
            reduce(self.opt_dis_patch)
            self._seg_dec()
        reduce(self.opt_dec)
        self._seg_fake()

    def _patch_stream(self):
        if self._pside is None:
            self._pside = torch.cuda.Stream(device=self.opt.flat.device, priority=-1)
        return self._pside

    def _det_chain(self, reduce):
        if not (self.overlap and self.split_detector):
            self._st.pop('feat', None)
            self._seg_det_backward()
            reduce(self.opt)
            self._seg_step()
            return
        lo = self._head_lo()
        cur, osd = torch.cuda.current_stream(), self._opt_stream()
        from . import tc_detector
        # the fc6 / fc7 weight gradients go to the opt stream too: the data-gradient chain (and with it the
        # backbone's backward) does not queue behind 478 MB of gradient writes
        osd.wait_stream(cur)
        tc_detector.WGRAD_STREAM = osd if self.wgrad_side else None
        try:
            self._seg_det_backward_head()
        finally:
            tc_detector.WGRAD_STREAM = None
        osd.wait_stream(cur)
        from . import timestamps as ts
        ts.mark("det head backward done")
        with torch.cuda.stream(osd):
            reduce(self.opt, lo, None)
            self._seg_step_head()
            ts.mark("head adam done")
        # a third bucket: conv4_1 .. conv5_3 + the RPN head (62 of the 69 MB in front of the head bucket) are
        # complete when the backward reaches conv3_3; they are reduced and stepped on the opt stream (behind the
        # head bucket, same communicator) while conv3_3 .. conv1_1 run.  What is left for the tail of the
        # iteration is the all-reduce and Adam of 7 MB.
        mid = self._mid_lo() if self.mid_bucket else None

        fired = []

        def mid_ready(main_stream, wgrad_stream):
            fired.append(True)
            osd.wait_stream(main_stream)
            if wgrad_stream is not None:
                osd.wait_stream(wgrad_stream)
            with torch.cuda.stream(osd):
                reduce(self.opt, mid, lo)
                self.opt.step_dev(lo=mid, hi=lo)

        tc_detector.BODY_BUCKET_HOOK = (self._mid_conv, mid_ready) if mid is not None else None
        try:
            self._seg_det_backward_body()
        finally:
            tc_detector.BODY_BUCKET_HOOK = None
        ts.mark("det body backward done")
        tail_hi = mid if fired else lo        # (a backbone form without the hook, e.g. the fp32-parity mode: two buckets)
        reduce(self.opt, 0, tail_hi)
        self.opt.step_dev(lo=0, hi=tail_hi)
        cur.wait_stream(osd)

    def _side_stream(self):
        """stream of the reconstruction / discriminator chain: the longer of the two chains,
        so it gets the higher priority when both have blocks waiting for an SM"""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.opt.flat.device, priority=-1)
        return self._side

    def _body_wgrad_stream(self):
        if self._bwside is None:
            self._bwside = torch.cuda.Stream(device=self.opt.flat.device)
        return self._bwside

    def _kmeans_streams(self):
        if self._ksides is None:
            self._ksides = tuple(torch.cuda.Stream(device=self.opt.flat.device) for _ in range(2))
        return self._ksides

    def _aux_stream(self):
        if self._aside is None:
            self._aside = torch.cuda.Stream(device=self.opt.flat.device)
        return self._aside

    def _target_stream(self):
        if self._tside is None:
            self._tside = torch.cuda.Stream(device=self.opt.flat.device)
        return self._tside

    def _body(self, reduce):
        """One iteration on the current stream (eager, or inside a capture).  With `overlap`
        the current stream forks after the detector forward: the side stream runs phases 1-3
        and the forward of phase 4 (hundreds of small cuDNN / elementwise kernels, latency
        bound), the current stream runs the detector backward and Adam (tensor-core / HBM
        bound), and they join before the losses are assembled."""
        from . import timestamps as ts
        gan_ops.PAIR_STREAMS = bool(self.overlap and self.pair_streams)
        gan_ops.WGRAD_SIDE = bool(self.overlap and self.gan_wgrad_side)
        from . import tc_detector
        tc_detector.MASK_SIDE_STREAMS = bool(self.overlap)
        ts.mark("start")
        self._seg_forward()
        ts.mark("forward done")
        if self.overlap:
            main = torch.cuda.current_stream()
            side = self._side_stream()
            if self._gan_go is not None:
                side.wait_event(self._gan_go)       # (recorded on `main` behind the crops, _on_clusters)
            else:
                side.wait_stream(main)
            with torch.cuda.stream(side):
                self._gan_chain(reduce)
                ts.mark("gan chain done")
            self._det_chain(reduce)
            ts.mark("det chain done")
            main.wait_stream(side)
        else:
            self._gan_chain(reduce)
            self._det_chain(reduce)
        self._seg_outputs()
        gan_ops.WGRAD_SIDE = False             # (module-level switches: only inside an iteration of this engine)
        ts.mark("end")

    def _reduce_fn(self):
        if self.world_size > 1:
            return lambda opt, lo=0, hi=None: opt.all_reduce(lo, hi, group=self._group_of(opt, lo))
        return lambda opt, lo=0, hi=None: None

    def _group_of(self, opt, lo=0):
        """One NCCL communicator per stream that issues collectives (`main`: backbone + RPN bucket of the
        detector, `opt`: its head bucket, `side`: image discriminator and decoder, `patch`: patch
        discriminator).  Collectives of ONE communicator must run in the same order on every rank; two
        forked branches of a captured graph carry no such order between them, which is what hung the
        captured form with a single communicator (profiles/r1_ddp2_check.txt, item 1).  With a
        communicator per branch every communicator sees one stream's program order, on every rank."""
        if not self._comm_per_stream or not self.overlap:
            return None
        import torch.distributed as dist
        if self._groups is None:
            if not (dist.is_available() and dist.is_initialized()):
                return None
            # created in the same order on every rank (the eager warm-up iteration gets here first).
            # SCDA_NCCL_OPT_CTAS=n caps the CTAs of the `opt` communicator: its one collective, the 478 MB head
            # bucket, runs beside the backbone's backward with ~3 ms of slack, so it can give SMs back to it.
            def make(name):
                cap = int(os.environ.get("SCDA_NCCL_OPT_CTAS", "0")) if name == "opt" else 0
                if cap > 0 and dist.get_backend() == "nccl":
                    opts = dist.ProcessGroupNCCL.Options()
                    opts.config.max_ctas = cap
                    opts.config.min_ctas = min(cap, 1)
                    return dist.new_group(backend="nccl", pg_options=opts)
                return dist.new_group(backend=dist.get_backend())
            self._groups = {k: make(k) for k in ("main", "opt", "side", "patch")}
        if opt is self.opt:
            return self._groups["opt" if (lo > 0 and self.split_detector) else "main"]
        if opt is self.opt_dis_patch:
            return self._groups["patch"]
        return self._groups["side"]

    def _segments(self):
        """world > 1 with the collectives NOT captured: the iteration cut at its four gradient
        all-reduces -> (name, stretch, stream it replays on, optimiser whose gradients are
        all-reduced after it).  With `overlap` three streams replay side by side: `main` (detector
        backward + Adam), `side` (phases 1, 3 and the forward of 4) and `patch` (the patch
        discriminator's forwards and phase 2); the collectives are issued from the host in one
        fixed order on every rank (patch, dis, dec, detector)."""
        if not self.overlap:
            return (('fwd', self._seg_forward, 'main', None),
                    ('dis', self._seg_dis, 'main', self.opt_dis),
                    ('dis_patch', self._seg_dis_patch, 'main', self.opt_dis_patch),
                    ('dec', self._seg_dec, 'main', self.opt_dec),
                    ('fake', self._seg_fake, 'main', None),
                    ('det_bwd', self._seg_det_backward, 'main', self.opt),
                    ('step', self._seg_step, 'main', None),
                    ('out', self._seg_outputs, 'main', None))
        head = (('fwd', self._seg_forward, 'main', None),
                ('patch_fwd', self._seg_patch_fwd, 'patch', None),
                ('patch_upd', self._patch_update, 'patch', self.opt_dis_patch),
                ('patch_step', self.opt_dis_patch.step_dev, 'patch', None),
                ('dis', lambda: self._seg_dis(patch_done=True), 'side', self.opt_dis),
                ('dis_step', self.opt_dis.step_dev, 'side', None),
                ('dec', self._dec_update, 'side', self.opt_dec),
                ('fake', self._seg_fake, 'side', None))
        if self.split_detector:
            # the detector's backward in two stretches (head bucket, then backbone + RPN bucket): the head
            # bucket's all-reduce and Adam go to the `opt` stream beside the second stretch
            lo = self._head_lo()
            det = (('det_bwd_head', self._seg_det_backward_head, 'main', None),
                   ('step_head', self._seg_step_head, 'opt', None),
                   ('det_bwd_body', self._seg_det_backward_body, 'main', (self.opt, 0, lo)),
                   ('step_body', self._seg_step_body, 'main', None))
        else:
            det = (('det_bwd', self._seg_det_backward, 'main', self.opt),
                   ('step', self._seg_step, 'main', None))
        return head + det + (('out', self._seg_outputs, 'main', None),)

    def _whole_graph(self):
        """world 1, or world > 1 with the all-reduces captured: ONE graph holds the iteration"""
        return (self.world_size == 1 or self.graph_collectives) and not self.force_cut

    def _capture(self):
        """Capture the iteration.  One graph (several streams inside, forked and joined within
        the capture) when `_whole_graph()`; otherwise one graph per stretch between two
        all-reduces.  Graphs that replay on the same stream share a memory pool (tensors handed
        from one stretch to the next stay put); the two streams' graphs replay concurrently and
        therefore allocate from different pools."""
        torch.cuda.synchronize()
        gan_ops.PAIR_STREAMS = bool(self.overlap and self.pair_streams)
        gan_ops.WGRAD_SIDE = bool(self.overlap and self.gan_wgrad_side)
        from . import tc_detector
        tc_detector.MASK_SIDE_STREAMS = bool(self.overlap)
        if self._whole_graph():
            reduce = self._reduce_fn()
            g = torch.cuda.CUDAGraph()
            # world > 1: NCCL's watchdog thread polls events while we capture; only THIS thread's calls
            # belong to the capture
            mode = "thread_local" if self.world_size > 1 else "global"
            with torch.cuda.graph(g, pool=torch.cuda.graph_pool_handle(), capture_error_mode=mode):
                self._body(reduce)
            return [g]
        pools = {k: torch.cuda.graph_pool_handle() for k in ('main', 'side', 'patch', 'opt')}
        graphs = []
        for _, fn, where, _ in self._segments():
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pools[where]):
                fn()
            graphs.append(g)
        gan_ops.WGRAD_SIDE = False
        return graphs

    def _replay_cut(self, graphs):
        """replay of the cut form: every stretch on its stream (and its all-reduce behind it on
        the same stream), the streams forked after the forward and joined before the losses are
        assembled; phase 1 waits for the patch forwards (it needs their mean), phase 3 for the
        whole patch stream (it needs the updated patch discriminator)"""
        main = torch.cuda.current_stream()
        segs = self._segments()
        by_name = {name: (g, opt) for g, (name, _, _, opt) in zip(graphs, segs)}

        def run(names):
            for n in names:
                g, opt = by_name[n]
                g.replay()
                if isinstance(opt, tuple):
                    opt[0].all_reduce(opt[1], opt[2], group=self._group_of(opt[0], opt[1]))
                elif opt is not None:
                    opt.all_reduce(group=self._group_of(opt))
        run(['fwd'])
        if not self.overlap:
            run(['dis', 'dis_patch', 'dec', 'fake', 'det_bwd', 'step', 'out'])
            return
        side, patch = self._side_stream(), self._patch_stream()
        patch.wait_stream(main)
        side.wait_stream(main)
        with torch.cuda.stream(patch):
            run(['patch_fwd'])
            have_mean = torch.cuda.Event()
            have_mean.record(patch)
            run(['patch_upd', 'patch_step'])
        with torch.cuda.stream(side):
            side.wait_event(have_mean)
            run(['dis', 'dis_step'])
            side.wait_stream(patch)
            run(['dec', 'fake'])
        if self.split_detector:
            osd = self._opt_stream()
            run(['det_bwd_head'])
            osd.wait_stream(main)
            with torch.cuda.stream(osd):
                self.opt.all_reduce(self._head_lo(), None, group=self._group_of(self.opt, self._head_lo()))
                run(['step_head'])
            run(['det_bwd_body', 'step_body'])
            main.wait_stream(osd)
        else:
            run(['det_bwd', 'step'])
        main.wait_stream(side)
        run(['out'])

    @staticmethod
    def _src_key(image, gts, target):
        return tuple((id(t), t.data_ptr(), tuple(t.shape), t.dtype) for t in (image, gts, target))

    def prefetch(self, image, gts, target):
        """Start the host -> device copy of the NEXT iteration's inputs (pinned host tensors) on a copy
        stream, so that it runs beside the iteration in flight (the role of the reference's DataLoader
        workers + `.cuda()` at the top of its loop, tools/faster_rcnn_train_val.py:528-533, made
        asynchronous).  `iteration()` called with the same tensors then takes the device copies instead of
        copying from the host.  Optional: without it `iteration()` copies its inputs itself."""
        dev = self.opt.flat.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        shape_key = (tuple(image.shape), tuple(target.shape), tuple(gts.shape), gts.dtype)
        st = self._stage.get(shape_key)
        if st is None:
            st = {'buf': {'image': torch.empty(image.shape, dtype=torch.float32, device=dev),
                          'target': torch.empty(target.shape, dtype=torch.float32, device=dev),
                          'gts': torch.empty(gts.shape, dtype=gts.dtype, device=dev)},
                  'free': torch.cuda.Event()}
            st['free'].record(torch.cuda.current_stream())
            self._stage[shape_key] = st
        cs = self._copy_stream
        cs.wait_event(st['free'])                # the previous iteration has taken its inputs out of `buf`
        with torch.cuda.stream(cs):
            st['buf']['image'].copy_(image, non_blocking=True)
            st['buf']['target'].copy_(target, non_blocking=True)
            st['buf']['gts'].copy_(gts, non_blocking=True)
            done = torch.cuda.Event()
            done.record(cs)
        self._prefetched = {'src': self._src_key(image, gts, target), 'buf': st['buf'], 'done': done,
                            'free': st['free']}

    def iteration(self, cfg, image, image_info, gts, target, lr=None):
        """One iteration; returns a dict of 0-dim loss tensors (no host sync here).

        image / target / gts may live on the host (pinned) or on the device: they are copied
        into fixed device buffers, which is what lets the iteration be replayed as a graph."""
        info_key = tuple(float(v) for v in torch.as_tensor(image_info).reshape(-1).tolist())
        key = (tuple(image.shape), tuple(target.shape), tuple(gts.shape), info_key, id(cfg))
        ent = self._by_shape.get(key)
        dev = self.opt.flat.device
        if ent is None:
            ent = {'static': {'cfg': cfg, 'info': image_info,
                              'image': torch.empty(image.shape, dtype=torch.float32, device=dev),
                              'target': torch.empty(target.shape, dtype=torch.float32, device=dev),
                              'gts': torch.empty(gts.shape, dtype=gts.dtype, device=dev)},
                   'graphs': None, 'calls': 0, 'st': {}}
            self._by_shape[key] = ent
        self._static, self._st = ent['static'], ent['st']
        b = self._static
        pf, self._prefetched = self._prefetched, None
        if pf is not None and pf['src'] == self._src_key(image, gts, target):
            # the inputs are already on the device (prefetch): wait for that copy, move them into the
            # graph's fixed buffers (12.6 MB device to device) and hand the staging buffers back
            main = torch.cuda.current_stream()
            main.wait_event(pf['done'])
            for k in ('image', 'target', 'gts'):
                b[k].copy_(pf['buf'][k], non_blocking=True)
            pf['free'].record(main)
        else:
            b['image'].copy_(image, non_blocking=True)
            b['target'].copy_(target, non_blocking=True)
            b['gts'].copy_(gts, non_blocking=True)
        for o in (self.opt_dis, self.opt_dis_patch, self.opt_dec, self.opt):
            o.begin_step(lr)
        if not self.use_graphs or ent['calls'] == 0 or self.taps is not None or self.rng is not None:
            self._body(self._reduce_fn())           # eager (also the warm-up before a capture)
        else:
            if ent['graphs'] is None:
                ent['graphs'] = self._capture()
            if self._whole_graph():
                ent['graphs'][0].replay()
            else:
                self._replay_cut(ent['graphs'])
        ent['calls'] += 1
        self._graphs = ent['graphs']
        return {k: v.clone() for k, v in self._st['out'].items()}


def crop_corners(centers, recon_size, new_w, new_h):
    """(x1, y1) of the recon_size windows: `get_corner_from_center`
    (tools/faster_rcnn_train_val.py:411-438) is a clamp of int(c) - recon_size // 2 to
    [0, size - recon_size].  The rule scda_crop_regions implements, as tensor ops (host logic,
    tests/test_engine_cpu.py)."""
    half = recon_size // 2
    x1 = (centers[:, 0].to(torch.int64) - half).clamp(0, new_w - recon_size)
    y1 = (centers[:, 1].to(torch.int64) - half).clamp(0, new_h - recon_size)
    return x1, y1


def crops_device(image, centers, recon_size, new_w, new_h):
    """recon_size windows around the cluster centres, shifted to stay inside the image (the rule
    of `crop_corners`), gathered by one kernel (csrc/proposal_ops.cu) from centres that never
    visited the host.  image [1,3,H,W] fp32 CUDA, centers [K,2] -> [K,3,R,R].  CUDA only."""
    if not torch.is_tensor(centers):
        centers = torch.as_tensor(np.asarray(centers), dtype=torch.float32, device=image.device)
    assert recon_size % 2 == 0 and image.shape[0] == 1
    if not (image.is_cuda and centers.is_cuda):
        raise RuntimeError("crops_device: CUDA tensors required (scda_b200 has no CPU path)")
    if tuple(image.shape[2:]) != (new_h, new_w):
        raise ValueError("crops_device: image is %s, expected (%d, %d)" % (tuple(image.shape[2:]), new_h, new_w))
    image = image.float().contiguous()
    cen = centers.float().contiguous()
    K, C = cen.shape[0], image.shape[1]
    out = torch.empty(K, C, recon_size, recon_size, dtype=torch.float32, device=image.device)
    with torch.cuda.device(image.device):
        check(load().scda_crop_regions(K, C, new_h, new_w, recon_size, image.data_ptr(), cen.data_ptr(),
                                       out.data_ptr(), stream_ptr(image.device)), "scda_crop_regions")
    return out


def builder_gan(cluster_num=4, threshold=128, recon_size=256, neww=64, newh=64):
    """The three GAN networks with the hyper-parameters of the reference's builder_gan
    (tools/faster_rcnn_train_val.py:255-273)."""
    from .models.faster_rcnn.faster_rcnn_adver_expansion_reweight_cluster import (
        GAN_decoder_AE, GAN_dis_AE, GAN_dis_AE_patch)
    size2layers = {256: 3, 512: 4, 128: 2}
    params_dec = {'ch': threshold, 'input_dim_a': 3, 'input_dim_b': 3, 'n_gen_res_blk': 3,
                  'n_gen_front_blk': size2layers[recon_size], 'res_dropout_ratio': 0.5,
                  'neww': neww, 'newh': newh, 'cluster_num': cluster_num, 'threshold': threshold}
    params_dis = {'input_dim_a': 3, 'input_dim_b': 3, 'ch': 32, 'n_gen_res_blk': 3,
                  'n_layer': size2layers[recon_size]}
    params_patch_dis = {'n_in': threshold, 'n_out': threshold * 2, 'cluster_num': cluster_num}
    return GAN_dis_AE(params_dis), GAN_decoder_AE(params_dec), GAN_dis_AE_patch(params_patch_dis)


def build_trainer(cfg, lr=1.25e-5, device="cuda", cluster_num=4, threshold=128, recon_size=256,
                  new_w=1024, new_h=512, world_size=1, seed=0, use_graphs=True, overlap=True,
                  graph_collectives=None, force_cut=False):
    from .models.faster_rcnn.vgg_adver_expansion_cluster import vgg16
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = vgg16(pretrained=False, cfg=cfg['shared']).to(device)
    dis_model, dec_model, dis_model_patch = builder_gan(cluster_num, threshold, recon_size)
    tr = SCDATrainer(model, dec_model.to(device), dis_model.to(device), dis_model_patch.to(device),
                     lr, cluster_num, threshold, recon_size, new_w, new_h, world_size,
                     use_graphs=use_graphs, overlap=overlap, graph_collectives=graph_collectives,
                     force_cut=force_cut)
    tr.train_mode()
    return tr
