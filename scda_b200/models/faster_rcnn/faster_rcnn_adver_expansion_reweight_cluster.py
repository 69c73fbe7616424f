"""Detector graph and the adversarial / reconstruction networks (mirrors
models/faster_rcnn/faster_rcnn_adver_expansion_reweight_cluster.py of the reference):
`FasterRCNN_AdEx` (:19-234), `smooth_l1_loss_with_sigma` (:238-246), `accuracy` (:249-267),
`GAN_dis_AE` (:270-308), `GAN_dis_AE_patch` (:312-333), `GAN_decoder_AE` (:336-399).

What differs from the reference is where the plumbing between the networks runs: the
reference hops to the host for anchors, proposals, NMS scan, targets and sampling
(~10 H2D + ~8 D2H synchronous copies per image); here those stages run on
the device through scda_b200.functions.* and the only host round trip left in the training
forward is the 512 x 5 RoI table for the k-means of compute_cluster_targets.
"""
import functools
import logging
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...functions.anchor_target import compute_anchor_targets
from ...functions.mask import cluster_targets_device, compute_cluster_targets, kmeans_regions_forked
from ...functions.predict_bbox import compute_predicted_bboxes
from ...functions.proposal_target import compute_proposal_targets, proposal_targets_device
from ...functions.rpn_proposal import compute_rpn_proposals, rpn_proposals_device
from ...gan_ops import run_pair
from ...loss_ops import rpn_fg_scores, smooth_l1_masked_sum, softmax_ce_acc
from .common_net import (ConvTranspose1x1, TanhAfterHead, INSResBlock, LeakyReLUConv2d, LeakyReLUConvTranspose2d_2,
                         LinUnsRes_cluster, ResDis_cluster, gaussian_weights_init)

logger = logging.getLogger('global')


def _host_info(image_info):
    if torch.is_tensor(image_info):
        return image_info.cpu().numpy() if image_info.is_cuda else image_info.numpy()
    return image_info


class FasterRCNN_AdEx(nn.Module):
    def __init__(self, gan_model_flag):
        super(FasterRCNN_AdEx, self).__init__()

    def feature_extractor(self, x):
        raise NotImplementedError

    def rpn(self, x):
        raise NotImplementedError

    def rcnn(self, x, rois):
        raise NotImplementedError

    def _add_rpn_loss(self, compute_anchor_targets_fn, rpn_pred_cls, rpn_pred_loc):
        cls_targets, loc_targets, loc_masks, loc_normalizer = \
            compute_anchor_targets_fn(rpn_pred_loc.size())
        rpn_pred_cls = rpn_pred_cls.permute(0, 2, 3, 1).contiguous().view(-1, 2)
        cls_targets = cls_targets.permute(0, 2, 3, 1).contiguous().view(-1)
        rpn_loss_loc = _smooth_l1_masked(rpn_pred_loc, loc_masks, loc_targets) / loc_normalizer
        if rpn_pred_cls.is_cuda:        # cross entropy (ignore -1) + top-1 accuracy: one pass (csrc/loss_ops.cu)
            rpn_loss_cls, acc = softmax_ce_acc(rpn_pred_cls, cls_targets, ignore_index=-1)
            return rpn_loss_cls, rpn_loss_loc, acc
        rpn_loss_cls = F.cross_entropy(rpn_pred_cls, cls_targets, ignore_index=-1)
        acc = accuracy(rpn_pred_cls.data, cls_targets.data)[0]
        return rpn_loss_cls, rpn_loss_loc, acc

    def _add_rcnn_loss(self, rcnn_pred_cls, rcnn_pred_loc, cls_targets, loc_targets, loc_weights):
        loc_normalizer = cls_targets.shape[0]
        rcnn_loss_loc = _smooth_l1_masked(rcnn_pred_loc, loc_weights, loc_targets) / loc_normalizer
        if rcnn_pred_cls.is_cuda:
            rcnn_loss_cls, acc = softmax_ce_acc(rcnn_pred_cls.contiguous(), cls_targets, ignore_index=-100)
            return rcnn_loss_cls, rcnn_loss_loc, acc
        rcnn_loss_cls = F.cross_entropy(rcnn_pred_cls, cls_targets)
        acc = accuracy(rcnn_pred_cls, cls_targets)[0]
        return rcnn_loss_cls, rcnn_loss_loc, acc

    def _pin_args_to_fn(self, cfg, ground_truth_bboxes, image_info, ignore_regions, rng=None):
        partial_fn = {}
        if self.training:
            partial_fn['anchor_target_fn'] = functools.partial(
                compute_anchor_targets, cfg=cfg['train_anchor_target_cfg'],
                ground_truth_bboxes=ground_truth_bboxes, ignore_regions=ignore_regions,
                image_info=image_info, rng=(rng or {}).get('anchor'))
            partial_fn['proposal_target_fn'] = functools.partial(
                compute_proposal_targets, cfg=cfg['train_proposal_target_cfg'],
                ground_truth_bboxes=ground_truth_bboxes, ignore_regions=ignore_regions,
                image_info=image_info)
            partial_fn['rpn_proposal_fn'] = functools.partial(
                compute_rpn_proposals, cfg=cfg['train_rpn_proposal_cfg'], image_info=image_info)
        else:
            partial_fn['rpn_proposal_fn'] = functools.partial(
                compute_rpn_proposals, cfg=cfg['test_rpn_proposal_cfg'], image_info=image_info)
            partial_fn['predict_bbox_fn'] = functools.partial(
                compute_predicted_bboxes, image_info=image_info, cfg=cfg['test_predict_bbox_cfg'])
        return partial_fn

    @staticmethod
    def _rpn_scores(rpn_pred_cls):
        """2-way softmax over each anchor's (bg, fg) channel pair, NCHW in / NCHW out (:153-155)."""
        x = rpn_pred_cls.permute(0, 2, 3, 1).contiguous()
        x = F.softmax(x.view(-1, 2), dim=1).view_as(x)
        return x.permute(0, 3, 1, 2)

    def _train_rois(self, cfg, proposals_per_image, ground_truth_bboxes, image_info, rng=None):
        """proposal targets straight from the device-side proposal buffers (no host hop)."""
        info = _host_info(image_info)
        outs = []
        for b, (boxes, n_keep) in enumerate(proposals_per_image):
            outs.append(proposal_targets_device(
                boxes, n_keep, ground_truth_bboxes[b].float(), cfg['train_proposal_target_cfg'],
                (float(info[b][0]), float(info[b][1])), batch_ix=b, rng=rng))
        return tuple(torch.cat([o[i] for o in outs], 0).contiguous() for i in range(4))

    def forward(self, input, target=None):
        '''
        input: dict with 'cfg', 'image' [b,3,h,w], 'ground_truth_bboxes' [b,max_num_gts,5] or
               None, 'image_info' [b,3], 'ignore_regions', and (training) 'cluster_num',
               'threshold'.
        target: the unlabelled target-domain image [b,3,h,w] (training only).
        Return: dict of losses, predict, accuracy (+ cluster_features / cluster_centers).
        '''
        cfg = input['cfg']
        x_input = input['image']
        ground_truth_bboxes = input['ground_truth_bboxes']
        image_info = input['image_info']
        ignore_regions = input['ignore_regions']
        if self.training and ground_truth_bboxes is not None and not ground_truth_bboxes.is_cuda:
            ground_truth_bboxes = ground_truth_bboxes.to(x_input.device, non_blocking=True)
        # tests only: input['rng'] = {'anchor': rng, 'proposal': rng} replays prescribed sampling keys
        # (functions/_sampling.ArrayRng); input['taps'] (a dict) receives the intermediate decisions
        rng = input.get('rng') or {}
        taps = input.get('taps')
        partial_fn = self._pin_args_to_fn(cfg, ground_truth_bboxes, image_info, ignore_regions, rng)
        if taps is not None and self.training:
            raw_fn, memo = partial_fn['anchor_target_fn'], {}

            def anchor_once(size):          # computed once per forward, also when the taps ask for it again
                key = tuple(int(v) for v in size)
                if key not in memo:
                    memo[key] = raw_fn(size)
                return memo[key]
            partial_fn['anchor_target_fn'] = anchor_once

        outputs = {'losses': [], 'predict': [], 'accuracy': []}
        if not self.training:
            x = self.feature_extractor(x_input)
            rpn_pred_cls, rpn_pred_loc = self.rpn(x)

        if self.training:
            pcfg = cfg['train_rpn_proposal_cfg']
            n_t = cfg['train_proposal_target_cfg']['batch_size']
            # input['device_clusters']: keep the cluster centres on the device (no host
            # synchronisation anywhere in this forward); default = the reference's return
            # type, centres as a host numpy array
            on_dev = bool(input.get('device_clusters', False))
            cluster_fn = cluster_targets_device if on_dev else compute_cluster_targets
            # input['side_streams'] = (s0, s1): the k-means of the source / target RoIs runs on s0 / s1 beside
            # the RCNN head (it reads the RoI table only), joined where the rows are gathered
            kstreams = input.get('side_streams') if (on_dev and taps is None) else None
            # input['on_clusters'](centres, features): called as soon as both are known, BEFORE the four
            # detection losses are joined — the engine crops the regions there, so the reconstruction chain
            # does not wait for the loss kernels
            on_clusters = input.get('on_clusters') if on_dev else None

            def run_target_backbone():
                with torch.no_grad():
                    x_gan = self.feature_extractor(target)
                    rpn_pred_cls_gan, rpn_pred_loc_gan = self.rpn(x_gan)
                return x_gan, rpn_pred_cls_gan, rpn_pred_loc_gan

            def run_target(dense=None):
                """RPN + RCNN on the target image: top-512 proposals, no ground truth (:171-189)
                (nothing downstream differentiates through this branch: its only product, the
                cluster features, is detached by compute_cluster_targets — functions/mask.py:234)"""
                x_gan, rpn_pred_cls_gan, rpn_pred_loc_gan = dense if dense is not None else run_target_backbone()
                props_gan = rpn_proposals_device(None, rpn_pred_loc_gan.data, pcfg, image_info,
                                                 fg_scores=rpn_fg_scores(rpn_pred_cls_gan))
                gan_rows = []
                for b, (boxes, n_keep) in enumerate(props_gan):
                    gan_rows.append((torch.cat([torch.full((boxes.shape[0], 1), float(b),
                                                           device=boxes.device), boxes[:, :4]], 1),
                                     n_keep))
                if len(gan_rows) == 1 and gan_rows[0][0].shape[0] >= n_t:
                    proposals_gan = gan_rows[0][0][:n_t].contiguous()
                    enough = gan_rows[0][1] >= n_t           # 0-dim device flag, read by the caller
                else:
                    ks = [int(n.item()) for _, n in gan_rows]
                    proposals_gan = torch.cat([r[:k] for (r, _), k in zip(gan_rows, ks)], 0)[:n_t].contiguous()
                    enough = torch.tensor(proposals_gan.shape[0] == n_t, device=x_gan.device)
                pre = None
                if kstreams is not None and proposals_gan.shape[0] == n_t:
                    pre = kmeans_regions_forked(proposals_gan, input['cluster_num'], input['threshold'], kstreams[1])
                with torch.no_grad():
                    x_fea_gan, _, _ = self.rcnn(x_gan, proposals_gan)
                clusters = None
                if on_dev and proposals_gan.shape[0] == n_t:
                    ktap = {} if taps is not None else None
                    clusters = cluster_targets_device(proposals_gan, x_fea_gan, N_cluster=input['cluster_num'],
                                                      threshold=input['threshold'], taps=ktap, pre=pre)
                    if taps is not None:
                        taps.update(cluster_tgt=ktap)
                if taps is not None:
                    taps.update(rois_gan=proposals_gan, enough=enough, feat_gan=x_gan, fc7_gan=x_fea_gan,
                                rpn_cls_gan=rpn_pred_cls_gan, rpn_loc_gan=rpn_pred_loc_gan)
                return x_gan, proposals_gan, enough, x_fea_gan, clusters

            # input['target_stream']: run the target branch on that stream, beside the source
            # branch (its dense backbone fills the SMs the source's latency-bound proposal /
            # target plumbing leaves idle); forked from and joined back into the current stream
            tstream = input.get('target_stream') if on_dev else None
            # (measured on B200: forking after the source backbone was issued schedules better than before it)
            early = os.environ.get("SCDA_EARLY_FORK", "0") == "1"
            # default: BOTH dense backbones back to back on the current stream, then the two latency-bound
            # proposal / RoI chains side by side.  (With the target's backbone beside the source's chain, the
            # chain's kernels — 48-150 KB of shared memory each — could not share an SM with the persistent
            # convolution CTAs and waited for the gaps between them: 2.02 -> 1.8x ms for the forward.)
            dense_first = tstream is not None and not early and os.environ.get("SCDA_TARGET_DENSE_FIRST", "1") == "1"
            target_dense = run_target_backbone() if dense_first else None
            if tstream is not None and early:
                cur_stream = torch.cuda.current_stream()
                tstream.wait_stream(cur_stream)
                with torch.cuda.stream(tstream):
                    tgt = run_target()

            # input['aux_stream']: the anchor targets depend on the ground truth and the feature-map
            # SIZE only, not on the network: computed on that stream beside the backbone
            astream = input.get('aux_stream') if on_dev else None
            conv_loc = getattr(getattr(self, 'rpn_head', None), 'conv_loc', None)
            anchor_pre = None
            if astream is not None and conv_loc is not None:
                # forked from HERE (an event in front of the backbone) but issued BEHIND the backbone:
                # a replayed CUDA graph dispatches its nodes roughly in creation order, and ~110 tiny
                # kernels created first held the backbone's first node back by 0.58 ms
                # (profiles/r2_timeline_a_forward.txt)
                main_stream = torch.cuda.current_stream()
                fork_point = torch.cuda.Event()
                fork_point.record(main_stream)

            x = self.feature_extractor(x_input)
            rpn_pred_cls, rpn_pred_loc = self.rpn(x)
            if astream is not None and conv_loc is not None:
                astream.wait_event(fork_point)
                with torch.cuda.stream(astream):
                    pre = partial_fn['anchor_target_fn'](rpn_pred_loc.size())
                anchor_pre = pre
                partial_fn['anchor_target_fn'] = lambda size: pre
            late = os.environ.get("SCDA_TARGET_LATE", "0")
            if tstream is not None and not early:
                cur_stream = torch.cuda.current_stream()
                after_rpn = torch.cuda.Event()
                after_rpn.record(cur_stream)
                if late == "0":
                    tstream.wait_event(after_rpn)
                    with torch.cuda.stream(tstream):
                        tgt = run_target(target_dense)
            from ... import timestamps as ts
            ts.mark("src rpn head done")
            fg = rpn_fg_scores(rpn_pred_cls)
            props = rpn_proposals_device(None, rpn_pred_loc.data, pcfg, image_info, fg_scores=fg)
            ts.mark("src proposals done")
            if tstream is not None and not early and late == "1":
                tstream.wait_event(after_rpn)
                with torch.cuda.stream(tstream):
                    tgt = run_target()
            rois, cls_targets, loc_targets, loc_weights = self._train_rois(
                cfg, props, ground_truth_bboxes, image_info, rng.get('proposal'))
            assert rois.shape[1] == 5
            ts.mark("src targets done")
            if os.environ.get("SCDA_DEBUG_TARGETS") == "1":
                self._dbg = dict(orig=(rois, cls_targets, loc_targets, loc_weights),
                                 early=tuple(t.clone() for t in (rois, cls_targets, loc_targets, loc_weights)))
            pre_src = None
            if kstreams is not None:
                pre_src = kmeans_regions_forked(rois, input['cluster_num'], input['threshold'], kstreams[0])
            x_fea, rcnn_pred_cls, rcnn_pred_loc = self.rcnn(x, rois)
            if tstream is not None and not early and late == "2":
                tstream.wait_event(after_rpn)
                with torch.cuda.stream(tstream):
                    tgt = run_target()
            if taps is not None and on_dev:
                ktap = {}
                x_cluster_fea, x_center_cluster = cluster_targets_device(
                    rois, x_fea, N_cluster=input['cluster_num'], threshold=input['threshold'], taps=ktap)
                taps.update(cluster_src=ktap,
                            rois_targets=(rois, cls_targets, loc_targets, loc_weights), feat=x,
                            rpn_cls=rpn_pred_cls, rpn_loc=rpn_pred_loc, fc7=x_fea, rcnn_cls=rcnn_pred_cls,
                            rcnn_loc=rcnn_pred_loc, proposals=props, fg_scores=fg)
            elif pre_src is not None:
                x_cluster_fea, x_center_cluster = cluster_targets_device(
                    rois, x_fea, N_cluster=input['cluster_num'], threshold=input['threshold'], pre=pre_src)
            else:
                x_cluster_fea, x_center_cluster = cluster_fn(
                    rois, x_fea, N_cluster=input['cluster_num'], threshold=input['threshold'])

            ts.mark("src rcnn+kmeans done")
            if tstream is not None:
                cur_stream.wait_stream(tstream)
            else:
                tgt = run_target()
            ts.mark("target joined")
            x_gan, proposals_gan, enough, x_fea_gan, clusters_gan = tgt
            assert x_gan.size() == x.size(), "gan_features does not match the backbone"

            # fewer than 512 surviving target proposals: reuse the source clusters (:207-215)
            if proposals_gan.shape[0] != n_t:
                logger.info("Different channels {} at target image".format(x_fea_gan.size(0)))
                outputs['cluster_features'] = [x_cluster_fea, x_cluster_fea]
                outputs['cluster_centers'] = [x_center_cluster, x_center_cluster]
            elif on_dev:
                # the same choice made by a device-side select on the `enough` flag
                fea_gan, center_gan = clusters_gan
                outputs['cluster_features'] = [x_cluster_fea, torch.where(enough, fea_gan, x_cluster_fea)]
                outputs['cluster_centers'] = [x_center_cluster,
                                              torch.where(enough, center_gan, x_center_cluster)]
            elif not bool(enough):
                logger.info("Different channels {} at target image".format(x_fea_gan.size(0)))
                outputs['cluster_features'] = [x_cluster_fea, x_cluster_fea]
                outputs['cluster_centers'] = [x_center_cluster, x_center_cluster]
            else:
                x_cluster_fea_gan, x_center_cluster_gan = compute_cluster_targets(
                    proposals_gan, x_fea_gan, N_cluster=input['cluster_num'],
                    threshold=input['threshold'])
                outputs['cluster_features'] = [x_cluster_fea, x_cluster_fea_gan]
                outputs['cluster_centers'] = [x_center_cluster, x_center_cluster_gan]
            if on_clusters is not None:
                on_clusters(outputs['cluster_centers'], outputs['cluster_features'])

            # the four losses are only needed by the backward: computed here, at the end of the forward,
            # so that the proposal / RoI stages above never wait for the anchor targets and the consumer of
            # the clusters (on_clusters) never waits for the loss kernels
            if anchor_pre is not None:
                torch.cuda.current_stream().wait_stream(astream)
                for t_ in anchor_pre:
                    if torch.is_tensor(t_):
                        t_.record_stream(torch.cuda.current_stream())
            if taps is not None:
                taps.update(anchor_targets=partial_fn['anchor_target_fn'](rpn_pred_loc.size()))
            rpn_loss_cls, rpn_loss_loc, rpn_acc = self._add_rpn_loss(
                partial_fn['anchor_target_fn'], rpn_pred_cls, rpn_pred_loc)
            rcnn_loss_cls, rcnn_loss_loc, rcnn_acc = self._add_rcnn_loss(
                rcnn_pred_cls, rcnn_pred_loc, cls_targets, loc_targets, loc_weights)
            if os.environ.get("SCDA_DEBUG_TARGETS") == "1":
                self._dbg['late'] = tuple(t.clone() for t in (rois, cls_targets, loc_targets, loc_weights))
                self._dbg['pred'] = (rcnn_pred_cls.detach().clone(), rcnn_pred_loc.detach().clone())
            outputs['losses'] = [rpn_loss_cls, rpn_loss_loc, rcnn_loss_cls, rcnn_loss_loc]
            outputs['accuracy'] = [rpn_acc, rcnn_acc]
            outputs['predict'] = [props]
            outputs['feature_map'] = x          # (engine: the detector's backward is cut here, see SCDATrainer)
        else:
            proposals = partial_fn['rpn_proposal_fn'](self._rpn_scores(rpn_pred_cls).data,
                                                      rpn_pred_loc.data)
            proposals = proposals[:, :5].cuda().contiguous()
            assert proposals.shape[1] == 5
            x_fea, rcnn_pred_cls, rcnn_pred_loc = self.rcnn(x, proposals)
            rcnn_pred_cls = F.softmax(rcnn_pred_cls, dim=1)
            bboxes = partial_fn['predict_bbox_fn'](proposals, rcnn_pred_cls, rcnn_pred_loc)
            outputs['predict'] = [proposals, bboxes]
        return outputs


def _smooth_l1_masked(pred, mask, targets, sigma=3.0):
    """smooth_l1_loss_with_sigma(pred * mask, targets): one fused kernel each way on CUDA
    (csrc/loss_ops.cu), the reference's chain of tensor ops elsewhere"""
    if pred.is_cuda and pred.dtype == torch.float32 and mask.dtype == torch.float32 and mask.shape == pred.shape:
        return smooth_l1_masked_sum(pred, mask, targets.float(), sigma)
    return smooth_l1_loss_with_sigma(pred * mask, targets, sigma)


def smooth_l1_loss_with_sigma(pred, targets, sigma=3.0):
    sigma_2 = sigma ** 2
    diff = pred - targets
    abs_diff = torch.abs(diff)
    smoothL1_sign = (abs_diff < 1. / sigma_2).detach().float()
    loss = torch.pow(diff, 2) * sigma_2 / 2. * smoothL1_sign \
        + (abs_diff - 0.5 / sigma_2) * (1. - smoothL1_sign)
    return torch.sum(loss)


def accuracy(output, target, topk=(1,), ignore_index=-1):
    """precision@k in percent over the rows whose target != ignore_index.  Same values as the
    reference's (:249-267) without its nonzero()/index host synchronisation."""
    keep = target != ignore_index
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.eq(target.view(-1, 1)) & keep.view(-1, 1)
    n = keep.sum().clamp(min=1).float()
    return [correct[:, :k].reshape(-1).float().sum(0, keepdim=True) * (100.0 / n) for k in topk]


class GAN_dis_AE(nn.Module):
    """Image-level discriminators, one per domain: n_layer stride-2 LeakyReLU convs and a
    1x1 conv to one channel; outputs are flattened to [clusters, H*W/4^n]."""

    def __init__(self, params):
        super(GAN_dis_AE, self).__init__()
        ch, input_dim_a, n_layer = params['ch'], params['input_dim_a'], params['n_layer']
        self.model_A = self._make_net(ch, input_dim_a, n_layer - 1)
        self.model_A.apply(gaussian_weights_init)
        self.model_B = self._make_net(ch, input_dim_a, n_layer - 1)
        self.model_B.apply(gaussian_weights_init)

    def _make_net(self, ch, input_dim, n_layer):
        model = [LeakyReLUConv2d(input_dim, ch, kernel_size=3, stride=2, padding=1)]
        tch = ch
        for _ in range(n_layer):
            model += [LeakyReLUConv2d(tch, tch * 2, kernel_size=3, stride=2, padding=1)]
            tch *= 2
        model += [nn.Conv2d(tch, 1, kernel_size=1, stride=1, padding=0)]
        return nn.Sequential(*model)

    @staticmethod
    def _run(seq, x):
        if x.is_cuda:
            from ... import disc_ops
            if disc_ops.image_dis_supported(seq, x):
                # one autograd node on hand-written kernels (scda_b200/disc_ops.py): direct first layer,
                # stride-2 tcgen05 convolutions with bias + LeakyReLU in the epilogue, dot-product head
                return disc_ops.image_dis(seq, x)
            # NHWC end to end: no cuDNN layout transposes around the convolutions
            x = x.contiguous(memory_format=torch.channels_last)
        return seq(x)

    def forward(self, x_aa, x_bb):
        out_A, out_B = run_pair(lambda: self._run(self.model_A, x_aa), lambda: self._run(self.model_B, x_bb))
        return out_A.reshape(out_A.size(0), -1), out_B.reshape(out_B.size(0), -1)


class GAN_dis_AE_patch(nn.Module):
    """Feature-level discriminator on the [clusters, threshold, 64, 64] RoI-feature image."""

    def __init__(self, params=None):
        super(GAN_dis_AE_patch, self).__init__()
        if params:
            self.n_in, self.n_out, cluster_num1 = params['n_in'], params['n_out'], params['cluster_num']
        else:
            self.n_in, self.n_out, cluster_num1 = 128, 256, 4
        self.model_A_patch = nn.Sequential(
            ResDis_cluster(n_in=self.n_in, n_out=self.n_out, kernel_size=3, stride=2, padding=1,
                           w=64, h=64, cluster_num=cluster_num1))

    def forward(self, rois_features):
        return torch.sigmoid(self.model_A_patch(rois_features))


class GAN_decoder_AE(nn.Module):
    """Two decoders (source / target): view to [clusters, ch, 64, 64] -> n_gen_res_blk IN
    res-blocks -> (n_gen_front_blk - 1) x [bilinear x2, conv3x3 halving channels, IN, LReLU]
    -> 1x1 ConvTranspose to 3 channels -> tanh."""

    def __init__(self, params):
        super(GAN_decoder_AE, self).__init__()
        input_dim_b, ch = params['input_dim_b'], params['ch']
        n_gen_res_blk, n_gen_front_blk = params['n_gen_res_blk'], params['n_gen_front_blk']
        res_dropout_ratio = params.get('res_dropout_ratio', 0)
        neww, newh = params.get('neww', 64), params.get('newh', 64)
        cluster_num = params.get('cluster_num', 4)

        def make():
            tch = ch
            dec = [LinUnsRes_cluster(ch, neww, newh, cluster_num)]
            for _ in range(n_gen_res_blk):
                dec += [INSResBlock(tch, tch, dropout=res_dropout_ratio)]
            for _ in range(n_gen_front_blk - 1):
                dec += [LeakyReLUConvTranspose2d_2(tch, tch // 2, kernel_size=3, stride=1,
                                                   padding=1, output_padding=0)]
                tch = tch // 2
            head = ConvTranspose1x1(tch, input_dim_b, kernel_size=1, stride=1, padding=0)
            head.fuse_tanh = True
            dec += [head]
            dec += [TanhAfterHead()]
            return nn.Sequential(*dec)

        # construction order B then A, as in the reference (it fixes the RNG stream of the init)
        self.decode_B = make()
        self.decode_B.apply(gaussian_weights_init)
        self.decode_A = make()
        self.decode_A.apply(gaussian_weights_init)

    def forward(self, x_aa, x_bb):
        return run_pair(lambda: self.decode_A(x_aa), lambda: self.decode_B(x_bb))
