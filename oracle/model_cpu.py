"""CPU restatement of the reference's training iteration — the "reference CPU-extension
path" of BASELINE.md §4.  Two uses, both test infrastructure:
  * the timed CPU baseline (bench.py --impl reference and the cpu_baseline leg);
  * the checker of the whole-iteration parity test (tests/test_iteration_parity_gpu.py):
    module and parameter names are the reference's, so one state dict loads into this and
    into the product's networks, and `iteration(..., forced=..., soft=...)` replays the
    data-dependent DISCRETE decisions (sampled anchors / RoIs, cluster membership) and the
    random soft labels handed to it, so that the continuous arithmetic — every loss value,
    every gradient, every Adam update — can be compared on identical decisions.

TEST INFRASTRUCTURE ONLY.  Nothing here imports scda_b200.

PARITY UNPINNED by the reference: it has no end-to-end CPU path (its functions hard-code
.cuda(), RoIAlign and focal loss ship no CPU source) and no tests, so this
path is assembled from:
  * plain torch.nn modules on CPU, fp32, all host threads, with the reference's layer
    lists and names (models/faster_rcnn/vgg_adver_expansion_cluster.py:27-60,101-120;
    models/head.py:3-32; faster_rcnn_adver_expansion_reweight_cluster.py:270-399;
    common_net.py:59-80,107-129,160-169,205-245,251-261,279-293);
  * the C restatements of the reference's CUDA ops (oracle/scda_oracle.c, OpenMP), pinned
    against the reference's own kernels by tests/test_oracle_cpu.py;
  * the numpy restatement of the reference's host plumbing (oracle/host.py, pinned against
    the reference's own Python by tests/test_oracle_host.py) and the reference's sklearn
    KMeans call;
  * Adam as torch 0.4.1 implements it (the version the reference pins, README.md:16;
    torch/optim/adam.py of that release: eps is added to sqrt(v) BEFORE the bias
    correction, which modern torch.optim.Adam does not do) x 4, and the four-phase update
    of tools/faster_rcnn_train_val.py:567-750.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import host, roi_pool_backward, roi_pool_forward

VGG16_D = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512]


class _RoIPoolCPU(torch.autograd.Function):
    """extensions/_roi_pooling/functions/roi_pool.py:6-42 over the C oracle."""

    @staticmethod
    def forward(ctx, features, rois, ph, pw, scale):
        out, arg = roi_pool_forward(features.detach().numpy(), rois.numpy(), ph, pw, scale)
        ctx.meta = (rois.numpy().copy(), arg, tuple(features.shape), scale)
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, grad_output):
        rois, arg, shape, scale = ctx.meta
        g = roi_pool_backward(grad_output.contiguous().numpy(), rois, arg, shape, scale)
        return torch.from_numpy(g), None, None, None, None


def _gauss(m):
    if m.__class__.__name__.find('Conv') == 0:
        m.weight.data.normal_(0.0, 0.02)


class _RpnHead(nn.Module):
    """models/head.py:3-32"""

    def __init__(self, inplanes, num_anchors):
        super().__init__()
        self.conv3x3 = nn.Conv2d(inplanes, 512, 3, 1, 1)
        self.conv_cls = nn.Conv2d(512, num_anchors * 2, 1)
        self.conv_loc = nn.Conv2d(512, num_anchors * 4, 1)

    def forward(self, x):
        x = F.relu(self.conv3x3(x))
        return self.conv_cls(x), self.conv_loc(x)


class Detector(nn.Module):
    """models/faster_rcnn/vgg_adver_expansion_cluster.py:27-98 (names: features.N,
    rpn_head.*, classifier.{0,3}, fc_rcnn_cls, fc_rcnn_loc)"""

    def __init__(self, shared_cfg):
        super().__init__()
        layers, cin = [], 3
        for v in VGG16_D:
            if v == 'M':
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(inplace=True)]
                cin = v
        self.features = nn.Sequential(*layers)
        A = len(shared_cfg['anchor_scales']) * len(shared_cfg['anchor_ratios'])
        self.rpn_head = _RpnHead(512, A)
        self.classifier = nn.Sequential(nn.Linear(512 * 49, 4096), nn.ReLU(True), nn.Dropout(),
                                        nn.Linear(4096, 4096), nn.ReLU(True), nn.Dropout())
        self.fc_rcnn_cls = nn.Linear(4096, shared_cfg['num_classes'])
        self.fc_rcnn_loc = nn.Linear(4096, shared_cfg['num_classes'] * 4)
        self.scale = 1.0 / shared_cfg['anchor_stride']
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, math.sqrt(2. / (m.kernel_size[0] * m.kernel_size[1] * m.out_channels)))
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()

    def rpn(self, x):
        return self.rpn_head(x)

    def rcnn(self, x, rois):
        p = _RoIPoolCPU.apply(x, rois, 7, 7, self.scale)
        f = self.classifier(p.view(p.size(0), -1))
        return f, self.fc_rcnn_cls(f), self.fc_rcnn_loc(f)


def _smooth_l1(pred, targets, sigma=3.0):
    s2 = sigma ** 2
    d = pred - targets
    a = d.abs()
    sign = (a < 1. / s2).detach().float()
    return (d.pow(2) * s2 / 2. * sign + (a - 0.5 / s2) * (1. - sign)).sum()


def _scores(cls):
    x = cls.permute(0, 2, 3, 1).contiguous()
    return F.softmax(x.view(-1, 2), dim=1).view_as(x).permute(0, 3, 1, 2)


def _gather_clusters(fea, index, k, th):
    return fea.detach()[torch.as_tensor(np.asarray(index), dtype=torch.int64)].view(k, th, -1)


def detector_forward(model, cfg, image, image_info, gts, target, cluster_num, threshold, forced=None,
                     taps=None):
    """FasterRCNN_AdEx.forward, training branch (…reweight_cluster.py:106-215).

    forced: optional dict of decisions to replay instead of drawing them here —
      'anchor_targets' (cls, loc, mask, normalizer), 'rois_targets' (rois, labels, loc_t, loc_w),
      'rois_gan' [512, 5], 'cluster_src' / 'cluster_tgt' (row index [K*threshold], centres [K, 2]).
    taps: optional dict that receives the continuous intermediates (rpn outputs, features)."""
    forced = forced or {}
    info = image_info.numpy()
    x = model.features(image)
    cls, loc = model.rpn(x)
    if 'anchor_targets' in forced:
        ct, lt, lm, norm = forced['anchor_targets']
    else:
        ct, lt, lm, norm = host.compute_anchor_targets(tuple(loc.shape), cfg['train_anchor_target_cfg'],
                                                       gts.numpy(), info)
    pc = cls.permute(0, 2, 3, 1).contiguous().view(-1, 2)
    tc = torch.as_tensor(ct).permute(0, 2, 3, 1).contiguous().view(-1)
    rpn_loss_cls = F.cross_entropy(pc, tc, ignore_index=-1)
    rpn_loss_loc = _smooth_l1(loc * torch.as_tensor(lm), torch.as_tensor(lt)) / float(norm)
    if 'rois_targets' in forced:
        rois, labels, loc_t, loc_w = [torch.as_tensor(a) for a in forced['rois_targets']]
    else:
        props = host.compute_rpn_proposals(_scores(cls).detach().numpy(), loc.detach().numpy(),
                                           cfg['train_rpn_proposal_cfg'], info)
        rois, labels, loc_t, loc_w = [torch.from_numpy(a) for a in host.compute_proposal_targets(
            props, cfg['train_proposal_target_cfg'], gts.numpy(), info)]
    fea, rc, rl = model.rcnn(x, rois)
    if 'cluster_src' in forced:
        src_centers = np.asarray(forced['cluster_src'][1])
        src_fea = _gather_clusters(fea, forced['cluster_src'][0], cluster_num, threshold)
    else:
        f, src_centers, _ = host.compute_cluster_targets(rois.numpy(), fea.detach().numpy(),
                                                         cluster_num, threshold)
        src_fea = torch.from_numpy(f)
    with torch.no_grad():       # nothing differentiates through the target branch (functions/mask.py:234)
        xg = model.features(target)
        cls_g, loc_g = model.rpn(xg)
        if 'rois_gan' in forced:
            rois_g = torch.as_tensor(forced['rois_gan'])
        else:
            props_g = host.compute_rpn_proposals(_scores(cls_g).numpy(), loc_g.numpy(),
                                                 cfg['train_rpn_proposal_cfg'], info)
            rois_g = torch.from_numpy(props_g[:512, :5].copy())
        fea_g, _, _ = model.rcnn(xg, rois_g)
    if fea_g.size(0) != 512:
        tgt_fea, tgt_centers = src_fea, src_centers
    elif 'cluster_tgt' in forced:
        tgt_centers = np.asarray(forced['cluster_tgt'][1])
        tgt_fea = _gather_clusters(fea_g, forced['cluster_tgt'][0], cluster_num, threshold)
    else:
        f, tgt_centers, _ = host.compute_cluster_targets(rois_g.numpy(), fea_g.numpy(), cluster_num, threshold)
        tgt_fea = torch.from_numpy(f)
    rcnn_loss_cls = F.cross_entropy(rc, labels)
    rcnn_loss_loc = _smooth_l1(rl * loc_w, loc_t) / labels.shape[0]
    if taps is not None:
        taps.update(feat=x.detach(), rpn_cls=cls.detach(), rpn_loc=loc.detach(), feat_gan=xg, rpn_cls_gan=cls_g,
                    rpn_loc_gan=loc_g, fc7=fea.detach(), rcnn_cls=rc.detach(), rcnn_loc=rl.detach(),
                    fc7_gan=fea_g)
    return ([rpn_loss_cls, rpn_loss_loc, rcnn_loss_cls, rcnn_loss_loc],
            (src_fea.detach(), tgt_fea.detach()), (src_centers, tgt_centers))


# ------------------------------------------------------------------ reconstruction networks
class _Reshape(nn.Module):
    """common_net.py:107-129 LinUnsRes_cluster: a view, no parameters (index 0 of each decoder)"""

    def __init__(self, ch, clusters):
        super().__init__()
        self.ch, self.clusters = ch, clusters

    def forward(self, x):
        return x.view(self.clusters, self.ch, 64, 64)


class _ResBlock(nn.Module):
    """common_net.py:59-80 INSResBlock"""

    def __init__(self, ch, dropout):
        super().__init__()
        m = [nn.Conv2d(ch, ch, 3, 1, 1), nn.InstanceNorm2d(ch), nn.ReLU(inplace=True),
             nn.Conv2d(ch, ch, 3, 1, 1), nn.InstanceNorm2d(ch)]
        if dropout > 0:
            m.append(nn.Dropout(p=dropout))
        self.model = nn.Sequential(*m)
        self.model.apply(_gauss)

    def forward(self, x):
        return self.model(x) + x


class _Interp(nn.Module):
    def forward(self, x):
        return F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)


class _Up(nn.Module):
    """common_net.py:279-293 LeakyReLUConvTranspose2d_2: upsample, conv (model.1), IN, LeakyReLU"""

    def __init__(self, cin, cout):
        super().__init__()
        self.model = nn.Sequential(_Interp(), nn.Conv2d(cin, cout, 3, 1, 1), nn.InstanceNorm2d(cout),
                                   nn.LeakyReLU(inplace=True))
        self.model.apply(_gauss)

    def forward(self, x):
        return self.model(x)


class Decoder(nn.Module):
    """…reweight_cluster.py:336-399 GAN_decoder_AE (decode_B constructed before decode_A)"""

    def __init__(self, ch=128, clusters=4, n_res=3, n_front=3, dropout=0.5):
        super().__init__()

        def make():
            layers = [_Reshape(ch, clusters)] + [_ResBlock(ch, dropout) for _ in range(n_res)]
            t = ch
            for _ in range(n_front - 1):
                layers.append(_Up(t, t // 2))
                t //= 2
            layers += [nn.ConvTranspose2d(t, 3, 1), nn.Tanh()]
            seq = nn.Sequential(*layers)
            seq.apply(_gauss)
            return seq
        self.decode_B = make()
        self.decode_A = make()

    def forward(self, a, b):
        return self.decode_A(a), self.decode_B(b)


class _LReLUConv(nn.Module):
    """common_net.py:251-261 LeakyReLUConv2d"""

    def __init__(self, cin, cout):
        super().__init__()
        self.model = nn.Sequential(nn.Conv2d(cin, cout, 3, 2, 1), nn.LeakyReLU(inplace=True))
        self.model.apply(_gauss)

    def forward(self, x):
        return self.model(x)


class ImageDis(nn.Module):
    """…reweight_cluster.py:270-308 GAN_dis_AE"""

    def __init__(self, ch=32, n_layer=3):
        super().__init__()

        def make():
            m, t = [_LReLUConv(3, ch)], ch
            for _ in range(n_layer - 1):
                m.append(_LReLUConv(t, t * 2))
                t *= 2
            m.append(nn.Conv2d(t, 1, 1))
            seq = nn.Sequential(*m)
            seq.apply(_gauss)
            return seq
        self.model_A = make()
        self.model_B = make()

    def forward(self, a, b):
        oa, ob = self.model_A(a), self.model_B(b)
        return oa.view(oa.size(0), -1), ob.view(ob.size(0), -1)


class _ResDis(nn.Module):
    """common_net.py:205-245 ResDis_cluster"""

    def __init__(self, n_in, clusters):
        super().__init__()
        self.n_in, self.clusters = n_in, clusters
        n_out = 2 * n_in
        self.model = nn.Sequential(
            nn.Conv2d(n_in, n_out, 3, 2, 1, bias=False), nn.BatchNorm2d(n_out), nn.LeakyReLU(inplace=True),
            nn.Conv2d(n_out, 2 * n_out, 3, 2, 1, bias=False), nn.BatchNorm2d(2 * n_out), nn.LeakyReLU(inplace=True),
            nn.Conv2d(2 * n_out, 2 * n_out, 3, 2, 1, bias=False))
        self.model.apply(_gauss)

    def forward(self, x):
        o = self.model(x.view(self.clusters, self.n_in, 64, 64))
        return torch.squeeze(F.avg_pool2d(o, o.size()[2:]))


class PatchDis(nn.Module):
    """…reweight_cluster.py:312-333 GAN_dis_AE_patch"""

    def __init__(self, n_in=128, clusters=4):
        super().__init__()
        self.model_A_patch = nn.Sequential(_ResDis(n_in, clusters))

    def forward(self, x):
        return torch.sigmoid(self.model_A_patch(x))


class Adam041(object):
    """torch.optim.Adam.step() of torch 0.4.1 (amsgrad off): weight decay folded into the
    gradient; denom = sqrt(exp_avg_sq) + eps; step_size = lr * sqrt(1 - b2^t) / (1 - b1^t)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.state = {}

    def zero_grad(self):
        for p in self.params:
            if p.grad is not None:
                p.grad.detach_()
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        b1, b2 = self.betas
        for p in self.params:
            if p.grad is None:
                continue
            g = p.grad
            st = self.state.setdefault(p, {'step': 0, 'm': torch.zeros_like(p), 'v': torch.zeros_like(p)})
            st['step'] += 1
            if self.wd != 0:
                g = g.add(p, alpha=self.wd)
            st['m'].mul_(b1).add_(g, alpha=1 - b1)
            st['v'].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = st['v'].sqrt().add_(self.eps)
            step_size = self.lr * math.sqrt(1 - b2 ** st['step']) / (1 - b1 ** st['step'])
            p.addcdiv_(st['m'], denom, value=-step_size)


def _soft(flag, like):
    """generate_soft_label (tools/faster_rcnn_train_val.py:440-448): numpy draw on the host"""
    lo, hi = (0.8, 1.0) if flag == 1 else (0.0, 0.3)
    return torch.from_numpy(np.random.uniform(lo, hi, size=tuple(like.shape))).float()


class CPUTrainer(object):
    def __init__(self, cfg, lr=1.25e-5, cluster_num=4, threshold=128, recon_size=256, new_w=1024,
                 new_h=512, seed=0, dropout=True):
        torch.manual_seed(seed)
        np.random.seed(seed)
        self.cfg, self.k, self.th, self.rs, self.w, self.h = cfg, cluster_num, threshold, recon_size, new_w, new_h
        self.model = Detector(cfg['shared']).train()
        self.dec = Decoder(threshold, cluster_num).train()
        self.dis = ImageDis().train()
        self.dis_patch = PatchDis(threshold, cluster_num).train()
        if not dropout:
            for net in self.nets():
                for m in net.modules():
                    if isinstance(m, nn.Dropout):
                        m.p = 0.0
        self.keep_grads, self.grads_at_step = False, {}
        mk = lambda m: Adam041(m.parameters(), lr=lr, weight_decay=0.0001)
        self.opt, self.opt_dec, self.opt_dis, self.opt_patch = mk(self.model), mk(self.dec), mk(self.dis), mk(self.dis_patch)

    def nets(self):
        """same order as scda_b200.engine.SCDATrainer.nets()"""
        return (self.model, self.dec, self.dis, self.dis_patch)

    def iteration(self, image, image_info, gts, target, forced=None, soft=None, taps=None):
        """tools/faster_rcnn_train_val.py:526-750, world_size 1.  Returns the total detector
        loss; `self.last` holds every logged component.
        soft: optional dict of the four soft-label rows ('score_1', 'score_0' [1, M];
        'score_0_patch', 'score_1_patch' [K, P]) instead of numpy draws."""
        bce = F.binary_cross_entropy
        soft = soft or {}
        losses, (sp, tp), (cs, ct_) = detector_forward(self.model, self.cfg, image, image_info, gts,
                                                       target, self.k, self.th, forced, taps)
        crop = lambda img, cen: torch.cat([img[:, :, y1:y2, x1:x2] for x1, y1, x2, y2 in
                                           host.get_corner_from_center(cen, self.rs, self.w, self.h)], 0)
        xs, ts = crop(image, cs), crop(target, ct_)
        s_rec, t_rec = self.dec(sp, tp)
        # (1) image discriminator (:567-616)
        self.opt_dis.zero_grad()
        s_dis, t_dis = [torch.sigmoid(o) for o in self.dis(s_rec, t_rec)]
        s_real, t_real = [torch.sigmoid(o) for o in self.dis(xs, ts)]
        one = torch.as_tensor(soft['score_1']) if 'score_1' in soft else _soft(1, s_real[:1])
        zero = torch.as_tensor(soft['score_0']) if 'score_0' in soft else _soft(0, s_dis[:1])
        t_pro = self.dis_patch(tp)
        t_mean = t_pro.mean(1)
        s_pro = self.dis_patch(sp)
        ad = 0.0
        for c in range(self.k):
            ad = ad + bce(s_dis[c:c + 1], one) + bce(s_real[c:c + 1], zero)
            ad = ad + t_mean[c] * bce(t_dis[c:c + 1], zero) + bce(t_real[c:c + 1], one)
        ad.backward(retain_graph=True)
        self._snap('image_dis', self.dis)
        self.opt_dis.step()
        # (2) patch discriminator (:623-635)
        self.opt_patch.zero_grad()
        z_p = torch.as_tensor(soft['score_0_patch']) if 'score_0_patch' in soft else _soft(0, t_pro)
        o_p = torch.as_tensor(soft['score_1_patch']) if 'score_1_patch' in soft else _soft(1, s_pro)
        pl = bce(s_pro, o_p) + bce(t_pro, z_p)
        pl.backward(retain_graph=True)
        self._snap('patch_dis', self.dis_patch)
        self.opt_patch.step()
        # (3) decoder (:642-704)
        self.opt_dec.zero_grad()
        s_dis, t_dis = [torch.sigmoid(o) for o in self.dis(s_rec, t_rec)]
        s_real, t_real = [torch.sigmoid(o) for o in self.dis(xs, ts)]
        t_mean2 = self.dis_patch(tp).mean(1)
        o1, z1 = torch.ones_like(t_dis[:1]), torch.zeros_like(t_dis[:1])
        rl = 0.0
        for c in range(self.k):
            rl = rl + t_mean2[c] * (bce(t_dis[c:c + 1], o1) + bce(t_real[c:c + 1], z1))
            rl = rl + bce(s_dis[c:c + 1], o1) + bce(s_real[c:c + 1], z1)
        rl.backward(retain_graph=True)
        self._snap('decoder', self.dec)
        self.opt_dec.step()
        # (4) detector (:716-750)
        s_rec2, t_rec2 = self.dec(tp, sp)
        s_d, t_d = self.dis(s_rec2, t_rec2)
        fs = torch.sigmoid(t_d)
        f_src = bce(fs, torch.ones_like(fs))
        f2 = torch.sigmoid(s_d)
        f_tgt = 0.0
        for c in range(self.k):
            f_tgt = f_tgt + t_mean2[c] * bce(f2[c:c + 1], torch.ones_like(f2[c:c + 1]))
        loss = sum(losses) + 0.1 * (f_src + f_tgt)
        self.opt.zero_grad()
        loss.backward()
        self._snap('detector', self.model)
        self.opt.step()
        fl = lambda t: float(t.detach())
        self.last = {'loss': fl(loss), 'rpn_cls': fl(losses[0]), 'rpn_loc': fl(losses[1]),
                     'rcnn_cls': fl(losses[2]), 'rcnn_loc': fl(losses[3]),
                     'fake_loss_source': fl(f_src), 'fake_loss': fl(f_tgt), 'dec_loss': fl(rl),
                     'dis_loss': fl(ad), 'dis_patch_loss': fl(pl)}
        return self.last['loss']

    def _snap(self, name, net):
        """gradients of `net` as its optimiser sees them (the reference's later backward passes add
        stray gradients to the other networks, which their next zero_grad() discards)"""
        if self.keep_grads:
            self.grads_at_step[name] = {n: p.grad.detach().clone() for n, p in net.named_parameters()
                                        if p.grad is not None}
