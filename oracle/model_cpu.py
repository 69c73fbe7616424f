"""CPU restatement of the reference's training iteration — the "reference CPU-extension
path" of BASELINE.md §4, used ONLY as the timed CPU baseline (bench.py --impl reference and
the cpu_baseline leg) and as a shape/semantics check in tests.

TEST INFRASTRUCTURE ONLY.  Nothing here imports scda_b200.

The reference has no end-to-end CPU path (its functions hard-code .cuda(), RoIAlign and
focal loss ship no CPU source), so the path is assembled from:
  * plain torch.nn modules on CPU, fp32, all host threads, with the reference's layer
    lists (models/faster_rcnn/vgg_adver_expansion_cluster.py:27-60,101-120;
    models/head.py:3-32; faster_rcnn_adver_expansion_reweight_cluster.py:270-399;
    common_net.py:59-80,107-129,160-169,205-245,251-261,279-293);
  * the C restatements of the reference's CUDA ops (oracle/scda_oracle.c, OpenMP);
  * the numpy restatement of the reference's host plumbing (oracle/host.py) and the
    reference's own sklearn KMeans call;
  * torch.optim.Adam(lr, weight_decay=1e-4) x 4 and the four-phase update of
    tools/faster_rcnn_train_val.py:567-750.
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


class Detector(nn.Module):
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
        self.rpn_conv = nn.Conv2d(512, 512, 3, padding=1)
        self.rpn_cls = nn.Conv2d(512, A * 2, 1)
        self.rpn_loc = nn.Conv2d(512, A * 4, 1)
        self.classifier = nn.Sequential(nn.Linear(512 * 49, 4096), nn.ReLU(True), nn.Dropout(),
                                        nn.Linear(4096, 4096), nn.ReLU(True), nn.Dropout())
        self.fc_cls = nn.Linear(4096, shared_cfg['num_classes'])
        self.fc_loc = nn.Linear(4096, shared_cfg['num_classes'] * 4)
        self.scale = 1.0 / shared_cfg['anchor_stride']
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, math.sqrt(2. / (m.kernel_size[0] * m.kernel_size[1] * m.out_channels)))
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()

    def rpn(self, x):
        x = F.relu(self.rpn_conv(x))
        return self.rpn_cls(x), self.rpn_loc(x)

    def rcnn(self, x, rois):
        p = _RoIPoolCPU.apply(x, rois, 7, 7, self.scale)
        f = self.classifier(p.view(p.size(0), -1))
        return f, self.fc_cls(f), self.fc_loc(f)


def _smooth_l1(pred, targets, sigma=3.0):
    s2 = sigma ** 2
    d = pred - targets
    a = d.abs()
    sign = (a < 1. / s2).detach().float()
    return (d.pow(2) * s2 / 2. * sign + (a - 0.5 / s2) * (1. - sign)).sum()


def _scores(cls):
    x = cls.permute(0, 2, 3, 1).contiguous()
    return F.softmax(x.view(-1, 2), dim=1).view_as(x).permute(0, 3, 1, 2)


def detector_forward(model, cfg, image, image_info, gts, target, cluster_num, threshold):
    """FasterRCNN_AdEx.forward, training branch (…reweight_cluster.py:106-215)."""
    info = image_info.numpy()
    x = model.features(image)
    cls, loc = model.rpn(x)
    ct, lt, lm, norm = host.compute_anchor_targets(tuple(loc.shape), cfg['train_anchor_target_cfg'],
                                                   gts.numpy(), info)
    pc = cls.permute(0, 2, 3, 1).contiguous().view(-1, 2)
    tc = torch.from_numpy(ct).permute(0, 2, 3, 1).contiguous().view(-1)
    rpn_loss_cls = F.cross_entropy(pc, tc, ignore_index=-1)
    rpn_loss_loc = _smooth_l1(loc * torch.from_numpy(lm), torch.from_numpy(lt)) / norm
    props = host.compute_rpn_proposals(_scores(cls).detach().numpy(), loc.detach().numpy(),
                                       cfg['train_rpn_proposal_cfg'], info)
    rois, labels, loc_t, loc_w = [torch.from_numpy(a) for a in host.compute_proposal_targets(
        props, cfg['train_proposal_target_cfg'], gts.numpy(), info)]
    fea, rc, rl = model.rcnn(x, rois)
    src_fea, src_centers, _ = host.compute_cluster_targets(rois.numpy(), fea.detach().numpy(),
                                                           cluster_num, threshold)
    xg = model.features(target)
    cls_g, loc_g = model.rpn(xg)
    props_g = host.compute_rpn_proposals(_scores(cls_g).detach().numpy(), loc_g.detach().numpy(),
                                         cfg['train_rpn_proposal_cfg'], info)
    rois_g = torch.from_numpy(props_g[:512, :5].copy())
    fea_g, _, _ = model.rcnn(xg, rois_g)
    if fea_g.size(0) != 512:
        tgt_fea, tgt_centers = src_fea, src_centers
    else:
        tgt_fea, tgt_centers, _ = host.compute_cluster_targets(rois_g.numpy(), fea_g.detach().numpy(),
                                                               cluster_num, threshold)
    rcnn_loss_cls = F.cross_entropy(rc, labels)
    rcnn_loss_loc = _smooth_l1(rl * loc_w, loc_t) / labels.shape[0]
    return ([rpn_loss_cls, rpn_loss_loc, rcnn_loss_cls, rcnn_loss_loc],
            (torch.from_numpy(src_fea), torch.from_numpy(tgt_fea)), (src_centers, tgt_centers))


class _ResBlock(nn.Module):
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


class _Up(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, 1, 1)
        self.norm = nn.InstanceNorm2d(cout)
        self.apply(_gauss)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)
        return F.leaky_relu(self.norm(self.conv(x)))


class Decoder(nn.Module):
    def __init__(self, ch=128, clusters=4, n_res=3, n_front=3, dropout=0.5):
        super().__init__()
        self.ch, self.clusters = ch, clusters

        def make():
            layers = [_ResBlock(ch, dropout) for _ in range(n_res)]
            t = ch
            for _ in range(n_front - 1):
                layers.append(_Up(t, t // 2))
                t //= 2
            layers += [nn.ConvTranspose2d(t, 3, 1), nn.Tanh()]
            seq = nn.Sequential(*layers)
            seq.apply(_gauss)
            return seq
        self.decode_B, self.decode_A = make(), make()

    def forward(self, a, b):
        v = lambda t: t.view(self.clusters, self.ch, 64, 64)
        return self.decode_A(v(a)), self.decode_B(v(b))


class ImageDis(nn.Module):
    def __init__(self, ch=32, n_layer=3):
        super().__init__()

        def make():
            m, t = [nn.Conv2d(3, ch, 3, 2, 1), nn.LeakyReLU(inplace=True)], ch
            for _ in range(n_layer - 1):
                m += [nn.Conv2d(t, t * 2, 3, 2, 1), nn.LeakyReLU(inplace=True)]
                t *= 2
            m.append(nn.Conv2d(t, 1, 1))
            seq = nn.Sequential(*m)
            seq.apply(_gauss)
            return seq
        self.model_A, self.model_B = make(), make()

    def forward(self, a, b):
        oa, ob = self.model_A(a), self.model_B(b)
        return oa.view(oa.size(0), -1), ob.view(ob.size(0), -1)


class PatchDis(nn.Module):
    def __init__(self, n_in=128, clusters=4):
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
        return torch.sigmoid(torch.squeeze(F.avg_pool2d(o, o.size()[2:])))


def _soft(flag, like):
    lo, hi = (0.8, 1.0) if flag == 1 else (0.0, 0.3)
    return torch.from_numpy(np.random.uniform(lo, hi, size=tuple(like.shape))).float()


class CPUTrainer(object):
    def __init__(self, cfg, lr=1.25e-5, cluster_num=4, threshold=128, recon_size=256, new_w=1024,
                 new_h=512, seed=0):
        torch.manual_seed(seed)
        np.random.seed(seed)
        self.cfg, self.k, self.th, self.rs, self.w, self.h = cfg, cluster_num, threshold, recon_size, new_w, new_h
        self.model = Detector(cfg['shared']).train()
        self.dec, self.dis, self.dis_patch = Decoder(threshold, cluster_num).train(), ImageDis().train(), PatchDis(threshold, cluster_num).train()
        mk = lambda m: torch.optim.Adam(m.parameters(), lr=lr, weight_decay=0.0001)
        self.opt, self.opt_dec, self.opt_dis, self.opt_patch = mk(self.model), mk(self.dec), mk(self.dis), mk(self.dis_patch)

    def iteration(self, image, image_info, gts, target):
        """tools/faster_rcnn_train_val.py:526-750, world_size 1."""
        bce = F.binary_cross_entropy
        losses, (sp, tp), (cs, ct_) = detector_forward(self.model, self.cfg, image, image_info, gts,
                                                       target, self.k, self.th)
        crop = lambda img, cen: torch.cat([img[:, :, y1:y2, x1:x2] for x1, y1, x2, y2 in
                                           host.get_corner_from_center(cen, self.rs, self.w, self.h)], 0)
        xs, ts = crop(image, cs), crop(target, ct_)
        s_rec, t_rec = self.dec(sp, tp)
        # (1)
        self.opt_dis.zero_grad()
        s_dis, t_dis = [torch.sigmoid(o) for o in self.dis(s_rec, t_rec)]
        s_real, t_real = [torch.sigmoid(o) for o in self.dis(xs, ts)]
        one, zero = _soft(1, s_real[:1]), _soft(0, s_dis[:1])
        t_pro = self.dis_patch(tp)
        t_mean = t_pro.mean(1)
        s_pro = self.dis_patch(sp)
        ad = 0.0
        for c in range(self.k):
            ad = ad + bce(s_dis[c:c + 1], one) + bce(s_real[c:c + 1], zero)
            ad = ad + t_mean[c] * bce(t_dis[c:c + 1], zero) + bce(t_real[c:c + 1], one)
        ad.backward(retain_graph=True)
        self.opt_dis.step()
        # (2)
        self.opt_patch.zero_grad()
        pl = bce(t_pro, _soft(0, t_pro)) + bce(s_pro, _soft(1, s_pro))
        pl.backward(retain_graph=True)
        self.opt_patch.step()
        # (3)
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
        self.opt_dec.step()
        # (4)
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
        self.opt.step()
        return float(loss.item())
