"""Training / validation driver on scda_b200.engine.SCDATrainer — the caller of the hot path
(mirrors tools/faster_rcnn_train_val.py:276-408 main, :461-770 train, :773-884 validate of the reference;
same option names, same checkpoint keys, same results-file format).

What differs from the reference's loop, on purpose:
  * one process per GPU under `torchrun` (RANK / WORLD_SIZE / MASTER_* from the environment) instead of SLURM
    variables; `--port` is accepted and ignored when the environment already names the rendezvous;
  * the four-phase update is `SCDATrainer.iteration` (one replayed CUDA graph per input shape), so the
    ground-truth tensor is padded to a FIXED number of rows (`--max_gts`, zero rows are padding for every target
    function) and images are resized to the fixed `--new_w x --new_h` the reference uses for Cityscapes;
  * the warm-up / step schedule is a number handed to `iteration(lr=...)` (utils/lr_helper.py);
  * `--synthetic N` trains on N seeded synthetic (source, target) pairs of the benchmark shape: no dataset on
    disk is needed to exercise the whole driver.

    torchrun --nproc-per-node 8 -m scda_b200.tools.faster_rcnn_train_val --config scda_b200/configs/config_512_merged.json \
        --dataset cityscapes --datadir /data/cityscapes --train_meta_file train.txt --target_meta_file foggy.txt \
        --val_meta_file val.txt --pretrained vgg16-397923af.pth --epochs 15 --warmup_epochs 1 --step_epochs 8,11
"""
import argparse
import json
import logging
import os
import time

import numpy as np
import torch

logger = logging.getLogger('global')
model_zoo = ['vgg16_FasterRCNN']


def str2bool(v):
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('Boolean value expected.')


def build_parser():
    p = argparse.ArgumentParser(description='SCDA Faster R-CNN training on scda_b200')
    p.add_argument('-j', '--workers', default=4, type=int)
    p.add_argument('--epochs', default=15, type=int)
    p.add_argument('--start-epoch', default=0, type=int)
    p.add_argument('-b', '--batch-size', default=1, type=int, help='images per GPU (the SCDA path trains with 1)')
    p.add_argument('--lr', '--learning-rate', default=1.25e-5, type=float)
    p.add_argument('--weight-decay', '--wd', default=1e-4, type=float)
    p.add_argument('--print-freq', '-p', default=10, type=int)
    p.add_argument('--resume', default='', type=str)
    p.add_argument('-e', '--evaluate', dest='evaluate', action='store_true')
    p.add_argument('--pretrained', dest='pretrained', default='')
    p.add_argument('--dist', dest='dist', type=int, default=1)
    p.add_argument('--backend', dest='backend', type=str, default='nccl')
    p.add_argument('--results_dir', dest='results_dir', default='results_dir')
    p.add_argument('--port', dest='port', default=None)
    p.add_argument('--save_dir', dest='save_dir', default='checkpoints')
    p.add_argument('--warmup_epochs', dest='warmup_epochs', type=int, default=0)
    p.add_argument('--step_epochs', dest='step_epochs', type=lambda x: list(map(int, x.split(','))), default=[-1])
    p.add_argument('--config', dest='config', required=True)
    p.add_argument('--arch', dest='arch', default='vgg16_FasterRCNN', choices=model_zoo)
    p.add_argument('--dataset', dest='dataset', default='cityscapes', choices=['cityscapes'])
    p.add_argument('--datadir', dest='datadir', default='')
    p.add_argument('--train_meta_file', dest='train_meta_file', default='')
    p.add_argument('--val_meta_file', dest='val_meta_file', default='')
    p.add_argument('--target_meta_file', dest='target_meta_file', default='')
    p.add_argument('--eval_interval', dest='eval_interval', type=int, default=1)
    p.add_argument('--new_w', dest='new_w', type=int, default=1024)
    p.add_argument('--new_h', dest='new_h', type=int, default=512)
    p.add_argument('--cluster_num', dest='cluster_num', type=int, default=4)
    p.add_argument('--threshold', dest='threshold', type=int, default=128)
    p.add_argument('--recon_size', dest='recon_size', type=int, default=256)
    # not in the reference
    p.add_argument('--max_gts', type=int, default=64, help='rows of the padded ground-truth tensor')
    p.add_argument('--synthetic', type=int, default=0, help='train on N synthetic pairs per epoch (no dataset)')
    p.add_argument('--iters', type=int, default=0, help='stop an epoch after this many iterations (0 = all)')
    p.add_argument('--no_graphs', action='store_true')
    p.add_argument('--seed', type=int, default=0)
    return p


def load_config(config_path):
    """sections inherit the `shared` keys (tools/faster_rcnn_train_val.py:186-192)"""
    assert os.path.exists(config_path), config_path
    cfg = json.load(open(config_path, 'r'))
    for key in cfg.keys():
        if key != 'shared':
            cfg[key].update(cfg['shared'])
    return cfg


def pad_gts(gts, rows):
    """[1, G, 5] -> [1, rows, 5] (zero rows = padding); more than `rows` boxes keep the first `rows`"""
    out = torch.zeros(gts.shape[0], rows, 5, dtype=torch.float32)
    n = min(rows, gts.shape[1])
    out[:, :n] = gts[:, :n, :5]
    return out


def build_loaders(args, rank, world):
    from ..datasets.example_dataset import ExampleDataset, ExampleTransform, SyntheticPairs, TargetDataset, collate
    if args.synthetic > 0:
        return SyntheticPairs(args.synthetic, args.new_h, args.new_w, seed=args.seed * 1000 + rank), None
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler
    keep = lambda batch: batch                      # the pixels are finished on the device by `collate`
    size = min(args.new_h, args.new_w)

    def loader(ds, shuffle):
        sampler = DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=shuffle) if world > 1 else None
        return DataLoader(ds, batch_size=args.batch_size, shuffle=shuffle and sampler is None, sampler=sampler,
                          num_workers=args.workers, collate_fn=keep, drop_last=shuffle)
    train = ExampleDataset(args.datadir, args.train_meta_file, ExampleTransform([size], max(args.new_h, args.new_w),
                                                                                flip=True))
    target = TargetDataset(args.datadir, args.target_meta_file, new_w=args.new_w, new_h=args.new_h)
    val = None
    if args.val_meta_file:
        val = loader(ExampleDataset(args.datadir, args.val_meta_file,
                                    ExampleTransform([size], max(args.new_h, args.new_w), flip=False)), False)

    def pairs():
        from ..datasets.example_dataset import prepare_image
        tl = iter(loader(target, True))
        for batch in loader(train, True):
            images, info, gts, _, _ = collate(batch)
            try:
                tb = next(tl)
            except StopIteration:
                tl = iter(loader(target, True))
                tb = next(tl)
            tgt = torch.cat([prepare_image(px, nh, nw) for px, (nh, nw) in tb], 0)
            yield images, info, gts, tgt
    return pairs, val


def lr_at(args, epoch, it, iters_per_epoch, world):
    """warm-up: the rate grows by world * batch over the warm-up epochs, once per iteration (:356-363);
    afterwards MultiStepLR(gamma 0.1) on the enlarged rate (:372-385)"""
    from ..utils.lr_helper import multistep_lr, warmup_gamma
    warm_iters = args.warmup_epochs * iters_per_epoch
    factor = float(world * args.batch_size)
    if epoch < args.warmup_epochs and warm_iters > 1:
        return args.lr * warmup_gamma(world, args.batch_size, warm_iters) ** (epoch * iters_per_epoch + it)
    base = args.lr * (factor if args.warmup_epochs > 0 else 1.0)
    return multistep_lr(base, args.step_epochs, epoch - args.warmup_epochs)


def train(args, cfg, tr, data, epoch, world, rank):
    tr.train_mode()
    n = args.synthetic if args.synthetic > 0 else None
    it, t0, hist = 0, time.time(), []
    source = data if args.synthetic > 0 else data()
    iters_per_epoch = n or 1000
    for image, info, gts, target in source:
        lr = lr_at(args, epoch, it, iters_per_epoch, world)
        out = tr.iteration(cfg, image, info, pad_gts(gts, args.max_gts), target, lr=lr)
        it += 1
        if it % args.print_freq == 0 or (args.iters and it == args.iters):
            vals = {k: float(v) for k, v in out.items()}           # the only host synchronisation of the loop
            hist.append(vals)
            if rank == 0:
                logger.info('Epoch: [%d][%d] %.1f img/s lr %.3g  loss %.4f (rpn %.3f/%.3f rcnn %.3f/%.3f) dis %.3f '
                            'patch %.3f dec %.3f fake %.3f  acc %.1f/%.1f', epoch, it,
                            world * args.print_freq / max(time.time() - t0, 1e-9), lr, vals['loss'], vals['rpn_cls'],
                            vals['rpn_loc'], vals['rcnn_cls'], vals['rcnn_loc'], vals['dis_loss'],
                            vals['dis_patch_loss'], vals['dec_loss'], vals['fake_loss'], vals['rpn_acc'],
                            vals['rcnn_acc'])
            t0 = time.time()
        if args.iters and it >= args.iters:
            break
    return hist


def validate(args, cfg, model, val_loader, rank, world):
    """detections of every validation image -> results.txt.rank<r> (`img x1 y1 x2 y2 score cls`, coordinates in
    the ORIGINAL image, :826-851), RPN recall, then Cal_MAP on rank 0"""
    from ..datasets.example_dataset import collate
    from ..utils import bbox_helper
    from ..utils.cal_mAP import Cal_MAP
    model.eval()
    os.makedirs(args.results_dir, exist_ok=True)
    total_rc = total_gt = 0
    with open(os.path.join(args.results_dir, 'results.txt.rank%d' % rank), 'w') as fout, torch.no_grad():
        for batch in val_loader:
            images, info, gts, _, names = collate(batch)
            x = {'cfg': cfg, 'image': images, 'image_info': info, 'ground_truth_bboxes': gts, 'ignore_regions': None}
            proposals, bboxes = model(x)['predict']
            proposals, bboxes, gts_np, info_np = (t.cpu().numpy() for t in (proposals, bboxes, gts, info))
            for b in range(images.shape[0]):
                img_id = names[b].rsplit('/', 1)[-1].rsplit('.', 1)[0]
                scale = info_np[b, -1]
                rc, ng = bbox_helper.compute_recall(proposals[proposals[:, 0] == b][:, 1:5], gts_np[b])
                total_rc, total_gt = total_rc + rc, total_gt + ng
                dts = bboxes[bboxes[:, 0] == b]
                dts = dts[dts[:, -2].argsort()[::-1][:100]]
                for c in range(1, cfg['shared']['num_classes']):
                    d = bbox_helper.clip_bbox(dts[dts[:, -1] == c][:, 1:-1], info_np[b, :2])
                    if len(d) > 0:
                        d[:, :4] = d[:, :4] / scale
                    for bx in d:
                        fout.write('{0} {1} {2}\n'.format(img_id, ' '.join(map(str, bx)), c))
    recall = total_rc / max(total_gt, 1)
    logger.info('rpn300 recall=%f', recall)
    if world > 1:
        torch.distributed.barrier()
    if rank == 0 and args.val_meta_file:
        Cal_MAP(args.results_dir, args.val_meta_file, int(cfg['shared']['num_classes']))
    return recall


def save_checkpoint(args, tr, epoch, best_recall):
    """the reference's keys (:401-408) for the detector + the three reconstruction networks beside it"""
    os.makedirs(args.save_dir, exist_ok=True)
    path = os.path.join(args.save_dir, 'checkpoint_e%d.pth' % (epoch + 1))
    torch.save({'epoch': epoch + 1, 'arch': args.arch, 'state_dict': tr.model.state_dict(), 'best_recall': best_recall,
                'optimizer': {}, 'dec_state_dict': tr.dec_model.state_dict(), 'dis_state_dict': tr.dis_model.state_dict(),
                'dis_patch_state_dict': tr.dis_model_patch.state_dict(),
                'adam_steps': [o.t for o in (tr.opt, tr.opt_dec, tr.opt_dis, tr.opt_dis_patch)]}, path)
    return path


def main(argv=None):
    args = build_parser().parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='[%(asctime)s %(levelname)s] %(message)s')
    from ..engine import build_trainer
    from ..utils.distributed_utils import broadcast_params, dist_init
    from ..utils.load_helper import load_pretrain, restore_from
    rank, world = 0, 1
    if args.dist and int(os.environ.get('WORLD_SIZE', '1')) > 1:
        rank, world = dist_init(args.port, backend=args.backend)
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    cfg = load_config(args.config)
    tr = build_trainer(cfg, lr=args.lr, cluster_num=args.cluster_num, threshold=args.threshold,
                       recon_size=args.recon_size, new_w=args.new_w, new_h=args.new_h, world_size=world,
                       seed=args.seed, use_graphs=not args.no_graphs)
    best_recall = 0.0
    if args.pretrained:
        load_pretrain(tr.model, args.pretrained)
    if args.resume:
        _, _, args.start_epoch, best_recall, _ = restore_from(tr.model, None, args.resume)
        extra = torch.load(args.resume, map_location='cpu', weights_only=False)
        for key, net in (('dec_state_dict', tr.dec_model), ('dis_state_dict', tr.dis_model),
                         ('dis_patch_state_dict', tr.dis_model_patch)):
            if key in extra:
                net.load_state_dict(extra[key])
    if world > 1:
        for net in tr.nets():
            broadcast_params(net)
    train_data, val_loader = build_loaders(args, rank, world)
    if args.evaluate:
        assert val_loader is not None, '--evaluate needs --val_meta_file'
        return validate(args, cfg, tr.model, val_loader, rank, world)
    history = []
    for epoch in range(args.start_epoch, args.epochs):
        history += train(args, cfg, tr, train_data, epoch, world, rank)
        if val_loader is not None and (epoch + 1) % args.eval_interval == 0:
            best_recall = max(best_recall, validate(args, cfg, tr.model, val_loader, rank, world))
        if rank == 0:
            logger.info('saved %s', save_checkpoint(args, tr, epoch, best_recall))
    if world > 1:
        tr.close()          # captured NCCL kernels pin their communicators: graphs first, then the groups
    return history


if __name__ == '__main__':
    main()
