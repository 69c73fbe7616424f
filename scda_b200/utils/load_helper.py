"""Checkpoint helpers (mirror utils/load_helper.py:6-54 of the reference): `load_pretrain` loads a
state dict BY NAME with strict=False after stripping the 'module.' prefix of DataParallel-era files
(ImageNet `vgg16-397923af.pth`: only `features.*` match the detector, the classifier / RPN / RCNN
heads keep their initialisation); `restore_from` resumes a training checkpoint
(tools/faster_rcnn_train_val.py:401-408 writes {'epoch', 'arch', 'state_dict', 'best_recall',
'optimizer'}).  Parameters of a network owned by engine.FlatAdam live in a flat buffer:
`load_state_dict` copies into those views in place, and the bf16 shadows are re-derived."""
import logging

import torch

logger = logging.getLogger('global')


def check_keys(model, pretrained_state_dict):
    ckpt_keys = set(pretrained_state_dict.keys())
    model_keys = set(model.state_dict().keys())
    used = model_keys & ckpt_keys
    logger.info('missing keys:{}'.format(len(model_keys - ckpt_keys)))
    logger.info('unused checkpoint keys:{}'.format(len(ckpt_keys - model_keys)))
    logger.info('used keys:{}'.format(len(used)))
    assert len(used) > 0, 'load NONE from pretrained checkpoint'
    return True


def remove_prefix(state_dict, prefix):
    '''old style checkpoints store every parameter name with the common prefix "module."'''
    f = lambda x: x.split(prefix, 1)[-1] if x.startswith(prefix) else x
    return {f(key): value for key, value in state_dict.items()}


def _touch(model):
    """parameters were overwritten in place: invalidate the derived operand copies (bf16 shadows, split /
    phase-decomposed weights) keyed on the optimiser epoch"""
    for p in model.parameters():
        p._scda_epoch = getattr(p, "_scda_epoch", 0) + 1
        if hasattr(p, "_scda_shadow_version"):
            p._scda_shadow_version = -1


def load_pretrain(model, pretrained_path, map_location=None):
    logger.info('load pretrained model from {}'.format(pretrained_path))
    if map_location is None:
        map_location = next(model.parameters()).device
    pretrained = torch.load(pretrained_path, map_location=map_location, weights_only=False)
    if isinstance(pretrained, dict) and 'state_dict' in pretrained:
        pretrained = pretrained['state_dict']
    pretrained = remove_prefix(pretrained, 'module.')
    check_keys(model, pretrained)
    model.load_state_dict(pretrained, strict=False)
    _touch(model)
    return model


def restore_from(model, optimizer, ckpt_path, map_location=None):
    """-> (model, optimizer (None, as the reference: the optimiser state is not restored), epoch,
    best_recall, arch)"""
    logger.info('restore from {}'.format(ckpt_path))
    if map_location is None:
        map_location = next(model.parameters()).device
    ckpt = torch.load(ckpt_path, map_location=map_location, weights_only=False)
    state = remove_prefix(ckpt['state_dict'], 'module.')
    check_keys(model, state)
    model.load_state_dict(state, strict=False)
    _touch(model)
    return model, None, ckpt['epoch'], ckpt['best_recall'], ckpt['arch']
