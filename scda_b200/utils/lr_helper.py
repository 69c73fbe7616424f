"""Learning-rate schedules of the reference driver as plain functions of the iteration / epoch
(utils/lr_helper.py:32-50 IterExponentialLR for the warm-up, torch's MultiStepLR(gamma=0.1) afterwards,
tools/faster_rcnn_train_val.py:356-385).  engine.FlatAdam takes the step size per iteration, so a schedule
is just a number handed to SCDATrainer.iteration(lr=...)."""


class IterExponentialLR(object):
    """lr(i) = base_lr * gamma ** i, stepped once per iteration"""

    def __init__(self, base_lr, gamma, last_iter=-1):
        self.base_lr, self.gamma, self.last_iter = float(base_lr), float(gamma), last_iter
        self.step()

    def step(self, it=None):
        self.last_iter = self.last_iter + 1 if it is None else it
        return self.get_lr()

    def get_lr(self):
        return self.base_lr * self.gamma ** self.last_iter


def warmup_gamma(world_size, batch_size, warmup_iters):
    """the reference enlarges the rate by world_size * batch_size over the warm-up iterations (:357-363)"""
    assert warmup_iters > 1
    return float(world_size * batch_size) ** (1.0 / (warmup_iters - 1))


def multistep_lr(base_lr, milestones, epoch, gamma=0.1):
    """torch.optim.lr_scheduler.MultiStepLR after `epoch` scheduler steps"""
    return float(base_lr) * gamma ** sum(1 for m in milestones if 0 <= m <= epoch)
