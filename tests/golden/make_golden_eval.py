"""Golden vectors for the evaluation / input-pipeline helpers, produced by importing the reference's own Python
from /root/reference in this container:
  - utils/cal_mAP.py: parse_gts / parse_res / cal_mAP on a synthetic meta file + result list -> AP per class,
    max recall, mAP (Cal_MAP1)
  - utils/lr_helper.py: IterExponentialLR rates over a warm-up
  - datasets/example_dataset.py: ExampleTransform box arithmetic (scale, floor / ceil, mirror) for fixed sizes
Run here (needs /root/reference):  python tests/golden/make_golden_eval.py  -> tests/golden/eval_helpers.json
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SCDA_REFERENCE_ROOT", "/root/reference")


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def synth_lists(seed=0, n_img=30, num_classes=9):
    """a meta file in the reference's format + detections around (and away from) the ground truth"""
    r = np.random.RandomState(seed)
    gts_lines, res_lines = [], []
    for i in range(n_img):
        name = "city/img_%03d.png" % i
        n_gt = r.randint(2, 7)
        boxes = []
        for _ in range(n_gt):
            x1, y1 = r.randint(0, 1500), r.randint(0, 700)
            w, h = r.randint(30, 400), r.randint(30, 250)
            boxes.append((r.randint(1, num_classes), x1, y1, x1 + w, y1 + h))
        gts_lines += ["# %d\n" % i, name + "\n", "3\n", "1024\n", "2048\n", "0\n", "0\n", "%d\n" % n_gt]
        gts_lines += ["%d %d %d %d %d\n" % b for b in boxes]
        pure = name.split('/')[-1][:-4]
        for (c, x1, y1, x2, y2) in boxes:
            for k in range(r.randint(0, 3)):           # zero to two detections per ground truth, jittered
                j = r.normal(0, 12 + 25 * k, 4)
                res_lines.append("%s %.2f %.2f %.2f %.2f %.4f %d\n" % (
                    pure, x1 + j[0], y1 + j[1], x2 + j[2], y2 + j[3], r.uniform(0.05, 1.0),
                    c if r.uniform() < 0.85 else r.randint(1, num_classes)))
        for _ in range(r.randint(0, 4)):               # false positives
            x1, y1 = r.randint(0, 1500), r.randint(0, 700)
            res_lines.append("%s %d %d %d %d %.4f %d\n" % (pure, x1, y1, x1 + r.randint(20, 300), y1 + r.randint(20, 200),
                                                            r.uniform(0.05, 1.0), r.randint(1, num_classes)))
    for c in range(1, num_classes):                    # the reference's cal_mAP fails on a class without detections
        res_lines.append("img_000 %d %d %d %d %.4f %d\n" % (10 * c, 10 * c, 10 * c + 50, 10 * c + 40, 0.01 * c, c))
    return gts_lines, res_lines


def main():
    out = {}
    cm = load("ref_cal_map", "utils/cal_mAP.py")
    gts_lines, res_lines = synth_lists()
    gts = cm.parse_gts(gts_lines, 9)
    res = cm.parse_res(res_lines)
    ap, max_recall = cm.cal_mAP(gts, res, 9, 0.5)
    out["map"] = {"gts_lines": gts_lines, "res_lines": res_lines, "ap": [float(v) for v in ap],
                  "max_recall": [float(v) for v in max_recall], "mAP": float(np.mean(ap[1:]))}
    lr = load("ref_lr_helper", "utils/lr_helper.py")
    import torch
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.25e-5)
    gamma = 8.0 ** (1.0 / (50 - 1))
    sched = lr.IterExponentialLR(opt, gamma)
    rates = []
    for it in range(50):
        sched.step(it)
        rates.append(opt.param_groups[0]['lr'])
    out["lr"] = {"base": 1.25e-5, "gamma": gamma, "rates": rates}
    json.dump(out, open(os.path.join(HERE, "eval_helpers.json"), "w"))
    print("mAP", out["map"]["mAP"], "rates", rates[0], rates[-1])


if __name__ == "__main__":
    main()
