"""Evaluation / input-pipeline helpers against vectors produced by the reference's own Python
(tests/golden/make_golden_eval.py -> eval_helpers.json): Cal_MAP arithmetic, warm-up learning rates, the meta
file parser and the box arithmetic of the dataset transform, checkpoint loading by name."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "eval_helpers.json")))


def test_cal_map_matches_reference():
    from scda_b200.utils import cal_mAP as cm
    m = G["map"]
    ap, max_recall = cm.cal_mAP(cm.parse_gts(m["gts_lines"], 9), cm.parse_res(m["res_lines"]), 9, 0.5)
    np.testing.assert_allclose(ap, m["ap"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(max_recall, m["max_recall"], rtol=0, atol=1e-12)
    assert abs(cm.Cal_MAP1(m["res_lines"], m["gts_lines"], 9) - m["mAP"]) < 1e-12


def test_cal_map_files_round_trip(tmp_path):
    from scda_b200.utils import cal_mAP as cm
    m = G["map"]
    half = len(m["res_lines"]) // 2
    (tmp_path / "results.txt.rank0").write_text("".join(m["res_lines"][:half]))
    (tmp_path / "results.txt.rank1").write_text("".join(m["res_lines"][half:]))
    meta = tmp_path / "val.txt"
    meta.write_text("".join(m["gts_lines"]))
    assert abs(cm.Cal_MAP(str(tmp_path), str(meta), 9) - m["mAP"]) < 1e-12
    assert (tmp_path / "results.txt").exists()


def test_warmup_rates_match_reference():
    from scda_b200.utils.lr_helper import IterExponentialLR, multistep_lr, warmup_gamma
    lr = G["lr"]
    assert abs(warmup_gamma(8, 1, 50) - lr["gamma"]) < 1e-15
    s = IterExponentialLR(lr["base"], lr["gamma"])
    rates = [s.step(it) for it in range(50)]
    np.testing.assert_allclose(rates, lr["rates"], rtol=1e-12)
    assert multistep_lr(1e-4, [8, 11], 7) == 1e-4
    assert abs(multistep_lr(1e-4, [8, 11], 8) - 1e-5) < 1e-18 and abs(multistep_lr(1e-4, [8, 11], 12) - 1e-6) < 1e-18


def test_meta_file_parser_and_transform(tmp_path):
    from scda_b200.datasets.example_dataset import ExampleTransform, parse_meta_file
    meta = tmp_path / "train.txt"
    meta.write_text("".join(G["map"]["gts_lines"]))
    metas = parse_meta_file(str(meta))
    assert len(metas) == 30 and metas[0][0] == "city/img_000.png" and metas[0][1] == 1024 and metas[0][2] == 2048
    n_gt = int(G["map"]["gts_lines"][7])
    assert metas[0][3].shape == (n_gt, 4) and metas[0][4].shape == (n_gt,)
    first = [float(v) for v in G["map"]["gts_lines"][8].split()]
    assert metas[0][4][0] == int(first[0]) and list(metas[0][3][0]) == first[1:]

    class Fixed(object):                         # the reference's draws: randint for the size, random() for the mirror
        def __init__(self, size, u):
            self.size, self.u = size, u

        def randint(self, lo, hi):
            assert lo <= self.size < hi
            return self.size

        def random(self):
            return self.u
    t = ExampleTransform([512], 1024, flip=True, rng=Fixed(512, 0.25))
    boxes = np.array([[100.4, 50.2, 300.7, 200.9]])
    new_w, new_h, scale, flip, nb, ni = t(2048, 1024, boxes, np.zeros((1, 4)))
    assert (new_w, new_h, scale, flip) == (1024, 512, 0.5, True)
    # scale: floor the top-left, ceil the bottom-right; mirror: x1' = w - x2, x2' = w - x1
    assert list(nb[0]) == [1024 - 151.0, 25.0, 1024 - 50.0, 101.0]
    t2 = ExampleTransform([512], 1024, flip=True, rng=Fixed(512, 0.75))
    assert list(t2(2048, 1024, boxes, np.zeros((1, 4)))[4][0]) == [50.0, 25.0, 151.0, 101.0]


def test_load_pretrain_by_name(tmp_path):
    import torch
    from scda_b200.utils.load_helper import load_pretrain, restore_from
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.ReLU(), torch.nn.Linear(3, 2))
    src = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.ReLU(), torch.nn.Linear(3, 2))
    sd = {"module." + k: v for k, v in src.state_dict().items() if k.startswith("0.")}
    sd["module.classifier.weight"] = torch.zeros(1)                 # unused key: ignored
    path = tmp_path / "pre.pth"
    torch.save(sd, str(path))
    before = net[2].weight.clone()
    load_pretrain(net, str(path), map_location="cpu")
    assert torch.equal(net[0].weight, src[0].weight) and torch.equal(net[2].weight, before)
    ck = tmp_path / "ck.pth"
    torch.save({"epoch": 3, "arch": "vgg16_FasterRCNN", "state_dict": src.state_dict(), "best_recall": 0.5,
                "optimizer": {}}, str(ck))
    _, _, epoch, best, arch = restore_from(net, None, str(ck), map_location="cpu")
    assert (epoch, best, arch) == (3, 0.5, "vgg16_FasterRCNN") and torch.equal(net[2].weight, src[2].weight)
    with pytest.raises(AssertionError):
        torch.save({"nothing.weight": torch.zeros(1)}, str(path))
        load_pretrain(net, str(path), map_location="cpu")
