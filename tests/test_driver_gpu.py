"""The training / validation driver (scda_b200/tools/faster_rcnn_train_val.py) end to end on the GPU: synthetic
pairs through SCDATrainer for 24 iterations (the detector's loss goes down, a checkpoint with the reference's
keys is written and restores), and the validation path on a tiny on-disk dataset in the reference's meta-file
format (results file written, Cal_MAP evaluated)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "scda_b200", "configs", "config_512_merged.json")


def test_train_on_synthetic_pairs_and_checkpoint(cuda_lib, tmp_path):
    import torch
    from scda_b200.tools import faster_rcnn_train_val as drv
    args = ["--config", CFG, "--synthetic", "24", "--epochs", "1", "--lr", "1e-4", "--print-freq", "1",
            "--new_w", "512", "--new_h", "256", "--save_dir", str(tmp_path / "ck"), "--max_gts", "32"]
    hist = drv.main(args)
    assert len(hist) == 24
    for h in hist:
        assert all(np.isfinite(v) for v in h.values())
    det = [h["rpn_cls"] + h["rpn_loc"] + h["rcnn_cls"] + h["rcnn_loc"] for h in hist]
    assert np.mean(det[-4:]) < 0.8 * det[0], det          # random init -> a trained-for-24-steps detector
    ck = torch.load(str(tmp_path / "ck" / "checkpoint_e1.pth"), map_location="cpu", weights_only=False)
    assert {"epoch", "arch", "state_dict", "best_recall", "optimizer"} <= set(ck)
    assert ck["epoch"] == 1 and ck["arch"] == "vgg16_FasterRCNN" and ck["adam_steps"] == [24, 24, 24, 24]
    assert any(k.startswith("features.") for k in ck["state_dict"])
    # resume: the detector and the three reconstruction networks come back
    hist2 = drv.main(args + ["--resume", str(tmp_path / "ck" / "checkpoint_e1.pth"), "--epochs", "2", "--iters", "4"])
    assert len(hist2) == 4 and np.isfinite(hist2[-1]["loss"])
    assert hist2[0]["rpn_cls"] < 0.8 * hist[0]["rpn_cls"]            # it continues from the trained weights


def test_validate_on_a_tiny_dataset(cuda_lib, tmp_path):
    from PIL import Image
    from scda_b200.tools import faster_rcnn_train_val as drv
    r = np.random.RandomState(0)
    lines = []
    for i in range(2):
        name = "img_%d.png" % i
        Image.fromarray(r.randint(0, 256, (256, 512, 3), dtype=np.uint8)).save(str(tmp_path / name))
        lines += ["# %d\n" % i, name + "\n", "3\n", "256\n", "512\n", "0\n", "0\n", "2\n",
                  "1 20 30 200 180\n", "3 250 60 480 220\n"]
    meta = tmp_path / "val.txt"
    meta.write_text("".join(lines))
    res = tmp_path / "res"
    recall = drv.main(["--config", CFG, "--evaluate", "--datadir", str(tmp_path), "--val_meta_file", str(meta),
                       "--train_meta_file", str(meta), "--target_meta_file", str(tmp_path / "t.txt"),
                       "--results_dir", str(res), "--new_w", "512", "--new_h", "256", "--workers", "0"]) \
        if (tmp_path / "t.txt").write_text("img_0.png\nimg_1.png\n") else None
    assert 0.0 <= recall <= 1.0
    out = (res / "results.txt.rank0").read_text().splitlines()
    assert (res / "results.txt").exists()
    for line in out[:5]:
        f = line.split()
        assert len(f) == 7 and f[0] in ("img_0", "img_1") and 1 <= int(f[6]) <= 8
