"""Device k-means of the RoI centres (csrc/kmeans.cu) against scikit-learn — the library the
reference calls (functions/mask.py:209, `KMeans(n_clusters, random_state=0)`), run here on
the host as the oracle.  scikit-learn's float32 path goes through BLAS reductions whose
summation order is unspecified, so agreement is exact except at float32 rounding knife
edges: the test pins (a) identical labels on at least 90 % of the seeded cases and (b) for
every case a valid k-means fixed point whose inertia is within 1e-3 of scikit-learn's."""
import numpy as np
import pytest

import _inputs

gpu = pytest.mark.gpu


def _sk(rois, k):
    from sklearn.cluster import KMeans
    c = np.vstack([(rois[:, 3] + rois[:, 1]) / 2.0, (rois[:, 4] + rois[:, 2]) / 2.0]).transpose()
    assert c.dtype == np.float32
    km = KMeans(n_clusters=k, random_state=0).fit(c)
    return c, km.labels_, km.cluster_centers_, km.inertia_


def _inertia(c, labels, centers):
    return float(((c.astype(np.float64) - centers.astype(np.float64)[labels]) ** 2).sum())


@gpu
@pytest.mark.parametrize("n,k", [(512, 4), (512, 2), (512, 8), (300, 4), (64, 4)])
def test_kmeans_matches_sklearn(cuda_lib, n, k):
    import torch
    from scda_b200.functions.mask import kmeans_regions_device
    same, total = 0, 0
    for seed in range(12):
        if seed % 3 == 2:      # clustered RoIs (what NMS survivors look like) besides uniform ones
            boxes = _inputs.clustered_boxes(n, seed)
            rois = np.concatenate([np.zeros((n, 1), np.float32), boxes[:, :4].astype(np.float32)], 1)
        else:
            rois = _inputs.rois_uniform(n, seed, img_w=1024, img_h=512, wh=(16, 300)).astype(np.float32)
        c, sk_labels, sk_centers, sk_inertia = _sk(rois, k)
        labels, centers, counts, index = kmeans_regions_device(torch.from_numpy(rois).cuda(), k, 128)
        labels, centers, counts = labels.cpu().numpy(), centers.cpu().numpy(), counts.cpu().numpy()
        assert labels.min() >= 0 and labels.max() < k
        assert np.array_equal(counts, np.bincount(labels, minlength=k))
        # a Lloyd fixed point: every point carries the label of its nearest centre
        d = ((c[:, None, :].astype(np.float64) - centers[None].astype(np.float64)) ** 2).sum(-1)
        near = d.argmin(1)
        gap = np.sort(d, 1)
        ok = (near == labels) | (gap[:, 1] - gap[:, 0] < 1e-2)
        assert ok.all()
        total += 1
        if np.array_equal(labels, sk_labels):
            same += 1
            np.testing.assert_allclose(centers, sk_centers, rtol=1e-5, atol=2e-3)
            assert abs(_inertia(c, labels, centers) - sk_inertia) <= 1e-4 * sk_inertia + 1e-2
    assert same >= int(0.9 * total), "labels identical to scikit-learn in %d of %d cases" % (same, total)


@gpu
def test_member_selection(cuda_lib):
    import torch
    from scda_b200.functions.mask import kmeans_regions_device
    rois = _inputs.rois_uniform(512, 1, img_w=1024, img_h=512, wh=(16, 300)).astype(np.float32)
    pick = torch.rand(4 * 128, generator=torch.Generator().manual_seed(0)).cuda()
    labels, centers, counts, index = kmeans_regions_device(torch.from_numpy(rois).cuda(), 4, 128,
                                                           pick_uniform=pick)
    labels, counts, index = labels.cpu().numpy(), counts.cpu().numpy(), index.cpu().numpy().reshape(4, 128)
    pick = pick.cpu().numpy().reshape(4, 128)
    for c in range(4):
        members = np.where(labels == c)[0]
        if counts[c] >= 128:
            assert np.array_equal(index[c], members[:128])
        else:
            want = members[np.minimum((pick[c] * np.float32(counts[c])).astype(np.int64), counts[c] - 1)]
            assert np.array_equal(index[c], want)


def test_draws_are_the_ones_sklearn_consumes():
    """kmeanspp_draws replays RandomState(0) the way KMeans(random_state=0) consumes it: seeding
    scikit-learn's own k-means++ with that state picks the same first centre."""
    pytest.importorskip("sklearn")
    from sklearn.cluster import kmeans_plusplus
    from scda_b200.functions.mask import kmeanspp_draws
    x = np.random.RandomState(3).standard_normal((512, 2)).astype(np.float32)
    _, idx = kmeans_plusplus(x, 4, random_state=np.random.RandomState(0))
    first, trials, uni = kmeanspp_draws(512, 4)
    assert idx[0] == first and trials == 3 and uni.shape == (3, 3)
