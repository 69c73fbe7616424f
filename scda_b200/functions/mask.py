"""Region grouping of RoI features (mirrors functions/mask.py:183-237 of the reference).

`compute_cluster_targets(proposals, features, N_cluster, threshold)`: k-means on the RoI
centres, then `threshold` fc7 rows per cluster (the first members, or members re-drawn with
replacement when a cluster is smaller), stacked to [N_cluster, threshold, 4096] and
DETACHED, plus the cluster centres.

The reference copies the RoIs and the 512 x 4096 feature block to the host (8 MB), runs
scikit-learn `KMeans(n_clusters, random_state=0)` there (functions/mask.py:209), gathers on
the host and copies 8 MB back — twice per iteration, each a device synchronisation.  Here
the clustering and the member selection are one device kernel (csrc/kmeans.cu, scikit-learn's
k-means++ / Lloyd restated for float32 data) and the rows are gathered in HBM:
`cluster_targets_device` never touches the host; `compute_cluster_targets` keeps the
reference's return type (centres as a numpy array) and therefore ends with one small D2H.
"""
import functools

import numpy as np
import torch

from .._lib import check, load, require_cuda, stream_ptr


def proposals_to_centers(proposals):
    cx = (proposals[:, 3] + proposals[:, 1]) / 2.0
    cy = (proposals[:, 4] + proposals[:, 2]) / 2.0
    return np.vstack([cx, cy]).transpose()


@functools.lru_cache(maxsize=None)
def kmeanspp_draws(n, k):
    """The random numbers `KMeans(random_state=0)` consumes for n samples and k clusters:
    check_random_state(0) = RandomState(0); k-means++ draws the first centre with
    `choice(n, p=w / w.sum())` (w = float32 ones) and then `uniform(size=2 + int(log k))`
    once per further centre — none of it depends on the data."""
    rs = np.random.RandomState(0)
    w = np.ones(n, dtype=np.float32)
    first = int(rs.choice(n, p=w / w.sum()))
    trials = 2 + int(np.log(k))
    uni = np.stack([rs.uniform(size=trials) for _ in range(1, k)]) if k > 1 else np.zeros((0, trials))
    return first, trials, uni.astype(np.float64)


_DRAWS_DEV = {}


def _draws_on(device, n, k):
    key = (str(device), n, k)
    if key not in _DRAWS_DEV:
        first, trials, uni = kmeanspp_draws(n, k)
        _DRAWS_DEV[key] = (first, trials, torch.from_numpy(uni.reshape(-1).copy()).to(device))
    return _DRAWS_DEV[key]


def kmeans_regions_device(proposals, N_cluster, threshold, max_iter=300, tol=1e-4, pick_uniform=None):
    """proposals [n, >=5] fp32 CUDA -> (labels int32 [n], centers fp32 [K, 2], counts int32
    [K], index int64 [K * threshold]); everything stays on the device."""
    require_cuda(proposals)
    assert proposals.dtype == torch.float32 and proposals.dim() == 2 and proposals.shape[1] >= 5
    rois = proposals if proposals.stride(1) == 1 else proposals.contiguous()
    n, dev = rois.shape[0], rois.device
    first, trials, uni = _draws_on(dev, n, N_cluster)
    labels = torch.empty(n, dtype=torch.int32, device=dev)
    centers = torch.empty(N_cluster, 2, dtype=torch.float32, device=dev)
    counts = torch.empty(N_cluster, dtype=torch.int32, device=dev)
    index = torch.empty(N_cluster * threshold, dtype=torch.int64, device=dev)
    if pick_uniform is None:
        pick_uniform = torch.rand(N_cluster * threshold, device=dev)
    lib = load()
    ws_bytes = lib.scda_kmeans_workspace_bytes(n, N_cluster)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.scda_kmeans_regions(rois.data_ptr(), rois.stride(0), n, N_cluster, first, uni.data_ptr(),
                                      trials, max_iter, tol, pick_uniform.data_ptr(), threshold,
                                      labels.data_ptr(), centers.data_ptr(), counts.data_ptr(),
                                      index.data_ptr(), ws.data_ptr(), ws_bytes, stream_ptr(dev)),
              "scda_kmeans_regions")
    return labels, centers, counts, index


def kmeans_regions_forked(proposals, N_cluster, threshold, stream):
    """The k-means of `cluster_targets_device` issued on `stream`, forked from the current stream: the
    clustering reads the RoI table only (functions/mask.py:196-209 of the reference clusters the box
    CENTRES), so it can run beside RoIPool -> fc6 -> fc7 instead of behind them (0.13-0.19 ms of a single
    CTA, profiles/r2_timeline_b_forward.txt).  -> a handle for `cluster_targets_device(..., pre=handle)`,
    which joins the stream."""
    cur = torch.cuda.current_stream()
    rois = proposals.detach().float()
    stream.wait_stream(cur)
    with torch.cuda.stream(stream):
        res = kmeans_regions_device(rois, N_cluster, threshold)
    rois.record_stream(stream)
    return res, stream


def cluster_targets_device(proposals, features, N_cluster=4, threshold=128, taps=None, pre=None):
    """-> batch_rois [N_cluster, threshold, F] (detached) and centres as a DEVICE fp32 [K, 2];
    `taps` (a dict, tests only) receives the row index and the labels the gather used;
    `pre`: the clustering already issued by `kmeans_regions_forked` on the same proposals"""
    assert features.is_cuda
    if pre is not None:
        (labels, centers, _, index), stream = pre
        cur = torch.cuda.current_stream()
        cur.wait_stream(stream)
        for t in (labels, centers, index):
            t.record_stream(cur)
    else:
        labels, centers, _, index = kmeans_regions_device(proposals.detach().float(), N_cluster, threshold)
    rows = features.detach().float().index_select(0, index)
    if taps is not None:
        taps.update(index=index, labels=labels, centers=centers)
    return rows.view(N_cluster, threshold, features.shape[1]), centers


def compute_cluster_targets(proposals, features, N_cluster=4, threshold=128):
    '''
    Args:
        proposals:[N, k], k>=5(b_ix, x1,y1,x2,y2, ...), N = 512
        features: [N, 4096]
    Return:
        batch_rois: [N_cluster, threshold, 4096] (CUDA, detached)
        batch_cluster_center: [N_cluster, 2], (center_x, center_y) numpy (host)
    '''
    if not torch.is_tensor(proposals):
        proposals = torch.from_numpy(np.asarray(proposals, dtype=np.float32)).to(features.device)
    batch_rois_cluster, centers = cluster_targets_device(proposals, features, N_cluster, threshold)
    return batch_rois_cluster.contiguous(), centers.cpu().numpy()
