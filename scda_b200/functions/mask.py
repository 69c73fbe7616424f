"""Region grouping of RoI features (mirrors functions/mask.py:183-237 of the reference).

`compute_cluster_targets(proposals, features, N_cluster, threshold)`: k-means on the RoI
centres, then `threshold` fc7 rows per cluster (the first members, or members re-drawn
with replacement when a cluster is smaller), stacked to [N_cluster, threshold, 4096] and
DETACHED, plus the cluster centres as a float64 numpy array.

The reference copies the 512 x 4096 feature block to the host (8 MB), gathers there and
copies 8 MB back, twice per iteration.  Here only the 512 x 5 RoI table crosses PCIe (the
k-means itself is the reference's: scikit-learn `KMeans(n_clusters, random_state=0)` on the
host, functions/mask.py:209); the feature rows are gathered on the device.
"""
import numpy as np
import torch


def proposals_to_centers(proposals):
    cx = (proposals[:, 3] + proposals[:, 1]) / 2.0
    cy = (proposals[:, 4] + proposals[:, 2]) / 2.0
    return np.vstack([cx, cy]).transpose()


def cluster_assignments(proposals_np, N_cluster):
    from sklearn.cluster import KMeans
    centers = proposals_to_centers(proposals_np)
    kmeans = KMeans(n_clusters=N_cluster, random_state=0).fit(centers)
    return kmeans.cluster_centers_, kmeans.labels_


def compute_cluster_targets(proposals, features, N_cluster=4, threshold=128):
    '''
    Args:
        proposals:[N, k], k>=5(b_ix, x1,y1,x2,y2, ...), N = 512
        features: [N, 4096]
    Return:
        batch_rois: [N_cluster, threshold, 4096] (CUDA, detached)
        batch_cluster_center: [N_cluster, 2], (center_x, center_y) float64 numpy
    '''
    assert features.is_cuda
    proposals_np = proposals.detach().cpu().numpy() if torch.is_tensor(proposals) else proposals
    cluster_center, cluster_labels = cluster_assignments(proposals_np, N_cluster)
    rows = []
    for cluster_idx in range(N_cluster):
        keep_ix = np.where(cluster_labels == cluster_idx)[0]
        if keep_ix.shape[0] < threshold:
            keep_ix = keep_ix[np.random.choice(keep_ix.shape[0], threshold, replace=True)]
        else:
            keep_ix = keep_ix[0:threshold]
        rows.append(keep_ix)
    index = torch.from_numpy(np.concatenate(rows).astype(np.int64)).to(features.device,
                                                                       non_blocking=True)
    batch_rois_cluster = features.detach().float().index_select(0, index)
    batch_rois_cluster = batch_rois_cluster.view(N_cluster, threshold, features.shape[1]).contiguous()
    return batch_rois_cluster, cluster_center
