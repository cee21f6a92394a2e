"""RDGCN's hard-negative search on the GPU (approaches/rdgcn.py:75-87 `get_neg`).

The reference pulls the whole output layer to the host, runs scipy `cdist(cityblock)` of the t seed rows against
all E rows in float64 (1.2·10¹² |a−b| terms and a 32 GB matrix at the 100K shape), full-argsorts every row and
keeps the k nearest (the seed itself comes first, distance 0).  Here it is one pass of the K3 tile kernel with the
L1 metric and a per-row top-k list: no matrix, no sort.
"""
import torch

from openea_b200 import finding as F


def get_neg(ILL, output_layer, k):
    """ILL: [t] entity ids (device or host); output_layer: [E, d] embeddings (CUDA tensor or array).
    Returns an int32 CUDA tensor [t·k]: for every seed the ids of its k L1-nearest entities, nearest first
    (ties: lower id first), flattened row-major exactly as the reference's `neg` list."""
    emb, d = F.to_device_rows(output_layer, False)
    ids = torch.as_tensor(ILL, dtype=torch.long, device=emb.device)
    sub = emb.index_select(0, ids).contiguous()
    res = F.topk(sub, emb, d, "manhattan", k, want=("idx",))     # similarity = 1 − L1 distance: top-k = k nearest
    return res["idx"].reshape(-1)
