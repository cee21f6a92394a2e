"""valid / test / early_stop with the reference's signatures (modules/finding/evaluation.py:6-33)."""
import numpy as np

from openea_b200.modules.finding.alignment import greedy_alignment


def _mapped(embeds1, mapping):
    if mapping is None:
        return embeds1
    import torch
    if isinstance(embeds1, torch.Tensor):
        return embeds1 @ torch.as_tensor(mapping, device=embeds1.device, dtype=embeds1.dtype)   # [n,d]·[d,d] (cuBLAS)
    return np.matmul(embeds1, mapping)


def valid(embeds1, embeds2, mapping, top_k, threads_num, metric='inner', normalize=False, csls_k=0, accurate=False):
    _, hits1_12, mr_12, mrr_12 = greedy_alignment(_mapped(embeds1, mapping), embeds2, top_k, threads_num,
                                                  metric, normalize, csls_k, accurate)
    return hits1_12, mrr_12


def test(embeds1, embeds2, mapping, top_k, threads_num, metric='inner', normalize=False, csls_k=0, accurate=True):
    alignment_rest_12, hits1_12, mr_12, mrr_12 = greedy_alignment(_mapped(embeds1, mapping), embeds2, top_k,
                                                                  threads_num, metric, normalize, csls_k, accurate)
    return alignment_rest_12, hits1_12, mrr_12


def early_stop(flag1, flag2, flag):
    """Stop when the validation score failed to improve twice in a row: flag <= flag2 <= flag1."""
    stop = flag <= flag2 <= flag1
    if stop:
        print("\n == should early stop == \n")
    return flag2, flag, stop
