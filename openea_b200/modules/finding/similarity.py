"""Pairwise similarity and CSLS with the reference's signatures (modules/finding/similarity.py); the
arithmetic runs on the GPU (openea_b200.finding → liboea.so).  NumPy in → NumPy out, as in the reference."""
import numpy as np

from openea_b200 import finding as _f
from openea_b200.modules.utils.util import task_divide


def sim(embed1, embed2, metric='inner', normalize=False, csls_k=0):
    """n1×n2 float32 similarity matrix: 'inner', 'cosine', 'euclidean' (1 − distance), 'manhattan' (1 − L1);
    optional row normalisation first and CSLS rescaling (2·S − r_i − c_j, means of the csls_k nearest)."""
    return _f.sim(embed1, embed2, metric=metric, normalize=normalize, csls_k=csls_k).cpu().numpy()


def csls_sim(sim_mat, k):
    """CSLS of an existing similarity matrix."""
    import torch
    s = torch.as_tensor(np.asarray(sim_mat, dtype=np.float32)).cuda()
    r = calculate_nearest_k(s, k, _device=True)
    c = calculate_nearest_k(s.t().contiguous(), k, _device=True)
    return ((2 * s - r[:, None]) - c[None, :]).cpu().numpy()


def calculate_nearest_k(sim_mat, k, _device=False):
    """Mean of the k largest entries of every row."""
    import ctypes as C
    import torch
    from openea_b200 import lib as L
    from openea_b200.engine import _ptr, _stream_ptr
    s = sim_mat if isinstance(sim_mat, torch.Tensor) else torch.as_tensor(np.asarray(sim_mat, dtype=np.float32)).cuda()
    s = s.contiguous()
    n, m = s.shape
    idx = torch.empty(n, k, dtype=torch.int32, device=s.device)
    lib = L.load()
    L.check(lib.oea_rows_select_topk(_ptr(s), s.stride(0), n, m, k, None, _ptr(idx), _stream_ptr()), "oea_rows_select_topk")
    mean = torch.gather(s, 1, idx.long()).mean(dim=1)
    return mean if _device else mean.cpu().numpy()


def csls_sim_multi_threads(sim_mat, k, nums_threads):
    return calculate_nearest_k(sim_mat, k)


def sim_multi_threads(embeds1, embeds2, threads_num=16):
    return sim(embeds1, embeds2, metric='inner')


def sim_multi_blocks(embeds1, embeds2, blocks_num=16):
    return sim(embeds1, embeds2, metric='inner')
