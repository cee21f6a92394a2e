"""f-4 on the device: Gale–Shapley over K3's top-`cut` lists (oea_gale_shapley) against the host restatement of
modules/finding/alignment.py:171-224 (galeshapley) fed with the full argsort lists of the same similarity matrix —
both list routes (fused sorted top-k for cut <= 32, materialise + radix select + gather-sort above), with and without
CSLS, rectangular problems, and the reference's `stable_alignment` result line."""
import contextlib
import io
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _host_matching(s, cut):
    from openea_b200.modules.finding.alignment import arg_sort, galeshapley
    n1, n2 = s.shape
    m = galeshapley(arg_sort(list(range(n1)), s, "x_", "y_"), arg_sort(list(range(n2)), s.T, "y_", "x_"), cut)
    out = -np.ones(n1, dtype=np.int64)
    for x, y in m.items():
        out[int(x[2:])] = int(y[2:])
    return out


@pytest.mark.parametrize("n1,n2,cut,csls", [(300, 300, 100, 0), (257, 311, 20, 0), (311, 257, 100, 5), (200, 200, 7, 10),
                                            (180, 180, 3, 0)])
def test_device_gale_shapley_equals_host(cuda_device, n1, n2, cut, csls):
    from openea_b200 import finding as F
    rng = np.random.default_rng(n1 + cut)
    # small-integer embeddings: every inner product is exact in fp32 whatever the summation order, so the fused top-k
    # kernel, the stored matrix and NumPy see the SAME numbers (and exact ties exercise the tie rules: lower column /
    # lower suitor index first); with CSLS the offsets are means of k such integers, still exactly representable sums / k
    e2 = rng.integers(-3, 4, (n2, 40)).astype(np.float32)
    base = e2[rng.integers(0, n2, n1)] if n1 != n2 else e2
    e1 = (base + rng.integers(-2, 3, (n1, 40))).astype(np.float32)             # noisy: many contested reviewers
    s = F.sim(e1, e2, "inner", False, csls)
    s = s.cpu().numpy() if hasattr(s, "cpu") else np.asarray(s)
    want = _host_matching(s, cut)
    match, rounds = F.stable_matching(e1, e2, "inner", False, csls, cut)
    got = match.cpu().numpy().astype(np.int64)
    assert 1 <= rounds <= cut
    assert np.array_equal(got, want), (int((got != want).sum()), rounds)
    held = got[got >= 0]
    assert len(np.unique(held)) == len(held), "a reviewer is held by one suitor"


def test_stable_alignment_prints_the_reference_line(cuda_device):
    from openea_b200.modules.finding.alignment import stable_alignment
    rng = np.random.default_rng(3)
    e2 = rng.standard_normal((400, 32)).astype(np.float32)
    e1 = (e2 + 0.3 * rng.standard_normal((400, 32))).astype(np.float32)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        assert stable_alignment(e1, e2, "inner", True, 0, 4) is None
    text = buf.getvalue()
    assert "generating candidate lists costs time" in text
    prec = float(re.findall(r"stable alignment precision = ([0-9.]+)%", text)[0])
    assert prec > 80.0, text
