"""Triple losses on gathered rows: tensor-in → scalar-tensor-out, as modules/base/losses.py of the reference
(get_loss_func :4, margin_loss :15, positive_loss :30, limited_loss :42, logistic_loss :59, mapping_loss :76).

Each loss is a torch.autograd.Function whose forward AND backward run in one liboea.so kernel
(oea_loss_rows / oea_mapping_fwd_bwd); PyTorch only carries the tensors.  The fused training step of
BasicModel does not come through here (it uses oea_triple_score_sampled / _fed); these are for approaches
that assemble their own graphs.
"""
import ctypes as C

import torch

from openea_b200 import lib as L
from openea_b200.engine import _ptr, _stream_ptr, loss_cfg, pitch_for


def _padded(t, pitch):
    t = t.to(torch.float32)
    if t.shape[1] == pitch and t.is_contiguous():
        return t
    out = torch.zeros(t.shape[0], pitch, dtype=torch.float32, device=t.device)
    out[:, :t.shape[1]] = t
    return out


class _TripleLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, phs, prs, pts, nhs, nrs, nts):
        lib = L.load()
        dim = phs.shape[1]
        pitch = pitch_for(dim)
        pos = [_padded(x, pitch) for x in (phs, prs, pts)]
        has_neg = nhs is not None and nhs.shape[0] > 0
        neg = [_padded(x, pitch) for x in (nhs, nrs, nts)] if has_neg else [None, None, None]
        gpos = [torch.empty_like(x) for x in pos]
        gneg = [torch.empty_like(x) for x in neg] if has_neg else [None, None, None]
        loss = torch.zeros(1, dtype=torch.float64, device=phs.device)
        n_pos, n_neg = pos[0].shape[0], (neg[0].shape[0] if has_neg else 0)
        L.check(lib.oea_loss_rows(_ptr(pos[0]), _ptr(pos[1]), _ptr(pos[2]), n_pos, _ptr(neg[0]), _ptr(neg[1]), _ptr(neg[2]),
                                  n_neg, dim, pitch, C.byref(cfg), _ptr(loss), _ptr(gpos[0]), _ptr(gpos[1]), _ptr(gpos[2]),
                                  _ptr(gneg[0]), _ptr(gneg[1]), _ptr(gneg[2]), _stream_ptr()), "oea_loss_rows")
        ctx.dim = dim
        ctx.has_neg = has_neg
        ctx.save_for_backward(*[g for g in gpos + gneg if g is not None])
        return loss.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, gout):
        saved = list(ctx.saved_tensors)
        d = ctx.dim
        grads = [g[:, :d] * gout for g in saved]
        if not ctx.has_neg:
            grads += [None, None, None]
        return (None, *grads)


def _run(loss, loss_norm, phs, prs, pts, nhs=None, nrs=None, nts=None, margin=0.0, neg_margin=0.0, balance=1.0):
    cfg = loss_cfg(loss, loss_norm, margin, neg_margin, balance)
    return _TripleLoss.apply(cfg, phs, prs, pts, nhs, nrs, nts)


def get_loss_func(phs, prs, pts, nhs, nrs, nts, args):
    triple_loss = None
    if args.loss == 'margin-based':
        triple_loss = margin_loss(phs, prs, pts, nhs, nrs, nts, args.margin, args.loss_norm)
    elif args.loss == 'logistic':
        triple_loss = logistic_loss(phs, prs, pts, nhs, nrs, nts, args.loss_norm)
    elif args.loss == 'limited':
        triple_loss = limited_loss(phs, prs, pts, nhs, nrs, nts, args.pos_margin, args.neg_margin, args.loss_norm)
    return triple_loss


def margin_loss(phs, prs, pts, nhs, nrs, nts, margin, loss_norm):
    return _run("margin-based", loss_norm, phs, prs, pts, nhs, nrs, nts, margin=margin)


def positive_loss(phs, prs, pts, loss_norm):
    return _run("positive", loss_norm, phs, prs, pts)


def limited_loss(phs, prs, pts, nhs, nrs, nts, pos_margin, neg_margin, loss_norm, balance=1.0):
    return _run("limited", loss_norm, phs, prs, pts, nhs, nrs, nts, margin=pos_margin, neg_margin=neg_margin, balance=balance)


def logistic_loss(phs, prs, pts, nhs, nrs, nts, loss_norm):
    return _run("logistic", loss_norm, phs, prs, pts, nhs, nrs, nts)


def alignment_loss(phs, prs, pts):
    """−Σ log σ(−‖h+r−t‖²) of approaches/bootea.py:197."""
    return _run("logsigmoid", "L2", phs, prs, pts)


class _MappingLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tes1, tes2, mapping):
        lib = L.load()
        n, dim = tes1.shape
        pitch = pitch_for(dim)
        e1, e2, M = _padded(tes1, pitch), _padded(tes2, pitch), _padded(mapping, pitch)
        g1, g2 = torch.zeros_like(e1), torch.zeros_like(e2)
        gM = torch.zeros_like(M)
        loss = torch.zeros(1, dtype=torch.float64, device=tes1.device)
        ws_bytes = lib.oea_mapping_workspace_bytes(dim)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=tes1.device)
        L.check(lib.oea_mapping_fwd_bwd(_ptr(e1), _ptr(e2), n, dim, pitch, _ptr(M), pitch, 1.0, _ptr(loss), _ptr(g1), _ptr(g2),
                                        _ptr(gM), _ptr(ws), ws_bytes, _stream_ptr()), "oea_mapping_fwd_bwd")
        ctx.dim = dim
        ctx.save_for_backward(g1, g2, gM)
        return loss.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, gout):
        g1, g2, gM = ctx.saved_tensors
        d = ctx.dim
        return g1[:, :d] * gout, g2[:, :d] * gout, gM[:, :d] * gout


def mapping_loss(tes1, tes2, mapping, eye=None):
    """Σ‖tes2 − tes1·M‖² + Σ(M·Mᵀ − I)²  (losses.py:76-80; `eye` accepted for signature compatibility)."""
    return _MappingLoss.apply(tes1, tes2, mapping)
