"""Evaluation metrics with the reference's signature (deephumor/experiments/metrics.py:4-9)."""
import torch

from ..runtime import ops


def perplexity(logits, targets, lengths, pad_index=0):
    """Per-sequence exp(-sum_t logp_t / len) with pads zeroed, batch mean.  The log-softmax + target gather over
    [bs*T, V] runs in dh_token_logprob (one pass over the logits); only [bs, T] floats reach the torch tail."""
    bs, T, V = logits.shape
    logits = logits.float().contiguous()
    lp = torch.empty(bs * T, dtype=torch.float32, device=logits.device)
    ops.token_logprob(logits.view(bs * T, V), targets.to(logits.device).contiguous().view(-1), lp)
    lp = lp.view(bs, T) / lengths.to(logits.device).unsqueeze(1)       # divide by lengths BEFORE masking (Q27)
    lp = lp.masked_fill(targets.to(logits.device) == pad_index, 0.)
    return (-lp.sum(dim=-1)).exp().mean()
