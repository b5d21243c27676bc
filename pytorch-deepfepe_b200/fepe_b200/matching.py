"""Batched, on-device mirror of the reference's match construction from keypoints + descriptors.

Reference (per sample, numpy on the host after `.cpu()` of both descriptor sets):
    deepFEPE/train_good_utils.py:649-724 get_matches_from_SP
        matching_mask = SP_tracker.nn_match_two_way(desc1.T, desc2.T, nn_thresh)            # [3, n]
        choice = utils_misc.crop_or_pad_choice(n, out_num_points, shuffle=True)              # utils_misc.py:139-161
        xs = cat(pts1[mask[0, choice]], pts2[mask[1, choice]], 1); offsets likewise; quality = mask[2:3, choice].T
Here the matching is one call of fepe_nn_match for the batch; the random crop / pad and the gathers are batched torch
indexing on the device (no host round trip, no Python loop over samples).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def crop_or_pad_choice_batch(counts: torch.Tensor, cap: int, out_num_points: int,
                             generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """utils_misc.crop_or_pad_choice(n, out_num_points, shuffle=True) for every sample at once.
    counts [B] (n per sample, <= cap).  Returns choice [B,out_num_points] int64 into 0..n-1: a random permutation
    prefix when n >= out_num_points, otherwise the permutation followed by a resample with replacement.  Samples
    with n == 0 (the reference would raise inside np.random.choice) get index 0."""
    B, dev = counts.shape[0], counts.device
    n = counts.to(torch.int64).clamp(min=0, max=cap)
    valid = torch.arange(cap, device=dev).unsqueeze(0) < n.unsqueeze(1)                  # [B,cap]
    keys = torch.rand(B, cap, device=dev, generator=generator).masked_fill(~valid, 2.0)   # invalid entries sort last
    perm = keys.argsort(dim=1)                                                            # valid indices first, shuffled
    take = min(out_num_points, cap)
    choice = perm[:, :take]
    if take < out_num_points:
        choice = torch.cat((choice, choice.new_zeros(B, out_num_points - take)), 1)
    pos = torch.arange(out_num_points, device=dev).unsqueeze(0)
    pad = (torch.rand(B, out_num_points, device=dev, generator=generator) * n.clamp(min=1).unsqueeze(1)).long()
    pad = torch.gather(perm, 1, pad.clamp(max=cap - 1))                                   # resample among the valid ones
    return torch.where(pos < n.unsqueeze(1), choice, pad)


def get_matches_from_descriptors(pts1: torch.Tensor, pts2: torch.Tensor, desc1: torch.Tensor, desc2: torch.Tensor,
                                 nn_thresh: float, out_num_points: int = 1000, res1: Optional[torch.Tensor] = None,
                                 res2: Optional[torch.Tensor] = None, n1: Optional[torch.Tensor] = None,
                                 n2: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None) -> dict:
    """The second half of get_matches_from_SP (train_good_utils.py:680-724) for a batch.

    pts1 [B,N1,2], pts2 [B,N2,2] keypoints (pixels); desc1 [B,N1,D], desc2 [B,N2,D]; res1/res2 sub-pixel offsets like
    pts (optional); n1/n2 [B] int32 valid counts.  Returns the reference's dict: 'xs' [B,out,4], 'offsets' [B,out,4]
    (zeros without res), 'quality' [B,out,1], 'num_matches' [B]."""
    idx1, idx2, score, count = ops.nn_match_two_way(desc1, desc2, nn_thresh, n1, n2)
    cap = idx1.shape[1]
    choice = crop_or_pad_choice_batch(count, cap, out_num_points, generator)
    i1 = torch.gather(idx1.long(), 1, choice)
    i2 = torch.gather(idx2.long(), 1, choice)
    has = (count > 0).view(-1, 1, 1)
    g = lambda t, i: torch.gather(t, 1, i.unsqueeze(-1).expand(-1, -1, t.shape[2]).clamp(0, t.shape[1] - 1))
    xs = torch.cat((g(pts1, i1), g(pts2, i2)), 2) * has
    if res1 is not None and res2 is not None:
        offsets = torch.cat((g(res1, i1), g(res2, i2)), 2) * has
    else:
        offsets = torch.zeros_like(xs)
    quality = (torch.gather(score, 1, choice) * has.view(-1, 1)).unsqueeze(-1)
    return {"xs": xs, "offsets": offsets, "quality": quality, "num_matches": count}
