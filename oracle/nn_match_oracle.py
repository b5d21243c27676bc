"""CPU ORACLE for the descriptor matching that feeds the hot path -- TEST INFRASTRUCTURE ONLY.

Call site in the reference (deepFEPE/train_good_utils.py:683-724, get_matches_from_SP):

    matching_mask = SP_tracker.nn_match_two_way(desc1.T, desc2.T, nn_thresh=SP_tracker.nn_thresh)   # [3, n]
    choice = utils_misc.crop_or_pad_choice(n, out_num_points=1000, shuffle=True)                    # np.random
    xs = cat(pts1[matching_mask[0, choice]], pts2[matching_mask[1, choice]]); quality = matching_mask[2:3, choice].T

`SP_tracker` is `superpoint.models.model_wrap.PointTracker` (train_good.py:222, nn_thresh 0.7 / 1.0 in the configs):
the `superpoint` package (eric-yyjau/pytorch-superpoint, un-vendored submodule, no pinned version) is NOT under
/root/reference, so **parity is unpinned** for this row.  Its published algorithm (identical to MagicLeap's
SuperPointPretrainedNetwork demo_superpoint.py `PointTracker.nn_match_two_way`) is restated below:

    dmat = sqrt(2 - 2 clip(desc1^T desc2, -1, 1));  idx = argmin(dmat, axis=1);  scores = dmat[arange, idx]
    keep = (scores < nn_thresh) & (arange == argmin(dmat, axis=0)[idx])              # mutual nearest neighbours
    matches = [arange[keep], idx[keep], scores[keep]]                                # ascending in the first index

crop_or_pad_choice (deepFEPE/dsac_tools/utils_misc.py:139-161) draws from numpy's global RNG and is not
reproducible by construction; only its invariants are testable (a permutation prefix, or all n plus a resample).
"""
from __future__ import annotations

import numpy as np


def nn_match_two_way(desc1: np.ndarray, desc2: np.ndarray, nn_thresh: float) -> np.ndarray:
    """desc1 [D,N1], desc2 [D,N2] (L2-normalised columns) -> matches [3,n] (idx1, idx2, score)."""
    assert desc1.shape[0] == desc2.shape[0]
    if desc1.shape[1] == 0 or desc2.shape[1] == 0:
        return np.zeros((3, 0))
    dmat = np.dot(desc1.T, desc2)
    dmat = np.sqrt(2 - 2 * np.clip(dmat, -1, 1))
    idx = np.argmin(dmat, axis=1)
    scores = dmat[np.arange(dmat.shape[0]), idx]
    keep = scores < nn_thresh
    idx2 = np.argmin(dmat, axis=0)
    keep = np.logical_and(keep, np.arange(len(idx)) == idx2[idx])
    m_idx1 = np.arange(desc1.shape[1])[keep]
    matches = np.zeros((3, int(keep.sum())))
    matches[0, :] = m_idx1
    matches[1, :] = idx[keep]
    matches[2, :] = scores[keep]
    return matches
