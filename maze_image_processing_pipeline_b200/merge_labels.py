"""Drop-in for ``maze_ipp/merge_labels.py`` computed on the GPU (``maze_merge_labels``).

Same signature, identity / aliasing behaviour and exception as the reference
(maze_ipp/merge_labels.py:29-113): returns ``labels`` itself when fewer than two labels are listed,
writes into ``labels_out`` (which may be ``labels`` -- the pipeline's call, loki/pipeline.py:452-457),
pops the processed entries from a caller-supplied ``index`` list and raises ``TypeError`` where
``find_objects`` yields ``None`` in the reference (:19-20).
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import MAZE_ERR_TYPEERROR
from .device import BatchGeometry, DeviceBatch


def merge_labels(labels, index=None, max_distance=None, path_tolerance=5, return_merge_distances=False,
                 labels_out=None):
    lab = np.asarray(labels)
    if lab.ndim != 2:
        raise ValueError("merge_labels is implemented for 2-D label images")
    if index is not None and len(index) < 2:
        return (labels, []) if return_merge_distances else labels
    if lab.size == 0:
        return (labels, []) if return_merge_distances else labels
    geom = BatchGeometry([lab.shape[0]], [lab.shape[1]])
    batch = DeviceBatch(geom)
    d_lab = batch.upload(geom.pack_host([lab.astype(np.int32, copy=False)], dtype=np.int32))
    bound = int(batch.max_label(d_lab).cpu()[0])
    d_index = d_index_off = None
    if index is not None:
        idx = np.asarray(list(index), dtype=np.int64)
        in_range = idx[(idx > 0) & (idx < 2 ** 31)]
        bound = max(bound, int(in_range.max()) if in_range.size else 0, len(idx))
        # labels outside int32 / non-positive can never match a pixel: map them to an absent label
        bound += 1
        idx32 = np.where((idx > 0) & (idx < 2 ** 31), idx, bound).astype(np.int32)
        d_index = torch.from_numpy(idx32).to(batch.device)
        d_index_off = torch.tensor([0, len(idx32)], dtype=torch.int32, device=batch.device)
    lab_off, n_obj = batch.lab_off_from_bounds([bound])
    aliased = labels_out is labels
    d_out = d_lab if aliased else (d_lab.clone() if labels_out is None else
                                   batch.upload(geom.pack_host([np.asarray(labels_out).astype(np.int32, copy=False)],
                                                               dtype=np.int32)))
    merge_dist, n_merge, index_state, status, obj_scratch = batch.merge_labels(
        d_lab, d_out, lab_off, n_obj, max_distance, path_tolerance, d_index, d_index_off)
    n_idx, popped = (int(v) for v in index_state.cpu()[:2])
    if index is not None and popped:
        order = obj_scratch.cpu().numpy()[:n_idx]
        # the reference pops from the caller's list (merge_labels.py:66, 84)
        original = list(index)
        remaining_pos = _remaining_positions(original, idx32, order, popped)
        index[:] = [original[p] for p in remaining_pos]
    if int(status.cpu()[0]) == MAZE_ERR_TYPEERROR:
        raise TypeError("'NoneType' object is not iterable")
    if n_idx < 2:
        return (labels, []) if return_merge_distances else labels
    res = geom.view(d_out.cpu().numpy(), 0)
    if labels_out is None:
        labels_out = res.astype(lab.dtype, copy=True)
    else:
        labels_out[...] = res
    if return_merge_distances:
        nm = int(n_merge.cpu()[0])
        return labels_out, [np.float64(v) for v in merge_dist.cpu().numpy()[:nm]]
    return labels_out


def _remaining_positions(original, idx32, order, popped):
    """Positions (in the caller's list) of the entries the loop did not pop, in list order."""
    remaining_vals = list(order[popped:])
    pos, out = 0, []
    for v in remaining_vals:
        while int(idx32[pos]) != int(v):
            pos += 1
        out.append(pos)
        pos += 1
    return out
