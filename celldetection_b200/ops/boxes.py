"""Box ops of the CPN inference path with the reference's names (/root/reference/celldetection/ops/boxes.py):
``filter_by_box_voting`` (:61-83, the ensemble vote of cpn_inference.py:419-424), ``contours2boxes`` (:86-98) and
``nms`` (torch.ops.torchvision.nms semantics).  CUDA tensors only; executed by the C ABI's kernels."""
import torch
from torch import Tensor

from .. import _lib as L
from .cpn import nms, _require_cuda

__all__ = ['filter_by_box_voting', 'box_votes', 'contours2boxes', 'nms']


def box_votes(boxes: Tensor, thresh: float) -> Tensor:
    """``get_iou_voting`` (ops/boxes.py:53-58): ``(iou * (iou > thresh)).sum(-1)`` over all box pairs, including the
    box itself, without the dense K x K matrix (``cpn_box_votes``)."""
    _require_cuda(boxes)
    lib = L.load()
    K = int(boxes.shape[0])
    votes = torch.empty((K,), dtype=torch.float32, device=boxes.device)
    if K:
        ws = torch.empty((int(lib.cpn_box_votes_workspace_bytes(K)),), dtype=torch.uint8, device=boxes.device)
        boxes = boxes.contiguous().float()                    # named: a converted copy must outlive the launch
        L.check(lib.cpn_box_votes(L.ptr(boxes), K, float(thresh), L.ptr(ws), L.ptr(votes),
                                  L.stream_ptr()), 'box_votes')
    return votes


def filter_by_box_voting(boxes, thresh, min_vote, return_votes: bool = False):
    """ops/boxes.py:61-83: keep indices (int32, ascending) of the boxes whose vote -- the sum of the IoUs above
    ``thresh`` with all boxes, itself included -- reaches ``min_vote``; optionally the votes of the kept boxes."""
    votes = box_votes(boxes, thresh)
    mask = votes >= min_vote
    keep = torch.nonzero(mask, as_tuple=False).reshape(-1).to(torch.int)
    if return_votes:
        return keep, votes[mask]
    return keep


def contours2boxes(contours, axis=-2):
    """ops/boxes.py:86-98"""
    return torch.cat((contours.min(axis).values, contours.max(axis).values), axis + (axis < 0))
