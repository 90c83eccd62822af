"""CPU restatement of the reference's heat-map decoders (TEST INFRASTRUCTURE).

* ``get_max_preds``     -- ``libs/common/img_proc.py:608-637`` (hard arg-max,
  numpy first-occurrence tie rule, ``maxval > 0`` mask).
* ``soft_arg_max``      -- ``libs/common/img_proc.py:678-707`` (torch, soft-max
  normalised, raw-map max, no mask).  The upstream function hard-codes
  ``torch.cuda.FloatTensor`` / ``torch.cuda.comm`` (lines 696-700) and cannot run
  on CPU; this follows it line by line with CPU tensors.
* ``soft_arg_max_np``   -- ``libs/common/img_proc.py:639-676`` (sum-normalised;
  the ``np.clip`` on line 656 is dead code upstream because its result is never
  used by the reshaped view; the in-place division of the caller's array is
  NOT reproduced -- inputs are left untouched).
"""
import numpy as np
import torch
import torch.nn.functional as F


def get_max_preds(batch_heatmaps):
    hm = np.asarray(batch_heatmaps)
    assert hm.ndim == 4
    B, K, H, W = hm.shape
    flat = hm.reshape(B, K, -1)
    idx = np.argmax(flat, 2)                       # first occurrence on ties
    maxvals = np.amax(flat, 2).reshape(B, K, 1)
    preds = np.zeros((B, K, 2), dtype=np.float32)
    preds[:, :, 0] = (idx % W).astype(np.float32)
    preds[:, :, 1] = np.floor(idx.astype(np.float32) / W)
    mask = (maxvals > 0.0).astype(np.float32)
    preds *= np.tile(mask, (1, 1, 2))
    return preds, maxvals, idx.astype(np.int32)


def soft_arg_max(batch_heatmaps):
    hm = torch.as_tensor(batch_heatmaps, dtype=torch.float32)
    assert hm.dim() == 4
    B, K, H, W = hm.shape
    flat = hm.reshape(B, K, -1)
    maxvals = flat.max(dim=2)[0].view(B, K, 1)
    p = F.softmax(flat, dim=2).view(B, K, H, W)
    x = p.sum(dim=2)                               # [B,K,W]
    y = p.sum(dim=3)                               # [B,K,H]
    x = x * torch.arange(W, dtype=torch.float32).view(1, 1, W)
    y = y * torch.arange(H, dtype=torch.float32).view(1, 1, H)
    preds = torch.cat([x.sum(dim=2, keepdim=True), y.sum(dim=2, keepdim=True)], dim=2)
    return preds.numpy(), maxvals.numpy()


def soft_arg_max_np(batch_heatmaps):
    hm = np.array(batch_heatmaps, dtype=np.float32, copy=True)
    assert hm.ndim == 4
    B, K, H, W = hm.shape
    flat = hm.reshape(B, K, -1)
    maxvals = np.amax(flat, 2).reshape(B, K, 1)
    flat = flat / flat.sum(axis=2, keepdims=True)   # float32 sum, as upstream
    p = flat.reshape(B, K, H, W)
    x = p.sum(axis=2) * np.arange(W, dtype=np.float32).reshape(1, 1, W)
    y = p.sum(axis=3) * np.arange(H, dtype=np.float32).reshape(1, 1, H)
    preds = np.concatenate([x.sum(axis=2, keepdims=True), y.sum(axis=2, keepdims=True)], axis=2)
    preds = preds * np.tile((maxvals > 0.0).astype(np.float32), (1, 1, 2))
    return preds, maxvals
