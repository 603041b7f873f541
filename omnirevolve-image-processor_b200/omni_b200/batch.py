"""Frame batches over several GPUs: the path shards by frame with NO data-path collective.

One process per GPU (torchrun).  Each rank takes a contiguous block of ceil(B/G) frames, runs the fused
colour+edge call per frame on its own device, keeps its masks/edges in its own (pinned) host buffers, and only
the per-frame counts (the numbers the reference logs: pixels, mask_nonzero, edge nz) are gathered to rank 0.
torch.distributed is plumbing here: gloo on CPU boxes (tests), nccl on GPU boxes.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_frames: int, world: int, rank: int) -> range:
    """Contiguous block of ceil(B/G) frames for `rank` (the last ranks may get fewer or none)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = -(-n_frames // world) if n_frames > 0 else 0
    lo = min(n_frames, rank * per)
    return range(lo, min(n_frames, lo + per))


def run_shard(frames, frame_ids, compute):
    """compute(frame) -> int64 array [K,3] of counts; returns {frame_id: counts} for this rank's frames."""
    return {int(i): np.asarray(compute(f), dtype=np.int64) for i, f in zip(frame_ids, frames)}


def gather_counts(local: dict, n_frames: int, K: int, dist=None) -> np.ndarray | None:
    """All ranks' {frame_id: counts[K,3]} -> on rank 0 one int64 array [n_frames, K, 3] in frame order."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        parts = [local]
        rank = 0
    else:
        rank = dist.get_rank()
        parts = [None] * dist.get_world_size() if rank == 0 else None
        dist.gather_object(local, parts, dst=0)
    if rank != 0:
        return None
    out = np.full((n_frames, K, 3), -1, np.int64)
    for part in parts:
        for i, c in part.items():
            if not (0 <= i < n_frames) or out[i, 0, 0] != -1:
                raise RuntimeError(f"frame {i} reported twice or out of range")
            out[i] = c
    if (out == -1).any():
        raise RuntimeError("some frames were not processed by any rank")
    return out


class FrameBatcher:
    """Runs omni_host_color_edge over this rank's frames with fixed centres (one Engine, pinned result buffers)."""

    def __init__(self, engine, centers, lut, edge_cfg, h: int, w: int):
        from .ops import pinned_empty
        self.eng, self.centers, self.lut, self.ec = engine, np.asarray(centers, np.float32), lut, edge_cfg
        K = self.centers.shape[0]
        self.masks = pinned_empty((K, h, w))
        self.edges = pinned_empty((K, h, w))

    def __call__(self, frame: np.ndarray) -> np.ndarray:
        r = self.eng.host_color_edge(frame, self.centers, self.lut, self.ec, want_labels=False,
                                     masks=self.masks, edges=self.edges, want_counts=True)
        return r["counts"]


def frame_groups(n_frames: int, K: int, max_planes: int = 32) -> list:
    """Split a rank's frames into groups whose n * K layers fit the plane dimension of one omni_color_edge_batch pass."""
    per = max(1, max_planes // max(1, K))
    return [range(lo, min(n_frames, lo + per)) for lo in range(0, n_frames, per)]


class DeviceFrameBatcher:
    """Device-resident frames of one rank through omni_color_edge_batch, a group of frames per call (their n * K layers are
    the plane dimension of one morphology / edge / hysteresis launch).  Returns (masks, edges) tensors [n,K,H,W]."""

    def __init__(self, engine, centers, lut, edge_cfg):
        self.eng, self.centers, self.lut, self.ec = engine, np.asarray(centers, np.float32), lut, edge_cfg

    def __call__(self, frames, masks=None, edges=None):
        return self.eng.color_edge_batch(frames, self.centers, self.lut, self.ec, masks=masks, edges=edges)
