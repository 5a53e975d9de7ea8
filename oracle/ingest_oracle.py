"""ORACLE — CPU restatement of the reference's raw-feature ingest.  TEST INFRASTRUCTURE ONLY (see oracle/nmn_oracle.py).

Follows video_nmn/dataset.py:134-172 line by line (one video at a time, the very torch / numpy calls of the reference):
  h5 path   :145-152   feat = resnet_features[id]; feat = feat[:max_len] if longer; torch.tensor(feat).mean(dim=1)
            :161-172   m = resnext_features[id]; m = m[:max_len] if longer; torch.cat([feat, torch.tensor(m)], dim=-1)
  npy path  :134-143   feat = np.load(...); feat = feat[np.arange(0, n, 2)]; feat = feat[:max_len] if longer; torch.tensor(feat).squeeze()
PARITY PIN: the reference ships no fixtures for this step and its loader needs h5py + the AGQA feature files (absent); the pin is
that these functions are the reference's own ATen / numpy calls on the same arrays.
"""
import numpy as np
import torch


def rx_video_features(appearance: np.ndarray, motion: np.ndarray | None, max_video_length: int) -> torch.Tensor:
    feat = appearance
    if feat.shape[0] > max_video_length:
        feat = feat[:max_video_length]
    feat = torch.tensor(feat).mean(dim=1)
    if motion is not None:
        m = motion
        if m.shape[0] > max_video_length:
            m = m[:max_video_length]
        feat = torch.cat([feat, torch.tensor(m)], dim=-1)
    return feat


def i3d_video_features(feats: np.ndarray, max_video_length: int) -> torch.Tensor:
    select_idx = np.arange(start=0, stop=feats.shape[0], step=2)
    feat = feats[select_idx, :]
    if feat.shape[0] > max_video_length:
        feat = feat[:max_video_length]
    return torch.tensor(feat).squeeze()
