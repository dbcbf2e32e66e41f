"""Event windows of a 10-second VidSitu video.

The reference cuts each video (300 frames at 30 fps, `{vseg}_{i:06d}.jpg`) into five 2-second
events and, per event, samples DATA.NUM_FRAMES frames every DATA.SAMPLING_RATE around the event
centre, clamping indices that run past either end of the video
(`VsituDS.set_comm_args` vidsitu_code/dat_loader.py:69-79, `get_frms_all` :454-472,
`get_sequence` utils/video_utils.py:18-38).  The tables built here drive the frame-pack kernel
directly on a device-resident `[n_videos, 300, H, W, 3]` uint8 tensor, so one upload per video
(45 MB) replaces five normalised fp32 clip uploads (120 MB).
"""
from __future__ import annotations

from typing import List

EVENTS_PER_VIDEO = 5
EVENT_SECONDS = 2
VIDEO_FRAMES = 300   # comm.max_frms, dat_loader.py:77


def window_indices(center: int, half_len: int, step: int, num_video_frames: int) -> List[int]:
    """Indices center-half_len, +step, ... (< center+half_len), each clamped into the video."""
    last = num_video_frames - 1
    return [min(max(i, 0), last) for i in range(center - half_len, center + half_len, step)]


def event_centers(fps: int = 30, events: int = EVENTS_PER_VIDEO) -> List[int]:
    """Centre frame of Ev1..Ev5: int((k + 1/2) * fps * 2) = 30, 90, 150, 210, 270 at 30 fps."""
    return [int((k + 0.5) * fps * EVENT_SECONDS) for k in range(events)]


def event_frame_indices(num_frames: int, sampling_rate: int, fps: int = 30,
                        num_video_frames: int = VIDEO_FRAMES) -> List[List[int]]:
    """Per event, the NUM_FRAMES indices of the fast / single pathway window."""
    half = (num_frames * sampling_rate) // 2
    return [window_indices(c, half, sampling_rate, num_video_frames) for c in event_centers(fps)]
