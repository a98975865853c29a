"""Python restatement of frame2note (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows MIR_ST500/utils.py:82-149 (identical body in all four recipe utils.py) statement by
statement, with the CPython `max(set(c), key=c.count)` expression kept verbatim because its
tie-break depends on CPython's set iteration order.  The independent C
restatement (oracle/frame2note_oracle.c) emulates that order and is cross-checked against
this file and against the imported reference in tests/.
"""
import numpy as np


def frame2note(frame_info, onset_thres, offset_thres, frame_size=1 / 49.8):
    result = []
    current_onset = None
    pitch_counter = []
    onset_seq = np.array([frame_info[i][0] for i in range(len(frame_info))])
    n = len(onset_seq)
    current_time = 0.0
    for i in range(len(frame_info)):
        current_time = frame_size * i
        info = frame_info[i]
        lo = max(i - 3, 0)  # utils.py:106-108
        hi = min(i + 4, n - 1)  # utils.py:110-112 (clamps to n-1: last frame never in a window)
        if info[0] >= onset_thres and onset_seq[i] == np.amax(onset_seq[lo:hi]):  # :115
            if current_onset is None:
                current_onset = current_time
            else:
                if len(pitch_counter) > 0:
                    result.append([current_onset, current_time, max(set(pitch_counter), key=pitch_counter.count) + 36])
                current_onset = current_time
                pitch_counter = []
        elif info[1] >= offset_thres:  # :129
            if current_onset is not None:
                if len(pitch_counter) > 0:
                    result.append([current_onset, current_time, max(set(pitch_counter), key=pitch_counter.count) + 36])
                current_onset = None
                pitch_counter = []
        if current_onset is not None:  # :138-142
            final_pitch = int(info[2] * 12 + info[3])
            if info[2] != 4 and info[3] != 12:
                pitch_counter.append(final_pitch)
    if current_onset is not None:  # :144-147
        if len(pitch_counter) > 0:
            result.append([current_onset, current_time, max(set(pitch_counter), key=pitch_counter.count) + 36])
    return result
