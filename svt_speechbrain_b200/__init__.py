"""svt_speechbrain_b200 -- B200-native (sm_100a) AMT inference hot path of guxm2021/SVT_SpeechBrain.

Reference-shaped modules (same names, constructor kwargs, forward signatures and state_dict keys):
    HuggingFaceWav2Vec2  <- MIR_ST500/huggingface_interface.py
    Linear               <- speechbrain/nnet/linear.py
    FusionRCA            <- N20EMv2/audio_visual/fusion.py
    FairseqAVHubertPretrain <- N20EMv2/video_only/fairseq_interface.py
    frame2note           <- MIR_ST500/utils.py
All arithmetic runs in the C-ABI library libsvt_b200.so (include/svt_b200.h); there is no CPU fallback.
"""
from ._lib import LIB_PATH, SvtError, lib  # noqa: F401
from .amt import (AMTHparams, AMTTranscriber, AVTranscriber, decode_logits, frame_info, split_song,  # noqa: F401
                  split_song_overlapped, stitch_plan)
from .fairseq_interface import FairseqAVHubertPretrain  # noqa: F401
from .fusion import FusionRCA  # noqa: F401
from .huggingface_interface import HuggingFaceWav2Vec2  # noqa: F401
from .linear import Linear  # noqa: F401
from .utils import decode_arrays, frame2note  # noqa: F401
from . import feature_cache  # noqa: F401,E402

__all__ = ["HuggingFaceWav2Vec2", "FairseqAVHubertPretrain", "AVTranscriber", "Linear", "FusionRCA", "frame2note", "decode_arrays", "AMTTranscriber", "AMTHparams",
           "split_song", "lib", "SvtError", "LIB_PATH"]
