"""B200-native batched PESQ and STOI/ESTOI scoring (drop-in for the hot path of
kcoost/fast_speech_enhancement_metrics).

    from fast_speech_enhancement_metrics_b200 import PESQ, STOI
    pesq = PESQ(sample_rate=16000, use_gpu=True)
    stoi = STOI(sample_rate=16000, use_gpu=True)
    pesq(clean, denoised)   # [{"PESQ": ...}, ...]
    stoi(clean, denoised)   # [{"STOI": ..., "ESTOI": ...}, ...]
"""
from .base import BaseMetric
from .PESQ import PESQ
from .STOI import STOI
from .LSD import LSD
from .SDR import SDR
from .fused import CapturedScorer, score_pesq_stoi, score_pesq_stoi_tensors

__all__ = ["BaseMetric", "PESQ", "STOI", "LSD", "SDR", "score_pesq_stoi", "score_pesq_stoi_tensors", "CapturedScorer"]
__version__ = "0.2.0"
