"""PESQ + STOI/ESTOI on ONE host->device upload.

Callers of the reference always score both metrics on the same pair of tensors
(README.md:29-30, benchmark_metrics.py:22,24).  With host tensors the PCIe copy dominates the
end-to-end time, so `score_pesq_stoi` hands the pair to the library once: every chunk is copied
once and both kernel pipelines run on it.  Results are identical to
calling the two metric objects separately.  (Round 1's single-READ first-pass kernel and two-stream
overlap modes measured slower than the two separate first kernels on B200 and were removed; DESIGN.md.)
"""
from __future__ import annotations

import ctypes as C
import warnings

import torch

from . import _lib
from .PESQ import PESQ
from .STOI import STOI


def score_pesq_stoi_tensors(pesq: PESQ, stoi: STOI, clean: torch.Tensor, deg: torch.Tensor, lengths=None):
    """Device-resident scoring of both metrics: [B, n] float32 / int16 / float16 CUDA tensors -> (scores[3, B] f32 = PESQ / STOI /
    ESTOI rows, pesq_status[B], kept_frames[B], stoi_status[B]) CUDA tensors; stream-ordered, no host sync
    (C ABI fsem_pesq_stoi_score_f32: the two kernel chains back to back on the current stream)."""
    clean, deg = pesq._on_device(clean), pesq._on_device(deg)
    if stoi.device != pesq.device:
        raise Exception("both metrics must live on the same CUDA device")
    b, n = clean.shape
    lens = pesq._lengths_tensor(lengths, b, n, clean.device)
    scores = torch.empty(3, b, dtype=torch.float32, device=clean.device)
    pst = torch.empty(b, dtype=torch.int32, device=clean.device)
    kept = torch.empty(b, dtype=torch.int32, device=clean.device)
    sst = torch.empty(b, dtype=torch.int32, device=clean.device)
    with torch.cuda.device(clean.device):
        if deg.stride(0) != clean.stride(0) and b > 1:
            deg = deg.contiguous(); clean = clean.contiguous()
        wp = pesq._get_workspace(pesq._lib.fsem_pesq_workspace_bytes(pesq._ctx, b, n))
        wsb = stoi._get_workspace(stoi._lib.fsem_stoi_workspace_bytes(stoi._ctx, b, n))
        batch = _lib.Batch(clean.data_ptr(), deg.data_ptr(), lens.data_ptr() if lens is not None else None,
                           b, n, clean.stride(0) if b > 1 else max(n, clean.stride(0)))
        PESQ._check_score(pesq._lib.fsem_pesq_stoi_score(
            pesq._ctx, stoi._ctx, C.byref(batch), _lib.dtype_code(clean.dtype), scores[0].data_ptr(), pst.data_ptr(), scores[1].data_ptr(),
            scores[2].data_ptr(), kept.data_ptr(), sst.data_ptr(), wp.data_ptr(), wp.numel(), wsb.data_ptr(),
            wsb.numel(), C.c_void_p(torch.cuda.current_stream(clean.device).cuda_stream)))
    pesq._last_shape = (b, n)
    stoi._last_shape = (b, n)
    stoi._last_lengths = lens
    return scores, pst, kept, sst


def score_pesq_stoi(pesq: PESQ, stoi: STOI, clean_speech: torch.Tensor, denoised_speech: torch.Tensor,
                    lengths=None) -> list[dict[str, float]]:
    """Returns [{"PESQ": ..., "STOI": ..., "ESTOI": ...}, ...], one dict per batch row.
    `stoi` must have been built for the tensors' sample rate (e.g. STOI(16000)) and `pesq` for 16 kHz."""
    if stoi.sample_rate != pesq.sample_rate:
        raise Exception("both metrics must be built for the sample rate of the audio")
    clean, deg = pesq.prepare_inputs(clean_speech, denoised_speech)
    assert clean is not None
    b, n = clean.shape
    if clean.is_cuda:
        # device-resident tensors: one library call for both kernel chains
        scores, pst, kept, _ = score_pesq_stoi_tensors(pesq, stoi, clean, deg, lengths)
        packed = torch.cat([scores, pst.to(torch.float32)[None], kept.to(torch.float32)[None]]).cpu()
        mos, sc, pst, kept = packed[0], packed[1:3], packed[3].to(torch.int32), packed[4].to(torch.int32)
    else:
        lens = pesq._lengths_tensor(lengths, b, n, "cpu")
        mos = torch.empty(b, dtype=torch.float32)
        pst = torch.empty(b, dtype=torch.int32)
        sc = torch.empty(2, b, dtype=torch.float32)
        kept = torch.empty(b, dtype=torch.int32)
        sst = torch.empty(b, dtype=torch.int32)
        if clean.stride(0) != deg.stride(0) and b > 1:
            clean, deg = clean.contiguous(), deg.contiguous()
        with torch.cuda.device(pesq.device):
            PESQ._check_score(pesq._lib.fsem_score_host(
                pesq._ctx, stoi._ctx, clean.data_ptr(), deg.data_ptr(), _lib.dtype_code(clean.dtype),
                lens.data_ptr() if lens is not None else None, b, n, clean.stride(0) if b > 1 else n,
                mos.data_ptr(), pst.data_ptr(), sc[0].data_ptr(), sc[1].data_ptr(), kept.data_ptr(), sst.data_ptr()))
    if lengths is not None and bool((pst == _lib.ITEM_TOO_SHORT).any()):
        raise RuntimeError("PESQ needs at least 20 frames of 512/256 samples for every item")
    stoi.last_kept_frames = kept
    if int(kept.max()) <= stoi.N + 1:
        warnings.warn("Not enough non-silent frames. Please check your sound files", RuntimeWarning, stacklevel=2)
        raise TypeError("iteration over a 0-d tensor")
    return [{"PESQ": p, "STOI": s, "ESTOI": e} for p, s, e in zip(mos.tolist(), sc[0].tolist(), sc[1].tolist())]
