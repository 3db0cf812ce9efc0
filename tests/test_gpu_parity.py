"""GPU parity tests (-m gpu): the CUDA path, called through the public classes and hence
through the C ABI of libfsem_b200.so, against

  * the golden fixtures (outputs of the real reference's CPU path, tests/golden/*.npz), and
  * the numpy oracle (oracle/) on the same seeded inputs.

Tolerances (BASELINE.json north_star): |dPESQ| <= 1e-3, |dSTOI|, |dESTOI| <= 1e-4, the
silent-frame mask and the kept-frame count K bit-exact.  Against the float64 oracle (which
does not carry the reference's own float32-IIR noise) PESQ is held to 2e-4.
"""
import json
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import pesq_oracle as po
from oracle import stoi_oracle as so
from tests.conftest import ROOT, unpack_masks
from tests.golden.cases import pesq_cases, pesq_rate_cases, stoi_cases

pytestmark = pytest.mark.gpu

PESQ_CASES = pesq_cases()
PESQ_RATE_CASES = pesq_rate_cases()
STOI_CASES = stoi_cases()
REPORT = {}


def _report(key, value):
    REPORT[key] = value
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_report.json"), "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


@pytest.fixture(scope="module")
def pesq():
    from fast_speech_enhancement_metrics_b200 import PESQ
    return PESQ(16000, use_gpu=True)


@pytest.fixture(scope="module")
def stoi_metrics():
    from fast_speech_enhancement_metrics_b200 import STOI
    cache = {}

    def get(fs):
        if fs not in cache:
            cache[fs] = STOI(fs, use_gpu=True)
        return cache[fs]
    return get


def _maxdiff(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)), (a, b)
    ok = ~np.isnan(b)
    return float(np.max(np.abs(a[ok] - b[ok]), initial=0.0))


# ------------------------------------------------------------------------------------------ PESQ
@pytest.mark.parametrize("name", sorted(PESQ_CASES))
def test_pesq_matches_reference_and_oracle(name, pesq, golden_pesq):
    clean, deg, lengths = PESQ_CASES[name]
    c = torch.from_numpy(clean).cuda()
    d = torch.from_numpy(deg).cuda()
    got = np.array([r["PESQ"] for r in pesq(c, d, lengths=lengths)])
    want = golden_pesq[name]
    oracle = po.pesq_batch(clean, deg, lengths)
    d_ref, d_orc = _maxdiff(got, want), _maxdiff(got, oracle)
    _report("pesq/" + name, {"max_abs_vs_reference": d_ref, "max_abs_vs_oracle": d_orc, "got": got.tolist()})
    assert d_ref <= 1e-3
    assert d_orc <= 2e-4


@pytest.mark.parametrize("name", sorted(PESQ_RATE_CASES))
def test_pesq_resample_on_ingest(name, golden_pesq):
    """PESQ(sample_rate != 16000): the polyphase resampler of base.py:13,19-20 runs as the first kernel."""
    from fast_speech_enhancement_metrics_b200 import PESQ
    clean, deg, lengths, fs = PESQ_RATE_CASES[name]
    metric = PESQ(fs, use_gpu=True)
    got = np.array([r["PESQ"] for r in metric(torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda(), lengths=lengths)])
    host = np.array([r["PESQ"] for r in metric(torch.from_numpy(clean), torch.from_numpy(deg), lengths=lengths)])
    d_ref = _maxdiff(got, golden_pesq["rate/" + name])
    d_orc = _maxdiff(got, po.pesq_batch(clean, deg, lengths, fs))
    _report("pesq_rate/" + name, {"max_abs_vs_reference": d_ref, "max_abs_vs_oracle": d_orc})
    assert d_ref <= 1e-3 and d_orc <= 2e-4
    assert np.array_equal(got, host)


def test_pesq_stage_taps(pesq):
    clean, deg, _ = PESQ_CASES["speech2s"]
    c = torch.from_numpy(clean).cuda()
    d = torch.from_numpy(deg).cuda()
    pesq.score_tensors(c, d)
    bark, power = pesq.debug_taps()
    bark, power = bark.cpu().numpy().astype(np.float64), power.cpu().numpy()
    n = clean.shape[1]
    worst_p, worst_b = 0.0, 0.0
    for i in range(clean.shape[0]):
        for s, x in enumerate((clean[i], deg[i])):
            p_or = po.band_power(x.astype(np.float64))
            worst_p = max(worst_p, abs(power[s, i] - p_or) / p_or)
            m = max(np.abs(clean[i]).max(), np.abs(deg[i]).max())
            b_or = po.bark_bands(x.astype(np.float64) / m)
            worst_b = max(worst_b, np.max(np.abs(bark[s, i, : b_or.shape[0]] - b_or)) / b_or.max())
    _report("pesq/taps", {"band_power_rel": worst_p, "bark_rel_of_max": worst_b})
    assert worst_p <= 3e-5
    assert worst_b <= 1e-4


def test_pesq_host_entry_equals_device_entry(pesq):
    clean, deg, _ = PESQ_CASES["speech2s"]
    dev = [r["PESQ"] for r in pesq(torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda())]
    host = [r["PESQ"] for r in pesq(torch.from_numpy(clean), torch.from_numpy(deg))]
    assert dev == host
    # unaligned rows (odd row pitch) take the scalar-load path and must agree to rounding
    pad_c = torch.zeros(clean.shape[0], clean.shape[1] + 3, device="cuda")
    pad_d = torch.zeros_like(pad_c)
    pad_c[:, 1:-2] = torch.from_numpy(clean).cuda()
    pad_d[:, 1:-2] = torch.from_numpy(deg).cuda()
    odd = [r["PESQ"] for r in pesq(pad_c[:, 1:-2], pad_d[:, 1:-2])]
    assert np.max(np.abs(np.array(odd) - np.array(dev))) <= 1e-5


def test_pesq_errors(pesq):
    x = torch.randn(2, 5120, device="cuda")
    with pytest.raises(RuntimeError):                       # 19 frames (PESQ.py:169)
        pesq(x, x)
    with pytest.raises(Exception, match="same shape"):      # base.py:26-27
        pesq(torch.randn(2, 8000, device="cuda"), torch.randn(2, 8001, device="cuda"))
    with pytest.raises(AssertionError):                     # PESQ.py:235
        pesq(None, torch.randn(2, 8000, device="cuda"))
    with pytest.raises(RuntimeError):                       # float64 input
        pesq(torch.randn(1, 8000, dtype=torch.float64), torch.randn(1, 8000, dtype=torch.float64))
    out = pesq(torch.randn(16000), torch.randn(16000))      # 1-D input -> one item (atleast_2d)
    assert len(out) == 1 and isinstance(out[0]["PESQ"], float)


# ------------------------------------------------------------------------------------------ STOI
@pytest.mark.parametrize("name", sorted(STOI_CASES))
def test_stoi_matches_reference_and_oracle(name, stoi_metrics, golden_stoi):
    clean, deg, lengths, fs = STOI_CASES[name]
    metric = stoi_metrics(fs)
    c = torch.from_numpy(clean).cuda()
    d = torch.from_numpy(deg).cuda()
    res = metric(c, d, lengths=lengths)
    s = np.array([r["STOI"] for r in res])
    e = np.array([r["ESTOI"] for r in res])
    k = metric.last_kept_frames.numpy().astype(np.int64)
    os_, oe, ok = so.stoi_batch(clean, deg, fs, lengths)
    rep = {
        "stoi_vs_reference": _maxdiff(s, golden_stoi[name + "/stoi"]),
        "estoi_vs_reference": _maxdiff(e, golden_stoi[name + "/estoi"]),
        "stoi_vs_oracle": _maxdiff(s, os_), "estoi_vs_oracle": _maxdiff(e, oe),
        "K": k.tolist(), "K_reference": golden_stoi[name + "/K"].tolist(),
    }
    # silent-frame mask, bit for bit
    taps = metric.debug_taps()
    words = taps["mask"].cpu().numpy().view(np.uint32)
    masks = unpack_masks(golden_stoi, name)
    flips = 0
    for i in range(clean.shape[0]):
        nfr = len(masks[i]) if lengths is None else None
        ref_mask = masks[i]
        t0 = so.silent_frame_mask(so.resample(clean[i, : (clean.shape[1] if lengths is None else lengths[i])], fs, so.FS))[0].shape[0]
        mine = np.array([(words[i, t >> 5] >> (t & 31)) & 1 for t in range(t0)], bool)
        flips += int(np.sum(mine != ref_mask[:t0]))
    rep["mask_bit_flips"] = flips
    _report("stoi/" + name, rep)
    assert np.array_equal(k, golden_stoi[name + "/K"])
    assert flips == 0
    assert rep["stoi_vs_reference"] <= 1e-4 and rep["estoi_vs_reference"] <= 1e-4
    assert rep["stoi_vs_oracle"] <= 1e-4 and rep["estoi_vs_oracle"] <= 1e-4


def test_stoi_stage_taps(stoi_metrics, golden_stoi):
    clean, deg, _, fs = STOI_CASES["speech16k_3s"]
    metric = stoi_metrics(fs)
    metric.score_tensors(torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda())
    taps = metric.debug_taps()
    res = taps["resampled"].cpu().numpy()
    d_res = float(np.max(np.abs(res[0, 0, :4000] - golden_stoi["tap_resampled_clean"])))
    ref_tob = golden_stoi["tap_tob"]
    tob = taps["tob"].cpu().numpy()
    u = ref_tob.shape[2]
    d_tob = float(max(np.max(np.abs(tob[0, 0, :, :u] - ref_tob[0])), np.max(np.abs(tob[1, 0, :, :u] - ref_tob[1]))) / ref_tob.max())
    _report("stoi/taps", {"resampled_abs": d_res, "tob_rel_of_max": d_tob})
    assert d_res <= 5e-7
    assert d_tob <= 2e-5


def test_stoi_host_entry_equals_device_entry(stoi_metrics):
    clean, deg, _, fs = STOI_CASES["speech16k_3s"]
    metric = stoi_metrics(fs)
    dev = metric(torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda())
    host = metric(torch.from_numpy(clean), torch.from_numpy(deg))
    assert dev == host


def test_fused_upload_equals_separate_calls(pesq, stoi_metrics):
    from fast_speech_enhancement_metrics_b200 import score_pesq_stoi
    clean, deg, _, fs = STOI_CASES["speech16k_3s"]
    st = stoi_metrics(16000)
    c, d = torch.from_numpy(clean), torch.from_numpy(deg)
    sep_p, sep_s = pesq(c, d), st(c, d)
    for cc, dd in ((c, d), (c.cuda(), d.cuda())):
        both = score_pesq_stoi(pesq, st, cc, dd)
        assert [r["PESQ"] for r in both] == [r["PESQ"] for r in sep_p]
        assert [(r["STOI"], r["ESTOI"]) for r in both] == [(r["STOI"], r["ESTOI"]) for r in sep_s]
    lens = [48000, 30000, 20001, 48000, 12345, 40000, 48000, 8000]
    both = score_pesq_stoi(pesq, st, c, d, lengths=lens)
    sep_p, sep_s = pesq(c, d, lengths=lens), st(c, d, lengths=lens)
    assert [r["PESQ"] for r in both] == [r["PESQ"] for r in sep_p]
    assert np.array_equal(np.array([r["STOI"] for r in both]), np.array([r["STOI"] for r in sep_s]), equal_nan=True)


@pytest.mark.parametrize("shape", [(64, 48000, False), (33, 40004, False), (40, 64000, True), (150, 5380, False)])
def test_fused_device_entry_against_oracle(shape, pesq, stoi_metrics):
    """fsem_pesq_stoi_score_f32 (both kernel chains, one call) against the float64 ORACLE on the same seeded inputs --
    not against the CUDA path itself: PESQ <= 2e-4, STOI/ESTOI <= 1e-4, K exact (sample of items for the big shapes)."""
    from fast_speech_enhancement_metrics_b200 import score_pesq_stoi_tensors
    from fast_speech_enhancement_metrics_b200.synth import synth_batch
    b, n, ragged = shape
    clean, deg, _ = synth_batch(4242 + b, b, n)
    c, d = torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda()
    lens = None
    if ragged:
        lens = np.random.default_rng(b).integers(5400, n + 1, size=b)
        lens[0], lens[-1] = n, 5400
        lens = lens.tolist()
    st = stoi_metrics(16000)
    scores, pst, kept, sst = score_pesq_stoi_tensors(pesq, st, c, d, lens)
    scores, kept = scores.cpu().numpy().astype(np.float64), kept.cpu().numpy()
    pick = sorted(set(np.linspace(0, b - 1, 12).round().astype(int).tolist()))
    pl = None if lens is None else [lens[i] for i in pick]
    want_p = po.pesq_batch(clean[pick], deg[pick], pl)
    ws, we, wk = so.stoi_batch(clean[pick], deg[pick], 16000, pl)
    dp = _maxdiff(scores[0][pick], want_p)
    assert dp <= 2e-4, dp
    assert _maxdiff(scores[1][pick], ws) <= 1e-4 and _maxdiff(scores[2][pick], we) <= 1e-4
    assert np.array_equal(kept[pick], wk)
    _report("fused_vs_oracle_pesq_%dx%d%s" % (b, n, "_ragged" if ragged else ""), dp)


def test_mask_margin_reports_how_close_the_silent_frame_decisions_were(stoi_metrics, golden_stoi):
    """STOI.mask_margin(): min_t |(max E - 40) - E_t| per item (dB), against the oracle's float64 energies; the mask of
    every fixture case is bit-exact AND its smallest margin is reported (ties are reported by margin)."""
    for name in ("speech10k_3s", "speech16k_3s", "ragged10k"):
        clean, deg, lengths, fs = STOI_CASES[name]
        metric = stoi_metrics(fs)
        c, d = torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda()
        metric.score_tensors(c, d, lengths)
        got = metric.mask_margin().cpu().numpy()
        win = so.window32().astype(np.float64)
        for i in range(clean.shape[0]):
            n = clean.shape[1] if lengths is None else lengths[i]
            x = so.resample(clean[i, :n], fs, 10000).astype(np.float64)
            if len(x) < 256:
                assert np.isinf(got[i])
                continue
            t0 = (len(x) - 256) // 128 + 1
            fr = np.stack([x[128 * t:128 * t + 256] * win for t in range(t0)])
            e = 20 * np.log10(np.linalg.norm(fr, axis=1) + 1e-9)
            want = np.min(np.abs(e.max() - 40 - e))
            assert abs(got[i] - want) <= 2e-4 + 1e-3 * want, (name, i, got[i], want)
        _report("mask_margin_min_db_" + name, float(np.min(got[np.isfinite(got)])))


def test_headline_workload_sample_against_oracle(pesq, stoi_metrics):
    """BASELINE configs[4] at FULL size (8192 x 10 s on one GPU, the tensors bench.py times): every item scored, a
    sample of 32 items spread over the batch checked against the oracle (and the staged reference, if present) through
    bench.py's own parity block."""
    import bench
    free, _ = torch.cuda.mem_get_info()
    batch = 8192 if free > 60e9 else 1024
    device = torch.device("cuda", torch.cuda.current_device())
    clean, deg = bench.make_shard(batch, 160000, 1000, device)
    st = stoi_metrics(16000)
    mos, _ = pesq.score_tensors(clean, deg)
    sc, _, _ = st.score_tensors(clean, deg)
    table = torch.stack([mos, sc[0], sc[1]], dim=1)
    par = bench.parity_block(pesq, st, clean, deg, table, 0, batch, 1, device, 32)
    _report("headline_%dx10s" % batch, par)
    assert par["ok"], par
    assert par["items"] >= 32 and par["k_equal"]
    del clean, deg
    pesq._workspace = None
    st._workspace = None
    torch.cuda.empty_cache()


def test_batches_beyond_the_grid_y_limit(stoi_metrics):
    """More than 32767 items per call (2 x batch used to sit in gridDim.y, capped at 65535)."""
    from fast_speech_enhancement_metrics_b200.synth import synth_batch
    clean, deg, _ = synth_batch(77, 8, 8000)
    b = 33000
    c = torch.from_numpy(clean).cuda().repeat(b // 8, 1).contiguous()
    d = torch.from_numpy(deg).cuda().repeat(b // 8, 1).contiguous()
    st = stoi_metrics(16000)
    scores, kept, _ = st.score_tensors(c, d)
    scores = scores.cpu().numpy()
    assert np.array_equal(scores[:, :8], scores[:, -8:], equal_nan=True)
    ws, we, wk = so.stoi_batch(clean, deg, 16000)
    assert _maxdiff(scores[0, :8], ws) <= 1e-4 and _maxdiff(scores[1, :8], we) <= 1e-4
    st._workspace = None
    torch.cuda.empty_cache()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_tensors_on_another_device_are_moved_like_the_reference_does(pesq, stoi_metrics):
    """A metric lives on the device that was current at construction; CUDA tensors on another device are moved there
    (base.py:18 `.to(self.device)`) instead of being dereferenced from the wrong GPU."""
    clean, deg, _ = PESQ_CASES["speech2s"]
    c0, d0 = torch.from_numpy(clean).cuda(0), torch.from_numpy(deg).cuda(0)
    c1, d1 = c0.to("cuda:1"), d0.to("cuda:1")
    assert pesq(c1, d1) == pesq(c0, d0)
    st = stoi_metrics(16000)
    assert st(c1, d1) == st(c0, d0)


def test_stoi_errors(stoi_metrics, golden_stoi):
    metric = stoi_metrics(10000)
    assert int(golden_stoi["err_no_segments"]) == 1
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with pytest.raises(TypeError):                     # STOI.py:163-165, 205
            metric(torch.randn(2, 3000, device="cuda"), torch.randn(2, 3000, device="cuda"))
        assert any(issubclass(x.category, RuntimeWarning) for x in w)
    with pytest.raises(Exception, match="same shape"):
        metric(torch.randn(2, 8000, device="cuda"), torch.randn(2, 8001, device="cuda"))
    with pytest.raises(AssertionError):
        metric(None, torch.randn(2, 8000, device="cuda"))


# ------------------------------------------------------------------------------------------ properties at size
def test_properties_at_full_size(pesq, stoi_metrics):
    """BASELINE configs[1]/[2] shapes (256 x 10 s PESQ, 1024 x 4 s STOI): size-independent
    properties -- batch invariance, padded+lengths == sliced, identical pair ->
    the reference's fixed points (4.6438887 / 1.0), monotone in SNR."""
    from fast_speech_enhancement_metrics_b200.synth import synth_batch
    g = torch.Generator(device="cuda").manual_seed(7)
    base_c, base_d, _ = synth_batch(300, 8, 160000)
    c = torch.from_numpy(base_c).cuda().repeat(32, 1)                 # 256 x 10 s
    noise = torch.randn(c.shape, device="cuda", generator=g)
    snr = torch.linspace(-5, 25, 256, device="cuda")[:, None]
    d = c + noise * (c.pow(2).mean(1, keepdim=True) / 10 ** (snr / 10)).sqrt()
    full = np.array([r["PESQ"] for r in pesq(c, d)])
    assert np.all(np.isfinite(full)) and full.min() > 1.0 and full.max() < 4.65
    sub = np.array([r["PESQ"] for r in pesq(c[100:108].clone(), d[100:108].clone())])
    # batch invariance: the IIR pass picks its time-chunking from the batch size, so the band power (and with it
    # the score) may move in the last float32 bits between batch sizes; STOI below is bit-identical
    assert np.max(np.abs(sub - full[100:108])) <= 5e-6
    same = np.array([r["PESQ"] for r in pesq(c[:16], c[:16].clone())])
    assert np.max(np.abs(same - 4.6438887)) <= 1e-4
    # higher SNR -> better score for the same clean item (items i and i+8k share the clean signal)
    per_clean = full.reshape(32, 8)
    assert np.all(per_clean[-1] > per_clean[0])
    # padded batch + lengths == per-item slices
    lens = [160000, 48000, 16000 + 37, 5376, 99999, 160000, 8192, 30000]
    padded = np.array([r["PESQ"] for r in pesq(c[:8], d[:8], lengths=lens)])
    sliced = np.array([pesq(c[i:i + 1, :n].contiguous(), d[i:i + 1, :n].contiguous())[0]["PESQ"] for i, n in enumerate(lens)])
    assert np.max(np.abs(padded - sliced)) <= 1e-5      # IIR chunking differs with the batch size

    st = stoi_metrics(16000)
    c4 = c[:, :64000].repeat(4, 1).contiguous()                      # 1024 x 4 s
    d4 = d[:, :64000].repeat(4, 1).contiguous()
    res = st(c4, d4)
    s = np.array([r["STOI"] for r in res]); e = np.array([r["ESTOI"] for r in res])
    assert np.all(np.isfinite(s)) and np.all(np.isfinite(e)) and s.max() <= 1.0 + 1e-6
    assert np.array_equal(s[:256], s[256:512]) and np.array_equal(e[:256], e[768:])
    sub = st(c4[300:304].clone(), d4[300:304].clone())
    assert [r["STOI"] for r in sub] == s[300:304].tolist()
    one = st(c4[:8], c4[:8].clone())
    assert max(abs(r["STOI"] - 1.0) for r in one) <= 1e-5 and max(abs(r["ESTOI"] - 1.0) for r in one) <= 1e-5
    lens = [64000, 30000, 20001, 64000, 12345, 50000, 64000, 40000]
    padded = st(c4[:8], d4[:8], lengths=lens)
    kp = st.last_kept_frames.clone()
    for i, n in enumerate(lens):
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r = st(c4[i:i + 1, :n].contiguous(), d4[i:i + 1, :n].contiguous())[0]
        except TypeError:            # a lone item without a 30-frame segment (STOI.py:163-165, 205)
            assert int(kp[i]) <= 31 and np.isnan(padded[i]["STOI"])
            continue
        assert int(st.last_kept_frames[0]) == int(kp[i])
        assert abs(r["STOI"] - padded[i]["STOI"]) <= 2e-6 and abs(r["ESTOI"] - padded[i]["ESTOI"]) <= 2e-6


def test_variable_length_batch_against_oracle(pesq, stoi_metrics):
    """BASELINE configs[3]: a padded batch of 1-30 s items with per-item lengths, PESQ + STOI against the
    oracle run on every slice (the reference called per item: legitimate by batch invariance)."""
    from fast_speech_enhancement_metrics_b200.synth import synth_item
    rng = np.random.default_rng(4321)
    lens = [int(x) for x in rng.integers(16000, 480001, size=24)]
    lens[0], lens[1] = 480000, 16000
    nmax = max(lens)
    clean = np.zeros((len(lens), nmax), np.float32)
    deg = np.zeros_like(clean)
    for i, n in enumerate(lens):
        clean[i, :n], deg[i, :n], _ = synth_item(rng, n)
        clean[i, n:] = 7.0          # garbage beyond the valid length must not matter
        deg[i, n:] = -3.0
    c, d = torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda()
    got_p = np.array([r["PESQ"] for r in pesq(c, d, lengths=lens)])
    st = stoi_metrics(16000)
    res = st(c, d, lengths=lens)
    got_s = np.array([r["STOI"] for r in res]); got_e = np.array([r["ESTOI"] for r in res])
    kept = st.last_kept_frames.numpy()
    want_p = po.pesq_batch(clean, deg, lens)
    want_s, want_e, want_k = so.stoi_batch(clean, deg, 16000, lens)
    rep = {"pesq": _maxdiff(got_p, want_p), "stoi": _maxdiff(got_s, want_s), "estoi": _maxdiff(got_e, want_e),
           "K_equal": bool(np.array_equal(kept, want_k))}
    _report("variable_length", rep)
    # PESQ's asymmetry factor jumps from 0 to 3 at ratio^1.2 = 3 (PESQ.py:215): a band/frame whose ratio lies within
    # float32 noise of that threshold can flip and move the score by a few 1e-4 (about one item in fifty here);
    # everything else agrees with the oracle to 2e-5.
    dp = np.abs(got_p - want_p)
    assert rep["pesq"] <= 1e-3 and int(np.sum(dp > 2e-4)) <= 1 and np.median(dp) <= 2e-5
    assert rep["stoi"] <= 1e-4 and rep["estoi"] <= 1e-4 and rep["K_equal"]
    # the host entry point (chunked uploads) gives the same numbers
    host_p = np.array([r["PESQ"] for r in pesq(torch.from_numpy(clean), torch.from_numpy(deg), lengths=lens)])
    assert np.array_equal(host_p, got_p)


def test_large_batch_sample_against_oracle(pesq, stoi_metrics):
    """BASELINE configs[1] / [2] at full batch (256 x 10 s PESQ, 1024 x 4 s STOI): every item is scored on the
    GPU, a seeded sample of items is checked against the oracle."""
    from fast_speech_enhancement_metrics_b200.synth import synth_batch
    base_c, base_d, _ = synth_batch(555, 16, 160000)
    c = torch.from_numpy(base_c).cuda().repeat(16, 1)
    d = torch.from_numpy(base_d).cuda().repeat(16, 1)
    gains = torch.linspace(0.2, 3.0, 256, device="cuda")[:, None]
    c, d = (c * gains).contiguous(), (d * gains).contiguous()            # level alignment must cancel the gain
    got = np.array([r["PESQ"] for r in pesq(c, d)])
    pick = [0, 17, 100, 255]
    want = po.pesq_batch(c[pick].cpu().numpy(), d[pick].cpu().numpy())
    assert _maxdiff(got[pick], want) <= 2e-4
    # the same pair at 16 different input gains: identical up to float32 rounding, except where that rounding flips one of
    # PESQ's hard thresholds (asymmetry factor, silent-frame flag) -- rare and bounded by a few 1e-4
    dg = np.abs(got.reshape(16, 16) - got[:16][None, :])
    assert dg.max() <= 1e-3 and np.median(dg) <= 1e-5 and np.mean(dg > 5e-5) <= 0.05
    st = stoi_metrics(16000)
    c4 = c[:, :64000].repeat(4, 1).contiguous()
    d4 = d[:, 32000:96000].repeat(4, 1).contiguous() * 0.5 + c4 * 0.5
    res = st(c4, d4)
    pick = [3, 300, 777, 1023]
    ws, we, wk = so.stoi_batch(c4[pick].cpu().numpy(), d4[pick].cpu().numpy(), 16000)
    assert _maxdiff(np.array([res[i]["STOI"] for i in pick]), ws) <= 1e-4
    assert _maxdiff(np.array([res[i]["ESTOI"] for i in pick]), we) <= 1e-4
    assert np.array_equal(st.last_kept_frames.numpy()[pick], wk)


def test_readme_example_anchors(pesq, stoi_metrics):
    """The reference's README example (4 x 10 s of torch.randn, README.md:20-34) scored by the reference's CPU path
    in the survey container (SURVEY.md 8c "smoke anchors"; torch 2.11 CPU generator, seed 0)."""
    torch.manual_seed(0)
    clean = torch.randn(4, 160000)
    noisy = torch.randn(4, 160000)
    noisy2 = clean + 0.3 * torch.randn(4, 160000)
    got = np.array([r["PESQ"] for r in pesq(clean, noisy)])
    assert np.max(np.abs(got - np.array([1.55451312, 1.53621908, 1.57332395, 1.54720455]))) <= 1e-3
    res = stoi_metrics(16000)(clean, noisy)
    assert np.max(np.abs(np.array([r["STOI"] for r in res]) - np.array([0.00901818, -0.00190712, -0.00193275, 0.01228937]))) <= 1e-4
    assert np.max(np.abs(np.array([r["ESTOI"] for r in res]) - np.array([0.00474734, -0.00370289, -0.00057269, 0.00886550]))) <= 1e-4
    # second pair: values of the reference's CPU path run in the build container on exactly these tensors
    # (the survey quoted them only approximately)
    got2 = np.array([r["PESQ"] for r in pesq(clean, noisy2)])
    assert np.max(np.abs(got2 - np.array([4.097141528615998, 4.073677243367962, 4.149384423594616, 4.110026437722367]))) <= 1e-3
    res2 = stoi_metrics(16000)(clean, noisy2)
    assert np.max(np.abs(np.array([r["STOI"] for r in res2]) - np.array([0.9098063111, 0.9109758735, 0.9116464257, 0.9117136002]))) <= 1e-4
    assert np.max(np.abs(np.array([r["ESTOI"] for r in res2]) - np.array([0.9027701020, 0.9035763741, 0.9050849676, 0.9054367542]))) <= 1e-4


def test_random_shapes_against_oracle(pesq, stoi_metrics):
    """Seeded sweep over awkward shapes: batches that are not multiples of 32, lengths that are not multiples of
    4 / 32 / 64 / 256, row pitches wider than n (non-contiguous views), misaligned bases, CPU and CUDA inputs,
    with and without per-item lengths.  Every item is checked against the oracle."""
    from fast_speech_enhancement_metrics_b200.synth import synth_item
    rng = np.random.default_rng(2024)
    st16 = stoi_metrics(16000)
    worst = {"pesq": 0.0, "stoi": 0.0, "estoi": 0.0}
    for trial in range(10):
        b = int(rng.integers(1, 40))
        n = int(rng.integers(5700, 30000))
        pitch = n + int(rng.integers(0, 7))
        off = int(rng.integers(0, 4))
        clean = np.zeros((b, n), np.float32); deg = np.zeros_like(clean)
        for i in range(b):
            clean[i], deg[i], _ = synth_item(rng, n)
        use_len = bool(trial % 2)
        lens = [int(x) for x in rng.integers(5700, n + 1, size=b)] if use_len else None
        if lens is not None:
            lens[0] = n
        # strided / misaligned device views of the same data
        buf_c = torch.zeros(b * pitch + 8, device="cuda"); buf_d = torch.zeros_like(buf_c)
        vc = buf_c[off:off + b * pitch].view(b, pitch)[:, :n]
        vd = buf_d[off:off + b * pitch].view(b, pitch)[:, :n]
        vc.copy_(torch.from_numpy(clean)); vd.copy_(torch.from_numpy(deg))
        if trial % 3 == 2:                       # host path
            vc, vd = torch.from_numpy(clean), torch.from_numpy(deg)
        got_p = np.array([r["PESQ"] for r in pesq(vc, vd, lengths=lens)])
        res = st16(vc, vd, lengths=lens)
        want_p = po.pesq_batch(clean, deg, lens)
        ws, we, wk = so.stoi_batch(clean, deg, 16000, lens)
        worst["pesq"] = max(worst["pesq"], _maxdiff(got_p, want_p))
        worst["stoi"] = max(worst["stoi"], _maxdiff(np.array([r["STOI"] for r in res]), ws))
        worst["estoi"] = max(worst["estoi"], _maxdiff(np.array([r["ESTOI"] for r in res]), we))
        assert np.array_equal(st16.last_kept_frames.numpy(), wk), (trial, b, n)
    _report("random_shapes", worst)
    assert worst["pesq"] <= 1e-3 and worst["stoi"] <= 1e-4 and worst["estoi"] <= 1e-4


# ------------------------------------------------------------------ ingest formats
def _as_dtype(x, dtype):
    """float32 synth signal -> (tensor of `dtype`, the float32 tensor holding exactly the same values)."""
    t = torch.from_numpy(x)
    narrow = (t * 20000.0).round().clamp(-32768, 32767).to(torch.int16) if dtype == torch.int16 else t.to(torch.float16)
    return narrow, narrow.to(torch.float32)


@pytest.mark.parametrize("dtype", [torch.int16, torch.float16])
def test_ingest_kernel_is_a_value_preserving_cast(dtype):
    import ctypes as C
    from fast_speech_enhancement_metrics_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    for rows, n, pad in ((3, 4096, 0), (5, 1001, 0), (4, 777, 3), (1, 5, 0), (2, 8, 8)):
        src = (torch.randn(rows, n + pad, generator=g) * 3000).to(dtype).cuda()
        view = src[:, :n]                                              # pitch n + pad elements: odd pitches take the scalar path
        out = torch.full((rows, n + 2), -7.0, dtype=torch.float32, device="cuda")
        _lib.check(lib.fsem_ingest_f32(view.data_ptr(), _lib.dtype_code(dtype), rows, n, view.stride(0) if rows > 1 else n + pad,
                                       out.data_ptr(), out.stride(0), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        assert torch.equal(out[:, :n], view.to(torch.float32))
        assert bool((out[:, n:] == -7.0).all())                        # nothing written beyond the row
    assert lib.fsem_ingest_f32(None, 1, 1, 1, 1, None, 1, None) == _lib.FSEM_E_INVALID
    x = torch.zeros(8, device="cuda")
    assert lib.fsem_ingest_f32(x.data_ptr(), 9, 1, 8, 8, x.data_ptr(), 8, None) == _lib.FSEM_E_INVALID


@pytest.mark.parametrize("dtype", [torch.int16, torch.float16])
def test_int16_and_fp16_inputs_score_like_their_float32_values(dtype, pesq, stoi_metrics):
    from fast_speech_enhancement_metrics_b200 import LSD, SDR, score_pesq_stoi
    clean, deg, _, fs = STOI_CASES["speech16k_3s"]                      # 8 x 48000
    st = stoi_metrics(16000)
    cn, cf = _as_dtype(clean, dtype)
    dn, df = _as_dtype(deg, dtype)
    lens = [48000, 30000, 20001, 48000, 12345, 40000, 48000, 8000]
    for lengths in (None, lens):
        want_p, want_s = pesq(cf, df, lengths=lengths), st(cf, df, lengths=lengths)
        for cc, dd in ((cn, dn), (cn.cuda(), dn.cuda())):               # host pipeline (2-byte upload) and device tensors
            assert pesq(cc, dd, lengths=lengths) == want_p
            got_s = st(cc, dd, lengths=lengths)
            assert np.array_equal(np.array([[r["STOI"], r["ESTOI"]] for r in got_s]),
                                  np.array([[r["STOI"], r["ESTOI"]] for r in want_s]), equal_nan=True)
            both = score_pesq_stoi(pesq, st, cc, dd, lengths=lengths)
            assert [r["PESQ"] for r in both] == [r["PESQ"] for r in want_p]
            assert np.array_equal(np.array([r["ESTOI"] for r in both]), np.array([r["ESTOI"] for r in want_s]),
                                  equal_nan=True)
    # rows that are not a multiple of 8 samples (ragged 16-byte groups on the way in) and a strided host batch
    c2, d2 = cn[:5, :40001], dn[:5, :40001]
    ref_p, ref_s = pesq(cf[:5, :40001], df[:5, :40001]), st(cf[:5, :40001], df[:5, :40001])
    got_p, got_s = pesq(c2, d2), st(c2, d2)
    assert _maxdiff([r["PESQ"] for r in got_p], [r["PESQ"] for r in ref_p]) <= 1e-5
    assert _maxdiff([r["STOI"] for r in got_s], [r["STOI"] for r in ref_s]) <= 1e-6
    # the adjacent metrics take the same formats
    for metric in (LSD(16000, use_gpu=True), SDR(16000, use_gpu=True)):
        assert metric(cn, dn) == metric(cf, df)
        assert metric(cn.cuda(), dn.cuda()) == metric(cf, df)
    with pytest.raises(Exception, match="same dtype"):
        pesq(cn, df)
    with pytest.raises(RuntimeError, match="expected scalar type Float"):
        pesq(cf.double(), df.double())


@pytest.mark.parametrize("dtype", [torch.int16, torch.float16])
def test_16_bit_device_tensors_are_read_by_the_first_kernels(dtype, pesq, stoi_metrics):
    """Device-resident int16 / float16 batches: no widening pass (`ingest_kernel` is not launched), the IIR pass and the
    resampler read the rows in their own dtype; scores are bit-identical to the float32 tensor holding the same values.
    Also the corners: batches below 16 items (thread-per-chunk IIR kernel), rows that are not 8-byte aligned (scalar
    loads), STOI on 10 kHz input (the one path that still widens, inside the library) and PESQ at 8 kHz (typed general
    resampler)."""
    from fast_speech_enhancement_metrics_b200 import PESQ, _lib, score_pesq_stoi_tensors
    from fast_speech_enhancement_metrics_b200.synth import synth_batch
    clean, deg, _ = synth_batch(31, 40, 24000)
    cn, cf = _as_dtype(clean, dtype)
    dn, df = _as_dtype(deg, dtype)
    cn, dn, cf, df = cn.cuda(), dn.cuda(), cf.cuda(), df.cuda()
    st = stoi_metrics(16000)
    _lib.profile_reset()
    _lib.profile_enable(True)
    got, _, kept, _ = score_pesq_stoi_tensors(pesq, st, cn, dn)
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    prof = _lib.profile_read()
    assert prof["ingest_kernel"][1] == 0, prof
    assert prof["pesq_filter_kernel"][1] == 1 and prof["stoi_resample_kernel"][1] == 1
    want, _, kept_f, _ = score_pesq_stoi_tensors(pesq, st, cf, df)
    assert torch.equal(got.view(torch.int32), want.view(torch.int32)) and torch.equal(kept, kept_f)
    # small batch (thread-per-(signal, chunk) IIR kernel) and unaligned rows (odd offset inside a wider buffer)
    for sl in (slice(0, 5), slice(0, 40)):
        for off in (0, 1):
            a, b = cn[sl, off:off + 20001], dn[sl, off:off + 20001]
            af, bf = cf[sl, off:off + 20001], df[sl, off:off + 20001]
            assert pesq(a, b) == pesq(af, bf)
            assert st(a, b) == st(af, bf)
    # 10 kHz input: STOI has no resampler in front, the library widens into its workspace
    s10 = stoi_metrics(10000)
    assert s10(cn[:6], dn[:6]) == s10(cf[:6], df[:6])
    # PESQ at 8 kHz: the general polyphase resampler reads the 16-bit rows
    p8 = PESQ(8000, use_gpu=True)
    assert p8(cn[:6, :16000], dn[:6, :16000]) == p8(cf[:6, :16000], df[:6, :16000])


def test_pesq_ragged_batch_with_empty_and_short_items(pesq):
    """Variable-length batches go through the length-sorted IIR pass and the frame-prefix work split of the
    spectrum kernel: items without a single frame (T = 0) and items below 20 frames must neither disturb their
    neighbours nor the unit bookkeeping, whatever their position in the batch."""
    from fast_speech_enhancement_metrics_b200 import _lib
    from fast_speech_enhancement_metrics_b200.synth import synth_batch
    b, n = 70, 48000
    clean, deg, _ = synth_batch(977, b, n)
    rng = np.random.default_rng(3)
    lens = rng.integers(5376, n + 1, size=b)
    lens[[0, 7, 8, 9, 33, 68, 69]] = [0, 100, 255, 0, 3000, 5120, 1]       # T = 0, 0, 0, 0, 10, 19, 0
    lens[[1, 34]] = [n, 5376]
    c, d = torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda()
    mos, status = pesq.score_tensors(c, d, lens.tolist())
    mos, status = mos.cpu().numpy(), status.cpu().numpy()
    padded = lens + lens % 256                                              # PESQ.py:128-133
    short = np.where(padded < 512, 0, 1 + (padded - 512) // 256) < 20
    assert short.sum() == 7
    assert np.array_equal(status == _lib.ITEM_TOO_SHORT, short)
    assert np.isnan(mos[short]).all() and np.isfinite(mos[~short]).all()
    for i in np.flatnonzero(~short)[::3]:
        one, _ = pesq.score_tensors(c[i:i + 1, :lens[i]].contiguous(), d[i:i + 1, :lens[i]].contiguous())
        assert abs(float(one[0]) - mos[i]) <= 1e-5, (i, lens[i], float(one[0]), mos[i])
    for i in np.flatnonzero(~short)[:8]:
        want = po.pesq_item(clean[i, :lens[i]], deg[i, :lens[i]])
        assert abs(want - mos[i]) <= 2e-4, (i, lens[i], want, mos[i])


@pytest.mark.parametrize("shape", [(4, 160000, False, 0), (6, 48000, True, 2), (40, 40004, False, 3), (70, 33000, True, 16)])
def test_captured_scorer_against_oracle_and_direct_calls(shape, pesq, stoi_metrics):
    """CUDA-graph path (fsem_graph_*: the PESQ and STOI chains as parallel branches of one graph): scores against the
    float64 ORACLE, bit-identical to the direct entry point, and replays follow in-place updates of the captured
    buffers (the README shape 4 x 10 s is the case the graph exists for)."""
    from fast_speech_enhancement_metrics_b200 import CapturedScorer, score_pesq_stoi_tensors
    from fast_speech_enhancement_metrics_b200.synth import synth_batch
    b, n, ragged, slices = shape
    st = stoi_metrics(16000)
    clean, deg, _ = synth_batch(777 + b, b, n)
    clean2, deg2, _ = synth_batch(888 + b, b, n)
    lens = None
    if ragged:
        lens = np.random.default_rng(b).integers(20000, n + 1, size=b).tolist()
        lens[0] = n
    c, d = torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda()
    scorer = CapturedScorer(pesq, st, c, d, lens, slices=slices)       # scores must not depend on the slicing
    assert scorer.slices == (slices or 1) and scorer.kernel_nodes >= 10 * scorer.slices
    for cc, dd in ((clean, deg), (clean2, deg2), (clean, deg)):
        c.copy_(torch.from_numpy(cc)); d.copy_(torch.from_numpy(dd))
        rows = scorer()
        direct, pst, kept, _ = score_pesq_stoi_tensors(pesq, st, c, d, lens)
        got = scorer.scores.cpu().numpy()
        assert np.array_equal(got, direct.cpu().numpy(), equal_nan=True)
        assert np.array_equal(scorer.kept_frames.cpu().numpy(), kept.cpu().numpy())
        want_p = po.pesq_batch(cc, dd, lens)
        ws, we, wk = so.stoi_batch(cc, dd, 16000, lens)
        assert _maxdiff(got[0].astype(np.float64), want_p) <= 2e-4
        assert _maxdiff(got[1].astype(np.float64), ws) <= 1e-4 and _maxdiff(got[2].astype(np.float64), we) <= 1e-4
        assert np.array_equal(scorer.kept_frames.cpu().numpy(), wk)
        assert [r["PESQ"] for r in rows] == got[0].tolist() and set(rows[0]) == {"PESQ", "STOI", "ESTOI"}
    scorer.close()
    with pytest.raises(Exception):
        scorer.replay()


def test_captured_scorer_single_metric_and_errors(pesq, stoi_metrics):
    from fast_speech_enhancement_metrics_b200 import CapturedScorer
    clean, deg, _, fs = STOI_CASES["speech16k_3s"]
    c, d = torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda()
    st = stoi_metrics(16000)
    only_p = CapturedScorer(pesq, None, c, d)
    only_s = CapturedScorer(None, st, c, d)
    assert only_p() == pesq(c, d)
    assert only_s() == st(c, d)
    assert only_p.kernel_nodes == 3 and only_s.kernel_nodes >= 6
    with pytest.raises(Exception):
        CapturedScorer(None, None, c, d)
    with pytest.raises(Exception):
        CapturedScorer(pesq, st, c.cpu(), d.cpu())                 # host tensors cannot be captured
    with pytest.raises(Exception):
        CapturedScorer(pesq, st, c[:, ::2], d[:, ::2])             # would be copied: not in place
    with pytest.raises(RuntimeError):
        CapturedScorer(pesq, None, c[:, :4000], d[:, :4000])       # < 20 PESQ frames, like the direct call
    after = pesq(c, d)                                             # a failed capture leaves the library usable
    assert after == only_p()


@pytest.mark.parametrize("shape", [(5, 24000, False), (33, 20004, False), (37, 17003, True)])
def test_kernels_stay_inside_their_buffers(shape, pesq, stoi_metrics):
    """compute-sanitizer is not available on the GPU pool, so bounds are checked by hand: the inputs are views into a
    NaN-filled buffer (NaN rows before and after, NaN in the pitch padding of every row) and the workspaces are views
    into canary-filled buffers.  Any read outside the rows that reaches a score turns it into NaN or moves it away from
    the oracle; any write outside the planned workspace breaks a canary."""
    from fast_speech_enhancement_metrics_b200.synth import synth_batch
    b, n, ragged = shape
    st = stoi_metrics(16000)
    clean, deg, _ = synth_batch(99 + b, b, n)
    lens = None
    if ragged:
        lens = np.random.default_rng(b).integers(6000, n + 1, size=b).tolist()
        lens[0], lens[-1] = n, 6000
    pitch = n + 52                                              # multiple of 4 (vector paths) for these n, not of 32
    pad_rows = 2

    def guarded(x):
        big = torch.full((b + 2 * pad_rows, pitch), float("nan"), dtype=torch.float32, device="cuda")
        view = big[pad_rows:pad_rows + b, :n]
        view.copy_(torch.from_numpy(x))
        return big, view

    bc, c = guarded(clean)
    bd, d = guarded(deg)
    G = 4096
    for metric, nbytes in ((pesq, pesq._lib.fsem_pesq_workspace_bytes(pesq._ctx, b, n)),
                           (st, st._lib.fsem_stoi_workspace_bytes(st._ctx, b, n))):
        big = torch.full((int(nbytes) + 2 * G,), 0xA5, dtype=torch.uint8, device="cuda")
        saved = metric._workspace
        try:
            metric._workspace = big[G:G + int(nbytes)]
            res = metric(c, d, lengths=lens)
            torch.cuda.synchronize()
        finally:
            metric._workspace = saved
        assert bool((big[:G] == 0xA5).all()) and bool((big[G + int(nbytes):] == 0xA5).all()), type(metric).__name__
        if metric is pesq:
            got = np.array([r["PESQ"] for r in res])
            assert _maxdiff(got, po.pesq_batch(clean, deg, lens)) <= 2e-4
        else:
            ws, we, wk = so.stoi_batch(clean, deg, 16000, lens)
            assert _maxdiff(np.array([r["STOI"] for r in res]), ws) <= 1e-4
            assert _maxdiff(np.array([r["ESTOI"] for r in res]), we) <= 1e-4
    # the inputs (and their NaN surroundings) are untouched
    assert bool(torch.isnan(bc[:pad_rows]).all()) and bool(torch.isnan(bc[pad_rows + b:]).all()) and bool(torch.isnan(bc[:, n:]).all())
    assert np.array_equal(c.cpu().numpy(), clean) and np.array_equal(d.cpu().numpy(), deg)
