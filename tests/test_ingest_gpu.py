"""GPU parity of the raw-feature ingest kernels (csrc/ingest.cu through the C ABI) against the oracle's restatement of
video_nmn/dataset.py:134-172.  fp32 out: |err| <= 1e-6 * max|ref| (the mean is summed in frame order, torch sums in another order);
bf16 out: equal to the fp32 reference rounded to bf16 up to 1 ulp.  Also the size-independent property at the full bench shape:
the pooled features of a constant-per-clip input are that constant, and motion columns are copied bit-exactly."""
import numpy as np
import pytest
import torch

from oracle import ingest_oracle as orc
from stair_b200 import ingest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('T,F,Da,Dm,max_len', [(8, 16, 2048, 2048, 8), (8, 16, 256, 128, 8), (10, 5, 64, 0, 8), (3, 1, 8, 8, 8), (8, 16, 2048, 2048, 150)])
def test_pool_concat_matches_reference_loader(T, F, Da, Dm, max_len):
    rng = np.random.default_rng(T * 1000 + F)
    B = 5
    app = np.abs(rng.standard_normal((B, T, F, Da))).astype(np.float32)
    mot = np.abs(rng.standard_normal((B, T, Dm))).astype(np.float32) if Dm else None
    want = torch.stack([orc.rx_video_features(app[b], mot[b] if Dm else None, max_len) for b in range(B)])
    a, m = torch.from_numpy(app).cuda(), (torch.from_numpy(mot).cuda() if Dm else None)
    got32 = ingest.pool_concat(a, m, out_dtype=torch.float32, max_video_length=max_len).cpu()
    assert got32.shape == want.shape
    assert float((got32 - want).abs().max()) <= 1e-6 * float(want.abs().max())
    if Dm:
        assert torch.equal(got32[..., Da:], want[..., Da:])                    # concat is a copy
    got16 = ingest.pool_concat(a, m, out_dtype=torch.bfloat16, max_video_length=max_len).float().cpu()
    ref16 = want.to(torch.bfloat16).float()
    assert float((got16 - ref16).abs().max()) <= 2 ** -7 * float(want.abs().max())
    # bf16 inputs (features stored in bf16 on the device)
    got_b = ingest.pool_concat(a.to(torch.bfloat16), m.to(torch.bfloat16) if Dm else None, out_dtype=torch.float32, max_video_length=max_len).cpu()
    want_b = torch.stack([orc.rx_video_features(a[b].to(torch.bfloat16).float().cpu().numpy(),
                                                m[b].to(torch.bfloat16).float().cpu().numpy() if Dm else None, max_len) for b in range(B)])
    assert float((got_b - want_b).abs().max()) <= 1e-6 * float(want_b.abs().max())


@pytest.mark.parametrize('n,D,max_len', [(128, 1024, 64), (37, 64, 64), (300, 1024, 64), (1, 8, 4)])
def test_subsample_matches_reference_loader(n, D, max_len):
    rng = np.random.default_rng(n)
    feats = rng.standard_normal((3, n, D)).astype(np.float32)
    want = torch.stack([orc.i3d_video_features(feats[b], max_len).reshape(-1, D) for b in range(3)])
    got = ingest.subsample(torch.from_numpy(feats).cuda(), max_len, out_dtype=torch.float32).cpu()
    assert torch.equal(got, want)


def test_full_size_properties():
    B, T, F, D = 512, 8, 16, 2048
    base = torch.rand(B, T, 1, D, device='cuda')
    app = base.expand(B, T, F, D).contiguous()                                 # constant over the frames of a clip
    mot = torch.rand(B, T, D, device='cuda')
    out = ingest.pool_concat(app, mot, out_dtype=torch.float32)
    assert float((out[..., :D] - base[:, :, 0]).abs().max()) <= 1e-6
    assert torch.equal(out[..., D:], mot)
    # linearity: pool(a + b) == pool(a) + pool(b) up to fp32 rounding
    a2 = torch.rand(B, T, F, D, device='cuda')
    lhs = ingest.pool_concat(app + a2, None, out_dtype=torch.float32)
    rhs = ingest.pool_concat(app, None, out_dtype=torch.float32) + ingest.pool_concat(a2, None, out_dtype=torch.float32)
    assert float((lhs - rhs).abs().max()) <= 1e-5


def test_ingest_feeds_the_model():
    """collate() accepts raw appearance / motion features and the forward equals the forward on host-pooled video_features."""
    from stair_b200 import VideoNMN, synthetic as syn, collate
    T, Da, Dm = 8, 64, 64
    cfg = syn.model_config(T=T, V=Da + Dm, hidden=64)
    torch.manual_seed(0)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='fp32').cuda().eval()
    qs = syn.make_questions(12, T, Da + Dm, seed=3, templates=list(syn.ALL_TEMPLATES))
    rng = np.random.default_rng(0)
    raw = []
    for d in qs:
        app = np.abs(rng.standard_normal((T, 16, Da))).astype(np.float32)
        mot = np.abs(rng.standard_normal((T, Dm))).astype(np.float32)
        d2 = dict(d)
        d2.pop('video_features')
        d2['appearance_features'], d2['motion_features'] = torch.from_numpy(app), torch.from_numpy(mot)
        raw.append(d2)
        d['video_features'] = orc.rx_video_features(app, mot, T)
    a = model(qs, return_res_by_step=False, test_mode=True)
    b = model(collate(raw, video_dtype=torch.float32), return_res_by_step=False, test_mode=True)
    torch.cuda.synchronize()
    assert float((a['logits'] - b['logits']).abs().max()) <= 2e-5 * float(a['logits'].abs().max())
    assert torch.equal(a['answers'], b['answers'])


def test_npy_directory_reader_on_the_device_equals_the_host_reader(tmp_path):
    """stair_b200.agqa_data: the I3D npy reader (dataset.py:134-143) with the stride-2 / truncate reduction on the GPU
    (stair_ingest_subsample) == the host reader, for files of several lengths."""
    from stair_b200 import agqa_data as AD
    rng = np.random.default_rng(1)
    ids = []
    for i, n in enumerate((130, 130, 40, 7, 200)):
        vid = 'v%d' % i
        ids.append(vid)
        np.save(tmp_path / (vid + '.npy'), rng.standard_normal((n, 1024)).astype(np.float32))
    host = AD.load_npy_features(str(tmp_path), ids, 64)
    dev = AD.load_npy_features_device(str(tmp_path), ids, 64, out_dtype=torch.float32)
    assert set(host) == set(dev)
    for k in host:
        assert torch.equal(dev[k].cpu(), host[k])
    dev16 = AD.load_npy_features_device(str(tmp_path), ids[:2], 64)
    assert dev16['v0'].dtype == torch.bfloat16 and torch.equal(dev16['v0'].cpu(), host['v0'].to(torch.bfloat16))
