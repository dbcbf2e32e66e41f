"""CPU-side tests of the callers either side of the forward: checkpoint
ingestion, the on-disk feature contract and the verb-prediction dict of EvalB."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from common import ROOT, build_model
from vidsitu_b200 import checkpoint as CK
from vidsitu_b200 import feat_io as FIO

PAIRS = json.load(open(os.path.join(ROOT, "tests", "golden", "c2_name_pairs.json")))


@pytest.mark.parametrize("name", sorted(PAIRS))
def test_caffe2_names_match_the_reference_converter(name):
    """Fixture = output of the reference's own regex table (tests/golden/make_c2_names.py)."""
    model, _, _ = build_model(name, seed=0, crop=64)
    keys = set(model.sf_mdl.state_dict().keys())
    for c2, want in PAIRS[name]["pairs"]:
        assert CK.convert_caffe2_name(c2) == want, c2
    covered = {w for _, w in PAIRS[name]["pairs"]}
    assert covered == {k for k in keys if not k.endswith("num_batches_tracked")}
    for c2, ref_out in PAIRS[name]["solver_blobs"]:
        # the reference maps these to strings that are not model keys and skips them (checkpoint.py:246-254)
        assert ref_out not in keys
        got = CK.convert_caffe2_name(c2)
        assert got is None or got not in keys


def _fake_caffe2_ckpt(model, path, seed=3):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    g = np.random.default_rng(seed)
    sd = model.sf_mdl.state_dict()
    inv = {want: c2 for c2, want in PAIRS["slow_fast_nl_r50_8x8"]["pairs"]}
    blobs, truth = {}, {}
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            continue
        a = g.standard_normal(tuple(v.shape)).astype(np.float32)
        if k.endswith("running_var"):
            a = np.abs(a) + 0.5
        blobs[inv[k]] = a
        truth[k] = a
    blobs["lr"] = np.float32(0.1)
    blobs["model_iter"] = np.int64(100)
    blobs["conv1_w_momentum"] = np.zeros((64, 3, 1, 7, 7), np.float32)
    blobs["pred_w"] = np.zeros((7, 2304), np.float32)       # wrong class count -> shape mismatch, not loaded
    truth.pop("head.projection.weight")
    with open(path, "wb") as f:
        pickle.dump({"blobs": blobs}, f, protocol=2)
    return truth


def test_caffe2_checkpoint_loads_into_sf_mdl(tmp_path):
    model, _, _ = build_model("slow_fast_nl_r50_8x8", seed=0, crop=64)
    before = model.sf_mdl.state_dict()["head.projection.weight"].clone()
    truth = _fake_caffe2_ckpt(model, tmp_path / "c2.pkl")
    rep = CK.load_caffe2_checkpoint(tmp_path / "c2.pkl", model.sf_mdl)
    sd = model.sf_mdl.state_dict()
    for k, a in truth.items():
        assert np.array_equal(sd[k].numpy(), a), k
    assert rep["mismatched"] == ["pred_w"]
    assert set(rep["skipped"]) == {"lr", "model_iter", "conv1_w_momentum"}
    assert torch.equal(sd["head.projection.weight"], before)
    assert all(k.endswith("num_batches_tracked") or k == "head.projection.weight" for k in rep["missing"])


def test_vidsitu_pth_with_module_prefix(tmp_path):
    src, _, _ = build_model("i3d_r50_8x8", seed=5, crop=64)
    dst, _, _ = build_model("i3d_r50_8x8", seed=6, crop=64)
    torch.save({"model_state_dict": {"module." + k: v for k, v in src.state_dict().items()}}, tmp_path / "m.pth")
    CK.load_vidsitu_checkpoint(tmp_path / "m.pth", dst)
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    with pytest.raises(IndexError):
        CK.strip_module_prefix("sf_mdl.s1.pathway0_stem.conv.weight")   # rem_mdl raises on un-prefixed keys too


def test_feature_files_are_what_the_reference_writes_and_reads(tmp_path):
    """feat_extractor.py:98-111 writes np.save(out_tdir/{vseg}_feats.npy, out_np[vix]) with out = [B,5,D] fp32;
    dat_loader.py:503-511 reads it back and asserts 5 events."""
    g = torch.Generator().manual_seed(0)
    feats = torch.randn((3 * 5, 2304), generator=g)
    names = ["v_aaa_seg_0_10", "v_bbb_seg_10_20", "v_ccc_seg_20_30"]
    with FIO.FeatureWriter(tmp_path, "slow_fast_nl_r50_8x8_feats") as w:
        w.put(feats, names)
        w.put(feats.view(3, 5, -1) * 2, [n + "_x" for n in names])
    out_dir = tmp_path / "slow_fast_nl_r50_8x8_feats"
    assert w.files_written == 6
    ref_np = feats.view(3, 5, -1).numpy()
    for v, n in enumerate(names):
        p = out_dir / f"{n}_feats.npy"
        # byte-identical to the reference's own np.save call
        import io
        bio = io.BytesIO()
        np.save(bio, ref_np[v])
        assert p.read_bytes() == bio.getvalue()
        t = FIO.read_frm_feats(out_dir, n)
        assert t.dtype == torch.float32 and tuple(t.shape) == (5, 2304)
        assert torch.equal(t, feats.view(3, 5, -1)[v])
        assert torch.equal(FIO.read_frm_feats(out_dir, n + "_x"), 2 * t)
    assert FIO.get_head_dim(str(out_dir)) == 2304
    assert FIO.get_head_dim("/data/vsitu_vid_feats/i3d_r50_8x8") == 2048
    assert FIO.get_head_dim("/x/sfast_feats") == 2304
    with pytest.raises(NotImplementedError):
        FIO.get_head_dim("/x/c2d")
    with pytest.raises(AssertionError):
        FIO.read_frm_feats(out_dir, "missing")
    with pytest.raises(ValueError):
        FIO.write_video_feats(out_dir, "bad", np.zeros((4, 2304), np.float32))


def test_feature_writer_rejects_wrong_shapes(tmp_path):
    w = FIO.FeatureWriter(tmp_path)
    with pytest.raises(ValueError):
        w.put(torch.zeros((7, 16)), ["a"])
    with pytest.raises(ValueError):
        w.put(torch.zeros((5, 16), dtype=torch.float64), ["a"])
    w.close()


def _write_video(tdir, vseg, total, rng):
    from PIL import Image
    d = tdir / vseg
    d.mkdir(parents=True)
    for ix in range(1, total + 1):
        arr = rng.integers(0, 256, size=(48, 64, 3), dtype=np.uint8)
        Image.fromarray(arr).save(d / f"{vseg}_{ix:06d}.jpg", quality=95)


def test_frame_ingest_matches_the_reference_reader(tmp_path):
    """frames_io: the `{vseg}_{i:06d}.jpg` naming (dat_loader.py:455-458), `read_img` (:183-191: open, RGB,
    resize with PIL's default filter), and the whole-video tensor holding exactly the frames the event windows
    select, at their own indices."""
    from PIL import Image
    from oracle import sf_oracle as O
    from vidsitu_b200 import frames_io as F
    rng = np.random.default_rng(0)
    _write_video(tmp_path, "v_abc_seg_1", 300, rng)
    paths = F.frame_paths(tmp_path, "v_abc_seg_1")
    assert len(paths) == 300 and paths[0].name == "v_abc_seg_1_000001.jpg" and paths[299].name == "v_abc_seg_1_000300.jpg"
    ref = np.array(Image.open(paths[7]).convert("RGB").resize((224, 224)))            # the reference's three calls
    assert np.array_equal(F.read_img(paths[7]), ref) and ref.shape == (224, 224, 3)
    need = F.needed_frames(32, 2)
    windows = O.event_frame_indices(32, 2)
    assert need == sorted({i for w in windows for i in w}) and len(need) <= 160
    vid = F.load_video(tmp_path, "v_abc_seg_1", need, size=32)
    assert vid.dtype == torch.uint8 and tuple(vid.shape) == (300, 32, 32, 3)
    for ev, w in enumerate(windows):                                                 # every window frame is there
        clip = np.stack([F.read_img(paths[i], 32) for i in w])
        assert np.array_equal(vid[w].numpy(), clip), ev
    unused = sorted(set(range(300)) - set(need))
    assert unused and not vid[unused].any()
    ds = F.VideoFrames(tmp_path, ["v_abc_seg_1"], 32, 2, size=32)
    frames, idxs = F.collate_videos([ds[0]])
    assert tuple(frames.shape) == (1, 300, 32, 32, 3) and idxs == [0]
    (tmp_path / "split.json").write_text('["v_abc_seg_1"]')
    assert F.read_vseg_list(tmp_path / "split.json") == ["v_abc_seg_1"]
    with pytest.raises(AssertionError):
        F.load_video(tmp_path, "missing", need, size=32)


def test_caffe2_checkpoint_loads_into_a_sub_batchnorm_model(tmp_path):
    """BN.NORM_TYPE sub_batchnorm: Caffe2 running statistics go to `<bn>.split_bn.running_*`, tiled NUM_SPLITS times
    (checkpoint.py:215-232, c2_normal_to_sub_bn :331-348), and the aggregated `<bn>.bn.*` the kernels fold follow."""
    import json
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_meta.json")))["sf50_subbn_n2_64"]
    plain, _, _ = build_model("slow_fast_nl_r50_8x8", seed=0, crop=64)
    truth = _fake_caffe2_ckpt(plain, tmp_path / "c2.pkl")
    sub, _, _ = build_model("slow_fast_nl_r50_8x8", seed=3, crop=64, sf_overrides=meta["sf_overrides"])
    rep = CK.load_caffe2_checkpoint(tmp_path / "c2.pkl", sub.sf_mdl)
    assert set(rep["skipped"]) == {"lr", "model_iter", "conv1_w_momentum"}
    sd = sub.sf_mdl.state_dict()
    n_split = meta["sf_overrides"]["BN"]["NUM_SPLITS"]
    checked = 0
    for k, a in truth.items():
        if ".running_" in k:
            pre, leaf = k.rsplit(".", 1)
            assert np.array_equal(sd[f"{pre}.split_bn.{leaf}"].numpy(), np.concatenate([a] * n_split)), k
            # identical splits: aggregated mean = the split mean, aggregated var = the split var
            assert np.allclose(sd[f"{pre}.bn.{leaf}"].numpy(), a, rtol=1e-6, atol=1e-6), k
            checked += 1
        else:
            assert np.array_equal(sd[k].numpy(), a), k
    assert checked > 200


# ---------------------------------------------------------------- frame ingest oracle against Pillow
def _pil_rgb(data: bytes):
    import io
    from PIL import Image
    return np.array(Image.open(io.BytesIO(data)).convert("RGB"))


@pytest.mark.parametrize("case", __import__("common").JPEG_CASES)
def test_jpeg_decode_oracle_is_bit_exact_with_pillow(case):
    """oracle/image_oracle.py::decode_jpeg (Huffman + islow IDCT + fancy upsampling + YCbCr tables) equals
    `Image.open(..).convert("RGB")` (libjpeg-turbo) bit for bit: 4:4:4 / 4:2:2 / 4:2:0, odd sizes, partial MCUs."""
    from common import jpeg_bytes
    from oracle import image_oracle as IO
    data = jpeg_bytes(*case)
    assert np.array_equal(IO.decode_jpeg(data), _pil_rgb(data))


def test_jpeg_decode_oracle_grayscale_restart_markers_and_refusals():
    import io
    from PIL import Image
    from common import jpeg_bytes, synthetic_image
    from oracle import image_oracle as IO
    data = jpeg_bytes(40, 56, 90, 2, "noisy", gray=True)
    assert np.array_equal(IO.decode_jpeg(data), _pil_rgb(data))
    buf = io.BytesIO()
    Image.fromarray(synthetic_image(70, 90, "noisy", 3)).save(buf, "JPEG", quality=88, subsampling=2, restart_marker_blocks=3)
    assert np.array_equal(IO.decode_jpeg(buf.getvalue()), _pil_rgb(buf.getvalue()))
    buf = io.BytesIO()
    Image.fromarray(synthetic_image(64, 64, "noisy", 4)).save(buf, "JPEG", quality=88, progressive=True)
    with pytest.raises(IO.JpegUnsupported):
        IO.decode_jpeg(buf.getvalue())


@pytest.mark.parametrize("shape", [(360, 640), (240, 320), (224, 224), (80, 100), (500, 333), (224, 640), (37, 53)])
def test_resize_oracle_is_bit_exact_with_pillow(shape):
    """resize_bicubic_u8 == `img.resize((224, 224))` (Pillow's default BICUBIC, fixed-point two-pass)."""
    from PIL import Image
    from common import synthetic_image
    from oracle import image_oracle as IO
    img = synthetic_image(*shape, "noisy", seed=shape[0])
    assert np.array_equal(IO.resize_bicubic_u8(img, 224, 224), np.array(Image.fromarray(img).resize((224, 224))))


def test_read_img_oracle_matches_the_reference_reader(tmp_path):
    """decode + resize == VsituDS.read_img (dat_loader.py:183-191) on a file, through the package's host reader."""
    from common import jpeg_bytes
    from oracle import image_oracle as IO
    from vidsitu_b200 import frames_io
    data = jpeg_bytes(360, 640, 96, 2, "noisy", seed=9)
    p = tmp_path / "f.jpg"
    p.write_bytes(data)
    assert np.array_equal(IO.read_img(data), frames_io.read_img(p))
