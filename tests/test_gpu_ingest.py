"""Frame ingest on the GPU (vsb_jpeg_*, csrc/jpeg_ingest.cu): bit-exact against Pillow - the
reference's reader `Image.open(p).convert("RGB").resize((224, 224))` (dat_loader.py:183-191) - and against the numpy
oracle, on 4:4:4 / 4:2:2 / 4:2:0 / grayscale files, odd sizes, restart markers; refusals are loud."""
import io

import numpy as np
import pytest
import torch

from common import JPEG_CASES, jpeg_bytes, synthetic_image

pytestmark = pytest.mark.gpu


def _pil_read_img(data: bytes, size: int = 224):
    from PIL import Image
    return np.array(Image.open(io.BytesIO(data)).convert("RGB").resize((size, size)))


@pytest.fixture(scope="module")
def decoder():
    from vidsitu_b200.jpeg import JpegDecoder
    return JpegDecoder(1920, 1088)


@pytest.mark.parametrize("case", JPEG_CASES)
def test_decode_resize_is_bit_exact_with_the_reference_reader(decoder, case):
    from oracle import image_oracle as IO
    data = jpeg_bytes(*case)
    out = torch.empty((224, 224, 3), dtype=torch.uint8, device="cuda")
    decoder.decode_resize(data, out)
    got = out.cpu().numpy()
    assert np.array_equal(got, _pil_read_img(data))
    assert np.array_equal(got, IO.read_img(data))


def test_native_size_decode_grayscale_and_restart_markers(decoder):
    from PIL import Image
    for data in (jpeg_bytes(40, 56, 90, 2, "noisy", gray=True), jpeg_bytes(224, 224, 97, 2, "noisy", seed=5)):
        w, h = Image.open(io.BytesIO(data)).size
        out = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
        decoder.decode_resize(data, out)           # no resampling: the decoder's pixels themselves
        assert np.array_equal(out.cpu().numpy(), np.array(Image.open(io.BytesIO(data)).convert("RGB")))
    buf = io.BytesIO()
    Image.fromarray(synthetic_image(170, 290, "noisy", 3)).save(buf, "JPEG", quality=88, subsampling=2, restart_marker_blocks=3)
    out = torch.empty((224, 224, 3), dtype=torch.uint8, device="cuda")
    decoder.decode_resize(buf.getvalue(), out)
    assert np.array_equal(out.cpu().numpy(), _pil_read_img(buf.getvalue()))


def test_back_to_back_frames_of_different_sizes(decoder):
    """The decoder's staging slots and resampling tables are reused across calls without a synchronisation by the
    caller: decode many frames of alternating sizes, check them all at the end."""
    cases = [(360, 640, 96, 2, "noisy"), (240, 426, 93, 2, "smooth"), (360, 640, 90, 2, "smooth"), (120, 160, 98, 2, "noisy")] * 3
    datas = [jpeg_bytes(*c, seed=i) for i, c in enumerate(cases)]
    out = torch.empty((len(datas), 224, 224, 3), dtype=torch.uint8, device="cuda")
    for i, d in enumerate(datas):
        decoder.decode_resize(d, out[i])
    got = out.cpu().numpy()
    for i, d in enumerate(datas):
        assert np.array_equal(got[i], _pil_read_img(d)), i


@pytest.mark.parametrize("shape", [(360, 640), (224, 640), (500, 333), (37, 53), (224, 224)])
def test_resize_alone_is_bit_exact_with_pillow(decoder, shape):
    from PIL import Image
    img = synthetic_image(*shape, "noisy", seed=shape[1])
    out = torch.empty((224, 224, 3), dtype=torch.uint8, device="cuda")
    decoder.resize(torch.from_numpy(img).cuda(), out)
    assert np.array_equal(out.cpu().numpy(), np.array(Image.fromarray(img).resize((224, 224))))


def test_unsupported_files_are_refused_loudly(decoder):
    from PIL import Image
    from vidsitu_b200.lib import VsbError
    out = torch.empty((224, 224, 3), dtype=torch.uint8, device="cuda")
    buf = io.BytesIO()
    Image.fromarray(synthetic_image(64, 64, "noisy", 4)).save(buf, "JPEG", quality=88, progressive=True)
    with pytest.raises(VsbError, match="baseline"):
        decoder.decode_resize(buf.getvalue(), out)
    with pytest.raises(VsbError, match="not a JPEG"):
        decoder.decode_resize(b"\x89PNG\r\n\x1a\n" + bytes(64), out)
    good = jpeg_bytes(64, 80, 95, 2, "smooth")
    with pytest.raises(VsbError):
        decoder.decode_resize(good[: len(good) // 3], out)          # truncated inside the headers / tables
    with pytest.raises(VsbError, match="decoder was created for"):
        from vidsitu_b200.jpeg import JpegDecoder
        JpegDecoder(32, 32).decode_resize(good, out)


def test_feature_dump_cli_with_gpu_decode_writes_the_same_files(tmp_path):
    """tools/extract_features.py --gpu-decode (JPEGs decoded and resized on the device by DeviceVideoLoader) writes
    byte-identical .npy files to the PIL-in-DataLoader path: the decoded frames are bit-equal, so are the features."""
    import json
    import os
    import sys
    from PIL import Image
    from common import ROOT, synthetic_image
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import extract_features as X
    from vidsitu_b200 import frames_io as F
    names = ["v_a_seg_0", "v_b_seg_1", "v_c_seg_2"]
    for vi, v in enumerate(names):
        d = tmp_path / "frames" / v
        d.mkdir(parents=True)
        for ix in range(1, 301):
            img = synthetic_image(90, 160, "noisy" if ix % 2 else "smooth", seed=vi * 1000 + ix)
            Image.fromarray(img).save(d / f"{v}_{ix:06d}.jpg", quality=94, subsampling=2)
    (tmp_path / "split.json").write_text(json.dumps(names))
    outs = {}
    for mode in ("host", "device", "hybrid"):
        torch.manual_seed(33)
        rc = X.main(["--frames-dir", str(tmp_path / "frames"), "--split-file", str(tmp_path / "split.json"),
                     "--out-dir", str(tmp_path / f"feats_{mode}"), "--mdl-name-used", "m", "--crop", "64",
                     "--videos-per-batch", "2", "--workers", "0" if mode == "host" else "3"]
                    + (["--gpu-decode", mode] if mode != "host" else []))
        assert rc == 0
        outs[mode] = {v: (tmp_path / f"feats_{mode}" / "m" / f"{v}_feats.npy").read_bytes() for v in names}
    assert outs["host"] == outs["device"] == outs["hybrid"]
    # and the loader's frames themselves equal the host reader's
    needed = F.needed_frames(32, 2)
    ref = F.load_video(tmp_path / "frames", names[1], needed, 64)
    for mode in ("device", "hybrid"):
        dl = F.DeviceVideoLoader(tmp_path / "frames", names, 32, 2, size=64, videos_per_batch=3, workers=2, mode=mode)
        (frames, idxs), = list(dl)
        assert idxs == [0, 1, 2] and dl.host_fallbacks == 0
        assert torch.equal(frames[1].cpu(), ref), mode
    # a progressive file in the middle of a video goes through the reference's reader, and is counted
    Image.fromarray(synthetic_image(90, 160, "noisy", 5)).save(tmp_path / "frames" / names[0] / f"{names[0]}_{needed[3] + 1:06d}.jpg",
                                                              quality=90, progressive=True)
    ref0 = F.load_video(tmp_path / "frames", names[0], needed, 64)
    for mode in ("device", "hybrid"):
        dl = F.DeviceVideoLoader(tmp_path / "frames", names[:1], 32, 2, size=64, videos_per_batch=1, workers=2, mode=mode)
        (frames, _), = list(dl)
        assert dl.host_fallbacks == 1 and torch.equal(frames[0].cpu(), ref0), mode


def test_fully_on_device_batch_decode_is_bit_exact_and_reports_refusals():
    """vsb_jpeg_batch_*: Huffman on the GPU too (one warp per frame).  A mixed batch - every sampling mode, odd sizes,
    grayscale, restart markers, plus a progressive file, a truncated file and garbage - against Pillow; the refused
    ones are reported and leave their outputs untouched."""
    from PIL import Image
    from vidsitu_b200.jpeg import JpegBatchDecoder
    datas = [jpeg_bytes(*c, seed=i) for i, c in enumerate(JPEG_CASES)]
    datas.append(jpeg_bytes(40, 56, 90, 2, "noisy", gray=True))
    buf = io.BytesIO()
    Image.fromarray(synthetic_image(170, 290, "noisy", 3)).save(buf, "JPEG", quality=88, subsampling=2, restart_marker_blocks=3)
    datas.append(buf.getvalue())
    n_good = len(datas)
    buf = io.BytesIO()
    Image.fromarray(synthetic_image(64, 64, "noisy", 4)).save(buf, "JPEG", quality=88, progressive=True)
    datas += [buf.getvalue(), datas[4][: len(datas[4]) // 2], b"GIF89a" + bytes(100)]
    out = torch.full((len(datas), 224, 224, 3), 7, dtype=torch.uint8, device="cuda")
    dec = JpegBatchDecoder()
    ok = dec.decode_resize(datas, list(out))
    assert ok == [True] * n_good + [False] * 3
    got = out.cpu().numpy()
    for i in range(n_good):
        assert np.array_equal(got[i], _pil_read_img(datas[i])), i
    assert (got[n_good:] == 7).all()
    # a second, larger batch through the same handle (workspaces grow), frames of one video
    datas = [jpeg_bytes(360, 640, 95, 2, "noisy" if i % 3 else "smooth", seed=100 + i) for i in range(40)]
    out = torch.empty((40, 224, 224, 3), dtype=torch.uint8, device="cuda")
    assert all(dec.decode_resize(datas, list(out)))
    got = out.cpu().numpy()
    for i in (0, 7, 39):
        assert np.array_equal(got[i], _pil_read_img(datas[i])), i


def test_corrupt_scan_data_never_crashes_the_device_decoder():
    """Random byte damage inside the entropy-coded segment: every file either decodes or is reported as refused, the
    device never faults, and a clean batch through the same handle afterwards is still bit-exact."""
    from vidsitu_b200.jpeg import JpegBatchDecoder, JpegDecoder
    rng = np.random.default_rng(11)
    base = [jpeg_bytes(*c, seed=50 + i) for i, c in enumerate(JPEG_CASES[:6])]
    damaged = []
    for i in range(48):
        d = bytearray(base[i % len(base)])
        lo = len(d) // 3                      # behind the headers
        for _ in range(int(rng.integers(1, 12))):
            d[int(rng.integers(lo, len(d) - 2))] = int(rng.integers(0, 256))
        if i % 5 == 0:
            del d[int(rng.integers(lo, len(d))):]     # and a truncation
        damaged.append(bytes(d))
    out = torch.zeros((len(damaged), 64, 64, 3), dtype=torch.uint8, device="cuda")
    dec = JpegBatchDecoder()
    ok = dec.decode_resize(damaged, list(out))
    torch.cuda.synchronize()
    assert len(ok) == len(damaged) and any(ok) and not all(ok)
    hyb = JpegDecoder(640, 640)
    from vidsitu_b200.lib import VsbError
    for i, d in enumerate(damaged[:16]):      # the hybrid path agrees on which files are decodable, and on their pixels
        o = torch.zeros((64, 64, 3), dtype=torch.uint8, device="cuda")
        try:
            hyb.decode_resize(d, o)
            good = True
        except VsbError:
            good = False
        assert good == ok[i], i
        if good:
            assert torch.equal(o, out[i]), i
    clean = torch.empty((len(base), 224, 224, 3), dtype=torch.uint8, device="cuda")
    assert all(dec.decode_resize(base, list(clean)))
    got = clean.cpu().numpy()
    for i, d in enumerate(base):
        assert np.array_equal(got[i], _pil_read_img(d)), i
