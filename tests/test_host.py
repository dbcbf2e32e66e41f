"""CPU-side tests: config / spec / state_dict contract, weight preparation, the C-ABI
library's exports, sharding logic (gloo, world_size 2).  No compute call needs a GPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

from common import ROOT, build_model
from vidsitu_b200 import lib as L
from vidsitu_b200.arch import build_spec
from vidsitu_b200.config import SF_MDL_PRESETS, make_cfg, make_comm
from vidsitu_b200.weights import fold_bn, pack_conv_weight, stem_quad_weight


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.load()
    assert lib.vsb_abi_version() == 8
    declared = set()
    for hdr in ("vidsitu_b200.h", "vidsitu_b200_debug.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"\b(vsb_[a-z0-9_]+)\s*\(", src))
    assert len(declared) >= 15
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"


def test_compute_entry_points_fail_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("only meaningful on a machine without a CUDA device")
    from vidsitu_b200 import ops
    from vidsitu_b200.lib import VsbError
    model, cfg, _ = build_model("slow_fast_nl_r50_8x8", crop=64)
    with pytest.raises(VsbError):
        model.extract_features(torch.zeros((1, 32, 64, 64, 3), dtype=torch.uint8))
    with pytest.raises(VsbError):
        ops.linear(torch.zeros(2, 4), torch.zeros(3, 4), None, torch.zeros(2, 3), False)
    # a device-less plan_create must be an error status with a message, not a crash
    d = L.ConvDesc()
    h = ctypes.c_void_p()
    assert L.load().vsb_conv3d_plan_create(ctypes.byref(d), ctypes.byref(h)) != 0
    assert L.load().vsb_last_error()


@pytest.mark.parametrize("name,nconv,dims", [
    ("slow_fast_nl_r50_8x8", 110, [2048, 256]), ("i3d_r50_8x8", 53, [2048]), ("i3d_r50_nl_8x8", 73, [2048]),
    ("slow_fast_r101_16x8", 212, [2048, 256]), ("slow_nl_r50_8x8", 53, [2048]), ("c2d_r50_8x8", 53, [2048])])
def test_spec_matches_survey_tables(name, nconv, dims):
    spec = build_spec(make_cfg(name).sf_mdl)
    assert len(spec.all_convs()) == nconv
    assert spec.feat_dims == dims


def test_sf50_conv_flops_match_survey():
    """100.615 conv GFLOP per SlowFast-R50 8x8 clip at 224x224."""
    spec = build_spec(make_cfg("slow_fast_nl_r50_8x8").sf_mdl)
    from oracle.flops import conv_gflop_per_clip
    assert abs(conv_gflop_per_clip(spec) - 100.615) < 0.01
    assert abs(conv_gflop_per_clip(build_spec(make_cfg("i3d_r50_8x8").sf_mdl)) - 56.814) < 0.01
    assert abs(conv_gflop_per_clip(build_spec(make_cfg("slow_fast_r101_16x8").sf_mdl)) - 326.187) < 0.01


def test_state_dict_key_contract():
    model, _, _ = build_model("slow_fast_nl_r50_8x8", crop=64)
    sd = model.state_dict()
    assert len(sd) == 666        # 662 sf_mdl.* + 4 proj_head.* 
    for k, shape in {
        "sf_mdl.s1.pathway0_stem.conv.weight": (64, 3, 1, 7, 7),
        "sf_mdl.s1.pathway1_stem.conv.weight": (8, 3, 5, 7, 7),
        "sf_mdl.s1_fuse.conv_f2s.weight": (16, 8, 7, 1, 1),
        "sf_mdl.s2.pathway0_res0.branch1.weight": (256, 80, 1, 1, 1),
        "sf_mdl.s4.pathway0_res0.branch2.a.weight": (256, 640, 3, 1, 1),
        "sf_mdl.s5.pathway1_res2.branch2.c_bn.running_var": (256,),
        "sf_mdl.head.projection.weight": (400, 2304),
        "proj_head.0.weight": (1152, 2304),
        "proj_head.2.bias": (1560,),
    }.items():
        assert tuple(sd[k].shape) == shape, k
    nl, _, _ = build_model("i3d_r50_nl_8x8", crop=64)
    keys = nl.state_dict().keys()
    assert "sf_mdl.s3.pathway0_nonlocal1.conv_theta.bias" in keys
    assert "sf_mdl.s4.pathway0_nonlocal5.bn.running_mean" in keys
    assert "sf_mdl.s3.pathway0_nonlocal0.conv_theta.weight" not in keys


def test_module_prefix_checkpoints_load():
    """feat_extractor.py:147-153 strips a DDP 'module.' prefix before load_state_dict."""
    model, _, _ = build_model("i3d_r50_8x8", crop=64)
    ddp = {"module." + k: v for k, v in model.state_dict().items()}
    stripped = {k[len("module."):]: v for k, v in ddp.items()}
    missing, unexpected = model.load_state_dict(stripped, strict=True)
    assert not missing and not unexpected


def test_training_mode_is_refused():
    model, _, _ = build_model("i3d_r50_8x8", crop=64)
    with pytest.raises(NotImplementedError):
        model.train()


def test_unknown_config_values_are_rejected():
    cfg = make_cfg("slow_fast_nl_r50_8x8")
    cfg.sf_mdl.BN.NORM_TYPE = "layernorm"              # get_norm raises for anything but the three BN flavours
    with pytest.raises(NotImplementedError):
        build_spec(cfg.sf_mdl)
    cfg.sf_mdl.BN.NORM_TYPE = "sync_batchnorm"         # NaiveSyncBatchNorm3d: a BatchNorm3d in eval mode
    assert build_spec(cfg.sf_mdl).norm_type == "sync_batchnorm"
    cfg = make_cfg("i3d_r50_8x8")
    cfg.sf_mdl.MODEL.ARCH = "x3d"
    with pytest.raises(NotImplementedError):
        build_spec(cfg.sf_mdl)
    with pytest.raises(KeyError):
        make_cfg("nope")
    assert set(SF_MDL_PRESETS) >= {"slow_fast_nl_r50_8x8", "slow_nl_r50_8x8", "c2d_r50_8x8", "i3d_r50_8x8",
                                   "i3d_r50_nl_8x8"}
    assert make_comm(make_cfg("i3d_r50_8x8").sf_mdl).path_type == "single"


def test_weight_packing_and_bn_fold():
    torch.manual_seed(0)
    w = torch.randn(24, 10, 3, 1, 1)
    p = pack_conv_weight(w, 16, 32, torch.float32)
    assert p.shape == (32, 3, 16) and float(p[24:].abs().max()) == 0 and float(p[:, :, 10:].abs().max()) == 0
    assert torch.equal(p[:24, :, :10].permute(0, 2, 1).reshape(24, 10, 3, 1, 1), w)
    g, b, m, v = torch.rand(8) + 0.5, torch.randn(8), torch.randn(8), torch.rand(8) + 0.5
    s, o = fold_bn(g, b, m, v, 1e-5, 16)
    x = torch.randn(4, 8, 2, 3, 3)
    ref = F.batch_norm(x, m, v, g, b, training=False, eps=1e-5)
    got = x * s[:8].view(1, -1, 1, 1, 1) + o[:8].view(1, -1, 1, 1, 1)
    assert torch.allclose(got, ref, atol=1e-5) and float(s[8:].abs().max()) == 0


@pytest.mark.parametrize("kt,cout", [(1, 64), (5, 8)])
def test_stem_quad_view_is_the_same_convolution(kt, cout):
    torch.manual_seed(1)
    n, t, h, w = 1, 3, 16, 16
    x = torch.randn(n, t, h, w, 3)
    wt = torch.randn(cout, 3, kt, 7, 7)
    ref = F.conv3d(x.permute(0, 4, 1, 2, 3), wt, stride=(1, 2, 2), padding=(kt // 2, 3, 3)).permute(0, 2, 3, 4, 1)
    x4 = torch.zeros(n, t, h, w, 4)
    x4[..., :3] = x
    wq = stem_quad_weight(wt, torch.float32).view(2 * cout, kt, 7, 3, 16).permute(0, 4, 1, 2, 3)
    y = F.conv3d(x4.view(n, t, h, w // 4, 16).permute(0, 4, 1, 2, 3), wq, stride=(1, 2, 1), padding=(kt // 2, 3, 1))
    y = y.permute(0, 2, 3, 4, 1).reshape(n, t, h // 2, w // 2, cout)
    assert torch.allclose(y, ref, atol=1e-4)


def test_shard_ranges_partition_the_videos():
    from vidsitu_b200.dist import shard_counts, shard_range
    for n in (0, 1, 7, 8, 1326):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(shard_counts(n, world)) - min(shard_counts(n, world)) <= 1


def test_gather_rows_world2_gloo(tmp_path):
    """N>1 host path: contiguous video shards + one all-gather, run as 2 gloo ranks on CPU."""
    script = tmp_path / "w.py"
    script.write_text(f"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {ROOT!r})
from vidsitu_b200.dist import shard_range, gather_rows
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
n_vid = 7
full = torch.arange(n_vid * 5 * 3, dtype=torch.float32).view(n_vid * 5, 3)
s, e = shard_range(n_vid, r, w)
out = gather_rows(full[s * 5:e * 5].clone(), n_vid * 5)
assert torch.equal(out, full), (r, out)
dist.destroy_process_group()
print('ok', r)
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         capture_output=True, text=True, timeout=240, env=env)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2


def test_event_windows_match_the_oracle_and_the_reference_facts():
    """vidsitu_b200.events (product) vs oracle.event_frame_indices (the reference restatement,
    dat_loader.py:69-79,454-472 + utils/video_utils.py:18-38), plus the SURVEY 8(a1) facts."""
    from oracle import sf_oracle as O
    from vidsitu_b200 import events as E
    assert E.event_centers(30) == [30, 90, 150, 210, 270]
    for nf, rate in ((32, 2), (8, 8), (64, 2), (16, 4)):
        for fps, total in ((30, 300), (30, 120), (25, 300), (30, 1)):
            assert E.event_frame_indices(nf, rate, fps, total) == O.event_frame_indices(nf, rate, fps, total)
    ev = E.event_frame_indices(32, 2)
    assert ev[0][:3] == [0, 0, 2] and ev[0][-1] == 60 and ev[4][-2:] == [298, 299] and all(len(e) == 32 for e in ev)
    assert E.event_frame_indices(8, 8)[0] == [0, 6, 14, 22, 30, 38, 46, 54]
    assert E.event_frame_indices(64, 2)[0][:18] == [0] * 18


def test_tuning_table_is_well_formed():
    """vidsitu_b200/tune_table.json (tools/autotune.py): shape signatures -> plan knobs that never change results."""
    import json
    import re
    from vidsitu_b200.engine import _PLAN_KNOBS
    p = os.path.join(ROOT, "vidsitu_b200", "tune_table.json")
    tab = json.load(open(p))
    assert tab["entries"] and all(m["gpu"] == "NVIDIA B200" for m in tab["meta"])
    sig = re.compile(r"^\d+>\d+\|k\d{3}\|s\d{3}\|o\d+x\d+x\d+\|p\d+>\d+\|r[01]\|x\d+$")
    for k, knobs in tab["entries"].items():
        assert sig.match(k), k
        assert knobs and set(knobs) <= set(_PLAN_KNOBS) | {"algo", "win_group"}, (k, knobs)
        assert knobs.get("epi_n", 32) in (16, 32, 64) and knobs.get("epi_bufs", 2) in (2, 3, 4)


def test_ctypes_conv_desc_matches_the_c_header(tmp_path):
    """The ctypes mirror of vsb_conv_desc (vidsitu_b200/lib.py) against the C header, compiled with gcc: total
    size and the offsets of fields spread over the struct (a silent mismatch would corrupt every plan)."""
    import ctypes as C
    import subprocess
    fields = ["dtype", "wgt", "cout", "scale", "out_pitch", "block_n", "algo", "kw_c_hi", "in2", "sw2", "epi_n",
              "epi_bufs", "flags", "out_f16", "wgt_clip_rows", "tile_wait", "grid_limit", "in_f16"]
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vidsitu_b200.h"\nint main(void){'
                   'printf("%zu", sizeof(vsb_conv_desc));'
                   + "".join(f'printf(" %zu", offsetof(vsb_conv_desc, {f}));' for f in fields) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got[0] == C.sizeof(L.ConvDesc)
    for f, off in zip(fields, got[1:]):
        assert getattr(L.ConvDesc, f).offset == off, f


def test_sub_batchnorm_layout_and_aggregation():
    """BN.NORM_TYPE = sub_batchnorm (SubBatchNorm3d, batchnorm_helper.py:37-109): the state_dict key set the real
    reference accepted with strict=True when the fixture was made (996 keys), the aggregation formula, and the
    key lookup the engine folds from."""
    import json
    from vidsitu_b200.model import SubBNParams
    from vidsitu_b200.weights import bn_tensor_keys
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_meta.json")))["sf50_subbn_n2_64"]
    model, cfg, _ = build_model("slow_fast_nl_r50_8x8", seed=meta["seed"], crop=64, sf_overrides=meta["sf_overrides"])
    sd = model.state_dict()
    assert len(sd) == meta["num_state_dict_keys"] == 996
    assert "sf_mdl.s2.pathway0_res0.branch2.a_bn.bn.running_var" in sd
    assert sd["sf_mdl.s2.pathway0_res0.branch2.a_bn.split_bn.running_mean"].numel() == 2 * 64
    bn = SubBNParams(6, 3)
    g = torch.Generator().manual_seed(0)
    bn.split_bn.running_mean.copy_(torch.randn(18, generator=g))
    bn.split_bn.running_var.copy_(torch.rand(18, generator=g) + 0.5)
    bn.aggregate_stats()
    m, v = bn.split_bn.running_mean.view(3, 6), bn.split_bn.running_var.view(3, 6)
    assert torch.allclose(bn.bn.running_mean, m.mean(0))
    assert torch.allclose(bn.bn.running_var, v.mean(0) + ((m - m.mean(0)) ** 2).mean(0))
    t = {k[len("sf_mdl."):]: v for k, v in sd.items() if k.startswith("sf_mdl.")}
    assert bn_tensor_keys(t, "s1.pathway0_stem.bn")[2:] == ("s1.pathway0_stem.bn.bn.running_mean",
                                                          "s1.pathway0_stem.bn.bn.running_var")
    plain, _, _ = build_model("slow_fast_nl_r50_8x8", seed=0, crop=64)
    tp = {k[len("sf_mdl."):]: v for k, v in plain.state_dict().items() if k.startswith("sf_mdl.")}
    assert bn_tensor_keys(tp, "s1.pathway0_stem.bn")[2] == "s1.pathway0_stem.bn.running_mean"


def test_ctypes_bottleneck_desc_matches_the_c_header(tmp_path):
    """ctypes mirror of vsb_bottleneck_desc against the C header (gcc): size and every field offset."""
    import ctypes as C
    fields = [f for f, _ in L.BottleneckDesc._fields_]
    src = tmp_path / "szb.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vidsitu_b200.h"\nint main(void){'
                   'printf("%zu", sizeof(vsb_bottleneck_desc));'
                   + "".join(f'printf(" %zu", offsetof(vsb_bottleneck_desc, {f}));' for f in fields) + "return 0;}\n")
    exe = tmp_path / "szb"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got[0] == C.sizeof(L.BottleneckDesc)
    for f, off in zip(fields, got[1:]):
        assert getattr(L.BottleneckDesc, f).offset == off, f
    # a device-less plan_create is an error status with a message, never a crash
    d, h = L.BottleneckDesc(), C.c_void_p()
    assert L.load().vsb_bottleneck_plan_create(C.byref(d), C.byref(h)) != 0 and L.load().vsb_last_error()


def test_presets_are_pinned_to_the_reference_yaml_files():
    """config.SF_MDL_PRESETS restates the reference's backbone YAMLs in Python: every hot-path key the YAML sets must
    come out of sf_mdl_cfg() with the YAML's value (tests/golden/yaml_presets.json is generated from the files by
    tests/golden/make_yaml_presets.py)."""
    import json
    from vidsitu_b200.config import sf_mdl_cfg
    pins = json.load(open(os.path.join(ROOT, "tests", "golden", "yaml_presets.json")))
    assert set(pins) == set(SF_MDL_PRESETS)
    unknown = []
    for name, rec in pins.items():
        cfg = sf_mdl_cfg(name)
        for sec, kv in rec["keys"].items():
            for k, v in kv.items():
                if sec not in cfg or k not in cfg[sec]:
                    unknown.append((name, sec, k))     # a key our config surface does not carry: must be irrelevant here
                    continue
                assert cfg[sec][k] == v, (name, rec["file"], sec, k, cfg[sec][k], v)
    # keys the YAMLs set that the hot path never reads (SURVEY.md 8b lists what is read)
    assert {(s, k) for _, s, k in unknown} <= {("MODEL", "LOSS_FUNC"), ("RESNET", "TRANS_FUNC"), ("DATA", "DECODING_BACKEND"),
                                               ("DATA", "PATH_PREFIX"), ("DATA", "PATH_LABEL_SEPARATOR"),
                                               ("DATA", "MULTI_LABEL"), ("DATA", "INV_UNIFORM_SAMPLE"),
                                               ("DATA", "ENSEMBLE_METHOD"), ("MODEL", "HEAD_ACT")} | set(), unknown


def test_spatial_dilation_is_refused_not_silently_wrong():
    """RESNET.SPATIAL_DILATIONS > 1 dilates the 3x3 conv in the reference (resnet_helper.py:196-207); the kernels have
    no dilation, so building such a spec must raise."""
    with pytest.raises(NotImplementedError):
        build_spec(make_cfg("i3d_r50_8x8", RESNET={"SPATIAL_DILATIONS": [[1], [1], [1], [2]]}).sf_mdl)
    with pytest.raises(NotImplementedError):
        build_spec(make_cfg("slow_fast_nl_r50_8x8", RESNET={"SPATIAL_DILATIONS": [[1, 1], [1, 1], [1, 2], [1, 1]]}).sf_mdl)


def test_loading_weights_through_sf_mdl_invalidates_prepared_engines():
    """The reference loads backbone checkpoints with load_checkpoint(model=mdl.sf_mdl, ...): every such path must drop
    the kernels' prepared weight copies."""
    model, _, _ = build_model("i3d_r50_8x8", seed=0, crop=64)
    model._prep[("bf16", "cuda:0")] = {"stale": 1}
    v0 = model._weights_version
    model.sf_mdl.load_state_dict(model.sf_mdl.state_dict())
    assert model._weights_version == v0 + 1 and not model._prep
    model._prep[("bf16", "cuda:0")] = {"stale": 1}
    model.sf_mdl.float()
    assert not model._prep


def test_program_handle_host_side_behaviour(tmp_path):
    """Clip programs (C ABI v7): the parts that need no GPU - create / destroy, counters, error reporting of the
    region lookup and of the file reader (a non-program file is refused with a message, never crashes)."""
    import ctypes as C
    lib = L.load()
    h = C.c_void_p()
    assert lib.vsb_program_create(C.byref(h)) == 0 and h.value
    assert lib.vsb_program_num_ops(h) == 0 and lib.vsb_program_num_launches(h) == 0
    assert lib.vsb_program_device_bytes(h) == 0
    ptr, nbytes = C.c_void_p(), C.c_ulonglong()
    assert lib.vsb_program_region(h, b"feats", C.byref(ptr), C.byref(nbytes)) != 0
    assert b"no region 'feats'" in lib.vsb_last_error()
    assert lib.vsb_program_add_region(h, b"x", None, 16, 0) != 0          # null pointer
    assert lib.vsb_program_add_sync(h, 0, 0) != 0                         # a sync joins two different lanes
    assert lib.vsb_program_add_maxpool3d(h, None, 1, 1, 1, 1, 1, 1, None, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 2, b"p") != 0
    assert b"lane" in lib.vsb_last_error()
    lib.vsb_program_destroy(h)
    bad = tmp_path / "bad.vsbprog"
    bad.write_bytes(b"not a program" * 10)
    assert lib.vsb_program_file_device_bytes(str(bad).encode(), C.byref(nbytes)) != 0
    assert b"not a vidsitu_b200 program file" in lib.vsb_last_error()
    out = C.c_void_p()
    assert lib.vsb_program_load(str(tmp_path / "missing").encode(), None, 0, C.byref(out)) != 0 and not out.value


def test_jpeg_header_parser_needs_no_gpu():
    """vsb_jpeg_info (the host half of the GPU frame ingest): geometry and sampling of the files it accepts, a
    message for the ones it refuses."""
    import io
    from PIL import Image
    from common import jpeg_bytes, synthetic_image
    from vidsitu_b200.jpeg import jpeg_info
    from vidsitu_b200.lib import VsbError
    assert jpeg_info(jpeg_bytes(77, 53, 90, 2, "noisy")) == (53, 77, 3, 2, 2)
    assert jpeg_info(jpeg_bytes(50, 70, 85, 1, "smooth")) == (70, 50, 3, 2, 1)
    assert jpeg_info(jpeg_bytes(48, 64, 75, 0, "noisy")) == (64, 48, 3, 1, 1)
    assert jpeg_info(jpeg_bytes(40, 56, 90, 2, "noisy", gray=True))[:3] == (56, 40, 1)
    buf = io.BytesIO()
    Image.fromarray(synthetic_image(64, 64, "noisy", 4)).save(buf, "JPEG", quality=88, progressive=True)
    with pytest.raises(VsbError, match="baseline"):
        jpeg_info(buf.getvalue())
    with pytest.raises(VsbError, match="not a JPEG"):
        jpeg_info(b"GIF89a" + bytes(32))


@pytest.mark.parametrize("mean,std", [([0.45, 0.45, 0.45], [0.225, 0.225, 0.225]),
                                      ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]),
                                      ([0.5, 0.43216, 0.1], [0.5, 0.22803, 1.0])])
def test_pack_fma_coefficients_reproduce_the_three_fp32_operations(mean, std):
    """The bf16 pack kernel computes bf16(fma(x, A, B)) instead of the reference's x / 255, - mean, / std
    (utils/video_utils.py:147-164) followed by the rounding to bf16: the host search (vsb_debug_pack_fma_coeffs, no GPU)
    must return coefficients that give the SAME bf16 value for every byte value and channel."""
    lib = L.load()
    f3 = ctypes.c_float * 3
    lib.vsb_debug_pack_fma_coeffs.restype = ctypes.c_int
    lib.vsb_debug_pack_fma_coeffs.argtypes = [ctypes.POINTER(ctypes.c_float)] * 4
    a, b = f3(), f3()
    assert lib.vsb_debug_pack_fma_coeffs(f3(*mean), f3(*std), a, b) == 1
    x = torch.arange(256, dtype=torch.float32)
    for c in range(3):
        want = (((x / 255.0) - torch.tensor(mean[c], dtype=torch.float32)) / torch.tensor(std[c], dtype=torch.float32))
        # fma in float64 is exact for an 8-bit x times a 24-bit coefficient; one rounding to fp32, one to bf16
        got = (x.double() * float(a[c]) + float(b[c])).float()
        assert torch.equal(got.to(torch.bfloat16), want.to(torch.bfloat16)), c


@pytest.mark.parametrize("ph,pw", [(56, 56), (24, 40), (8, 8), (16, 24), (64, 8)])   # crop / 4: multiples of 8
def test_fused_stem_seam_zeroing_covers_exactly_the_red_max_pixels(ph, pw):
    """stem_pool_sm100.cu: epilogue_tile writes a pooled pixel with a plain store when one 8 x 16 conv tile holds its whole
    3x3/s2 window and with red.max when two or four tiles contribute (`seam`); zero_seams_kernel gives only the latter their
    initial value.  Replay both index computations: every pooled pixel is either stored exactly once and never touched by
    red.max, or zeroed and only ever combined by red.max (max-pool after ReLU, stem_helper.py:157-178: values >= 0)."""
    tiles_h, tiles_w = ph // 4, pw // 8            # 8 x 16 conv tiles = 4 x 8 pooled pixels (+ the seam row / column)
    stores = [[0] * pw for _ in range(ph)]
    reds = [[0] * pw for _ in range(ph)]
    for th in range(tiles_h):
        for tw in range(tiles_w):
            for pl in range(5):
                for ql in range(9):
                    pg, qg = th * 4 + pl, tw * 8 + ql
                    if pg >= ph or qg >= pw:
                        continue
                    seam = (pl == 0 and th > 0) or pl == 4 or (ql == 0 and tw > 0) or ql == 8
                    (reds if seam else stores)[pg][qg] += 1
    nseam = (pw - 1) >> 3
    zeroed = [[False] * pw for _ in range(ph)]
    for pg in range(ph):                            # one warp per pooled row
        if pg >= 4 and pg % 4 == 0:
            zeroed[pg] = [True] * pw
        else:
            for s in range(nseam):
                zeroed[pg][(s + 1) * 8] = True
    for pg in range(ph):
        for qg in range(pw):
            if zeroed[pg][qg]:
                assert stores[pg][qg] == 0 and reds[pg][qg] in (2, 4), (pg, qg)
            else:
                assert stores[pg][qg] == 1 and reds[pg][qg] == 0, (pg, qg)
