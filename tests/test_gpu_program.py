"""Clip programs (C ABI v7, include/vidsitu_b200.h: vsb_program_*): the whole forward as one C handle.

The program path is checked for BIT equality against the Python launch loop (same plans, same kernels, same
order: any difference is a recording / relocation bug), through three hosts: the engine's own replay, a program
file loaded back into this process, and a C host process (examples/run_program.c) that never sees Python."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from common import ROOT, build_model, synthetic_frames

pytestmark = pytest.mark.gpu

CASES = [("slow_fast_nl_r50_8x8", 32), ("i3d_r50_nl_8x8", 8), ("i3d_r50_8x8", 8)]


def _engine(name: str, n: int = 2, crop: int = 64, seed: int = 3, precision: str = "bf16"):
    model, cfg, _ = build_model(name, seed=seed, crop=crop, precision=precision)
    model = model.cuda()
    eng = model._engine(n, torch.device("cuda"))
    frames = synthetic_frames(n, cfg.sf_mdl.DATA.NUM_FRAMES, crop, seed=seed + 100).cuda()
    return model, eng, frames


def _python_loop(eng, frames):
    eng.load_frames(frames)
    eng.run()                       # launch by launch from Python
    torch.cuda.synchronize()
    return eng.feats.clone(), eng.logits.clone()


@pytest.mark.parametrize("name,t", CASES)
def test_program_replay_is_bit_equal_to_the_python_launch_loop(name, t):
    model, eng, frames = _engine(name)
    feats, logits = _python_loop(eng, frames)
    assert eng.replay_mode == "program"
    eng.feats.zero_(), eng.logits.zero_()
    eng.replay()                    # vsb_program_run on the graph captured inside the library
    eng.replay()
    torch.cuda.synchronize()
    assert eng._programs is not None and eng._programs[0].num_launches == eng.num_launches
    assert torch.equal(eng.feats, feats) and torch.equal(eng.logits, logits)
    # an un-captured program (plain launches on the two lanes) gives the same bits
    prog = eng.build_program()
    eng.feats.zero_()
    prog.run()
    torch.cuda.synchronize()
    assert torch.equal(eng.feats, feats)


def test_program_replay_matches_torch_graph_replay_fp32():
    model, eng, frames = _engine("slow_fast_nl_r50_8x8", precision="fp32")
    feats, logits = _python_loop(eng, frames)
    eng.replay()
    torch.cuda.synchronize()
    assert torch.equal(eng.feats, feats) and torch.equal(eng.logits, logits)
    eng.replay_mode = "torch"
    eng.feats.zero_()
    eng.replay()
    torch.cuda.synchronize()
    assert torch.equal(eng.feats, feats)


@pytest.mark.parametrize("name,t", CASES[:2])
def test_saved_program_reloads_and_runs_bit_equal(name, t, tmp_path):
    from cuda import cudart
    from vidsitu_b200 import ops

    model, eng, frames = _engine(name)
    feats, logits = _python_loop(eng, frames)
    path = str(tmp_path / "m.vsbprog")
    saved = eng.export_program(path)
    assert os.path.getsize(path) > 1 << 20
    prog = ops.Program.load(path)
    assert prog.num_launches == saved.num_launches == eng.num_launches + len(eng.inputs)   # + the pack launches
    ptr, nbytes = prog.region("frames")
    assert nbytes == frames.numel()
    torch.cuda.synchronize()
    (err,) = cudart.cudaMemcpy(ptr, frames.data_ptr(), nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
    assert int(err) == 0
    prog.run()
    torch.cuda.synchronize()
    for region, ref in (("feats", feats), ("logits", logits)):
        ptr, nbytes = prog.region(region)
        got = torch.empty_like(ref)
        assert nbytes == got.numel() * 4
        (err,) = cudart.cudaMemcpy(got.data_ptr(), ptr, nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
        assert int(err) == 0
        torch.cuda.synchronize()
        assert torch.equal(got, ref), region
    # a program file is refused by a library of another ABI / a truncated file is refused
    blob = open(path, "rb").read()
    open(path, "wb").write(blob[: len(blob) // 2])
    with pytest.raises(Exception, match="truncated|not a vidsitu_b200 program"):
        ops.Program.load(path)


def test_c_host_runs_a_saved_program_without_python(tmp_path):
    """examples/run_program.c: gcc + include/vidsitu_b200.h + libvidsitu_b200.so + cudart, nothing else."""
    model, eng, frames = _engine("slow_fast_nl_r50_8x8")
    feats, logits = _python_loop(eng, frames)
    prog_path, frames_path = str(tmp_path / "m.vsbprog"), str(tmp_path / "frames.u8")
    eng.export_program(prog_path)
    frames.cpu().numpy().tofile(frames_path)
    exe = str(tmp_path / "run_program")
    libdir = os.path.join(ROOT, "vidsitu_b200")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                    os.path.join(ROOT, "examples", "run_program.c"), "-o", exe, "-L", libdir, "-lvidsitu_b200",
                    "-L", "/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{libdir}"], check=True)
    out_f, out_l = str(tmp_path / "feats.f32"), str(tmp_path / "logits.f32")
    r = subprocess.run([exe, prog_path, frames_path, out_f, out_l], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got_f = np.fromfile(out_f, dtype=np.float32).reshape(feats.shape)
    got_l = np.fromfile(out_l, dtype=np.float32).reshape(logits.shape)
    assert np.array_equal(got_f, feats.cpu().numpy())
    assert np.array_equal(got_l, logits.cpu().numpy())
