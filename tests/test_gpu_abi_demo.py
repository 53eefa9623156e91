"""The C ABI is self-sufficient: a plain C++ host program (tests/abi_demo/abi_demo.cpp: include/emia.h + the CUDA runtime, no
Python, no torch) runs the fused path and must produce exactly what the Python host layer produces on the same input."""
import os
import subprocess

import numpy as np
import pytest
import torch

from deepemia_b200 import engine, synthetic as syn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "abi_demo", "abi_demo")


def test_c_host_program_matches_python_host_layer(cuda_device, tmp_path):
    if not os.path.exists(EXE):
        import __graft_entry__ as ge
        ge.build()
    assert os.path.exists(EXE), "tests/abi_demo/abi_demo was not built by __graft_entry__.build()"
    H, W = 384, 416
    tiles = [syn.synthetic_heads(9000 + t, 60 + 7 * t, H, W, duplicate_frac=0.3, rmin=6, rmax=22, margin=25) for t in range(5)]
    probs = np.concatenate([t[0] for t in tiles]); boxes = np.concatenate([t[1] for t in tiles])
    scores = np.concatenate([t[2] for t in tiles]); classes = np.concatenate([t[3] for t in tiles]).astype(np.int32)
    offs = np.concatenate([[0], np.cumsum([len(t[0]) for t in tiles])]).astype(np.int32)
    n, T = len(probs), len(tiles)
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        np.array([T, H, W, n], np.int32).tofile(f); offs.tofile(f)
        probs.astype(np.float32).tofile(f); boxes.astype(np.float32).tofile(f); scores.astype(np.float32).tofile(f); classes.tofile(f)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "deepemia_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([EXE, str(inp), str(out)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(out, "rb").read()
    klen = np.frombuffer(raw, np.int32, T, 0)
    kidx = np.frombuffer(raw, np.int32, n, 4 * T)
    n_rec = int(np.frombuffer(raw, np.int64, 1, 4 * (T + n))[0])
    rec = np.frombuffer(raw, np.float64, n_rec * 16, 4 * (T + n) + 8).reshape(n_rec, 16)
    rinst = np.frombuffer(raw, np.int32, n_rec, 4 * (T + n) + 8 + 8 * 16 * n_rec)
    t = [torch.as_tensor(np.ascontiguousarray(a), device=cuda_device) for a in (probs, boxes, scores, classes)]
    iset, kept, meas = engine.run_tiles(*t, offs, H, W, um_pix=0.5, rules=syn.POLYHIPES_RULES, dedup_iou=0.7)
    kl = kept.to_lists()
    assert [kidx[offs[g]:offs[g] + klen[g]].tolist() for g in range(T)] == kl
    assert n_rec == meas.n_records
    assert np.array_equal(rec, meas.records.cpu().numpy()) and np.array_equal(rinst, meas.rec_inst.cpu().numpy())
