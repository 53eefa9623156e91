"""CPU, world_size 3 over gloo: the data-path exchange of the split-micrograph flow (deepemia_b200/distributed.py) — ONE size
all-gather + ONE byte-packed payload all-gather of bit-packed instance records — reproduces the rank-major concatenation on
every rank, including an empty rank, ragged sizes and the per-class list lengths that travel with the sizes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_instance_set(masks, scores, classes):
    """Host-side packing of byte masks into the InstanceSet layout (test helper; the product packs on the GPU)."""
    from deepemia_b200 import engine
    n = len(masks)
    H, W = masks[0].shape if n else (64, 64)
    meta = np.zeros((n, 8), np.int32); bbox = np.full((n, 4), -1, np.int32); area = np.zeros(n, np.int32)
    words, off = [], [0]
    for i, m in enumerate(masks):
        ys, xs = np.nonzero(m)
        if len(ys):
            y0, y1, x0, x1 = ys.min(), ys.max(), xs.min(), xs.max()
            wc0, wc1 = x0 >> 5, x1 >> 5
            meta[i] = [y0, wc0, y1 - y0 + 1, wc1 - wc0 + 1, x0, x1 + 1, 1, 0]
            bbox[i] = [y0, x0, y1, x1]; area[i] = len(ys)
            sub = np.zeros((y1 - y0 + 1, (wc1 - wc0 + 1) * 32), np.uint8)
            sub[:, :min(W, (wc1 + 1) * 32) - wc0 * 32] = m[y0:y1 + 1, wc0 * 32:(wc1 + 1) * 32]
            w = np.packbits(sub.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)
            words.append(w.view(np.int32))
        else:
            meta[i, 6] = 1
        off.append(off[-1] + (len(words[-1]) if len(ys) else 0))
    crops = np.concatenate(words) if words else np.zeros(1, np.int32)
    t = torch.as_tensor
    return engine.InstanceSet(n=n, H=H, W=W, meta=t(meta), crop_off=t(np.array(off, np.int64)), crops=t(crops), bbox=t(bbox), area=t(area),
                              scores=t(np.asarray(scores, np.float32)), classes=t(np.asarray(classes, np.int32)), total_crop_words=off[-1])


def _rank_data(rank):
    rng = np.random.default_rng(100 + rank)
    n = [5, 0, 9][rank]
    masks = []
    for _ in range(n):
        m = np.zeros((64, 96), np.uint8)
        y, x = rng.integers(0, 40), rng.integers(0, 60)
        m[y:y + rng.integers(3, 20), x:x + rng.integers(3, 34)] = 1
        masks.append(m)
    return masks, rng.random(n).astype(np.float32), rng.integers(0, 2, n).astype(np.int32)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from deepemia_b200 import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    masks, scores, classes = _rank_data(rank)
    iset = _cpu_instance_set(masks, scores, classes) if masks else _cpu_instance_set([], [], [])
    iset.H, iset.W = 64, 96
    calls = {"n": 0}
    real = dist.all_gather

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)
    dist.all_gather = counting
    allset, sizes = D.all_gather_packed(iset, extra=[rank + 1, 7])
    dist.all_gather = real
    out.put((rank, allset.n, allset.meta.numpy(), allset.crop_off.numpy(), allset.crops[:allset.total_crop_words].numpy(),
             allset.scores.numpy(), allset.classes.numpy(), allset.bbox.numpy(), sizes.tolist(), calls["n"], allset.area.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_variable_length_all_gather_three_ranks():
    sys.path.insert(0, ROOT)
    from deepemia_b200 import distributed as D, engine
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parts = [_cpu_instance_set(*_rank_data(r)) for r in range(world) if len(_rank_data(r)[0])]
    cat = lambda name: np.concatenate([getattr(p, name).numpy()[:p.n] for p in parts])
    ref_crops = np.concatenate([p.crops.numpy()[:p.total_crop_words] for p in parts])
    bases = np.concatenate([[0], np.cumsum([p.total_crop_words for p in parts])])
    ref_off = np.concatenate([p.crop_off.numpy()[:p.n] + b for p, b in zip(parts, bases)] + [[bases[-1]]])
    for rank, n, meta, crop_off, crops, scores, classes, bbox, sizes, n_coll, area in results:
        assert [s[0] for s in sizes] == [5, 0, 9] and n == 14 and n_coll == 2          # exactly two collectives
        assert [s[2:] for s in sizes] == [[1, 7], [2, 7], [3, 7]]
        assert np.array_equal(meta, cat("meta")) and np.array_equal(crop_off, ref_off)
        assert np.array_equal(crops, ref_crops)
        assert np.array_equal(scores, cat("scores")) and np.array_equal(classes, cat("classes"))
        assert np.array_equal(bbox, cat("bbox")) and np.array_equal(area, cat("area"))
    # contiguous bands: rank-major order is the global tile order
    assert [D.band_of_rank(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    assert [D.band_of_rank(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
